// fz_elem.cu — HBM-bound kernels of the UNet step: GroupNorm (joint-frame statistics), LayerNorm, nearest upsample,
// channel concat, input im2col / output temporal conv, time embedding, temporal attention, CFG + DDIM + latent blend,
// and the cross-attention blend mask.  All activations are fp16 NHWC ([B*F, H*W, C] == token-major), statistics fp32/fp64,
// 16-byte vector loads/stores, grids sized to cover the 148 SMs.
#include "fz_common.cuh"

#include <algorithm>
#include <cmath>

#include "../../include/fatezero_b200.h"

namespace fz {

static inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct alignas(16) Half8 {
  __half v[8];
};

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm.  x: [NB, HW, C]; statistics group sidx = (nb / frames_per_stat) * G + g.
// resnet.py:338,369 and unet_3d_condition.py:439 call nn.GroupNorm on the 5-D tensor => frames_per_stat = F (joint);
// models/attention.py:112 normalises "(b f) c h w" => frames_per_stat = 1.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGnThreads = 256;
constexpr int kGnMaxSlots = 4;

// Statistics pass: every CTA reduces its pixel chunk of one image to per-group partial (sum, sumsq) WITHOUT atomics
// (v0 used ~4k contended shared-memory atomics per CTA): registers -> smem [TY][C] -> per channel -> per group -> partial[nb][chunk][g].
template <int SLOTS>
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const __half* __restrict__ x, int HW, int C, int G, int TX,
                                                             int px_per_cta, float2* __restrict__ partial) {
  constexpr int slots = SLOTS;  // register arrays are sized by the template: 1 slot for C <= 2048 keeps occupancy (bytes in flight) high
  extern __shared__ float gn_smem[];  // [TY][C] sums, [TY][C] sumsq
  const int nb = blockIdx.y;
  const int CV = C / 8;
  const int cpg = C / G;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int TY = kGnThreads / TX;
  float* s_sum = gn_smem;
  float* s_sq = gn_smem + TY * C;
  float acc[SLOTS][8], acc2[SLOTS][8];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s)
#pragma unroll
    for (int e = 0; e < 8; ++e) { acc[s][e] = 0.f; acc2[s][e] = 0.f; }
  const int p0 = blockIdx.x * px_per_cta;
  const int p1 = min(HW, p0 + px_per_cta);
  if (ty < TY) {
    const __half* xb = x + (static_cast<long long>(nb) * HW) * C;
#pragma unroll 4
    for (int p = p0 + ty; p < p1; p += TY) {
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        const int cv = tx + s * TX;
        if (s < slots) {
          const Half8 h = *reinterpret_cast<const Half8*>(xb + static_cast<long long>(p) * C + cv * 8);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float f = __half2float(h.v[e]);
            acc[s][e] += f;
            acc2[s][e] += f * f;
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int cv = tx + s * TX;
      if (s < slots) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s_sum[ty * C + cv * 8 + e] = acc[s][e];
          s_sq[ty * C + cv * 8 + e] = acc2[s][e];
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kGnThreads) {
    float a = 0.f, b = 0.f;
    for (int t = 0; t < TY; ++t) { a += s_sum[t * C + c]; b += s_sq[t * C + c]; }
    s_sum[c] = a;
    s_sq[c] = b;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    float a = 0.f, b = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) { a += s_sum[c]; b += s_sq[c]; }
    partial[(static_cast<long long>(nb) * gridDim.x + blockIdx.x) * G + threadIdx.x] = make_float2(a, b);
  }
}

template <int SLOTS>
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y, int HW, int C, int G,
                                                             int frames_per_stat, int TX, int px_per_cta, int chunks,
                                                             const float2* __restrict__ partial, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int silu) {
  constexpr int slots = SLOTS;
  __shared__ double s_red[kGnThreads][2];
  __shared__ float s_mean[64], s_rstd[64];
  const int nb = blockIdx.y;
  const int CV = C / 8;
  const int cpg = C / G;
  // group statistics of this image's stat set: sum the partials of its frames_per_stat images x chunks (fp64)
  {
    const int first_nb = (nb / frames_per_stat) * frames_per_stat;
    const int n_part = frames_per_stat * chunks;  // partials per group, contiguous per image: [(nb*chunks + chunk)*G + g]
    const int g = threadIdx.x % G, lane_j = threadIdx.x / G, n_j = kGnThreads / G;
    double a = 0.0, b = 0.0;
    if (lane_j < n_j) {
      for (int i = lane_j; i < n_part; i += n_j) {
        const float2 v = partial[(static_cast<long long>(first_nb) * chunks + i) * G + g];
        a += v.x;
        b += v.y;
      }
    }
    s_red[threadIdx.x][0] = a;
    s_red[threadIdx.x][1] = b;
    __syncthreads();
    if (threadIdx.x < G) {
      double sa = 0.0, sb = 0.0;
      for (int j = 0; j < n_j; ++j) { sa += s_red[j * G + threadIdx.x][0]; sb += s_red[j * G + threadIdx.x][1]; }
      const double cnt = static_cast<double>(cpg) * HW * frames_per_stat;
      const double mean = sa / cnt;
      double var = sb / cnt - mean * mean;
      if (var < 0) var = 0;
      s_mean[threadIdx.x] = static_cast<float>(mean);
      s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
  }
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int TY = kGnThreads / TX;
  if (ty >= TY) return;
  float sc[SLOTS][8], sh[SLOTS][8];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int cv = tx + s * TX;
    if (s < slots) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = cv * 8 + e;
        const int g = c / cpg;
        sc[s][e] = s_rstd[g] * gamma[c];
        sh[s][e] = beta[c] - s_mean[g] * s_rstd[g] * gamma[c];
      }
    }
  }
  const int p0 = blockIdx.x * px_per_cta;
  const int p1 = min(HW, p0 + px_per_cta);
  const long long base = (static_cast<long long>(nb) * HW) * C;
#pragma unroll 4
  for (int p = p0 + ty; p < p1; p += TY) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int cv = tx + s * TX;
      if (s < slots) {
        const long long off = base + static_cast<long long>(p) * C + cv * 8;
        const Half8 h = *reinterpret_cast<const Half8*>(x + off);
        Half8 o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = __half2float(h.v[e]) * sc[s][e] + sh[s][e];
          if (silu) v = v / (1.0f + __expf(-v));
          o.v[e] = __float2half_rn(v);
        }
        *reinterpret_cast<Half8*>(y + off) = o;
      }
    }
  }
}

static void gn_geometry(int C, int HW, int NB, int* TX, int* slots, int* px_per_cta, int* chunks) {
  const int CV = C / 8;
  int s = (CV + kGnThreads - 1) / kGnThreads;
  while (CV % s || s == 3) ++s;  // TX * slots == CV exactly, slots in {1, 2, 4}
  *slots = s;
  *TX = CV / s;
  const int TY = kGnThreads / *TX;
  int want = std::max(1, ((s == 1 ? 8 : 4) * sm_count()) / std::max(1, NB));  // 1-slot kernels use ~40 registers: 8 CTAs per SM resident
  int ppc = std::max(TY, (HW + want - 1) / want);
  *px_per_cta = ppc;
  *chunks = (HW + ppc - 1) / ppc;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the channel axis of token rows (models/attention.py:281,303,320,331), one warp per row.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 8;  // C <= 8*32*8 = 2048

// NV = 16-byte vectors per lane (ceil(C/256)), ROWS = rows per warp in flight (memory-level parallelism); both compile-time so the
// row buffer stays in registers and occupancy follows the real channel count.
template <int NV, int ROWS>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long M, int C,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row0 = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= M) return;
  const int CV = C / 8;
  Half8 buf[ROWS][NV];
  float sum[ROWS], sq[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    sum[r] = 0.f;
    const long long row = min(row0 + r, M - 1);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int cv = lane + i * 32;
      if (cv < CV) buf[r][i] = *reinterpret_cast<const Half8*>(x + row * C + cv * 8);
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < CV) {
#pragma unroll
        for (int e = 0; e < 8; ++e) sum[r] += __half2float(buf[r][i].v[e]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const float mean = sum[r] / C;
    sq[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < CV) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = __half2float(buf[r][i].v[e]) - mean;
          sq[r] += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], o);
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const long long row = row0 + r;
    if (row >= M) break;
    const float mean = sum[r] / C;
    const float rstd = rsqrtf(sq[r] / C + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int cv = lane + i * 32;
      if (cv < CV) {
        Half8 o;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = __float2half_rn((__half2float(buf[r][i].v[e]) - mean) * rstd * gg[e] + bb[e]);
        *reinterpret_cast<Half8*>(y + row * C + cv * 8) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// nearest 2x upsample (resnet.py:145) and channel concat (unet_3d_blocks.py:522,611), NHWC
// ---------------------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const Half8* __restrict__ x, Half8* __restrict__ y, int NB, int H, int W, int CV) {
  const long long total = static_cast<long long>(NB) * (2 * H) * (2 * W) * CV;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = i % CV;
    long long r = i / CV;
    const int ox = r % (2 * W);
    r /= (2 * W);
    const int oy = r % (2 * H);
    const int nb = r / (2 * H);
    y[i] = x[((static_cast<long long>(nb) * H + oy / 2) * W + ox / 2) * CV + cv];
  }
}

__global__ void concat2_kernel(const Half8* __restrict__ a, int CVa, const Half8* __restrict__ b, int CVb, Half8* __restrict__ y, long long rows) {
  const int CV = CVa + CVb;
  const long long total = rows * CV;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = i % CV;
    const long long r = i / CV;
    y[i] = (cv < CVa) ? a[r * CVa + cv] : b[r * CVb + (cv - CVa)];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// UNet input: latents [B, Cl, F, H, W] fp32 -> im2col rows [B*F*H*W, 64] fp16 (col = tap*Cl + c, tap = ky*3+kx, zero padded)
// so conv_in (unet_3d_condition.py:375) runs as a K=64 GEMM.
// ---------------------------------------------------------------------------------------------------------------
__global__ void im2col_in_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int Cl, int F, int H, int W) {
  const long long rows = static_cast<long long>(B) * F * H * W;
  const long long total = rows * 8;  // 8 vectors of 8 halfs per row
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = i % 8;
    long long r = i / 8;
    const int xx = r % W;
    long long t = r / W;
    const int yy = t % H;
    t /= H;
    const int f = t % F;
    const int b = t / F;
    Half8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = v * 8 + e;
      float val = 0.f;
      if (col < 9 * Cl) {
        const int tap = col / Cl, c = col % Cl;
        const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) val = x[(((static_cast<long long>(b) * Cl + c) * F + f) * H + sy) * W + sx];
      }
      o.v[e] = __float2half_rn(val);
    }
    *reinterpret_cast<Half8*>(out + r * 64 + v * 8) = o;
  }
}

// conv_out tail: y [B*F*HW, ldy] fp16 (Co valid channels, conv bias already added) -> temporal conv over frames
//   lora : out = y + up(down(y))   (down [R, Co, 3], up [Co, R, 3], lora.py:46-54)       (w_full == null)
//   full : out = bias + W * y      (W [Co, Co, 3], resnet.py:42-55)                        (w_full != null)
// and scatter to eps [B, Co, F, H, W] fp32 (the layout the DDIM step consumes).
__global__ void out_temporal_kernel(const __half* __restrict__ y, int ldy, float* __restrict__ eps, int B, int Co, int F, int HW,
                                    const float* __restrict__ down, const float* __restrict__ up, int R, const float* __restrict__ w_full,
                                    const float* __restrict__ b_full) {
  const long long total = static_cast<long long>(B) * F * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = i % HW;
    const int f = (i / HW) % F;
    const int b = i / (static_cast<long long>(HW) * F);
    auto ld = [&](int ff, int c) -> float {
      if (ff < 0 || ff >= F) return 0.f;
      return __half2float(y[((static_cast<long long>(b) * F + ff) * HW + p) * ldy + c]);
    };
    float outv[8];
    if (w_full) {
      for (int c = 0; c < Co; ++c) {
        float a = b_full ? b_full[c] : 0.f;
        for (int ci = 0; ci < Co; ++ci)
          for (int t = 0; t < 3; ++t) a += w_full[(c * Co + ci) * 3 + t] * ld(f + t - 1, ci);
        outv[c] = a;
      }
    } else if (down) {
      // mid[r][g] for frames g = f-1, f, f+1 (zero outside), as the fp16-rounded intermediate of the reference autocast path
      float mid[4][3];
      for (int r = 0; r < R; ++r)
        for (int dg = 0; dg < 3; ++dg) {
          const int g = f + dg - 1;
          float a = 0.f;
          if (g >= 0 && g < F)
            for (int ci = 0; ci < Co; ++ci)
              for (int t = 0; t < 3; ++t) a += down[(r * Co + ci) * 3 + t] * ld(g + t - 1, ci);
          mid[r][dg] = (g >= 0 && g < F) ? __half2float(__float2half_rn(a)) : 0.f;
        }
      for (int c = 0; c < Co; ++c) {
        float a = ld(f, c);
        for (int r = 0; r < R; ++r)
          for (int t = 0; t < 3; ++t) a += up[(c * R + r) * 3 + t] * mid[r][t];
        outv[c] = a;
      }
    } else {
      for (int c = 0; c < Co; ++c) outv[c] = ld(f, c);
    }
    for (int c = 0; c < Co; ++c) eps[((static_cast<long long>(b) * Co + c) * F + f) * HW + p] = outv[c];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// small dense layers on a single row (time embedding path, unet_3d_condition.py:356-362; resnet.py:355):
//   y[n] = bias[n] + sum_k act(x[k]) * W[n,k]     act = identity | SiLU ;  one warp per output
// ---------------------------------------------------------------------------------------------------------------
__global__ void rowvec_linear_kernel(const float* __restrict__ x, const __half* __restrict__ W, const float* __restrict__ bias,
                                     float* __restrict__ y, int N, int K, int silu_in) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float a = 0.f;
  for (int k = lane * 2; k < K; k += 64) {
    float x0 = x[k], x1 = (k + 1 < K) ? x[k + 1] : 0.f;
    if (silu_in) { x0 = x0 / (1.f + __expf(-x0)); x1 = x1 / (1.f + __expf(-x1)); }
    const __half2 w = *reinterpret_cast<const __half2*>(W + static_cast<long long>(n) * K + k);
    a += x0 * __low2float(w) + x1 * __high2float(w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) y[n] = a + (bias ? bias[n] : 0.f);
}

// Timesteps(C0, flip_sin_to_cos, freq_shift) — diffusers embeddings.get_timestep_embedding
__global__ void timestep_sinusoid_kernel(float t, float* __restrict__ out, int C0, int flip, float freq_shift) {
  const int half = C0 / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const float e = expf(-logf(10000.f) * static_cast<float>(i) / (static_cast<float>(half) - freq_shift));
  const float a = t * e;
  const float s = sinf(a), c = cosf(a);
  if (flip) { out[i] = c; out[half + i] = s; }
  else { out[i] = s; out[half + i] = c; }
}

// ---------------------------------------------------------------------------------------------------------------
// temporal attention over the frame axis (models/attention.py:327-337): qkv [B*F*HW, 3C] fp16 -> out [B*F*HW, C] fp16
// one warp per (b, pixel, head); probabilities are rounded to fp16 before PV like the reference's `.to(value.dtype)`.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTaMaxF = 32;
__global__ void __launch_bounds__(256) temporal_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int B, int F, int HW,
                                                           int heads, int d, float scale) {
  extern __shared__ __half ta_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int C = heads * d;
  __half* sq = ta_smem + static_cast<size_t>(warp) * (3 * F * d + F * F * 2);
  __half* sk = sq + F * d;
  __half* sv = sk + F * d;
  float* sp = reinterpret_cast<float*>(sv + F * d);
  const long long items = static_cast<long long>(B) * HW * heads;
  for (long long it = static_cast<long long>(blockIdx.x) * wpb + warp; it < items; it += static_cast<long long>(gridDim.x) * wpb) {
    const int h = it % heads;
    const int p = (it / heads) % HW;
    const int b = it / (static_cast<long long>(heads) * HW);
    const int d2 = d / 2;
    for (int i = lane; i < F * d2; i += 32) {
      const int f = i / d2, j = i % d2;
      const long long row = (static_cast<long long>(b) * F + f) * HW + p;
      const __half2* src = reinterpret_cast<const __half2*>(qkv + row * 3 * C + h * d) + j;
      reinterpret_cast<__half2*>(sq)[i] = src[0];
      reinterpret_cast<__half2*>(sk)[i] = src[C / 2];
      reinterpret_cast<__half2*>(sv)[i] = src[C];
    }
    __syncwarp();
    for (int i = lane; i < F * F; i += 32) {
      const int f = i / F, g = i % F;
      float a = 0.f;
      for (int j = 0; j < d2; ++j) {
        const float2 qa = __half22float2(reinterpret_cast<const __half2*>(sq)[f * d2 + j]);
        const float2 ka = __half22float2(reinterpret_cast<const __half2*>(sk)[g * d2 + j]);
        a += qa.x * ka.x + qa.y * ka.y;
      }
      sp[i] = a * scale;
    }
    __syncwarp();
    if (lane < F) {
      float mx = -INFINITY;
      for (int g = 0; g < F; ++g) mx = fmaxf(mx, sp[lane * F + g]);
      float sum = 0.f;
      for (int g = 0; g < F; ++g) { const float e = __expf(sp[lane * F + g] - mx); sp[lane * F + g] = e; sum += e; }
      const float inv = 1.f / sum;
      for (int g = 0; g < F; ++g) sp[lane * F + g] = __half2float(__float2half_rn(sp[lane * F + g] * inv));
    }
    __syncwarp();
    for (int i = lane; i < F * d; i += 32) {
      const int f = i / d, dd = i % d;
      float a = 0.f;
      for (int g = 0; g < F; ++g) a += sp[f * F + g] * __half2float(sv[g * d + dd]);
      const long long row = (static_cast<long long>(b) * F + f) * HW + p;
      out[row * C + h * d + dd] = __float2half_rn(a);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// DDIM steps on fp32 latents [*, 4, F, H, W] (p2p_ddim_spatial_temporal.py:150-161 and :400-407 + diffusers DDIMScheduler.step
// eta=0) with the latent blend of spatial_blend.py:116-122 fused in.
// ---------------------------------------------------------------------------------------------------------------
__global__ void ddim_invert_kernel(float* __restrict__ x, const float* __restrict__ eps, long long n, float sqrt_a_prev, float sqrt_1m_a_prev,
                                   float sqrt_a_next, float sqrt_1m_a_next) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float e = eps[i];
    const float x0 = (x[i] - sqrt_1m_a_prev * e) / sqrt_a_prev;
    x[i] = sqrt_a_next * x0 + sqrt_1m_a_next * e;
  }
}

// eps2 = [uncond | cond] each n elements. mask (optional): [F*H*W] floats per frame pixel, broadcast over channels.
__global__ void cfg_ddim_kernel(float* __restrict__ x, const float* __restrict__ eps2, long long n, float guidance, float sqrt_a_t,
                                float sqrt_1m_a_t, float sqrt_a_prev, float sqrt_1m_a_prev, const float* __restrict__ x_inv,
                                const float* __restrict__ mask_a, const float* __restrict__ mask_b, long long fhw, int apply_blend) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float eu = eps2[i], ec = eps2[n + i];
    const float e = eu + guidance * (ec - eu);
    const float x0 = (x[i] - sqrt_1m_a_t * e) / sqrt_a_t;
    float xn = sqrt_a_prev * x0 + sqrt_1m_a_prev * e;
    if (apply_blend) {
      const long long q = i % fhw;
      float m = mask_a[q];
      if (mask_b) m = fmaxf(m, mask_b[q]);
      const float xi = x_inv[i];
      xn = xi + m * (xn - xi);
    }
    x[i] = xn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Blend mask (spatial_blend.py:24-39, 78-111): mean over (layers, heads) of sum_n map[f,hd,p,n]*w[n] on the r x r grid,
// 3x3 max-pool (stride 1, pad 1), nearest resize to (h, w), divide by the per-frame max, compare with th.
// maps: up to 8 pointers to [F, heads, r*r, ldm] fp16 (or fp32 running sums when maps_f32 != 0). One CTA per frame.
// ---------------------------------------------------------------------------------------------------------------
struct MaskParams {
  const void* maps[8];
  int num_maps;
  int maps_f32;
  int F, heads, r, ldm, ntok;
  float w[80];
  float th;
  int h, w_out;
  float* out;  // [F, h, w] 0/1
};

__global__ void __launch_bounds__(256) blend_mask_kernel(const __grid_constant__ MaskParams p) {
  extern __shared__ float mk_smem[];
  float* agg = mk_smem;              // [r*r]
  float* pooled = mk_smem + p.r * p.r;  // [r*r]
  __shared__ float s_max;
  const int f = blockIdx.x;
  const int rr = p.r * p.r;
  for (int px = threadIdx.x; px < rr; px += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < p.num_maps; ++l)
      for (int hd = 0; hd < p.heads; ++hd) {
        const long long off = ((static_cast<long long>(f) * p.heads + hd) * rr + px) * p.ldm;
        float s = 0.f;
        if (p.maps_f32) {
          const float* m = static_cast<const float*>(p.maps[l]) + off;
          for (int n = 0; n < p.ntok; ++n) s += m[n] * p.w[n];
        } else {
          const __half* m = static_cast<const __half*>(p.maps[l]) + off;
          for (int n = 0; n < p.ntok; ++n) s += __half2float(m[n]) * p.w[n];
        }
        a += s;
      }
    agg[px] = a / static_cast<float>(p.num_maps * p.heads);
  }
  if (threadIdx.x == 0) s_max = 0.f;
  __syncthreads();
  for (int px = threadIdx.x; px < rr; px += blockDim.x) {
    const int y = px / p.r, x = px % p.r;
    float m = -INFINITY;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < p.r && xx >= 0 && xx < p.r) m = fmaxf(m, agg[yy * p.r + xx]);
      }
    pooled[px] = m;
  }
  __syncthreads();
  // max over the RESIZED grid == max over the source pixels that the nearest resize actually samples
  float lm = 0.f;
  for (int i = threadIdx.x; i < p.h * p.w_out; i += blockDim.x) {
    const int y = i / p.w_out, x = i % p.w_out;
    const int sy = min(p.r - 1, (y * p.r) / p.h), sx = min(p.r - 1, (x * p.r) / p.w_out);
    lm = fmaxf(lm, pooled[sy * p.r + sx]);
  }
  atomicMax(reinterpret_cast<int*>(&s_max), __float_as_int(fmaxf(lm, 0.f)));
  __syncthreads();
  const float mx = s_max;
  for (int i = threadIdx.x; i < p.h * p.w_out; i += blockDim.x) {
    const int y = i / p.w_out, x = i % p.w_out;
    const int sy = min(p.r - 1, (y * p.r) / p.h), sx = min(p.r - 1, (x * p.r) / p.w_out);
    const float v = pooled[sy * p.r + sx];
    // reference: (v / mx) > th, with 0/0 = NaN -> False
    p.out[(static_cast<long long>(f) * p.h + y) * p.w_out + x] = (mx > 0.f && (v / mx) > p.th) ? 1.f : 0.f;
  }
}

static inline int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(std::max<long long>(1, std::min(g, cap)));
}

}  // namespace fz

using namespace fz;

extern "C" int fz_groupnorm_nhwc_f16(const void* x, void* y, int NB, int HW, int C, int groups, int frames_per_stat, const float* gamma,
                                     const float* beta, float eps, int silu, void* workspace_f64, cudaStream_t stream) {
  FZ_CHECK_ARG(x && y && gamma && beta && workspace_f64, "fz_groupnorm: null pointer");
  FZ_CHECK_ARG(C % 8 == 0 && C % groups == 0 && groups <= 64, "fz_groupnorm: C=%d groups=%d unsupported", C, groups);
  FZ_CHECK_ARG(frames_per_stat >= 1 && NB % frames_per_stat == 0, "fz_groupnorm: NB %% frames_per_stat != 0");
  int TX, slots, ppc, chunks;
  gn_geometry(C, HW, NB, &TX, &slots, &ppc, &chunks);
  FZ_CHECK_ARG(slots <= kGnMaxSlots, "fz_groupnorm: C=%d too large", C);
  FZ_CHECK_ARG(static_cast<size_t>(NB) * chunks * groups * sizeof(float2) <= (1u << 20), "fz_groupnorm: workspace (1 MiB) too small");
  const int TY = kGnThreads / TX;
  const size_t smem = static_cast<size_t>(2) * TY * C * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    FZ_CUDA(cudaFuncSetAttribute(gn_stats_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    FZ_CUDA(cudaFuncSetAttribute(gn_stats_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    FZ_CUDA(cudaFuncSetAttribute(gn_stats_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  dim3 grid(chunks, NB);
  float2* partial = static_cast<float2*>(workspace_f64);
  const __half* xh = static_cast<const __half*>(x);
  __half* yh = static_cast<__half*>(y);
#define FZ_GN_LAUNCH(SL)                                                                                                          \
  do {                                                                                                                             \
    gn_stats_kernel<SL><<<grid, kGnThreads, smem, stream>>>(xh, HW, C, groups, TX, ppc, partial);                                  \
    gn_apply_kernel<SL><<<grid, kGnThreads, 0, stream>>>(xh, yh, HW, C, groups, frames_per_stat, TX, ppc, chunks, partial, gamma, \
                                                         beta, eps, silu);                                                         \
  } while (0)
  if (slots == 1) FZ_GN_LAUNCH(1);
  else if (slots == 2) FZ_GN_LAUNCH(2);
  else FZ_GN_LAUNCH(4);
#undef FZ_GN_LAUNCH
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_layernorm_f16(const void* x, void* y, long long M, int C, const float* gamma, const float* beta, float eps,
                                cudaStream_t stream) {
  FZ_CHECK_ARG(x && y && gamma && beta, "fz_layernorm: null pointer");
  FZ_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * kLnMaxVec, "fz_layernorm: C=%d unsupported", C);
  const int nv = (C / 8 + 31) / 32;
  const __half* xh = static_cast<const __half*>(x);
  __half* yh = static_cast<__half*>(y);
#define FZ_LN_LAUNCH(NV, ROWS) \
  layernorm_kernel<NV, ROWS><<<static_cast<unsigned>((M + 8 * ROWS - 1) / (8 * ROWS)), 256, 0, stream>>>(xh, yh, M, C, gamma, beta, eps)
  switch (nv) {
    case 1: FZ_LN_LAUNCH(1, 4); break;
    case 2: FZ_LN_LAUNCH(2, 4); break;
    case 3: FZ_LN_LAUNCH(3, 2); break;
    case 4: FZ_LN_LAUNCH(4, 2); break;
    case 5: FZ_LN_LAUNCH(5, 2); break;
    case 6: FZ_LN_LAUNCH(6, 2); break;
    case 7: FZ_LN_LAUNCH(7, 1); break;
    default: FZ_LN_LAUNCH(8, 1); break;
  }
#undef FZ_LN_LAUNCH
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_upsample2x_nhwc_f16(const void* x, void* y, int NB, int H, int W, int C, cudaStream_t stream) {
  FZ_CHECK_ARG(x && y && C % 8 == 0, "fz_upsample2x: bad args");
  const long long total = static_cast<long long>(NB) * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const Half8*>(x), static_cast<Half8*>(y), NB, H, W, C / 8);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_concat_channels_f16(const void* a, int Ca, const void* b, int Cb, void* y, long long rows, cudaStream_t stream) {
  FZ_CHECK_ARG(a && b && y && Ca % 8 == 0 && Cb % 8 == 0, "fz_concat_channels: bad args");
  const long long total = rows * ((Ca + Cb) / 8);
  concat2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const Half8*>(a), Ca / 8, static_cast<const Half8*>(b), Cb / 8,
                                                           static_cast<Half8*>(y), rows);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_im2col_latents_f16(const float* x, void* out, int B, int Cl, int F, int H, int W, cudaStream_t stream) {
  FZ_CHECK_ARG(x && out && Cl * 9 <= 64, "fz_im2col_latents: bad args");
  const long long total = static_cast<long long>(B) * F * H * W * 8;
  im2col_in_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, static_cast<__half*>(out), B, Cl, F, H, W);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_out_temporal_f32(const void* y, int ldy, float* eps, int B, int Co, int F, int HW, const float* down, const float* up, int rank,
                                   const float* w_full, const float* b_full, cudaStream_t stream) {
  FZ_CHECK_ARG(y && eps && Co <= 8 && rank <= 4, "fz_out_temporal: bad args");
  const long long total = static_cast<long long>(B) * F * HW;
  out_temporal_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __half*>(y), ldy, eps, B, Co, F, HW, down, up, rank, w_full,
                                                                b_full);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_rowvec_linear(const float* x, const void* W_f16, const float* bias, float* y, int N, int K, int silu_in, cudaStream_t stream) {
  FZ_CHECK_ARG(x && W_f16 && y && K % 2 == 0, "fz_rowvec_linear: bad args");
  rowvec_linear_kernel<<<(N + 7) / 8, 256, 0, stream>>>(x, static_cast<const __half*>(W_f16), bias, y, N, K, silu_in);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_timestep_sinusoid(float t, float* out, int C0, int flip_sin_to_cos, float freq_shift, cudaStream_t stream) {
  FZ_CHECK_ARG(out && C0 % 2 == 0, "fz_timestep_sinusoid: bad args");
  timestep_sinusoid_kernel<<<(C0 / 2 + 127) / 128, 128, 0, stream>>>(t, out, C0, flip_sin_to_cos, freq_shift);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_temporal_attn_f16(const void* qkv, void* out, int B, int F, int HW, int heads, int d, float scale, cudaStream_t stream) {
  FZ_CHECK_ARG(qkv && out && F <= kTaMaxF && d % 2 == 0, "fz_temporal_attn: F=%d d=%d unsupported", F, d);
  const int wpb = 8;
  const size_t per_warp = static_cast<size_t>(3 * F * d + F * F * 2) * sizeof(__half);
  const size_t smem = per_warp * wpb;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    FZ_CUDA(cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  const long long items = static_cast<long long>(B) * HW * heads;
  const int grid = static_cast<int>(std::min<long long>((items + wpb - 1) / wpb, static_cast<long long>(sm_count()) * 4));
  temporal_attn_kernel<<<grid, wpb * 32, smem, stream>>>(static_cast<const __half*>(qkv), static_cast<__half*>(out), B, F, HW, heads, d, scale);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_ddim_invert_step(float* x, const float* eps, long long n, float a_prev, float a_next, cudaStream_t stream) {
  FZ_CHECK_ARG(x && eps, "fz_ddim_invert_step: null pointer");
  ddim_invert_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, eps, n, sqrtf(a_prev), sqrtf(1.f - a_prev), sqrtf(a_next), sqrtf(1.f - a_next));
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_cfg_ddim_step(float* x, const float* eps2, long long n, float guidance, float a_t, float a_prev, const float* x_inv,
                                const float* mask_a, const float* mask_b, long long fhw, int apply_blend, cudaStream_t stream) {
  FZ_CHECK_ARG(x && eps2, "fz_cfg_ddim_step: null pointer");
  FZ_CHECK_ARG(!apply_blend || (x_inv && mask_a && fhw > 0), "fz_cfg_ddim_step: blend needs x_inv and mask");
  cfg_ddim_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, eps2, n, guidance, sqrtf(a_t), sqrtf(1.f - a_t), sqrtf(a_prev), sqrtf(1.f - a_prev),
                                                        x_inv, mask_a, mask_b, fhw, apply_blend);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_blend_mask(const void* const* maps, int num_maps, int maps_f32, int F, int heads, int r, int ldm, int ntok, const float* word_w,
                             float th, int h, int w, float* out, cudaStream_t stream) {
  FZ_CHECK_ARG(maps && word_w && out && num_maps >= 1 && num_maps <= 8 && ntok <= 80, "fz_blend_mask: bad args");
  MaskParams p;
  for (int i = 0; i < 8; ++i) p.maps[i] = i < num_maps ? maps[i] : nullptr;
  p.num_maps = num_maps; p.maps_f32 = maps_f32; p.F = F; p.heads = heads; p.r = r; p.ldm = ldm; p.ntok = ntok;
  for (int i = 0; i < 80; ++i) p.w[i] = i < ntok ? word_w[i] : 0.f;  // word_w is a HOST array (77 floats, built once per edit)
  p.th = th; p.h = h; p.w_out = w; p.out = out;
  const size_t smem = static_cast<size_t>(2 * r * r) * sizeof(float);
  blend_mask_kernel<<<F, 256, smem, stream>>>(p);
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}
