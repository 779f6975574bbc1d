// fz_gemm.cu — the "tap-GEMM": one persistent, warp-specialised tcgen05 kernel that serves every dense contraction of the
// UNet step except attention:
//     D[M, N] = sum_{tap} A_tap[M, K] * W_tap[N, K]^T   (+ bias, + per-batch time-embedding row, + residual, GEGLU, V^T store)
//   * nn.Linear / 1x1 conv ............ 1 tap, A = [M, K] row-major tokens
//   * 3x3 conv (stride 1 / stride 2) .. 9 taps, A = NHWC activation addressed through a 4-D / 5-D TMA map; the tap moves the
//                                       box by (dy, dx) and TMA zero-fills the halo (implicit GEMM, no im2col in HBM)
//   * temporal LoRA Conv1d(k=3) ....... 3 taps along the frame axis of [B, F, HW, C]
// Reference ops replaced: models/resnet.py:57-80 (PseudoConv3d.forward), models/lora.py:46-54, nn.Linear call sites of
// prompt_attention/attention_register.py:81,99-100,124,156-160,214 and diffusers FeedForward/GEGLU (models/attention.py:320).
//
// Structure (per CTA, 320 threads, 1 CTA / SM, grid = min(#tiles, #SMs), static round-robin tile schedule):
//   warp 0 : TMA producer   — A box {64ch, rows} + W box {64ch, BLOCK_N} per stage, SWIZZLE_128B, mbarrier complete_tx
//   warp 1 : MMA issuer     — tcgen05.mma.cta_group::1.kind::f16  M=128 x N=BLOCK_N x K=16, fp32 accumulators in TMEM,
//                             double-buffered accumulators (2 x BLOCK_N columns) so the epilogue overlaps the next tile
//   warps 2-9 : epilogue    — tcgen05.ld 32x32b -> registers -> fused epilogue -> 16-byte global stores (two warps per TMEM
//                             lane quadrant, alternating 32-column chunks)
#include "fz_common.cuh"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>

#include "../../include/fatezero_b200.h"

namespace fz {

constexpr int kMaxTaps = 9;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB

struct TapGemmParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmC;    // output [M, N] as (cols, rows), box {32 cols, 32 rows}, SWIZZLE_64B (epilogue TMA stores)
  CUtensorMap tmC16;  // same output with box {16 cols, 32 rows}, no swizzle (GEGLU epilogue with 16-column chunks)
  CUtensorMap tmR[2]; // skip tensors [M, N] folded into the accumulator as extra k-blocks: box {64 cols, 128 rows}
  CUtensorMap tmE;    // identity blocks E[j][n][k] = (n == 64 j + k): (k : 64, n : 256, j : 4), box {64, BLOCK_N, 1}
  CUtensorMap tmVt;   // V^T output [BF * heads * d rows, vt_ld] as (s, row), box {32 s, 32 rows}: transposed chunks leave through smem + TMA
  int vt_tma;         // 1 = the V^T columns take the staged TMA path (vt_S % 32 == 0), 0 = per-element stores
  int n_res;          // number of folded skip tensors (0..2); the epilogue then sees residual == residual2 == null
  int res_kblocks;    // ceil(BLOCK_N / 64) extra k-blocks per folded skip tensor
  int use_tma_store;  // 0 = direct 16-byte stores (fallback for odd geometries)
  int a_rank;         // 2..5
  int M, N;           // valid output rows / columns (for GEGLU: N = number of OUTPUT columns = half the GEMM columns)
  int rows_per_tile;  // <= 128; tile t covers output rows [t*rows_per_tile, ...)
  int a_box_bytes;    // bytes TMA writes for one A box (rows_per_tile * 128)
  int k_blocks;       // per tap
  int num_taps;
  int n_tiles, m_tiles;
  int ndecomp;        // how many of A's dims 1.. take part in the row decomposition
  int dimsz[4];       // their sizes
  int tap_off[kMaxTaps][5];  // coordinate offsets per tap for A dims 0..4
  const float* bias;         // [gemm columns] fp32 (GEGLU: packed like the weights) or null
  const float* group_bias;   // [M / rows_per_group, N] fp32 or null (time-embedding projection per batch element)
  int rows_per_group;
  const __half* residual;    // [M, ldr] or null
  long long ldr;
  const __half* residual2;   // second skip tensor [M, ldr2] or null
  long long ldr2;
  __half* out;
  long long ldo;
  int mode;          // FZ_EPI_ROWMAJOR / FZ_EPI_GEGLU
  int vt_col_start;  // gemm columns >= this are written transposed (V^T) instead of row-major; INT_MAX = off
  __half* out_vt;    // [BF, heads, d, S]
  int vt_S, vt_d, vt_heads, vt_ld;
};

template <int BLOCK_N>
struct TapGemmCfg {
  static constexpr int kBTileBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kBTilePad = (kBTileBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kATileBytes + kBTilePad;
  static constexpr int kEpiBytes = 8 * 2 * 2048;  // 8 epilogue warps x 2 staging slots of 32 rows x 64 B
  static constexpr int kStagesRaw = (227 * 1024 - kEpiBytes - 1280) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// exact-erf GELU (F.gelu default).  erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the fp16 output rounding):
// 1 MUFU.RCP + 1 MUFU.EX2 + 8 FMA instead of the ~40-instruction erff() — the GEGLU epilogue was XU/ALU-bound (ncu: 24 % tensor pipe).
__device__ __forceinline__ float gelu_erf(float x) {
  // single-instruction MUFU forms: __frcp_rn / exp2f expand to IEEE-exact sequences (~10 extra instructions each) and made the GEGLU
  // epilogue latency-bound (measured 4900 cycles per 32x32 chunk, 84 % of the epilogue warp's time)
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Optional in-kernel cycle accounting (compile with -DFZ_GEMM_PROFILE, read with fz_debug_gemm_counters): CTA 0 only.
//  [0] epilogue warp 2: loop total  [1] wait tfull  [2] TMEM load  [3] math  [4] wait for a free staging slot  [5] stage + TMA store
//  [6] chunks  [7] tiles   [8] MMA thread: total  [9] wait tempty  [10] wait full   [12] producer: total  [13] wait empty
__device__ long long g_gemm_dbg[16];
#ifdef FZ_GEMM_PROFILE
#define GP_NOW() clock64()
#define GP_ADD(slot, expr) do { if (gp_on) gp[slot] += (expr); } while (0)
#else
#define GP_NOW() 0LL
#define GP_ADD(slot, expr) do { } while (0)
#endif

// EPI_GROUPS = epilogue warps per TMEM lane quadrant: 2 (320 threads) for every row-major epilogue; 4 (576 threads, 16-column chunks so that
// the kernel fits the smaller register budget) for the GEGLU epilogue, which is ALU-bound: it was 3x the mainloop with two warps per scheduler.
template <int BLOCK_N, int EPI_GROUPS = 2>
__global__ void __launch_bounds__(64 + 128 * EPI_GROUPS, 1) tapgemm_kernel(const __grid_constant__ TapGemmParams p) {
  using Cfg = TapGemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  pdl_launch_dependents();
#ifdef FZ_GEMM_PROFILE
  const bool gp_on = blockIdx.x == 0;
  long long gp[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.num_taps * p.k_blocks + p.n_res * p.res_kblocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.use_tma_store) tma_prefetch_desc(&p.tmC);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4 * EPI_GROUPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; operands / outputs are touched only from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const long long gp_t0 = GP_NOW();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        int rem = mt * p.rows_per_tile;
        int base[5] = {0, 0, 0, 0, 0};
        for (int d = 0; d < p.ndecomp; ++d) {
          base[d + 1] = rem % p.dimsz[d];
          rem /= p.dimsz[d];
        }
        const int n0 = nt * BLOCK_N;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int c1 = base[1] + p.tap_off[tap][1], c2 = base[2] + p.tap_off[tap][2];
          const int c3 = base[3] + p.tap_off[tap][3], c4 = base[4] + p.tap_off[tap][4];
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            const long long gp_a = GP_NOW();
            mbar_wait(&empty[stage], phase ^ 1);
            GP_ADD(13, GP_NOW() - gp_a);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + kATileBytes;
            mbar_expect_tx(&full[stage], p.a_box_bytes + Cfg::kBTileBytes);
            const int c0 = kb * kBlockK + p.tap_off[tap][0];
            switch (p.a_rank) {
              case 2: tma_load_2d(sa, &p.tmA, &full[stage], c0, c1); break;
              case 3: tma_load_3d(sa, &p.tmA, &full[stage], c0, c1, c2); break;
              case 4: tma_load_4d(sa, &p.tmA, &full[stage], c0, c1, c2, c3); break;
              default: tma_load_5d(sa, &p.tmA, &full[stage], c0, c1, c2, c3, c4); break;
            }
            tma_load_3d(sb, &p.tmB, &full[stage], kb * kBlockK, n0, tap);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
        // Skip connections ride the same pipeline as extra k-blocks: D += R[:, n0 + 64 j ...] * E_j^T with E_j a 0/1 selection block
        // (exact: fp16 x 1.0 accumulated in fp32 after the last tap, the same order as an epilogue add).  The epilogue of a
        // K = 320 linear used to wait ~1 us of exposed global-load latency per 32-column chunk for the skip tensor
        // (65536 x 320 x 320 + skip: 40 us against 21 us without); here the loads are prefetched by the TMA ring like any operand.
        for (int r = 0; r < p.n_res; ++r) {
          for (int j = 0; j < p.res_kblocks; ++j) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + kATileBytes;
            mbar_expect_tx(&full[stage], kATileBytes + Cfg::kBTileBytes);
            tma_load_2d(sa, &p.tmR[r], &full[stage], n0 + j * kBlockK, mt * p.rows_per_tile);
            tma_load_3d(sb, &p.tmE, &full[stage], 0, 0, j);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
#ifdef FZ_GEMM_PROFILE
      if (gp_on) { g_gemm_dbg[12] = GP_NOW() - gp_t0; g_gemm_dbg[13] = gp[13]; }
#endif
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N);
      const uint64_t desc_hi = umma_desc_k_sw128(0);  // constant descriptor fields; only the 14-bit start address varies
      const uint32_t smem_lo0 = (smem_u32(smem) & 0x3FFFF) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      const long long gp_m0 = GP_NOW();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        const uint32_t use = static_cast<uint32_t>(local >> 1);
        const long long gp_a = GP_NOW();
        mbar_wait(&tempty[buf], (use & 1) ^ 1);
        GP_ADD(9, GP_NOW() - gp_a);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
        for (int it = 0; it < k_iters; ++it) {
          const long long gp_b = GP_NOW();
          mbar_wait(&full[stage], phase);
          GP_ADD(10, GP_NOW() - gp_b);
          tc_fence_after();
          const uint32_t a_lo = smem_lo0 + stage * (Cfg::kStageBytes >> 4);
          const uint32_t b_lo = a_lo + (kATileBytes >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_f16_ss(d_tmem, desc_hi | (a_lo + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (it | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);
      }
#ifdef FZ_GEMM_PROFILE
      if (gp_on) { g_gemm_dbg[8] = GP_NOW() - gp_m0; g_gemm_dbg[9] = gp[9]; g_gemm_dbg[10] = gp[10]; }
#endif
    }
  } else {
    // ===================== epilogue warps =====================
    // Straight-line fast path per 32-column chunk (all option tests are warp-uniform and hoisted out of the element loops):
    // the v0 epilogue spent ~46 instructions per output element on per-element bound / option branches (ncu: 7 % tensor pipe
    // on K=320 GEMMs); edge tiles fall back to the generic masked path.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;  // the EPI_GROUPS warps of a quadrant take alternating column chunks
    const int row_in_tile = quad * 32 + lane;
    // Coalesced output path: each warp stages its 32 rows x 32 columns (64 B rows, 64-byte swizzle => conflict-free 16-byte
    // st.shared) and one lane issues a TMA store; direct per-row 16-byte stores touch 32 lines per instruction and made the
    // epilogue LSU-bound (~2.4 us per 128x160 tile).
    constexpr int kSlotBytes = (EPI_GROUPS == 4) ? 1024 : 2048;  // 32 rows x (16 | 32) fp16 columns
    uint8_t* my_epi = epi_smem + (warp - 2) * 2 * kSlotBytes;
    uint32_t epi_count = 0;
    auto store_chunk32 = [&](const uint4 (&o)[4], long long m_row, int col, int m_warp0, uint8_t* acquired) {
      if (p.use_tma_store) {
        uint8_t* slot = acquired;
        const long long gp_a = GP_NOW();
        if (slot == nullptr) {
          slot = my_epi + (epi_count & 1) * 2048;
          if (epi_count >= 2) {
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
          }
        }
        const long long gp_b = GP_NOW();
        GP_ADD(4, gp_b - gp_a);
        uint8_t* rowp = slot + lane * 64;
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = o[j];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tmC, slot, col, m_warp0);
          tma_store_commit();
        }
        ++epi_count;
        GP_ADD(5, GP_NOW() - gp_b);
        GP_ADD(6, 1);
      } else {
        uint4* op = reinterpret_cast<uint4*>(p.out + m_row * p.ldo + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = o[j];
      }
    };
    // 16-column variant (GEGLU with 16 epilogue warps): 32-byte rows, no swizzle, tensor map tmC16 with box {16, 32}
    auto store_chunk16 = [&](const uint4 (&o)[2], long long m_row, int col, int m_warp0) {
      if (p.use_tma_store) {
        uint8_t* slot = my_epi + (epi_count & 1) * kSlotBytes;
        const long long gp_a = GP_NOW();
        if (epi_count >= 2) {
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
        }
        const long long gp_b = GP_NOW();
        GP_ADD(4, gp_b - gp_a);
        *reinterpret_cast<uint4*>(slot + lane * 32) = o[0];
        *reinterpret_cast<uint4*>(slot + lane * 32 + 16) = o[1];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tmC16, slot, col, m_warp0);
          tma_store_commit();
        }
        ++epi_count;
        GP_ADD(5, GP_NOW() - gp_b);
        GP_ADD(6, 1);
      } else {
        uint4* op = reinterpret_cast<uint4*>(p.out + m_row * p.ldo + col);
        op[0] = o[0];
        op[1] = o[1];
      }
    };
    // Coalesced residual fetch: lane l reads the 16-byte piece (l & 3) of rows (l >> 2) + 8 i of the warp's 32 x 32 block (8 lines
    // per instruction instead of 32), the block is transposed through the staging slot and every lane picks up its own row.
    const int r_piece = lane & 3, r_row0 = lane >> 2;
    auto residual_fetch = [&](const __half* R, long long ld, int m_warp0, int col, uint4 (&pre)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long rr = static_cast<long long>(m_warp0) + r_row0 + 8 * i;
        pre[i] = (rr < p.M) ? *reinterpret_cast<const uint4*>(R + rr * ld + col + r_piece * 8) : make_uint4(0, 0, 0, 0);
      }
    };
    auto residual_add = [&](const uint4 (&pre)[4], uint8_t* slot, float (&v)[32]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r_row0 + 8 * i;
        *reinterpret_cast<uint4*>(slot + r * 64 + ((r_piece ^ ((r >> 1) & 3)) << 4)) = pre[i];
      }
      __syncwarp();
      const int sw = (lane >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 x = *reinterpret_cast<const uint4*>(slot + lane * 64 + ((j ^ sw) << 4));
        const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(h[q]); v[8 * j + 2 * q] += f.x; v[8 * j + 2 * q + 1] += f.y; }
      }
      __syncwarp();
    };
    auto acquire_slot = [&]() -> uint8_t* {
      if (epi_count >= 2) {
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
      }
      return my_epi + (epi_count & 1) * 2048;
    };
    int local = 0;
    const long long gp_e0 = GP_NOW();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t use = static_cast<uint32_t>(local >> 1);
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const long long m = static_cast<long long>(mt) * p.rows_per_tile + row_in_tile;
      const bool row_ok = row_in_tile < p.rows_per_tile && m < p.M;
      const long long gp_w = GP_NOW();
      mbar_wait(&tfull[buf], use & 1);
      GP_ADD(1, GP_NOW() - gp_w);
      GP_ADD(7, 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * BLOCK_N;
      if (p.mode == FZ_EPI_GEGLU) {
        constexpr int HALF = BLOCK_N / 2;
        constexpr int CH = (HALF >= 32 && EPI_GROUPS == 2) ? 32 : 16;
        if constexpr (HALF >= 16) {
#pragma unroll 1
          for (int c = half * CH; c < HALF; c += EPI_GROUPS * CH) {
            uint32_t xa[CH], ga[CH];
            const long long gp_l = GP_NOW();
            if constexpr (CH == 32) {
              tmem_ld_32x32b_x32(t_row + c, reinterpret_cast<uint32_t(&)[32]>(xa));
              tmem_ld_32x32b_x32(t_row + HALF + c, reinterpret_cast<uint32_t(&)[32]>(ga));
            } else {
              tmem_ld_32x32b_x16(t_row + c, reinterpret_cast<uint32_t(&)[16]>(xa));
              tmem_ld_32x32b_x16(t_row + HALF + c, reinterpret_cast<uint32_t(&)[16]>(ga));
            }
            tmem_ld_wait();
            const long long gp_c = GP_NOW();
            GP_ADD(2, gp_c - gp_l);
            const int ocol0 = nt * HALF + c;
            const int m_warp0 = mt * p.rows_per_tile + quad * 32;
            const bool warp_ok = quad * 32 < p.rows_per_tile && m_warp0 < p.M;
            if (warp_ok && ocol0 + CH <= p.N && (row_ok || (p.use_tma_store && (CH == 32 || EPI_GROUPS == 4)))) {
              float xv[CH], gv[CH];
#pragma unroll
              for (int e = 0; e < CH; ++e) { xv[e] = __uint_as_float(xa[e]); gv[e] = __uint_as_float(ga[e]); }
              // (prefetching these 64 bias values before the TMEM load was measured: the extra live registers spill under the
              //  168-register cap of a 10-warp CTA and the epilogue gets 1.6x slower)
              if (p.bias) {
                const float4* bx = reinterpret_cast<const float4*>(p.bias + nt * BLOCK_N + c);
                const float4* bg = reinterpret_cast<const float4*>(p.bias + nt * BLOCK_N + HALF + c);
#pragma unroll
                for (int j = 0; j < CH / 4; ++j) {
                  const float4 a = __ldg(bx + j), b = __ldg(bg + j);
                  xv[4 * j] += a.x; xv[4 * j + 1] += a.y; xv[4 * j + 2] += a.z; xv[4 * j + 3] += a.w;
                  gv[4 * j] += b.x; gv[4 * j + 1] += b.y; gv[4 * j + 2] += b.z; gv[4 * j + 3] += b.w;
                }
              }
              uint4 o[CH / 8];
#pragma unroll
              for (int j = 0; j < CH / 8; ++j) {
                __half2 h0 = __floats2half2_rn(xv[8 * j + 0] * gelu_erf(gv[8 * j + 0]), xv[8 * j + 1] * gelu_erf(gv[8 * j + 1]));
                __half2 h1 = __floats2half2_rn(xv[8 * j + 2] * gelu_erf(gv[8 * j + 2]), xv[8 * j + 3] * gelu_erf(gv[8 * j + 3]));
                __half2 h2 = __floats2half2_rn(xv[8 * j + 4] * gelu_erf(gv[8 * j + 4]), xv[8 * j + 5] * gelu_erf(gv[8 * j + 5]));
                __half2 h3 = __floats2half2_rn(xv[8 * j + 6] * gelu_erf(gv[8 * j + 6]), xv[8 * j + 7] * gelu_erf(gv[8 * j + 7]));
                o[j].x = *reinterpret_cast<uint32_t*>(&h0); o[j].y = *reinterpret_cast<uint32_t*>(&h1);
                o[j].z = *reinterpret_cast<uint32_t*>(&h2); o[j].w = *reinterpret_cast<uint32_t*>(&h3);
              }
              GP_ADD(3, GP_NOW() - gp_c);
              if constexpr (CH == 32) {
                store_chunk32(o, m, ocol0, m_warp0, nullptr);
              } else if constexpr (EPI_GROUPS == 4) {
                store_chunk16(o, m, ocol0, m_warp0);
              } else {
                uint4* op = reinterpret_cast<uint4*>(p.out + m * p.ldo + ocol0);
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) op[j] = o[j];
              }
            } else if (row_ok) {
#pragma unroll
              for (int e = 0; e < CH; ++e) {
                const int oc = ocol0 + e;
                if (oc < p.N) {
                  float xv = __uint_as_float(xa[e]), gv = __uint_as_float(ga[e]);
                  if (p.bias) { xv += p.bias[nt * BLOCK_N + c + e]; gv += p.bias[nt * BLOCK_N + HALF + c + e]; }
                  p.out[m * p.ldo + oc] = __float2half_rn(xv * gelu_erf(gv));
                }
              }
            }
          }
        }
      } else {
        constexpr int CH = (BLOCK_N >= 32 && EPI_GROUPS == 2) ? 32 : 16;
        const float* gb = p.group_bias ? p.group_bias + (m / p.rows_per_group) * p.N : nullptr;
        uint4 preA[4], preB[4];
        bool have_pre = false;
#pragma unroll 1
        for (int c = half * CH; c < BLOCK_N; c += EPI_GROUPS * CH) {
          uint32_t acc[CH];
          const long long gp_l = GP_NOW();
          const int col0 = nt * BLOCK_N + c;
          // the bias row of this chunk is fetched BEFORE the TMEM load so that its global-load latency overlaps it
          float4 bvec[CH / 4];
          const bool bias_pre = p.bias != nullptr && col0 + CH <= p.N;
          if (bias_pre) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) bvec[j] = __ldg(bp + j);
          }
          if constexpr (CH == 32) tmem_ld_32x32b_x32(t_row + c, reinterpret_cast<uint32_t(&)[32]>(acc));
          else tmem_ld_32x32b_x16(t_row + c, reinterpret_cast<uint32_t(&)[16]>(acc));
          tmem_ld_wait();
          const long long gp_c = GP_NOW();
          GP_ADD(2, gp_c - gp_l);
          const int m_warp0 = mt * p.rows_per_tile + quad * 32;
          const bool warp_ok = quad * 32 < p.rows_per_tile && m_warp0 < p.M;
          if (!warp_ok || col0 >= p.N) continue;
          const bool fast = col0 + CH <= p.N && col0 + CH <= p.vt_col_start;
          if (fast && (row_ok || (p.use_tma_store && (CH == 32 || EPI_GROUPS == 4)))) {
            // ---------------- fast path: full chunk, row-major output ----------------
            float v[CH];
#pragma unroll
            for (int e = 0; e < CH; ++e) v[e] = __uint_as_float(acc[e]);
            if (bias_pre) {
#pragma unroll
              for (int j = 0; j < CH / 4; ++j) {
                const float4 b = bvec[j];
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            if (gb && row_ok) {
              const float4* bp = reinterpret_cast<const float4*>(gb + col0);
#pragma unroll
              for (int j = 0; j < CH / 4; ++j) {
                const float4 b = __ldg(bp + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            uint8_t* slot = nullptr;
            if constexpr (CH == 32) {
              if (p.use_tma_store && (p.residual || p.residual2)) {
                slot = acquire_slot();
                if (!have_pre) {
                  if (p.residual) residual_fetch(p.residual, p.ldr, m_warp0, col0, preA);
                  if (p.residual2) residual_fetch(p.residual2, p.ldr2, m_warp0, col0, preB);
                }
                if (p.residual) residual_add(preA, slot, v);
                if (p.residual2) residual_add(preB, slot, v);
                // software pipeline: issue the next chunk's residual loads now, consume them after its TMEM load
                const int c_next = c + 2 * CH;
                const int col_next = nt * BLOCK_N + c_next;
                have_pre = c_next < BLOCK_N && col_next + CH <= p.N && col_next + CH <= p.vt_col_start;
                if (have_pre) {
                  if (p.residual) residual_fetch(p.residual, p.ldr, m_warp0, col_next, preA);
                  if (p.residual2) residual_fetch(p.residual2, p.ldr2, m_warp0, col_next, preB);
                }
              }
            }
            if (slot == nullptr) {
              if (p.residual && row_ok) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.ldr + col0);
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) {
                  const uint4 r = rp[j];
                  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                  for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(h[q]); v[8 * j + 2 * q] += f.x; v[8 * j + 2 * q + 1] += f.y; }
                }
              }
              if (p.residual2 && row_ok) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.residual2 + m * p.ldr2 + col0);
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) {
                  const uint4 r = rp[j];
                  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                  for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(h[q]); v[8 * j + 2 * q] += f.x; v[8 * j + 2 * q + 1] += f.y; }
                }
              }
            }
            uint4 o[CH / 8];
#pragma unroll
            for (int j = 0; j < CH / 8; ++j) {
              __half2 h0 = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]), h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
              __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]), h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
              o[j].x = *reinterpret_cast<uint32_t*>(&h0); o[j].y = *reinterpret_cast<uint32_t*>(&h1);
              o[j].z = *reinterpret_cast<uint32_t*>(&h2); o[j].w = *reinterpret_cast<uint32_t*>(&h3);
            }
            GP_ADD(3, GP_NOW() - gp_c);
            if constexpr (CH == 32) {
              store_chunk32(o, m, col0, m_warp0, slot);
            } else if constexpr (EPI_GROUPS == 4) {
              store_chunk16(o, m, col0, m_warp0);
            } else {
              uint4* op = reinterpret_cast<uint4*>(p.out + m * p.ldo + col0);
#pragma unroll
              for (int j = 0; j < CH / 8; ++j) op[j] = o[j];
            }
          } else if (CH == 32 && p.vt_tma && col0 >= p.vt_col_start && col0 + CH <= p.N) {
            // ---------------- V^T store, staged: the warp's 32 tokens x 32 columns are transposed through its staging slot ([column][token],
            // 64-byte rows) and leave as ONE bulk tensor store into out_vt viewed as [BF * heads * d, S] (the per-element variant below
            // issued 32 two-byte global stores per thread: 65536 x 960 x 320 ran 90 us against 30 us for a row-major GEMM of 1/3 the columns)
            if constexpr (CH == 32) {
              uint8_t* slot = acquire_slot();
              // 2 x 2 transposes between neighbouring lanes: an even lane keeps column 2j of tokens (lane, lane + 1), the odd lane column 2j + 1
              // of tokens (lane - 1, lane): 16 shuffles + 16 conflict-free 4-byte st.shared per thread instead of 32 two-byte stores
              const bool odd = lane & 1;
              uint32_t* wp = reinterpret_cast<uint32_t*>(slot) + (lane >> 1);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float v0 = __uint_as_float(acc[2 * j]), v1 = __uint_as_float(acc[2 * j + 1]);
                if (p.bias) { v0 += __ldg(p.bias + col0 + 2 * j); v1 += __ldg(p.bias + col0 + 2 * j + 1); }
                const float got = __shfl_xor_sync(0xffffffffu, odd ? v0 : v1, 1);
                wp[(2 * j + (odd ? 1 : 0)) * 16] = odd ? pack_h2(got, v1) : pack_h2(v0, got);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                const int bfi = m_warp0 / p.vt_S;
                tma_store_2d(&p.tmVt, slot, m_warp0 - bfi * p.vt_S, bfi * (p.N - p.vt_col_start) + (col0 - p.vt_col_start));
                tma_store_commit();
              }
              ++epi_count;
            }
          } else if (col0 >= p.vt_col_start) {
            // ---------------- V^T store: out_vt[((bf*heads + h)*d + dd)*ld + s]; lanes = consecutive s -> 64-byte segments ----------------
            if (!row_ok) continue;
            const long long bf = m / p.vt_S;
            const int s = static_cast<int>(m % p.vt_S);
            const int cv0 = col0 - p.vt_col_start;
            int h = cv0 / p.vt_d, dd = cv0 % p.vt_d;
            __half* vbase = p.out_vt + (bf * p.vt_heads) * static_cast<long long>(p.vt_d) * p.vt_ld + s;
#pragma unroll
            for (int e = 0; e < CH; ++e) {
              if (col0 + e < p.N) {
                float v = __uint_as_float(acc[e]);
                if (p.bias) v += p.bias[col0 + e];
                vbase[(static_cast<long long>(h) * p.vt_d + dd) * p.vt_ld] = __float2half_rn(v);
              }
              if (++dd == p.vt_d) { dd = 0; ++h; }
            }
          } else {
            // ---------------- generic masked path (right-edge tiles, chunks straddling vt_col_start) ----------------
            if (!row_ok) continue;
#pragma unroll
            for (int e = 0; e < CH; ++e) {
              const int col = col0 + e;
              if (col >= p.N) continue;
              float v = __uint_as_float(acc[e]);
              if (p.bias) v += p.bias[col];
              if (col >= p.vt_col_start) {
                const long long bf = m / p.vt_S;
                const int s = static_cast<int>(m % p.vt_S);
                const int cv = col - p.vt_col_start;
                p.out_vt[((bf * p.vt_heads + cv / p.vt_d) * p.vt_d + cv % p.vt_d) * p.vt_ld + s] = __float2half_rn(v);
              } else {
                if (gb) v += gb[col];
                if (p.residual) v += __half2float(p.residual[m * p.ldr + col]);
                if (p.residual2) v += __half2float(p.residual2[m * p.ldr2 + col]);
                p.out[m * p.ldo + col] = __float2half_rn(v);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    // the staging slots must stay valid until the bulk stores have READ them; their global writes complete with the grid
    // (kernel boundary / griddepcontrol.wait of the dependent), as in CUTLASS' store_tail
    if ((p.use_tma_store || p.vt_tma) && lane == 0) tma_store_wait_read<0>();
#ifdef FZ_GEMM_PROFILE
    if (gp_on && warp == 2 && lane == 0) {
      g_gemm_dbg[0] = GP_NOW() - gp_e0;
      for (int i = 1; i < 8; ++i) g_gemm_dbg[i] = gp[i];
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int EG>
static int launch_tapgemm_eg(const TapGemmParams& p, cudaStream_t stream) {
  using Cfg = TapGemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    FZ_CUDA(cudaFuncSetAttribute(tapgemm_kernel<BN, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = std::min(tiles, num_sms());
  FZ_CUDA(launch_pdl(tapgemm_kernel<BN, EG>, dim3(grid), dim3(64 + 128 * EG), Cfg::kSmemBytes, stream, p));
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

template <int BN>
static int launch_tapgemm(const TapGemmParams& p, cudaStream_t stream) {
  if constexpr (BN >= 64) {
    // 16 epilogue warps (16-column chunks, 96 registers) for the ALU-bound GEGLU epilogue.  Measured and rejected for the short-K
    // row-major GEMMs: their chunks are latency-bound (~860 cycles per chunk whether it is 16 or 32 columns wide), so twice the warps on
    // half-size chunks gain nothing (65536x320x320: epilogue warp 29.1k -> 39.5k cycles).
    if (p.use_tma_store == 2 && p.mode == FZ_EPI_GEGLU) return launch_tapgemm_eg<BN, 4>(p, stream);
  }
  return launch_tapgemm_eg<BN, 2>(p, stream);
}

static int pick_block_n(int gemm_cols, int mode, int forced, int m_tiles) {
  if (forced > 0) return forced;
  static const int cands[] = {256, 160, 128, 64, 32, 16};
  if (mode == FZ_EPI_GEGLU) return (gemm_cols % 256 == 0) ? 256 : ((gemm_cols % 160 == 0) ? 160 : ((gemm_cols % 128 == 0) ? 128 : ((gemm_cols % 64 == 0) ? 64 : 32)));
  // cost model: a tile costs ~ BLOCK_N MMA columns (+ a fixed part), tiles run in waves of one per SM.
  const int sms = num_sms();
  int best = 16;
  double best_cost = 1e30;
  for (int bn : cands) {
    const long long n_tiles = (gemm_cols + bn - 1) / bn;
    const long long tiles = n_tiles * m_tiles;
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = static_cast<double>(waves) * (bn + 24.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// Selection blocks for folding skip tensors into the MMA pipeline: E[j][n][k] = 1 if n == 64 j + k (static device memory, filled once).
__device__ __half g_eye[4][256][64];
__global__ void eye_init_kernel() {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * 256 * 64; i += gridDim.x * blockDim.x) {
    const int k = i % 64, n = (i / 64) % 256, j = i / (64 * 256);
    (&g_eye[0][0][0])[i] = __float2half_rn(n == 64 * j + k ? 1.f : 0.f);
  }
}

// One-time fill of the selection table.  It synchronises the stream, which is illegal under stream capture: fz_init() runs it up front
// (engine construction), so that the first GEMM with a skip tensor may already sit inside a CUDA-graph capture.
static int eye_table(__half** out, cudaStream_t stream) {
  static __half* eye = nullptr;
  if (!eye) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    FZ_CUDA(cudaStreamIsCapturing(stream, &cs));
    FZ_CHECK_ARG(cs == cudaStreamCaptureStatusNone, "fz_init() must run once before the library is used under CUDA-graph capture");
    __half* e = nullptr;
    FZ_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&e), g_eye));
    eye_init_kernel<<<64, 256, 0, stream>>>();
    FZ_CUDA(cudaGetLastError());
    // the table is read by TMA of kernels on ANY stream afterwards: make the one-time fill visible before returning
    FZ_CUDA(cudaStreamSynchronize(stream));
    eye = e;
  }
  *out = eye;
  return FZ_OK;
}

static int fold_residuals(TapGemmParams& p, int bn, cudaStream_t stream) {
  p.n_res = 0;
  p.res_kblocks = 0;
  if (p.mode != FZ_EPI_ROWMAJOR || p.vt_col_start != INT_MAX || (!p.residual && !p.residual2)) return FZ_OK;
  const __half* rs[2] = {p.residual, p.residual2};
  const long long lds[2] = {p.ldr, p.ldr2};
  for (int i = 0; i < 2; ++i)
    if (rs[i] && ((reinterpret_cast<uintptr_t>(rs[i]) & 15) != 0 || lds[i] % 8 != 0)) return FZ_OK;  // not TMA-addressable: epilogue path
  __half* eye = nullptr;
  if (int rc = eye_table(&eye, stream)) return rc;
  {
    uint64_t dims[3] = {64, 256, 4};
    uint64_t strides[2] = {64, 64 * 256};
    uint32_t box[3] = {64, static_cast<uint32_t>(bn), 1};
    if (int rc = encode_tmap_f16(&p.tmE, eye, 3, dims, strides, box, true)) return rc;
  }
  for (int i = 0; i < 2; ++i) {
    if (!rs[i]) continue;
    uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lds[i])};
    uint32_t box[2] = {kBlockK, kBlockM};
    if (int rc = encode_tmap_f16(&p.tmR[p.n_res], rs[i], 2, dims, strides, box, true)) return rc;
    ++p.n_res;
  }
  p.res_kblocks = (bn + kBlockK - 1) / kBlockK;
  p.residual = nullptr;
  p.residual2 = nullptr;
  return FZ_OK;
}

static int dispatch_tapgemm(TapGemmParams& p, int gemm_cols, int forced_bn, cudaStream_t stream) {
  if (int rc = check_single_device()) return rc;
  // output tensor map for the TMA-store epilogue ([M, N_out] row-major, row stride ldo)
  p.use_tma_store = 0;
  if (p.rows_per_tile % 32 == 0 && p.ldo % 8 == 0 && p.N % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
      std::min(p.N, p.vt_col_start) >= 32) {
    uint64_t dims[2] = {static_cast<uint64_t>(std::min(p.N, p.vt_col_start)), static_cast<uint64_t>(p.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(p.ldo)};
    uint32_t box[2] = {32, 32};
    if (int rc = encode_tmap_f16_sw(&p.tmC, p.out, 2, dims, strides, box, 64)) return rc;
    p.use_tma_store = 1;
    if (p.N % 16 == 0 && (p.vt_col_start == INT_MAX || p.vt_col_start % 16 == 0)) {
      uint32_t box16[2] = {16, 32};
      if (int rc = encode_tmap_f16_sw(&p.tmC16, p.out, 2, dims, strides, box16, 0)) return rc;
      p.use_tma_store = 2;  // both maps valid: the GEGLU launch may take the 16-epilogue-warp instantiation
    }
  }
  p.vt_tma = 0;
  if (p.out_vt && p.vt_S % 32 == 0 && p.rows_per_tile % 32 == 0 && p.vt_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out_vt) & 15) == 0 &&
      p.vt_col_start % 32 == 0 && p.N - p.vt_col_start == p.vt_heads * p.vt_d && p.M % p.vt_S == 0) {
    uint64_t dims[2] = {static_cast<uint64_t>(p.vt_S), static_cast<uint64_t>(p.M / p.vt_S) * (p.N - p.vt_col_start)};
    uint64_t strides[1] = {static_cast<uint64_t>(p.vt_ld)};
    uint32_t box[2] = {32, 32};
    if (int rc = encode_tmap_f16_sw(&p.tmVt, p.out_vt, 2, dims, strides, box, 0)) return rc;
    p.vt_tma = 1;
  }
  const int bn = pick_block_n(gemm_cols, p.mode, forced_bn, p.m_tiles);
  p.n_tiles = (gemm_cols + bn - 1) / bn;
  if (int rc = fold_residuals(p, bn, stream)) return rc;
  switch (bn) {
    case 256: return launch_tapgemm<256>(p, stream);
    case 160: return launch_tapgemm<160>(p, stream);
    case 128: return launch_tapgemm<128>(p, stream);
    case 64: return launch_tapgemm<64>(p, stream);
    case 32: return launch_tapgemm<32>(p, stream);
    case 16: return launch_tapgemm<16>(p, stream);
    default: set_error("unsupported BLOCK_N %d", bn); return FZ_ERR_INVALID;
  }
}

static int fill_epilogue(TapGemmParams& p, const fz_epilogue_t* e, int M, int gemm_cols) {
  p.bias = nullptr; p.group_bias = nullptr; p.rows_per_group = 1; p.residual = nullptr; p.ldr = 0; p.residual2 = nullptr; p.ldr2 = 0;
  p.mode = FZ_EPI_ROWMAJOR; p.vt_col_start = INT_MAX; p.out_vt = nullptr; p.vt_S = p.vt_d = p.vt_heads = p.vt_ld = 1;
  p.N = gemm_cols;
  if (!e) return FZ_OK;
  p.bias = e->bias; p.group_bias = e->group_bias; p.rows_per_group = e->rows_per_group > 0 ? e->rows_per_group : 1;
  p.residual = static_cast<const __half*>(e->residual); p.ldr = e->ldr;
  p.residual2 = static_cast<const __half*>(e->residual2); p.ldr2 = e->ldr2;
  p.mode = e->mode;
  if (e->mode == FZ_EPI_GEGLU) {
    FZ_CHECK_ARG(gemm_cols % 2 == 0, "GEGLU needs an even number of GEMM columns");
    p.N = gemm_cols / 2;
  }
  if (e->out_vt) {
    FZ_CHECK_ARG(e->vt_S > 0 && e->vt_d > 0 && e->vt_heads > 0 && M % e->vt_S == 0, "bad V^T geometry");
    p.vt_col_start = e->vt_col_start; p.out_vt = static_cast<__half*>(e->out_vt);
    p.vt_S = e->vt_S; p.vt_d = e->vt_d; p.vt_heads = e->vt_heads; p.vt_ld = e->vt_ld > 0 ? e->vt_ld : e->vt_S;
  }
  return FZ_OK;
}

}  // namespace fz

using namespace fz;

// Development aid (not part of include/fatezero_b200.h): cycle counters of the last tap-GEMM launch, zeros unless built with
// -DFZ_GEMM_PROFILE (tools/profile_gemm_epilogue.py).
extern "C" int fz_debug_gemm_counters(long long* host16) {
  FZ_CUDA(cudaDeviceSynchronize());
  FZ_CUDA(cudaMemcpyFromSymbol(host16, g_gemm_dbg, sizeof(long long) * 16));
  return FZ_OK;
}

// One-time device-side initialisation (constant tables).  Idempotent; must have run before the first call under stream capture.
extern "C" int fz_init(cudaStream_t stream) {
  if (int rc = check_single_device()) return rc;
  __half* eye = nullptr;
  return eye_table(&eye, stream);
}

// D[M,N] = A[M,K] * W[N,K]^T (+epilogue).  A, W fp16 row-major (lda, ldw in elements, multiples of 8).
extern "C" int fz_gemm_f16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const fz_epilogue_t* epi,
                           void* out, long long ldo, int force_block_n, cudaStream_t stream) {
  FZ_CHECK_ARG(A && W && out, "fz_gemm_f16: null pointer");
  FZ_CHECK_ARG(M > 0 && N > 0 && K > 0, "fz_gemm_f16: bad shape %d %d %d", M, N, K);
  FZ_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && K % 8 == 0, "fz_gemm_f16: lda/ldw/K must be multiples of 8 (16-byte TMA strides)");
  TapGemmParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda)};
    uint32_t box[2] = {kBlockK, kBlockM};
    if (int rc = encode_tmap_f16(&p.tmA, A, 2, dims, strides, box, true)) return rc;
  }
  const int rc0 = fill_epilogue(p, epi, M, N);
  if (rc0) return rc0;
  p.m_tiles = (M + kBlockM - 1) / kBlockM;
  const int bn = pick_block_n(N, p.mode, force_block_n, p.m_tiles);
  {
    uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N), 1};
    uint64_t strides[2] = {static_cast<uint64_t>(ldw), static_cast<uint64_t>(ldw) * N};
    uint32_t box[3] = {kBlockK, static_cast<uint32_t>(bn), 1};
    if (int rc = encode_tmap_f16(&p.tmB, W, 3, dims, strides, box, true)) return rc;
  }
  p.a_rank = 2; p.M = M; p.rows_per_tile = kBlockM; p.a_box_bytes = kATileBytes;
  p.k_blocks = (K + kBlockK - 1) / kBlockK; p.num_taps = 1;
  p.m_tiles = (M + kBlockM - 1) / kBlockM;
  p.ndecomp = 1; p.dimsz[0] = INT_MAX;
  p.out = static_cast<__half*>(out); p.ldo = ldo;
  return dispatch_tapgemm(p, N, bn, stream);
}

// 3x3 convolution, padding 1, stride 1 or 2, NHWC fp16.  x: [NB, H, W, Cin] (pixel stride ldx >= Cin),
// w: [9][Cout][Cin] (tap-major, tap = ky*3+kx), out: [NB, Ho, Wo, Cout] row-major with row stride ldo.
// asym_pad (stride 2 only): the input is padded by one pixel on the right / bottom ONLY (diffusers Downsample2D with padding = 0 in the VAE
// encoder: F.pad(x, (0, 1, 0, 1)) then a stride-2 conv without padding) instead of symmetrically.
static int conv3x3_impl(const void* x, long long ldx, int NB, int H, int W, int Cin, const void* w, int Cout, int stride, int asym_pad,
                        const fz_epilogue_t* epi, void* out, long long ldo, int force_block_n, cudaStream_t stream) {
  FZ_CHECK_ARG(x && w && out, "fz_conv3x3: null pointer");
  FZ_CHECK_ARG(stride == 1 || stride == 2, "fz_conv3x3: stride must be 1 or 2");
  FZ_CHECK_ARG(Cin % 8 == 0 && ldx % 8 == 0, "fz_conv3x3: Cin/ldx must be multiples of 8");
  FZ_CHECK_ARG(stride == 1 || (H % 2 == 0 && W % 2 == 0), "fz_conv3x3: stride 2 needs even H, W");
  const int Ho = H / stride, Wo = W / stride;
  FZ_CHECK_ARG(Wo <= 128 || Wo % 128 == 0, "fz_conv3x3: output width %d must be <= 128 or a multiple of 128", Wo);
  FZ_CHECK_ARG(!asym_pad || stride == 2, "fz_conv3x3: asymmetric padding is the stride-2 downsample variant");
  // box over (x, y, n): the full output width (or 128-pixel row segments of wider images: VAE resolutions), as many rows / images as fit
  // 128 GEMM rows
  int bw = std::min(Wo, 128), bh = (bw == Wo) ? std::min(Ho, 128 / bw) : 1;
  while (Ho % bh) --bh;
  int bn_img = (bh == Ho) ? std::min(NB, 128 / (bw * bh)) : 1;
  while (NB % bn_img) --bn_img;
  TapGemmParams p;
  memset(&p, 0, sizeof(p));
  const int M = NB * Ho * Wo;
  const int rc0 = fill_epilogue(p, epi, M, Cout);
  if (rc0) return rc0;
  if (stride == 1) {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t strides[3] = {(uint64_t)ldx, (uint64_t)ldx * W, (uint64_t)ldx * W * H};
    uint32_t box[4] = {kBlockK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn_img};
    if (int rc = encode_tmap_f16(&p.tmA, x, 4, dims, strides, box, true)) return rc;
    p.a_rank = 4;
    for (int t = 0; t < 9; ++t) {
      p.tap_off[t][0] = 0; p.tap_off[t][1] = t % 3 - 1; p.tap_off[t][2] = t / 3 - 1; p.tap_off[t][3] = 0; p.tap_off[t][4] = 0;
    }
  } else {
    // stride 2: view the input as (c|px : 2*ldx, x' : W/2, y' : H/2, n, py : 2); tap (ky,kx) reads phase (py,px) shifted by {-1,0}
    uint64_t dims[5] = {(uint64_t)(ldx + Cin), (uint64_t)Wo, (uint64_t)Ho, (uint64_t)NB, 2};
    uint64_t strides[4] = {(uint64_t)2 * ldx, (uint64_t)2 * ldx * W, (uint64_t)ldx * W * H, (uint64_t)ldx * W};
    uint32_t box[5] = {kBlockK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn_img, 1};
    if (int rc = encode_tmap_f16(&p.tmA, x, 5, dims, strides, box, true)) return rc;
    p.a_rank = 5;
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      // symmetric padding 1: input row 2y + ky - 1 -> phase (ky != 1), shift -1 for ky = 0; right/bottom-only padding: input row 2y + ky ->
      // phase ky & 1, shift +1 for ky = 2 (the zero row / column beyond the edge is TMA's out-of-bounds fill)
      const int py = asym_pad ? (ky & 1) : ((ky == 1) ? 0 : 1), px = asym_pad ? (kx & 1) : ((kx == 1) ? 0 : 1);
      p.tap_off[t][0] = px * (int)ldx;
      p.tap_off[t][1] = asym_pad ? (kx >> 1) : ((kx == 0) ? -1 : 0);
      p.tap_off[t][2] = asym_pad ? (ky >> 1) : ((ky == 0) ? -1 : 0);
      p.tap_off[t][3] = 0;
      p.tap_off[t][4] = py;
    }
  }
  p.rows_per_tile = bw * bh * bn_img;
  p.m_tiles = M / p.rows_per_tile;
  const int bn = pick_block_n(Cout, p.mode, force_block_n, p.m_tiles);
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
    uint64_t strides[2] = {(uint64_t)Cin, (uint64_t)Cin * Cout};
    uint32_t box[3] = {kBlockK, (uint32_t)bn, 1};
    if (int rc = encode_tmap_f16(&p.tmB, w, 3, dims, strides, box, true)) return rc;
  }
  p.M = M; p.rows_per_tile = bw * bh * bn_img; p.a_box_bytes = p.rows_per_tile * 128;
  p.k_blocks = (Cin + kBlockK - 1) / kBlockK; p.num_taps = 9;
  p.m_tiles = M / p.rows_per_tile;
  p.ndecomp = 3; p.dimsz[0] = Wo; p.dimsz[1] = Ho; p.dimsz[2] = NB;
  p.out = static_cast<__half*>(out); p.ldo = ldo;
  return dispatch_tapgemm(p, Cout, bn, stream);
}

extern "C" int fz_conv3x3_nhwc_f16(const void* x, long long ldx, int NB, int H, int W, int Cin, const void* w, int Cout, int stride,
                                   const fz_epilogue_t* epi, void* out, long long ldo, int force_block_n, cudaStream_t stream) {
  return conv3x3_impl(x, ldx, NB, H, W, Cin, w, Cout, stride, 0, epi, out, ldo, force_block_n, stream);
}

extern "C" int fz_conv3x3_down_asym_nhwc_f16(const void* x, long long ldx, int NB, int H, int W, int Cin, const void* w, int Cout,
                                             const fz_epilogue_t* epi, void* out, long long ldo, int force_block_n, cudaStream_t stream) {
  return conv3x3_impl(x, ldx, NB, H, W, Cin, w, Cout, 2, 1, epi, out, ldo, force_block_n, stream);
}

// Temporal Conv1d(k=3, padding 1, no bias) over the frame axis: x [B, F, HW, Cin] fp16 (row stride ldx), w [3][Cout][Cin].
// out[b,f,p,:] = sum_t w[t] * x[b, f+t-1, p, :]  (+ epilogue: residual = identity skip of the LoRA, group_bias = time embedding).
// halo = 1 (frame-sharded execution): x is [B, F+2, HW, Cin] whose frames 0 and F+1 hold the neighbour ranks' boundary frames (zeros at
// the ends of the clip, which is what the zero padding of the un-sharded conv reads); the F output frames are the interior ones.
static int tconv3_impl(const void* x, long long ldx, int B, int F, int HW, int Cin, const void* w, int Cout, const fz_epilogue_t* epi,
                       void* out, long long ldo, int force_block_n, int halo, cudaStream_t stream) {
  FZ_CHECK_ARG(x && w && out, "fz_tconv3: null pointer");
  FZ_CHECK_ARG(Cin % 8 == 0 && ldx % 8 == 0, "fz_tconv3: Cin/ldx must be multiples of 8");
  int bp = std::min(HW, 128);
  while (HW % bp) --bp;
  int bf = (bp == HW) ? std::min(F, 128 / bp) : 1;
  while (F % bf) --bf;
  TapGemmParams p;
  memset(&p, 0, sizeof(p));
  const int M = B * F * HW;
  const int rc0 = fill_epilogue(p, epi, M, Cout);
  if (rc0) return rc0;
  const int Fx = F + 2 * halo;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)HW, (uint64_t)Fx, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)ldx, (uint64_t)ldx * HW, (uint64_t)ldx * HW * Fx};
    uint32_t box[4] = {kBlockK, (uint32_t)bp, (uint32_t)bf, 1};
    if (int rc = encode_tmap_f16(&p.tmA, x, 4, dims, strides, box, true)) return rc;
  }
  p.a_rank = 4;
  for (int t = 0; t < 3; ++t) { p.tap_off[t][2] = t - 1 + halo; }
  p.rows_per_tile = bp * bf;
  p.m_tiles = M / p.rows_per_tile;
  const int bn = pick_block_n(Cout, p.mode, force_block_n, p.m_tiles);
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 3};
    uint64_t strides[2] = {(uint64_t)Cin, (uint64_t)Cin * Cout};
    uint32_t box[3] = {kBlockK, (uint32_t)bn, 1};
    if (int rc = encode_tmap_f16(&p.tmB, w, 3, dims, strides, box, true)) return rc;
  }
  p.M = M; p.rows_per_tile = bp * bf; p.a_box_bytes = p.rows_per_tile * 128;
  p.k_blocks = (Cin + kBlockK - 1) / kBlockK; p.num_taps = 3;
  p.m_tiles = M / p.rows_per_tile;
  p.ndecomp = 3; p.dimsz[0] = HW; p.dimsz[1] = F; p.dimsz[2] = B;
  p.out = static_cast<__half*>(out); p.ldo = ldo;
  return dispatch_tapgemm(p, Cout, bn, stream);
}

extern "C" int fz_tconv3_f16(const void* x, long long ldx, int B, int F, int HW, int Cin, const void* w, int Cout, const fz_epilogue_t* epi,
                             void* out, long long ldo, int force_block_n, cudaStream_t stream) {
  return tconv3_impl(x, ldx, B, F, HW, Cin, w, Cout, epi, out, ldo, force_block_n, 0, stream);
}

extern "C" int fz_tconv3_halo_f16(const void* x_ext, long long ldx, int B, int F, int HW, int Cin, const void* w, int Cout,
                                  const fz_epilogue_t* epi, void* out, long long ldo, int force_block_n, cudaStream_t stream) {
  return tconv3_impl(x_ext, ldx, B, F, HW, Cin, w, Cout, epi, out, ldo, force_block_n, 1, stream);
}
