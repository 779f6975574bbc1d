// Microbenchmark: cost of back-to-back tcgen05.mma (kind::f16, M=128, K=16) as a function of N, accumulator dependency,
// A-operand source (shared memory descriptor vs TMEM) and concurrent tcgen05.ld traffic.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fatezero_b200/csrc tools/micro/umma_rate.cu -o tools/micro/bin/umma_rate -lcuda
#include <cstdio>
#include "fz_common.cuh"
using namespace fz;

template <int N, int NACC, bool TS, bool LD>
__global__ void __launch_bounds__(160, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint64_t desc_hi = umma_desc_k_sw128(0);
      const uint32_t a_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
      const uint32_t b_lo = (smem_u32(smem + 16384) & 0x3FFFF) >> 4;
      const uint32_t idesc = umma_idesc_f16(128, N);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem_base + (it % NACC) * N;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) umma_f16_ts(d, tmem_base + 448 + k * 8, desc_hi | (b_lo + 2 * k), idesc, 1u);
          else umma_f16_ss(d, desc_hi | (a_lo + 2 * k), desc_hi | (b_lo + 2 * k), idesc, 1u);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
      *reinterpret_cast<volatile int*>(&stop) = 1;
    }
  } else if (LD) {
    // warps 1..4: keep reading 64 columns of an unrelated TMEM region (like softmax warps do)
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    while (!*reinterpret_cast<volatile int*>(&stop)) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + 384, r);
      tmem_ld_wait();
      acc += r[0] ^ r[31];
    }
    if (acc == 0x12345678u) out[200] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

template <int N, int NACC, bool TS, bool LD>
void run(const char* name, long long* d_out, int iters) {
  auto kfn = rate_kernel<N, NACC, TS, LD>;
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) kfn<<<148, 160, 100 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  const double per = double(mx) / (iters * 4.0);
  printf("%-34s N=%3d acc=%d  %7.1f clk/MMA (min-CTA %7.1f)  math %5.1f clk  -> %5.1f%% of 8192 flop/clk  [%s]\n", name, N, NACC, per,
         double(mn) / (iters * 4.0), N / 2.0, 100.0 * (N / 2.0) / per, cudaGetErrorString(e));
}


// GEMM-mainloop shaped issue stream: per k-block [try_wait on a completed barrier, fence, KMMA x UMMA, commit]; the data never
// changes, so this isolates the issuing thread's overhead from memory effects.
template <int N, int KMMA, bool WARP>
__global__ void __launch_bounds__(160, 1) mainloop_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, full_bar, empty_bar[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&full_bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&empty_bar[i], 1);
    fence_mbar_init();
    mbar_arrive(&full_bar);  // phase 0 complete: waits on parity 0 succeed immediately
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    const bool leader = elect_one();
    if (WARP || leader) {
      const uint64_t desc_hi = umma_desc_k_sw128(0);
      const uint32_t a_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
      const uint32_t b_lo = (smem_u32(smem + 32768) & 0x3FFFF) >> 4;
      const uint32_t idesc = umma_idesc_f16(128, N);
      int stage = 0;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar, 0);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < KMMA; ++k)
            umma_f16_ss(tmem_base, desc_hi | (a_lo + stage * 64 + 2 * (k & 3) + (k >> 2) * 1024), desc_hi | (b_lo + stage * 64 + 2 * (k & 3) + (k >> 2) * 2048), idesc, 1u);
          umma_commit(&empty_bar[stage]);
        }
        if (WARP) __syncwarp();
        if (++stage == 8) stage = 0;
      }
      if (leader) {
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

template <int N, int KMMA, bool WARP>
void run_mainloop(const char* name, long long* d_out, int iters) {
  auto kfn = mainloop_kernel<N, KMMA, WARP>;
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) kfn<<<148, 160, 200 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = double(mx) / iters;
  printf("%-40s N=%3d %d UMMA/k-block: %7.1f clk per k-block, math %6.1f -> %5.1f%%  [%s]\n", name, N, KMMA, per, KMMA * N / 2.0,
         100.0 * KMMA * N / 2.0 / per, cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 4096);
  cudaMemset(d_out, 0, 4096);
  const int it = 2000;
  run<256, 1, false, false>("SS dependent", d_out, it);
  run<128, 1, false, false>("SS dependent", d_out, it);
  run<128, 2, false, false>("SS 2 accumulators", d_out, it);
  run<64, 1, false, false>("SS dependent", d_out, it);
  run<64, 2, false, false>("SS 2 accumulators", d_out, it);
  run<64, 4, false, false>("SS 4 accumulators", d_out, it);
  run<48, 1, false, false>("SS dependent", d_out, it);
  run<48, 2, false, false>("SS 2 accumulators", d_out, it);
  run<48, 1, true, false>("TS dependent", d_out, it);
  run<48, 2, true, false>("TS 2 accumulators", d_out, it);
  run<64, 2, true, false>("TS 2 accumulators", d_out, it);
  run<128, 1, true, false>("TS dependent", d_out, it);
  run<128, 1, false, true>("SS dependent + tcgen05.ld traffic", d_out, it);
  run<48, 2, true, true>("TS 2 acc + tcgen05.ld traffic", d_out, it);
  run<256, 1, false, true>("SS dependent + tcgen05.ld traffic", d_out, it);
  run_mainloop<160, 4, false>("mainloop, single thread", d_out, it);
  run_mainloop<160, 8, false>("mainloop, single thread", d_out, it);
  run_mainloop<256, 4, false>("mainloop, single thread", d_out, it);
  run_mainloop<256, 8, false>("mainloop, single thread", d_out, it);
  run_mainloop<160, 4, true>("mainloop, convergent warp", d_out, it);
  run_mainloop<128, 3, true>("mainloop, convergent warp", d_out, it);
  run_mainloop<128, 3, false>("mainloop, single thread", d_out, it);
  run_mainloop<64, 4, false>("mainloop, single thread", d_out, it);
  return 0;
}
