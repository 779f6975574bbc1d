// Microbenchmark: throughput of exp2 variants per SM (all 4 schedulers busy, 8 warps per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/micro/exp_rate.cu -o tools/micro/bin/exp_rate
#include <cstdio>
#include <cuda_fp16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float poly2(float y) {  // 2^y for y <= 0 on the FMA/ALU pipes (degree-3 minimax on the fraction)
  float yr = __fadd_rd(y, 12582912.f);
  yr = fmaxf(yr, 12582912.f - 125.f);
  const float fl = yr - 12582912.f;
  const float f = y - fl;
  float p = fmaf(0.0790198f, f, 0.2241248f);
  p = fmaf(p, f, 0.6967632f);
  p = fmaf(p, f, 0.9998881f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(yr) << 23));
}
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -seed * (threadIdx.x % 7 + i) * 0.1f;
  unsigned h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = 0xb800b400u + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.5f;                       // 8 MUFU (+8 FADD)
      if (MODE == 1 && i < 4) h[i] = ex2h2(h[i]) ^ 0x80008000u;      // 4 packed = 8 exps
      if (MODE == 2) a[i] = poly2(a[i]) - 1.5f;                      // 8 polynomial
      if (MODE == 3) a[i] = ((i & 3) == 3 ? poly2(a[i]) : ex2f(a[i])) - 1.5f;  // 1/4 polynomial
      if (MODE == 4) a[i] = ((i & 1) ? poly2(a[i]) : ex2f(a[i])) - 1.5f;       // 1/2 polynomial
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += __uint_as_float(h[i]);
  if (s == 1234.5f) out[0] = s;
}
template <int MODE>
void run(const char* name, float* d) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 1024>>>(d, iters, 1.f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(d, iters, 1.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double exps = 148.0 * 2 * 1024 * iters * 8;
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%-28s %7.3f ms  %6.2f exp/clk/SM (at %d MHz nominal)\n", name, ms, exps / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
}
int main() {
  float* d; cudaMalloc(&d, 64);
  run<0>("ex2.approx.ftz.f32", d);
  run<1>("ex2.approx.f16x2", d);
  run<2>("polynomial (FMA pipe)", d);
  run<3>("3/4 MUFU + 1/4 polynomial", d);
  run<4>("1/2 MUFU + 1/2 polynomial", d);
  return 0;
}
