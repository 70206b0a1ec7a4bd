// Microbenchmark: MUFU throughput per SM for tanh.approx / ex2.approx / rcp.approx (independent chains, 16 warps per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.1f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      else y = fmaf(x[i], 1.0001f, 0.5f);
      x[i] = y;
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* o; cudaMalloc(&o, 148 * 512 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"tanh.approx", "ex2.approx", "rcp.approx", "fma"};
  for (int m = 0; m < 4; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<148, 512>>>(o, iters);
      if (m == 1) k<1><<<148, 512>>>(o, iters);
      if (m == 2) k<2><<<148, 512>>>(o, iters);
      if (m == 3) k<3><<<148, 512>>>(o, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops_per_clk = 512.0 * 8 * iters / (ms * 1e-3 * 1.9e9);
    printf("%-12s %8.3f ms  -> %.1f ops / clk / SM (at 1.9 GHz)\n", names[m], ms, ops_per_clk);
  }
  return 0;
}
