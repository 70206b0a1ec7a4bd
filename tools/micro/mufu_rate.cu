// Microbenchmark: MUFU / packed-half throughput per SM (independent chains, 16 warps per SM).
//   0 tanh.approx.f32   1 ex2.approx.ftz.f32   2 rcp.approx.ftz.f32   3 fma.f32
//   4 tanh.approx.f16x2 (two results per instruction)   5 ex2.approx.f16x2   6 fma.rn.f16x2
//   7 cvt.rn.f16x2.f32 + tanh.approx.f16x2 (the epilogue form: fp32 accumulator pair -> packed gate)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
  float x[8];
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i * 0.1f; h[i] = 0x3c003800u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = x[i];
      unsigned g = h[i];
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      else if (OP == 3) y = fmaf(x[i], 1.0001f, 0.5f);
      else if (OP == 4) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(g) : "r"(h[i]));
      else if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(g) : "r"(h[i]));
      else if (OP == 6) asm volatile("fma.rn.f16x2 %0, %1, %1, %1;" : "=r"(g) : "r"(h[i]));
      else {
        unsigned p;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(g) : "r"(p));
        y = x[i] + __uint_as_float(g & 0x3f800000u) * 1e-9f;
      }
      x[i] = y;
      h[i] = g;
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(float* o, int iters, const char* name, double per) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<OP><<<148, 512>>>(o, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = 512.0 * 8 * iters / (ms * 1e-3 * 1.9e9);
  printf("%-28s %8.3f ms  -> %.1f instr / clk / SM, %.1f results / clk / SM (at 1.9 GHz)\n", name, ms, inst, inst * per);
}
int main() {
  float* o; cudaMalloc(&o, 148 * 512 * 4);
  const int iters = 20000;
  run<0>(o, iters, "tanh.approx.f32", 1);
  run<1>(o, iters, "ex2.approx.f32", 1);
  run<2>(o, iters, "rcp.approx.f32", 1);
  run<3>(o, iters, "fma.f32", 1);
  run<4>(o, iters, "tanh.approx.f16x2", 2);
  run<5>(o, iters, "ex2.approx.f16x2", 2);
  run<6>(o, iters, "fma.rn.f16x2", 2);
  run<7>(o, iters, "cvt.f16x2 + tanh.f16x2", 2);
  return 0;
}
