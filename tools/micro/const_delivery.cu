// Microbenchmark: cost of delivering warp-uniform per-column constants to a row-per-thread computation.
// Variants: LDG.128 (uniform address, L1 hit), LDS.128 (broadcast), LDC (register-indexed __constant__),
// immediate constant-bank operands (compile-time index).  Build: nvcc -arch=sm_100a -O3 -o const_delivery const_delivery.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float4 c_tab[2048];
struct P { float4 t[192]; };
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float4* __restrict__ gtab, float* __restrict__ out, int iters, int off, const __grid_constant__ P prm) {
  __shared__ float4 stab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) stab[i] = gtab[i];
  __syncthreads();
  const int cq = (threadIdx.x >> 5) >> 2;
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  float o0 = 0, o1 = 0, o2 = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float4 w;
        if (MODE == 0) w = __ldg(gtab + 64 * cq + 16 * c + i);
        else if (MODE == 1) w = stab[64 * cq + 16 * c + i];
        else if (MODE == 2) w = c_tab[off + 64 * cq + 16 * c + i];
        else if (MODE == 3) w = c_tab[16 * c + i];                 // compile-time index
        else w = prm.t[16 * c + i];                                // kernel-parameter bank, compile-time index
        const float h = x[i] + w.x;
        o0 = fmaf(h, w.y, o0); o1 = fmaf(h, w.z, o1); o2 = fmaf(h, w.w, o2);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] += o0 * 1e-9f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = o0 + o1 + o2;
}
int main() {
  float4* g; float* o;
  cudaMalloc(&g, 256 * 16); cudaMalloc(&o, 148 * 512 * 4);
  cudaMemset(g, 0, 256 * 16);
  P p; for (int i = 0; i < 192; ++i) p.t[i] = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[5] = {"LDG.128 uniform", "LDS.128 broadcast", "LDC reg-index", "const imm index", "param imm index"};
  for (int m = 0; m < 5; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<148, 512>>>(g, o, iters, 0, p);
      if (m == 1) k<1><<<148, 512>>>(g, o, iters, 0, p);
      if (m == 2) k<2><<<148, 512>>>(g, o, iters, 256, p);
      if (m == 3) k<3><<<148, 512>>>(g, o, iters, 0, p);
      if (m == 4) k<4><<<148, 512>>>(g, o, iters, 0, p);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 16 warps x iters x 64 float4-constants
    double cyc = ms * 1e-3 * 1.9e9 / (16.0 * iters * 64);
    printf("%-20s %8.3f ms   ~%.2f cycles per (warp, float4 constant + 4 FP ops)  err=%s\n", names[m], ms, cyc, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
