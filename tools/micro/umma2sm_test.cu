// Validates the cta_group::2 recipe of csrc/common.cuh on the device: D[256 x N] = A[256 x K] B[N x K]^T with the A rows and
// the B rows (N) split over the two CTAs of a cluster, accumulators in each CTA's own tensor memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma2sm_test umma2sm_test.cu && ./umma2sm_test
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../jodo_b200/csrc/common.cuh"
using namespace jodo;
constexpr int N = 128, K = 128;      // K = 2 chunks of 64
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(const __half* A, const __half* B, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;                       // [K/64][128][128 B]
  uint8_t* Bs = smem + (K / 64) * 16384;    // this CTA's half: [K/64][N/2 rows][128 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + (K / 64) * (N / 2) * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int t = threadIdx.x, rank = cluster_ctarank();
  // operand images: row t of this CTA's A rows; B rows [rank * N/2, +N/2)
  for (int c = 0; c < K; c += 8) {
    uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)(rank * 128 + t) * K + c);
    *reinterpret_cast<uint4*>(As + img_piece(t, c >> 6, (c & 63) >> 3, 16384)) = v;
    if (t < N / 2) {
      uint4 w = *reinterpret_cast<const uint4*>(B + (size_t)(rank * (N / 2) + t) * K + c);
      *reinterpret_cast<uint4*>(Bs + img_piece(t, c >> 6, (c & 63) >> 3, (N / 2) * 128)) = w;
    }
  }
  if (t == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 2); fence_barrier_init(); }
  if (t < 32) tmem_alloc_2sm<128>(slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = *slot;
  // every CTA tells the leader that its operands are in place
  if (t == 0) mbar_arrive_cluster(&bar[1], 0);
  if (rank == 0 && t == 0) {
    mbar_wait_cluster<false>(&bar[1], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16_2sm(N);
    for (int k = 0; k < K / 16; ++k)
      umma_f16_2sm(tm, umma_desc_sw128(smem_u32(As) + (k >> 2) * 16384 + (k & 3) * 32),
                   umma_desc_sw128(smem_u32(Bs) + (k >> 2) * (N / 2) * 128 + (k & 3) * 32), idesc, k ? 1u : 0u);
    umma_commit_2sm(&bar[0], 3);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 32) {
    float v[32];
    tmem_ld32(tmem_addr(tm, c), v);
    for (int i = 0; i < 32; ++i) D[(size_t)(rank * 128 + t) * N + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (t < 32) tmem_dealloc_2sm<128>(tm);
}
int main() {
  std::vector<__half> A(256 * K), B(N * K);
  for (int i = 0; i < 256 * K; ++i) A[i] = __float2half((float)((i * 7 + (i / K) * 3) % 9 - 4));
  for (int i = 0; i < N * K; ++i) B[i] = __float2half((float)((i * 5 + (i / K)) % 7 - 3));
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 256 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  const int smem = (K / 64) * 16384 + (K / 64) * (N / 2) * 128 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<2, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  std::vector<float> D(256 * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int kk = 0; kk < K; ++kk) ref += __half2float(A[r * K + kk]) * __half2float(B[n * K + kk]);
      if (ref != D[r * N + n]) { if (bad < 5) printf("mismatch r=%d n=%d ref=%g got=%g\n", r, n, ref, D[r * N + n]); ++bad; }
    }
  printf("cta_group::2 M=256 N=%d K=%d: %d mismatches\n", N, K, bad);
  return bad != 0;
}
