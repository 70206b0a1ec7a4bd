// Fused reverse-SDE update of the ancestral sampler (the caller of the denoiser).
//
// reference sampling.py:569-589: x_mean = c_x x + c_p pred;  x = x_mean + sigma * z_node;  the same for the dense
// bond tensor with symmetric noise.  The noise construction of reference models/utils.py:67-99 is fused in:
//   z_node = [ (raw_pos * mask) - mean over the molecule's atoms , raw_feat * mask ]
//   z_edge[b,i,j,c] = raw_edge[b,c,max(i,j),min(i,j)] for i != j (tril(-1) + transpose), times the edge mask.
// The raw standard-normal draws come from the caller (torch.randn in the reference's call order), so the random
// stream is the reference's.  Products and sums are rounded separately (no FMA contraction), like the torch ops.
#include "common.cuh"
#include "kernels.h"

namespace jodo {

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per molecule
__global__ void k_ancestral_nodes(const float* __restrict__ x, const float* __restrict__ pred, const float* __restrict__ raw_pos,
                                  const float* __restrict__ raw_feat, const float* __restrict__ node_mask, int B, int N, int F,
                                  float c_x, float c_p, float sigma, const float* __restrict__ coef,
                                  float* __restrict__ x_new, float* __restrict__ x_mean) {
  if (coef) { c_x = coef[0]; c_p = coef[1]; sigma = coef[2]; }     // per-step coefficients of a replayed CUDA graph
  const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* m = node_mask + (size_t)b * N;
  float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
  for (int i = lane; i < N; i += 32) {
    const float mk = m[i];
    const float* r = raw_pos + ((size_t)b * N + i) * 3;
    sx += r[0] * mk; sy += r[1] * mk; sz += r[2] * mk;
    cnt += mk;
  }
  sx = warp_sum_f(sx); sy = warp_sum_f(sy); sz = warp_sum_f(sz); cnt = warp_sum_f(cnt);
  const float mx = sx / cnt, my = sy / cnt, mz = sz / cnt;
  const int nf = F - 3;
  for (int i = lane; i < N; i += 32) {
    const float mk = m[i];
    const size_t o = ((size_t)b * N + i) * F;
    const float* r = raw_pos + ((size_t)b * N + i) * 3;
    const float zc[3] = {__fsub_rn(__fmul_rn(r[0], mk), __fmul_rn(mx, mk)), __fsub_rn(__fmul_rn(r[1], mk), __fmul_rn(my, mk)),
                         __fsub_rn(__fmul_rn(r[2], mk), __fmul_rn(mz, mk))};
    for (int k = 0; k < F; ++k) {
      const float z = k < 3 ? zc[k] : __fmul_rn(raw_feat[((size_t)b * N + i) * nf + (k - 3)], mk);
      const float mean = __fadd_rn(__fmul_rn(c_x, x[o + k]), __fmul_rn(c_p, pred[o + k]));
      x_mean[o + k] = mean;
      x_new[o + k] = __fadd_rn(mean, __fmul_rn(sigma, z));
    }
  }
}

__global__ void k_ancestral_edges(const float* __restrict__ ex, const float* __restrict__ epred, const float* __restrict__ raw,
                                  const float* __restrict__ edge_mask, int B, int N, int ch, float c_x, float c_p, float sigma,
                                  const float* __restrict__ coef, float* __restrict__ e_new, float* __restrict__ e_mean) {
  if (coef) { c_x = coef[0]; c_p = coef[1]; sigma = coef[2]; }
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [B, N, N]
  const long long total = (long long)B * N * N;
  if (idx >= total) return;
  const int j = (int)(idx % N);
  const long long t = idx / N;
  const int i = (int)(t % N);
  const long long b = t / N;
  const float mk = edge_mask[idx];
  const int hi = i > j ? i : j, lo = i > j ? j : i;
  for (int c = 0; c < ch; ++c) {
    const float rz = i == j ? 0.f : raw[((b * ch + c) * N + hi) * N + lo];
    const float z = __fmul_rn(rz, mk);
    const size_t o = (size_t)idx * ch + c;
    const float mean = __fadd_rn(__fmul_rn(c_x, ex[o]), __fmul_rn(c_p, epred[o]));
    e_mean[o] = mean;
    e_new[o] = __fadd_rn(mean, __fmul_rn(sigma, z));
  }
}

}  // namespace

cudaError_t launch_ancestral_update(const float* x, const float* pred, const float* raw_pos, const float* raw_feat,
                                    const float* node_mask, const float* ex, const float* epred, const float* raw_edge,
                                    const float* edge_mask, int B, int N, int F, int ch, float c_x, float c_p, float sigma,
                                    const float* coef, float* x_new, float* x_mean, float* e_new, float* e_mean,
                                    cudaStream_t st) {
  k_ancestral_nodes<<<(B + 7) / 8, 256, 0, st>>>(x, pred, raw_pos, raw_feat, node_mask, B, N, F, c_x, c_p, sigma, coef, x_new, x_mean);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long total = (long long)B * N * N;
  k_ancestral_edges<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ex, epred, raw_edge, edge_mask, B, N, ch, c_x, c_p, sigma,
                                                                    coef, e_new, e_mean);
  return cudaGetLastError();
}

}  // namespace jodo
