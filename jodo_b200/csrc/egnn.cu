// Row kernels of the property classifier (reference cond_gen/model.py:26-220: EGNN of E_GCL_mask layers, run once per
// batch of conditional samples, sampling.py:363-367).  Every linear layer runs on the tensor-core GEMMs of this library
// (jodo_rowlinear on packed atoms, jodo_imglinear on the plan's directed edge rows); these kernels do the row-local work
// between them.  The reference enumerates the FULL n x n graph of every padded molecule (cond_gen/utils.py:18-40) and
// multiplies the masked rows by zero (model.py:207); here only real ordered pairs i != j exist (the plan's rows).
//
//   edge_mlp.0 is hoisted:  W0 cat[h_i, h_j, radial] + b0 = P[i] + Q[j] + w_r * radial   (P = W0[:, :H] h + b0, Q = W0[:, H:2H] h)
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

namespace jodo {
namespace {

__device__ __forceinline__ uint4* eg_img(void* img, int row, int col8, int K) {
  const int tile = row >> 7, r = row & 127, chunk = col8 >> 6, piece = (col8 & 63) >> 3;
  return reinterpret_cast<uint4*>(static_cast<uint8_t*>(img) + ((size_t)tile * (K >> 6) + chunk) * (128 * 128) +
                                  (size_t)r * 128 + ((piece ^ (r & 7)) << 4));
}
__device__ __forceinline__ void eg_ld8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ float eg_silu(float x) { return x / (1.0f + __expf(-x)); }

// ---- SiLU(P[i] + Q[j] + w_r |x_i - x_j|^2) -> fp16 operand image [rows][H] (the A operand of edge_mlp.2).
// Row (g, j) of the plan is the edge row = g, col = j (model.py:164-167: radial of coord[row] - coord[col]).
// 16 lanes per row (8 columns each, strided), two rows per warp.
__global__ void __launch_bounds__(256) k_egnn_edge_in(Plan p, const float4* __restrict__ pos, const float* __restrict__ PQ,
                                                      int ldpq, int H, const float* __restrict__ wr, void* img) {
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + ((threadIdx.x >> 4) & 1), lane = threadIdx.x & 15;
  if (row >= p.n_tiles * 128) return;
  const int g = p.row_g[row];
  if (g < 0) {
    for (int q = lane; q < (H >> 3); q += 16) *eg_img(img, row, 8 * q, H) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int j = p.row_j[row];
  const float4 a = pos[g], b = pos[j];
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  const float radial = dx * dx + dy * dy + dz * dz;
  for (int q = lane; q < (H >> 3); q += 16) {
    float pi[8], qj[8], w[8];
    eg_ld8(PQ + (size_t)g * ldpq + 8 * q, pi);
    eg_ld8(PQ + (size_t)j * ldpq + H + 8 * q, qj);
    eg_ld8(wr + 8 * q, w);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = eg_silu(pi[i] + qj[i] + w[i] * radial);
    uint4 o;
    o.x = pack_h2(v[0], v[1]); o.y = pack_h2(v[2], v[3]); o.z = pack_h2(v[4], v[5]); o.w = pack_h2(v[6], v[7]);
    *eg_img(img, row, 8 * q, H) = o;
  }
}

// ---- attention gate + aggregation onto the row atom (model.py:132-139, 207): m = SiLU(edge_mlp.2 ..) rows (fp16, written by
// the GEMM), agg[g] = sum over the atom's rows of m * sigmoid(w_a . m + b_a).  One warp per atom; lane owns 8 columns per
// 256 (H <= 256: at most one piece).
__global__ void __launch_bounds__(256) k_egnn_agg(const int* __restrict__ grp_row0, const int* __restrict__ grp_len,
                                                  const uint16_t* __restrict__ M, int ldm, int H, const float* __restrict__ wa,
                                                  float ba, float* __restrict__ agg, int ldagg, int Nn) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= Nn) return;
  const int c0 = 8 * lane;
  const bool on = c0 < H;
  float w[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { w[i] = 0.f; acc[i] = 0.f; }
  if (on && wa) eg_ld8(wa + c0, w);
  const int r0 = grp_row0[g], gl = grp_len[g];
  for (int k = 0; k < gl; ++k) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (on) {
      const uint4 u = *reinterpret_cast<const uint4*>(M + (size_t)(r0 + k) * ldm + c0);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    float gate = 1.0f;
    if (wa) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(v[i], w[i], d);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      gate = 1.0f / (1.0f + __expf(-(d + ba)));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], gate, acc[i]);
  }
  if (on) {
    float* dst = agg + (size_t)g * ldagg + c0;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ---- per-molecule sum over its packed atoms (model.py:66-68: h * node_mask, view, sum over nodes)
__global__ void k_mol_sum(const float* __restrict__ x, int ldx, int W, const int* __restrict__ mol_start, int B,
                          float* __restrict__ out, int ldo) {
  const int b = blockIdx.x;
  const int a0 = mol_start[b], a1 = mol_start[b + 1];
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float s = 0.f;
    for (int v = a0; v < a1; ++v) s += x[(size_t)v * ldx + c];
    out[(size_t)b * ldo + c] = s;
  }
}

}  // namespace

cudaError_t launch_egnn_edge_in(const Plan& p, const float* pos4, const float* PQ, int ldpq, int H, const float* wr, void* img,
                                cudaStream_t st) {
  k_egnn_edge_in<<<p.n_tiles * 128 / 16, 256, 0, st>>>(p, reinterpret_cast<const float4*>(pos4), PQ, ldpq, H, wr, img);
  return cudaGetLastError();
}
cudaError_t launch_egnn_agg(const int* grp_row0, const int* grp_len, const void* M16, int ldm, int H, const float* wa, float ba,
                            float* agg, int ldagg, int Nn, cudaStream_t st) {
  k_egnn_agg<<<(Nn + 7) / 8, 256, 0, st>>>(grp_row0, grp_len, static_cast<const uint16_t*>(M16), ldm, H, wa, ba, agg, ldagg, Nn);
  return cudaGetLastError();
}
cudaError_t launch_mol_sum(const float* x, int ldx, int W, const int* mol_start, int B, float* out, int ldo, cudaStream_t st) {
  k_mol_sum<<<B, 128, 0, st>>>(x, ldx, W, mol_start, B, out, ldo);
  return cudaGetLastError();
}

}  // namespace jodo
