// Per-atom / per-molecule elementwise kernels around the tensor-core row-linears.
// All tensors are in the packed (varlen) atom layout of kernels.h::Plan unless noted "dense".
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

namespace jodo {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// feat[b, 0:17] = (nl, sin(nl*w*2pi) x8, cos(nl*w*2pi) x8), padded with zeros to 64 columns.
// LearnedSinusodialposEmb.forward, reference models/layers.py:283-288.
__global__ void k_time_features(const float* __restrict__ nl, const float* __restrict__ w, float* __restrict__ feat, int B) {
  int b = blockIdx.x * blockDim.x / 32 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  float x = nl[b];
  float v = 0.f;
  if (lane == 0) v = x;
  else if (lane <= 16) {
    float f = x * w[(lane - 1) & 7];
    f = f * 2.0f;
    f = f * 3.14159265358979323846f;
    v = (lane <= 8) ? sinf(f) : cosf(f);
  }
  feat[(size_t)b * 64 + lane] = v;
  feat[(size_t)b * 64 + 32 + lane] = 0.f;
}

// c1[row, d] = GELU(ctx[row] * w0[d] + b0[d]);  cond_mlp.0 + GELU, reference models/mol_gnn.py:679-681,729-730
__global__ void k_cond_in(const float* __restrict__ ctx, const float* __restrict__ w0, const float* __restrict__ b0,
                          float* __restrict__ out, int rows, int D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  int r = i / D, d = i - r * D;
  out[i] = gelu_f(ctx[r] * w0[d] + b0[d]);
}

// Packed node inputs from the dense padded batch: xin[v] = (h[b,i,:], cond_h[b,i,:], 0...), pos[v] = xh[b,i,0:3].
// reference models/mol_gnn.py:509-510, 528-530.
// mol_bad != null: non-finite inputs become 0 and mark their molecule (NaN isolation, include/jodo_b200.h).
__global__ void k_gather_nodes(const float* __restrict__ xh, const float* __restrict__ cond_x,
                               const int* __restrict__ node_dense, const int* __restrict__ node_mol, int Nn, int inn, int kin,
                               float* __restrict__ xin, float4* __restrict__ pos, int* __restrict__ mol_bad) {
  int v = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (v >= Nn) return;
  const int row = node_dense[v];
  const int w = 3 + inn;
  const float* x = xh + (size_t)row * w;
  const float* c = cond_x ? cond_x + (size_t)row * w : nullptr;
  bool bad = false;
  for (int k = lane; k < kin; k += 32) {
    float val = 0.f;
    if (k < inn) val = x[3 + k];
    else if (k < 2 * inn) val = c ? c[3 + k - inn] : 0.f;
    if (mol_bad && !isfinite(val)) { val = 0.f; bad = true; }
    xin[(size_t)v * kin + k] = val;
  }
  if (lane == 0) {
    float4 p = make_float4(x[0], x[1], x[2], 0.f);
    if (mol_bad && !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { p = make_float4(0.f, 0.f, 0.f, 0.f); bad = true; }
    pos[v] = p;
  }
  if (bad) atomicOr(mol_bad + node_mol[v], 1);
}

// out[v,:] = LN(x[v,:] + gate[mol]*y[v,:]) * (1 + scale[mol]) + shift[mol]      (one warp per atom)
// norm1_node + modulate (reference models/mol_gnn.py:296) when y == null;
// gated residual + norm2_node + modulate (:307-308) otherwise.  LayerNorm eps = 1e-6, no affine (:234,240).
template <int D>
__global__ void k_ln_mod(const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy,
                         const float* __restrict__ tab, int ld_tab, int off_gate, int off_shift, int off_scale,
                         const int* __restrict__ node_mol, int Nn, float* __restrict__ out, int ldo) {
  constexpr int PER = D / 32;
  int v = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (v >= Nn) return;
  const float* t = tab + (size_t)node_mol[v] * ld_tab;
  float a[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int c = lane + 32 * i;
    float val = x[(size_t)v * ldx + c];
    if (y) val += t[off_gate + c] * y[(size_t)v * ldy + c];
    a[i] = val;
    s += val;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { float d = a[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-6f);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    int c = lane + 32 * i;
    out[(size_t)v * ldo + c] = (a[i] - mean) * rstd * t[off_scale + c] + t[off_shift + c];
  }
}

// Same math as k_ln_mod for D = 256, one warp per atom, lane owns 8 consecutive columns = one 16-byte piece of the
// fp16 operand image [tile][4 chunks][128 rows][128 B] that the persistent GEMM (imglinear.cu) bulk-copies.
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  o.x = pack_h2(v[0], v[1]); o.y = pack_h2(v[2], v[3]); o.z = pack_h2(v[4], v[5]); o.w = pack_h2(v[6], v[7]);
  return o;
}
__global__ void k_ln_mod_img(const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy,
                             const float* __restrict__ tab, int ld_tab, int off_gate, int off_shift, int off_scale,
                             const int* __restrict__ node_mol, int Nn, float* __restrict__ out32, int ldo,
                             uint8_t* __restrict__ out_img, uint8_t* __restrict__ y_img, const int* __restrict__ nonuni) {
  constexpr int D = 256;
  const int v = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);     // rows up to the end of the last tile
  const int lane = threadIdx.x & 31;
  const int tile = v >> 7, row = v & 127;
  const size_t ioff = (size_t)tile * (D / 64) * CHUNK_BYTES_A + img_piece(row, lane >> 3, lane & 7, CHUNK_BYTES_A);
  if (v >= Nn) {                                                         // padding rows of the last tile
    *reinterpret_cast<uint4*>(out_img + ioff) = make_uint4(0u, 0u, 0u, 0u);
    if (y_img) *reinterpret_cast<uint4*>(y_img + ioff) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float* t = tab + (size_t)((nonuni && *nonuni == 0) ? 0 : node_mol[v]) * ld_tab;     // uniform conditioning: row 0
  const int c0 = 8 * lane;
  float a[8];
  {
    const float4 x0 = *reinterpret_cast<const float4*>(x + (size_t)v * ldx + c0);
    const float4 x1 = *reinterpret_cast<const float4*>(x + (size_t)v * ldx + c0 + 4);
    a[0] = x0.x; a[1] = x0.y; a[2] = x0.z; a[3] = x0.w; a[4] = x1.x; a[5] = x1.y; a[6] = x1.z; a[7] = x1.w;
  }
  if (y) {
    const float4 y0 = *reinterpret_cast<const float4*>(y + (size_t)v * ldy + c0);
    const float4 y1 = *reinterpret_cast<const float4*>(y + (size_t)v * ldy + c0 + 4);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(t + off_gate + c0));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(t + off_gate + c0 + 4));
    const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += gg[i] * yy[i];
    if (y_img) *reinterpret_cast<uint4*>(y_img + ioff) = pack8(yy);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = a[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-6f);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(t + off_scale + c0));
  const float4 s1 = __ldg(reinterpret_cast<const float4*>(t + off_scale + c0 + 4));
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(t + off_shift + c0));
  const float4 h1 = __ldg(reinterpret_cast<const float4*>(t + off_shift + c0 + 4));
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = (a[i] - mean) * rstd * sc[i] + sh[i];            // the table stores 1 + scale
  if (out32) {
    *reinterpret_cast<float4*>(out32 + (size_t)v * ldo + c0) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(out32 + (size_t)v * ldo + c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
  *reinterpret_cast<uint4*>(out_img + ioff) = pack8(o);
}

// fp16 operand image of act(rows[M, K]) (K % 64 == 0): [ceil(M/128)][K/64][128 rows][128 B]; padding rows are zero.
// One thread per 16-byte piece.
__global__ void k_act_image(const float* __restrict__ rows, int ld, int M, int K, int act, uint8_t* __restrict__ img) {
  const int pieces_per_row = K / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)((M + 127) / 128 * 128) * pieces_per_row;
  if (i >= total) return;
  const int r = (int)(i / pieces_per_row), p = (int)(i - (long long)r * pieces_per_row);
  float v[8];
  if (r < M) {
    const float4 a = *reinterpret_cast<const float4*>(rows + (size_t)r * ld + 8 * p);
    const float4 b = *reinterpret_cast<const float4*>(rows + (size_t)r * ld + 8 * p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = act == ACT_SILU ? silu_f(v[k]) : (act == ACT_GELU ? gelu_f(v[k]) : v[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
  }
  const int tile = r >> 7, row = r & 127;
  *reinterpret_cast<uint4*>(img + (size_t)tile * (K / 64) * CHUNK_BYTES_A + img_piece(row, p >> 3, p & 7, CHUNK_BYTES_A)) = pack8(v);
}

// nonuni |= any element of rows[B, T] differs bitwise from row 0
__global__ void k_uniform_flag(const float* __restrict__ rows, int B, int T, int* __restrict__ nonuni) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T) return;
  const int c = (int)(i % T);
  if (__float_as_uint(rows[i]) != __float_as_uint(rows[c])) *nonuni = 1;
}

// Row 0 of  C = act_in(A) W^T + bias  as a matrix-vector product, run only while *run_if_zero == 0 (uniform conditioning:
// every molecule's AdaLN row equals row 0, which is all the consumers read).  W is the fp16 operand image of the GEMM
// kernels, [N / nt][K / 64][nt rows][128 B] with the 16-byte pieces of a row XOR-swizzled by (row & 7); one warp per output
// column, a lane takes one piece (8 weights) of four consecutive K chunks per step, x = act_in(A[0]) staged in shared memory.
// Replaces a 128-row tensor-core tile per 128 columns whose K loop is a latency chain (0.07 - 0.1 ms for the [1024 x 19712] table).
__global__ void __launch_bounds__(256) k_row0_linear(const float* __restrict__ A, int K, const uint8_t* __restrict__ Wimg, int nt, int N,
                                                     const float* __restrict__ bias, int act_in, int act_out,
                                                     const float* __restrict__ aux, float* __restrict__ out,
                                                     const int* __restrict__ run_if_zero) {
  if (run_if_zero && *run_if_zero != 0) return;
  extern __shared__ float xs[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = A[k];
    // rounded to fp16 exactly as the GEMM path stores its activation operand: the two paths then differ by accumulation order only
    xs[k] = __half2float(__float2half_rn(fminf(fmaxf(act_in == ACT_SILU ? v / (1.0f + __expf(-v)) : v, -65504.f), 65504.f)));
  }
  __syncthreads();
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const int tile = n / nt, r = n - tile * nt, nkc = K >> 6;
  const uint8_t* wrow = Wimg + ((size_t)tile * nkc * nt + r) * 128;
  const int piece = lane & 7;
  float acc = 0.f;
  for (int kc = lane >> 3; kc < nkc; kc += 4) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)kc * nt * 128 + ((piece ^ (r & 7)) << 4)));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    const float* x = xs + 64 * kc + 8 * piece;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 w = __half22float2(h[i]);
      acc = fmaf(w.x, x[2 * i], fmaf(w.y, x[2 * i + 1], acc));
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (bias ? bias[n] : 0.f);
    if (act_out == ACT_GELU) v = gelu_f(v);
    out[n] = v + (aux ? aux[n] : 0.f);
  }
}

// Per-molecule centre-of-mass removal of the block's coordinate update, in place on pos_new
// (remove_mean_with_mask, reference models/utils.py:38-45; call site models/mol_gnn.py:565-566).
// A single-atom molecule has no edges, so no kernel wrote pos_new for it: x - mean(x) = 0.
__global__ void k_com(float4* __restrict__ pos_new, const int* __restrict__ mol_start, int B) {
  int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int s = mol_start[b], e = mol_start[b + 1];
  const int n = e - s;
  if (n == 1) {
    if (lane == 0) pos_new[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int v = s + lane; v < e; v += 32) { float4 p = pos_new[v]; sx += p.x; sy += p.y; sz += p.z; }
  sx = warp_sum(sx) / n; sy = warp_sum(sy) / n; sz = warp_sum(sz) / n;
  for (int v = s + lane; v < e; v += 32) {
    float4 p = pos_new[v];
    pos_new[v] = make_float4(p.x - sx, p.y - sy, p.z - sz, 0.f);
  }
}

// flag |= any NaN in pos   (batch-global guard, reference models/mol_gnn.py:587)
__global__ void k_nan_flag(const float4* __restrict__ pos, int Nn, const int* __restrict__ mol_bad, int B, int* __restrict__ flag) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (mol_bad && v < B && mol_bad[v] != 0) atomicOr(flag, 1);      // a non-finite input: NaN positions in the reference
  if (v >= Nn) return;
  float4 p = pos[v];
  if (isnan(p.x) || isnan(p.y) || isnan(p.z)) atomicOr(flag, 1);
}

// Dense node output [B,N,3+inn]: positions (zeroed if the NaN flag is set, then CoM-free again) and atom
// logits; padded atoms stay 0 (the buffer is zero-filled by the caller).  reference models/mol_gnn.py:573,582-594.
__global__ void k_node_out(const float4* __restrict__ pos, const float* __restrict__ atom_pred, int ldp,
                           const int* __restrict__ mol_start, const int* __restrict__ node_dense,
                           const int* __restrict__ nan_flag, const int* __restrict__ mol_bad, int B, int inn,
                           float* __restrict__ out) {
  int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int s = mol_start[b], e = mol_start[b + 1];
  const int n = e - s;
  const bool zero = *nan_flag != 0;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (!zero)
    for (int v = s + lane; v < e; v += 32) { float4 p = pos[v]; sx += p.x; sy += p.y; sz += p.z; }
  sx = warp_sum(sx) / n; sy = warp_sum(sy) / n; sz = warp_sum(sz) / n;
  const int w = 3 + inn;
  for (int v = s + lane; v < e; v += 32) {
    float4 p = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : pos[v];
    float* o = out + (size_t)node_dense[v] * w;
    o[0] = p.x - sx; o[1] = p.y - sy; o[2] = p.z - sz;
    const bool bad = mol_bad != nullptr && mol_bad[b] != 0;      // the reference's logits of such a molecule are NaN
    for (int k = 0; k < inn; ++k) o[3 + k] = bad ? __int_as_float(0x7fc00000) : atom_pred[(size_t)v * ldp + k];
  }
}

// out[b,i,j,:] = 0.5 * (tmp[b,i,j,:] + tmp[b,j,i,:])      (reference models/mol_gnn.py:579)
__global__ void k_sym_edges(const float* __restrict__ tmp, float* __restrict__ out, int B, int N, int ch) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * N * N * ch;
  if (i >= total) return;
  int c = (int)(i % ch);
  long long e = i / ch;
  int jj = (int)(e % N);
  long long r = e / N;
  int ii = (int)(r % N);
  long long b = r / N;
  out[i] = 0.5f * (tmp[i] + tmp[((b * N + jj) * N + ii) * ch + c]);
}

}  // namespace

#define LAUNCH_OK() cudaGetLastError()

cudaError_t launch_time_features(const float* nl, const float* w, float* feat, int B, cudaStream_t st) {
  k_time_features<<<(B + 3) / 4, 128, 0, st>>>(nl, w, feat, B);
  return LAUNCH_OK();
}
cudaError_t launch_cond_in(const float* ctx, const float* w0, const float* b0, float* out, int rows, int D, cudaStream_t st) {
  k_cond_in<<<(rows * D + 255) / 256, 256, 0, st>>>(ctx, w0, b0, out, rows, D);
  return LAUNCH_OK();
}
cudaError_t launch_gather_nodes(const float* xh, const float* cond_x, const Plan& p, int inn, int kin, float* xin,
                                float* pos, int* mol_bad, cudaStream_t st) {
  if (mol_bad) {
    cudaError_t e = cudaMemsetAsync(mol_bad, 0, sizeof(int) * p.B, st);
    if (e != cudaSuccess) return e;
  }
  k_gather_nodes<<<(p.Nn + 7) / 8, 256, 0, st>>>(xh, cond_x, p.node_dense, p.node_mol, p.Nn, inn, kin, xin,
                                                  reinterpret_cast<float4*>(pos), mol_bad);
  return LAUNCH_OK();
}
cudaError_t launch_ln_mod(int D, const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab,
                          int off_gate, int off_shift, int off_scale, const Plan& p, float* out, int ldo,
                          cudaStream_t st) {
  dim3 grid((p.Nn + 7) / 8);
  if (D == 256)
    k_ln_mod<256><<<grid, 256, 0, st>>>(x, ldx, y, ldy, tab, ld_tab, off_gate, off_shift, off_scale, p.node_mol, p.Nn, out, ldo);
  else if (D == 384)
    k_ln_mod<384><<<grid, 256, 0, st>>>(x, ldx, y, ldy, tab, ld_tab, off_gate, off_shift, off_scale, p.node_mol, p.Nn, out, ldo);
  else
    return cudaErrorInvalidValue;
  return LAUNCH_OK();
}
cudaError_t launch_ln_mod_img(const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab, int off_gate,
                              int off_shift, int off_scale, const Plan& p, float* out32, int ldo, void* out_img,
                              void* y_img, const int* nonuni, cudaStream_t st) {
  const int rows = (p.Nn + 127) / 128 * 128;
  k_ln_mod_img<<<rows / 8, 256, 0, st>>>(x, ldx, y, ldy, tab, ld_tab, off_gate, off_shift, off_scale, p.node_mol, p.Nn,
                                         out32, ldo, static_cast<uint8_t*>(out_img), static_cast<uint8_t*>(y_img), nonuni);
  return LAUNCH_OK();
}
cudaError_t launch_act_image(const float* rows, int ld, int M, int K, int act, void* img, cudaStream_t st) {
  const long long total = (long long)((M + 127) / 128 * 128) * (K / 8);
  k_act_image<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rows, ld, M, K, act, static_cast<uint8_t*>(img));
  return LAUNCH_OK();
}
cudaError_t launch_uniform_flag(const float* rows, int B, int T, int* nonuni, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(nonuni, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const long long total = (long long)B * T;
  k_uniform_flag<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rows, B, T, nonuni);
  return LAUNCH_OK();
}
cudaError_t launch_row0_linear(const float* A, int K, const void* Wimg, int nt, int N, const float* bias, int act_in, int act_out,
                               const float* aux, float* out, const int* run_if_zero, cudaStream_t st) {
  k_row0_linear<<<(N + 7) / 8, 256, K * sizeof(float), st>>>(A, K, static_cast<const uint8_t*>(Wimg), nt, N, bias, act_in, act_out, aux, out,
                                                             run_if_zero);
  return LAUNCH_OK();
}
cudaError_t launch_com(float* pos_new, const Plan& p, cudaStream_t st) {
  k_com<<<(p.B + 7) / 8, 256, 0, st>>>(reinterpret_cast<float4*>(pos_new), p.mol_start, p.B);
  return LAUNCH_OK();
}
cudaError_t launch_nan_flag(const float* pos, int Nn, const int* mol_bad, int B, int* flag, cudaStream_t st) {
  const int n = Nn > B ? Nn : B;
  k_nan_flag<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(pos), Nn, mol_bad, B, flag);
  return LAUNCH_OK();
}
cudaError_t launch_node_out(const float* pos, const float* atom_pred, int ldp, const Plan& p, const int* nan_flag,
                            const int* mol_bad, int inn, float* out, cudaStream_t st) {
  k_node_out<<<(p.B + 7) / 8, 256, 0, st>>>(reinterpret_cast<const float4*>(pos), atom_pred, ldp, p.mol_start,
                                             p.node_dense, nan_flag, mol_bad, p.B, inn, out);
  return LAUNCH_OK();
}
cudaError_t launch_sym_edges(const float* tmp, float* out, int B, int N, int ch, cudaStream_t st) {
  long long total = (long long)B * N * N * ch;
  k_sym_edges<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tmp, out, B, N, ch);
  return LAUNCH_OK();
}

}  // namespace jodo
