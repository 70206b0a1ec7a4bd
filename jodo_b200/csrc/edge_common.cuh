// Helpers shared by the edge-tile kernels: row metadata, operand-row I/O, GBF, LayerNorm.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

namespace jodo {

constexpr int ET = 128;                   // threads per edge CTA == rows per tile
constexpr int D_ = 256;                   // node width the edge kernels are built for (config.model.nf)
constexpr int ED_ = 64;                   // edge width  (nf / 4)
constexpr int E_TILE_BYTES = 2 * CHUNK_BYTES_A;     // one [128 x 64] fp32 edge tile image (32 KB)

struct RowInfo {
  int g, j;          // packed atoms (clamped to 0 on padding rows so that loads stay in bounds)
  int gs, gl, gi;    // group start row, group length, group index inside the tile
  int mol;
  int pr;            // row of the unordered pair {g, j} in the pair plan (0 on padding rows or when the plan has no pair map)
  bool valid;
};

__device__ __forceinline__ RowInfo load_row(const Plan& p, int tile, int t) {
  RowInfo r;
  const int R = tile * TILE_ROWS + t;
  const int g = p.row_g[R];
  r.valid = g >= 0;
  r.g = r.valid ? g : 0;
  r.j = r.valid ? p.row_j[R] : 0;
  const uint32_t m = p.row_meta[R];
  r.gs = m & 255u; r.gl = (m >> 8) & 255u; r.gi = (m >> 16) & 255u;
  r.mol = p.row_mol[R];
  r.pr = (p.row_pair != nullptr && r.valid) ? p.row_pair[R] : 0;
  return r;
}

// ---- fp16 operand rows of a directed tile, gathered from the pair-row store ------------------------------------------
// The store is a sequence of K-major SWIZZLE_128B tile images of 64 columns: pair row P occupies the 128 bytes at
// P * 128, its 16-byte slot s holding piece s ^ (P & 7).  Row `row` of the destination chunk wants piece p at slot
// p ^ (row & 7), i.e. source slot s goes to slot s ^ (P & 7) ^ (row & 7).  A warp whose lane l holds the metadata of
// tile row row0 + l (fetched one tile ahead anyway) copies NR of those rows, [row0 + sub0, row0 + sub0 + NR), four rows
// = four whole 128-byte lines per cp.async instruction (eight lanes per row; P travels by shuffle), so no registers are
// held across the copy and no line is touched twice; padding rows are zero-filled so that they stay finite.  Completion:
// cp_async_wait_all() by every thread, then the usual fence.proxy.async + barrier in front of the MMA that reads the chunk.
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int NR>
__device__ __forceinline__ void gather_e16_warp(uint8_t* dst_chunk, const void* e16, int row0, int sub0, bool valid, int P, int lane) {
  static_assert(NR % 4 == 0, "four rows per instruction");
  const int s = lane & 7;
  const int mine = valid ? P : -1;
#pragma unroll
  for (int k = 0; k < NR / 4; ++k) {
    const int sub = sub0 + 4 * k + (lane >> 3);
    const int Pk = __shfl_sync(0xffffffffu, mine, sub);
    const int row = row0 + sub;
    if (Pk >= 0)
      cp_async16(smem_u32(dst_chunk) + row * 128 + ((s ^ ((Pk ^ row) & 7)) << 4), static_cast<const uint8_t*>(e16) + (size_t)Pk * 128 + s * 16);
    else
      *reinterpret_cast<uint4*>(dst_chunk + row * 128 + (s << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
}

__device__ __forceinline__ void sync_tc() {   // CTA barrier that also orders tcgen05 traffic around it
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// store 64 consecutive columns (2 chunks starting at chunk kc0) of row `row` into an operand image in smem
template <bool ROUND>
__device__ __forceinline__ void st_row64(uint8_t* img, int row, int kc0, const float (&v)[64]) {
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    float4 o;
    o.x = ROUND ? to_tf32(v[4 * p]) : v[4 * p];
    o.y = ROUND ? to_tf32(v[4 * p + 1]) : v[4 * p + 1];
    o.z = ROUND ? to_tf32(v[4 * p + 2]) : v[4 * p + 2];
    o.w = ROUND ? to_tf32(v[4 * p + 3]) : v[4 * p + 3];
    *reinterpret_cast<float4*>(img + img_piece(row, kc0 + (p >> 3), p & 7, CHUNK_BYTES_A)) = o;
  }
}
__device__ __forceinline__ void ld_row64(const uint8_t* img, int row, int kc0, float (&v)[64]) {
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float4 o = *reinterpret_cast<const float4*>(img + img_piece(row, kc0 + (p >> 3), p & 7, CHUNK_BYTES_A));
    v[4 * p] = o.x; v[4 * p + 1] = o.y; v[4 * p + 2] = o.z; v[4 * p + 3] = o.w;
  }
}
// 32 consecutive columns = one chunk
template <bool ROUND>
__device__ __forceinline__ void st_row32(uint8_t* img, int row, int kc, const float (&v)[32]) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    float4 o;
    o.x = ROUND ? to_tf32(v[4 * p]) : v[4 * p];
    o.y = ROUND ? to_tf32(v[4 * p + 1]) : v[4 * p + 1];
    o.z = ROUND ? to_tf32(v[4 * p + 2]) : v[4 * p + 2];
    o.w = ROUND ? to_tf32(v[4 * p + 3]) : v[4 * p + 3];
    *reinterpret_cast<float4*>(img + img_piece(row, kc, p, CHUNK_BYTES_A)) = o;
  }
}

// 8 packed fp16 (one 16-byte load of a per-atom fp16 row) -> fp32
__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
  const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// 32 consecutive fp16 columns of a per-atom row = 4 x 16-byte loads
struct H32 { uint4 u[4]; };
__device__ __forceinline__ H32 ldg_h32(const uint16_t* p) {
  H32 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.u[i] = __ldg(reinterpret_cast<const uint4*>(p) + i);
  return r;
}

// ---- fp16 operand rows: NH consecutive groups of 8 columns (16-byte pieces) starting at piece p0 of chunk kc
template <int N>
__device__ __forceinline__ void st_rowh(uint8_t* img, int row, int kc, int p0, const float (&v)[N]) {
  static_assert(N % 8 == 0, "whole 16-byte pieces");
#pragma unroll
  for (int p = 0; p < N / 8; ++p) {
    uint4 o;
    o.x = pack_h2(v[8 * p], v[8 * p + 1]);
    o.y = pack_h2(v[8 * p + 2], v[8 * p + 3]);
    o.z = pack_h2(v[8 * p + 4], v[8 * p + 5]);
    o.w = pack_h2(v[8 * p + 6], v[8 * p + 7]);
    const int pp = p0 + p;
    *reinterpret_cast<uint4*>(img + img_piece(row, kc + (pp >> 3), pp & 7, CHUNK_BYTES_A)) = o;
  }
}
// 32 consecutive fp32 columns [32*hh, 32*hh+32) of row `row` of a [128 x 64] fp32 image (2 chunks of 32 columns)
__device__ __forceinline__ void ld_row32(const uint8_t* img, int row, int kc, float (&v)[32]) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float4 o = *reinterpret_cast<const float4*>(img + img_piece(row, kc, p, CHUNK_BYTES_A));
    v[4 * p] = o.x; v[4 * p + 1] = o.y; v[4 * p + 2] = o.z; v[4 * p + 3] = o.w;
  }
}

// CondGaussianLayer (reference models/layers.py:291-295,328-334): x = d*(1+scale)+shift;
// out = [x, exp(-0.5((x-mu_k)/sg_k)^2) / (a*sg_k)], k < 63.
// c = {mu[64], sqrt(0.5*log2(e))/sg[64], 1/(a*sg)[64]} (packer), so that exp(-0.5 z^2) = 2^(-(w*w)), w = (x-mu)*c1.
__device__ __forceinline__ float gbf_one(float x, const float* __restrict__ c, int k) {
  const float w = (x - c[k]) * c[64 + k];
  return ex2_fast(-(w * w)) * c[128 + k];
}
__device__ __forceinline__ void gbf_eval(float d, float scale, float shift, const float* __restrict__ c, float (&out)[64]) {
  const float x = fmaf(d, scale, shift);               // `scale` is the table's 1 + scale
  out[0] = x;
#pragma unroll
  for (int k = 0; k < 63; ++k) out[1 + k] = gbf_one(x, c, k);
}
// columns [32*half, 32*half + 32) of the same feature row (two threads share one edge row)
__device__ __forceinline__ void gbf_eval_half(float d, float scale, float shift, const float* __restrict__ c, int half,
                                              float (&out)[32]) {
  const float x = fmaf(d, scale, shift);               // `scale` is the table's 1 + scale
  if (half == 0) {
    out[0] = x;
#pragma unroll
    for (int k = 0; k < 31; ++k) out[1 + k] = gbf_one(x, c, k);
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) out[k] = gbf_one(x, c, 31 + k);
  }
}

// Same features from the float4 table g4[c] = {mu_k, c1_k, c2_k, 0} with k = c - 1 (entry 0 unused): columns
// [col0, col0 + N) of the 64-wide feature row.
template <int N>
__device__ __forceinline__ void gbf_eval_cols(float d, float scale, float shift, const float4* __restrict__ g4, int col0,
                                              float (&out)[N]) {
  const float x = fmaf(d, scale, shift);               // `scale` is the table's 1 + scale
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float4 c = g4[col0 + i];
    const float w = (x - c.x) * c.y;
    out[i] = ex2_fast(-(w * w)) * c.z;
  }
  if (col0 == 0) out[0] = x;
}
// Piece-major per-atom operands: [N/8 pieces][ld rows][8 halves] -- the 16-byte piece p of atom v lives at
// (p * ld + v) * 16 bytes, so lanes that gather consecutive atoms (the partners of one group) read whole lines.
// 4 consecutive pieces (32 columns) of atom v:
__device__ __forceinline__ H32 ldg_pm32(const void* base, int ld, int v, int piece0) {
  H32 r;
  const uint4* b = static_cast<const uint4*>(base) + (size_t)piece0 * ld + v;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.u[i] = __ldg(b + (size_t)i * ld);
  return r;
}

// 16 consecutive fp16 columns of a per-atom row = 2 x 16-byte loads
struct H16 { uint4 u[2]; };
__device__ __forceinline__ H16 ldg_h16(const uint16_t* p) {
  H16 r;
  r.u[0] = __ldg(reinterpret_cast<const uint4*>(p));
  r.u[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  return r;
}

// N consecutive floats of a per-molecule table row (16-byte aligned) through 16-byte loads
template <int N>
__device__ __forceinline__ void ldg_row(const float* __restrict__ p, float (&v)[N]) {
  static_assert(N % 4 == 0, "whole float4s");
#pragma unroll
  for (int k = 0; k < N / 4; ++k) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p) + k);
    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
  }
}

// LayerNorm (no affine, eps 1e-6) + modulate over 64 thread-local values
__device__ __forceinline__ void ln_mod64(float (&x)[64], const float* __restrict__ shift, const float* __restrict__ scale) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += x[i];
  const float mean = s * (1.0f / 64.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) { const float d = x[i] - mean; q += d * d; }
  const float rstd = rsqrtf(q * (1.0f / 64.0f) + 1e-6f);
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const float4 sh = *reinterpret_cast<const float4*>(shift + i);
    const float4 sc = *reinterpret_cast<const float4*>(scale + i);
    x[i] = (x[i] - mean) * rstd * sc.x + sh.x;
    x[i + 1] = (x[i + 1] - mean) * rstd * sc.y + sh.y;
    x[i + 2] = (x[i + 2] - mean) * rstd * sc.z + sh.z;
    x[i + 3] = (x[i + 3] - mean) * rstd * sc.w + sh.w;
  }
}

__device__ __forceinline__ float sq_dist(const float4 a, const float4 b) {
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}

// trap if the dynamic shared memory window is not 1024-byte aligned (SWIZZLE_128B atoms need it)
__device__ __forceinline__ void require_smem_alignment(const void* base) {
  if ((smem_u32(base) & 1023u) != 0u) __trap();
}

}  // namespace jodo
