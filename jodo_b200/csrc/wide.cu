// Wide path: the DGT edge stages for hidden sizes the fused edge-tile kernels are not built for (model.nf = 384,
// the reference's "large" GEOM-Drugs configuration, README.md:156,168).  At nf = 384 one block's edge weights
// (lin_edge0/1, input_lin, coord_mlp.0: 96x378, 96x384, 192x384, 384x384 fp16) no longer fit the shared memory of one
// SM next to a tile, so this path keeps every linear layer on the persistent tcgen05 GEMM (jodo_imglinear, fp16
// operand images streamed by TMA) and runs the row-local work between two GEMMs -- GBF features, LayerNorm +
// modulation, the per-target softmax and message sum, the coordinate sum -- as the row kernels below.  Rows are the
// plan's edge rows (tiles of 128, a group = all partners of one atom, contiguous), per-edge intermediates live in HBM.
// Every size is a run-time argument.
//
// Reference lines restated: models/mol_gnn.py:270-322 (EquivariantMixBlock), :71-94 (MultiCondEquiUpdate),
// models/layers.py:131-186 (TransMixLayer), :291-295,328-334 (CondGaussianLayer), models/mol_gnn.py:517-557, 571-579.
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

namespace jodo {
namespace {

constexpr int WCHUNK = 128 * 128;      // bytes of one image chunk: [128 rows][64 fp16]

template <int LANES>
__device__ __forceinline__ float wsum(float v, unsigned mask) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
// address of the 16-byte piece holding columns [col8, col8 + 8) of `row` in an image with K columns
__device__ __forceinline__ uint4* wimg(void* img, int row, int col8, int K) {
  const int tile = row >> 7, r = row & 127, chunk = col8 >> 6, piece = (col8 & 63) >> 3;
  return reinterpret_cast<uint4*>(static_cast<uint8_t*>(img) + ((size_t)tile * (K >> 6) + chunk) * WCHUNK +
                                  (size_t)r * 128 + ((piece ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint4 wpack8(const float* v) {
  uint4 o;
  o.x = pack_h2(v[0], v[1]); o.y = pack_h2(v[2], v[3]); o.z = pack_h2(v[4], v[5]); o.w = pack_h2(v[6], v[7]);
  return o;
}
__device__ __forceinline__ void wload8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void wload8h(const void* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
// 8 columns of row `r` of an fp32 or fp16 row-major matrix (ld in elements)
__device__ __forceinline__ void wload8x(const float* base, bool f16, size_t r, int ld, int c, float* v) {
  if (f16) wload8h(reinterpret_cast<const uint16_t*>(base) + r * ld + c, v);
  else wload8(base + r * ld + c, v);
}
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

// 8 consecutive floats through the read-only path (16-byte aligned)
__device__ __forceinline__ void wldg8(const float* __restrict__ p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// CondGaussianLayer features of one squared distance, columns [c0, c0 + 8): column 0 is x = d (1 + scale) + shift,
// column c >= 1 is exp(-0.5 ((x - mu) / sg)^2) / (a sg) = 2^(-((x - mu) c1)^2) c2 of Gaussian c - 1.  The constant
// tables gbf = {mu, c1, c2} x ldg are indexed by COLUMN (entry 0 unused), so that a piece reads them as float4s.
__device__ __forceinline__ void wgbf8(float x, const float* __restrict__ gbf, int ldg, int c0, int ed, float* v) {
  float mu[8], c1[8], c2[8];
  wldg8(gbf + c0, mu);
  wldg8(gbf + ldg + c0, c1);
  wldg8(gbf + 2 * ldg + c0, c2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    const float u = (x - mu[i]) * c1[i];
    v[i] = c == 0 ? x : (c < ed ? exp2f(-u * u) * c2[i] : 0.f);
  }
}

// ---- any squared cond distance over real ordered pairs != 0 (batch-global branch, models/mol_gnn.py:544)
__global__ void k_wide_dist_flag(Plan p, const float* __restrict__ cond_x, int w, int* __restrict__ flag) {
  const int R = blockIdx.x * blockDim.x + threadIdx.x;
  if (R >= p.n_tiles * 128) return;
  const int g = p.row_g[R];
  if (g < 0) return;
  const float* a = cond_x + (size_t)p.node_dense[g] * w;
  const float* b = cond_x + (size_t)p.node_dense[p.row_j[R]] * w;
  const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  if (dx * dx + dy * dy + dz * dz != 0.0f) atomicOr(flag, 1);
}

// ---- model-level edge inputs (models/mol_gnn.py:517-557): image [dist0 (ed) | edge_x (ch) | cond_edge_x (ch) | 0]
// with K columns, and the two adjacency-head bits per row.  Two rows per warp, 16 lanes each (K <= 128: 16 pieces).
__global__ void k_wide_embed_in(WideEmbedArgs a) {
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + ((threadIdx.x >> 4) & 1), lane = threadIdx.x & 15;
  if (row >= a.p.n_tiles * 128) return;
  const int g = a.p.row_g[row];
  const int np = a.K >> 3;
  if (g < 0) {
    for (int p = lane; p < np; p += 16) *wimg(a.img, row, 8 * p, a.K) = make_uint4(0u, 0u, 0u, 0u);
    if (lane == 0) a.extra[row] = 0;
    return;
  }
  const int j = a.p.row_j[row], N = a.p.N, w = 3 + a.inn, ch = a.ch, ed = a.ed;
  const int dg = a.p.node_dense[g], dj = a.p.node_dense[j];
  const int b = dg / N, ig = dg - b * N, ij = dj - b * N;
  const size_t eoff = (((size_t)b * N + ij) * N + ig) * ch;                 // edge_x[b, r = j, c = g]
  float d0 = 0.f;
  if (a.cond_x) {
    const float* cg = a.cond_x + (size_t)dg * w;
    const float* cj = a.cond_x + (size_t)dj * w;
    const float dx = cj[0] - cg[0], dy = cj[1] - cg[1], dz = cj[2] - cg[2];
    d0 = dx * dx + dy * dy + dz * dz;
  }
  const bool use_gbf = a.cond_x && *a.dist_flag;
  const float* tr = a.tab + (size_t)a.p.row_mol[row] * a.ld_tab;
  const float x = d0 * tr[0] + tr[1];                                        // the table stores 1 + scale
  for (int p = lane; p < np; p += 16) {
    float v[8];
    const int c0 = 8 * p;
    if (c0 < ed) {
      if (use_gbf) wgbf8(x, a.gbf, a.ld_gbf, c0, ed, v);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = c0 + i - ed;
        v[i] = k < ch ? a.edge_x[eoff + k] : ((k < 2 * ch && a.cond_edge_x) ? a.cond_edge_x[eoff + k - ch] : 0.f);
      }
    }
    *wimg(a.img, row, c0, a.K) = wpack8(v);
  }
  if (lane == 0) {
    // adjacency heads: cond_adj_2d (models/mol_gnn.py:520-525), cond_adj_spatial (models/utils.py:111-119)
    const bool a2d = a.cond_edge_x ? (a.cond_edge_x[eoff] >= a.edge_th) : true;
    const bool asp = d0 <= a.spatial_cut;
    a.extra[row] = (uint8_t)((a2d ? 1 : 0) | (asp ? 2 : 0));
  }
}

// ---- fp32 rows -> fp16 pieces at a column offset of up to three images (the edge part of the next block's
// [dist | e] operand, of this block's [e | dist] operand, and this block's slot of the edge heads' operand).  W % 8 == 0.
__global__ void k_wide_put(const float* __restrict__ src, int ld, int M, int W, const int* __restrict__ valid,
                           void* img1, int K1, int col1, void* img2, int K2, int col2, void* img3, int K3, int col3) {
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + ((threadIdx.x >> 4) & 1), lane = threadIdx.x & 15;
  if (row >= M) return;                                          // two rows per warp, 16 lanes each
  const bool live = !valid || valid[row] >= 0;
  for (int p = lane; p < (W >> 3); p += 16) {
    float v[8];
    if (live) wload8(src + (size_t)row * ld + 8 * p, v);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    const uint4 o = wpack8(v);
    if (img1) *wimg(img1, row, col1 + 8 * p, K1) = o;
    if (img2) *wimg(img2, row, col2 + 8 * p, K2) = o;
    if (img3) *wimg(img3, row, col3 + 8 * p, K3) = o;
  }
}

// ---- per-block distance features (models/mol_gnn.py:284-286): written as columns [col1, col1 + ed) of the
// [dist | e] operand and [col2, col2 + ed) of the [e | dist] operand.
__global__ void k_wide_dist(Plan p, const float4* __restrict__ pos, const float* __restrict__ tab, int ld_tab,
                            int off_gbf, const float* __restrict__ gbf, int ld_gbf, int ed, void* img1, int K1, int col1,
                            void* img2, int K2, int col2) {
  // four rows per warp, 8 lanes each (ed / 8 <= 16 pieces: one or two per lane): the kernel waits on two dependent round
  // trips per row (atom indices, then positions), so rows in flight per warp are what it is bound by
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + ((threadIdx.x >> 3) & 3), lane = threadIdx.x & 7;
  if (row >= p.n_tiles * 128) return;
  const int g = p.row_g[row];
  float x = 0.f;
  if (g >= 0) {
    const float4 a = pos[g], b = pos[p.row_j[row]];
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    const float* tr = tab + (size_t)p.row_mol[row] * ld_tab + off_gbf;
    x = (dx * dx + dy * dy + dz * dz) * tr[0] + tr[1];
  }
  for (int q = lane; q < (ed >> 3); q += 8) {
    float v[8];
    if (g >= 0) wgbf8(x, gbf, ld_gbf, 8 * q, ed, v);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    const uint4 o = wpack8(v);
    *wimg(img1, row, col1 + 8 * q, K1) = o;
    if (img2) *wimg(img2, row, col2 + 8 * q, K2) = o;
  }
}

// ---- LayerNorm (eps 1e-6, no affine) + modulation of  a = x + gate * (y[yi] + y2[y2i] + ybias), one warp per row,
// W <= 512 columns.  Serves norm1/norm2 of atoms and edges (models/mol_gnn.py:296-297, 307-308, 313-314) and the
// input_lin LayerNorm of the coordinate update (:73-79, with the hoisted per-atom parts as y, y2).
// V = 0: every option is a run-time flag.  V = 1 / 2 / 3 fix the options of the three per-edge uses at compile time
// (coordinate branch: fp16 x + two gathered fp16 addends, image only; norm2_edge: fp32 x + two gathered fp32 addends +
// bias, gated, fp32 rows + image; norm1_edge: fp32 x, image only).
// LANES = 32: one warp per row (up to 64 pieces); LANES = 16: two rows per warp (up to 16 pieces each: the per-edge
// rows of width ed <= 128, where a full warp would leave most lanes idle).
// The compiled two-rows-per-warp variants are held to 40 registers (six 256-thread blocks per SM): the kernel waits on two
// dependent round trips per row (row indices, then the gathered rows), so residency buys more than the few spilled values
// cost (coordinate-branch LayerNorm 0.557 -> 0.488 ms per launch at GEOM nf = 384).
template <int V, int LANES, int KP = 2>      // KP pieces of 8 columns per lane: LANES * KP * 8 >= Kimg
__global__ void __launch_bounds__(256, (V >= 1 && LANES == 16) ? 6 : 1) k_wide_ln(WideLnArgs a) {
  constexpr int RPW = 32 / LANES;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + ((threadIdx.x & 31) / LANES);
  const int lane = threadIdx.x & (LANES - 1);
  const unsigned hm = LANES == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));     // the lanes of this row
  const int rows_pad = (a.M + 127) / 128 * 128;
  if (row >= rows_pad) return;
  const int npw = a.W >> 3, npk = a.Kimg >> 3;
  const bool x16 = V == 0 ? a.x_f16 != 0 : V == 1;
  const bool y16 = V == 0 ? a.y_f16 != 0 : V == 1;
  const bool has_y = V == 0 ? a.y != nullptr : V != 3;
  const bool has_y2 = V == 0 ? a.y2 != nullptr : V != 3;
  const bool has_bias = V == 0 ? a.ybias != nullptr : V == 2;
  const bool has_gate = V == 0 ? a.off_gate >= 0 : V == 2;
  const bool has_o32 = V == 0 ? a.out32 != nullptr : V == 2;
  const bool has_yimg = V == 0 ? a.y_img != nullptr : false;
  const bool has_img = V == 0 ? a.out_img != nullptr : true;
  // every per-row index is loaded up front (independent loads: one round trip instead of a dependent chain)
  const bool inb = row < a.M;
  const int vld = (inb && a.valid) ? __ldg(a.valid + row) : 0;
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;     // uniform conditioning: every molecule's table row is row 0
  const int mol = (inb && !uni) ? __ldg(a.row_mol + row) : 0;
  const int iy = (inb && has_y) ? (a.yi ? __ldg(a.yi + row) : row) : 0;
  const int iy2 = (inb && has_y2) ? (a.y2i ? __ldg(a.y2i + row) : row) : 0;
  const int ix = (inb && a.xi) ? max(__ldg(a.xi + row), 0) : row;        // padding rows carry -1
  const bool live = inb && vld >= 0;
  if (!live) {
    for (int p = lane; p < npk; p += LANES) {
      if (has_img) *wimg(a.out_img, row, 8 * p, a.Kimg) = make_uint4(0u, 0u, 0u, 0u);
      if (has_yimg) *wimg(a.y_img, row, 8 * p, a.Kimg) = make_uint4(0u, 0u, 0u, 0u);
      if (has_o32 && row < a.M && 8 * p < a.ldo) {
        *reinterpret_cast<float4*>(a.out32 + (size_t)row * a.ldo + 8 * p) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(a.out32 + (size_t)row * a.ldo + 8 * p + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }
  const float* t = a.tab + (size_t)mol * a.ld_tab;
  float v[KP][8];
  float s = 0.f, q = 0.f;                   // one pass: sum and sum of squares (as the fused edge kernels do)
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const int p = lane + LANES * k;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[k][i] = 0.f;
    if (p < npw) {
      if (V == 1) {
        // coordinate branch: three fp16 rows.  The two hoisted parts are pre-added as half2 (as in the fused kernel) and the
        // halves enter the fp32 sum through FHADD -- 20 instructions per piece instead of 24 conversions + 16 adds
        const uint4 ux = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.x) + (size_t)ix * a.ldx + 8 * p);
        uint4 uy = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.y) + (size_t)iy * a.ldy + 8 * p);
        const uint4 uz = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.y2) + (size_t)iy2 * a.ldy2 + 8 * p);
        __half2* hy = reinterpret_cast<__half2*>(&uy);
        const __half2* hz = reinterpret_cast<const __half2*>(&uz);
        const uint32_t* wx = reinterpret_cast<const uint32_t*>(&ux);
        const uint32_t* wy = reinterpret_cast<const uint32_t*>(&uy);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          hy[i] = __hadd2(hy[i], hz[i]);
          v[k][2 * i] = fhadd_lo(wx[i], fhadd_lo(wy[i], 0.f));
          v[k][2 * i + 1] = fhadd_hi(wx[i], fhadd_hi(wy[i], 0.f));
        }
      } else {
        wload8x(a.x, x16, ix, a.ldx, 8 * p, v[k]);
        if (has_y) {
          float y[8];
          wload8x(a.y, y16, iy, a.ldy, 8 * p, y);
          if (has_yimg) *wimg(a.y_img, row, 8 * p, a.Kimg) = wpack8(y);
          if (has_y2) {
            float y2[8];
            wload8x(a.y2, y16, iy2, a.ldy2, 8 * p, y2);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] += y2[i];
          }
          if (has_bias) {
            float yb[8];
            wldg8(a.ybias + 8 * p, yb);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] += yb[i];
          }
          if (has_gate) {
            float gt[8];
            wldg8(t + a.off_gate + 8 * p, gt);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[k][i] = fmaf(gt[i], y[i], v[k][i]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[k][i] += y[i];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { s += v[k][i]; q = fmaf(v[k][i], v[k][i], q); }
    } else if (p < npk && has_yimg) {
      *wimg(a.y_img, row, 8 * p, a.Kimg) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const float inv_w = 1.0f / (float)a.W;
  const float mean = wsum<LANES>(s, hm) * inv_w;
  const float rstd = rsqrtf(fmaxf(wsum<LANES>(q, hm) * inv_w - mean * mean, 0.f) + 1e-6f);
  const float nmr = -mean * rstd;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const int p = lane + LANES * k;
    if (p >= npk) continue;
    float o[8];
    if (p < npw) {
      float sc[8], sh[8];
      wldg8(t + a.off_scale + 8 * p, sc);     // the table stores 1 + scale
      wldg8(t + a.off_shift + 8 * p, sh);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(v[k][i], rstd, nmr), sc[i], sh[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
    if (has_o32 && 8 * p < a.ldo) {
      *reinterpret_cast<float4*>(a.out32 + (size_t)row * a.ldo + 8 * p) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(a.out32 + (size_t)row * a.ldo + 8 * p + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (has_img) *wimg(a.out_img, row, 8 * p, a.Kimg) = wpack8(o);
  }
}

// ---- TransMixLayer.message + aggregation (models/layers.py:157-186) for one target atom per CTA.
// logits over the sources of the atom's group: extra heads first (adjacency bit ? 1 : -1e10), then
// a[s] = sum_ch q[c,s,ch] k[r,s,ch] g0[s,ch] / sqrt(C); PyG softmax exp(a - max) / (sum + 1e-16); message
// v[r] * g1 * alpha summed over the sources.  g0 | g1 = tanh(lin_edge0 | lin_edge1) come from one GEMM.
constexpr int WA_THREADS = 128;
__global__ void __launch_bounds__(WA_THREADS) k_wide_attn(WideAttnArgs a) {
  extern __shared__ float wsm[];
  const int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gl = a.grp_len[g], row0 = a.grp_row0[g];
  const int D = a.D, H = a.H, X = a.X, S = H - X, sc = a.sc, qk = S * sc, C = D / H;
  const int qkp = (qk + 31) & ~31;
  float* qs = wsm;                          // [qkp] q of the target, pre-scaled
  float* prod = qs + qkp;                   // [4 warps][qkp]
  float* lg = prod + 4 * qkp;               // [max_gl][H] logits -> alpha
  int* js = reinterpret_cast<int*>(lg + a.max_gl * H);   // [max_gl] source atoms
  if (gl <= 0) {
    for (int c = tid; c < D; c += WA_THREADS) a.hnode[(size_t)g * D + c] = 0.f;
    return;
  }
  const float inv = rsqrtf((float)C);
  const uint16_t* qrow = a.qkv + (size_t)g * a.ldq;
  for (int c = tid; c < qkp; c += WA_THREADS) qs[c] = c < qk ? h2f(qrow[c]) * inv : 0.f;
  int* ps = js + a.max_gl;                  // [max_gl] row of G / extra: the edge row itself or its pair's row
  for (int i = tid; i < gl; i += WA_THREADS) {
    js[i] = a.row_j[row0 + i];
    ps[i] = a.row_pair ? a.row_pair[row0 + i] : row0 + i;
  }
  __syncthreads();
  for (int i = warp; i < gl; i += 4) {
    const __half2* krow = reinterpret_cast<const __half2*>(a.qkv + (size_t)js[i] * a.ldq + a.k_off);
    const __half2* grow = reinterpret_cast<const __half2*>(a.G + (size_t)ps[i] * a.ldg);
    float* pr = prod + warp * qkp;
    for (int c = lane; c < ((qk + 1) >> 1); c += 32) {            // an odd qk reads one zero pad column of the rows
      const float2 kk = __half22float2(krow[c]), gg = __half22float2(grow[c]);
      pr[2 * c] = qs[2 * c] * kk.x * gg.x;
      pr[2 * c + 1] = qs[2 * c + 1] * kk.y * gg.y;
    }
    __syncwarp();
    if (lane < S) {
      float s = 0.f;
      for (int k = 0; k < sc; ++k) s += pr[lane * sc + k];
      lg[i * H + X + lane] = s;
    } else if (lane - S < X) {
      const int x = lane - S;
      lg[i * H + x] = ((a.extra[ps[i]] >> x) & 1) ? 1.0f : -1e10f;
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid < H) {
    float m = -INFINITY;
    for (int i = 0; i < gl; ++i) m = fmaxf(m, lg[i * H + tid]);
    float s = 0.f;
    for (int i = 0; i < gl; ++i) { const float e = expf(lg[i * H + tid] - m); lg[i * H + tid] = e; s += e; }
    const float rs = 1.0f / (s + 1e-16f);
    for (int i = 0; i < gl; ++i) lg[i * H + tid] *= rs;
  }
  __syncthreads();
  for (int c = 4 * tid; c < D; c += 4 * WA_THREADS) {             // 4 columns of one head per thread (C % 4 == 0)
    const int h = c / C;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < gl; ++i) {
      const uint2 vu = *reinterpret_cast<const uint2*>(a.qkv + (size_t)js[i] * a.ldq + a.v_off + c);
      const uint2 gu = *reinterpret_cast<const uint2*>(a.G + (size_t)ps[i] * a.ldg + a.g1_off + c);
      const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&vu.x)), v1 = __half22float2(*reinterpret_cast<const __half2*>(&vu.y));
      const float2 g0 = __half22float2(*reinterpret_cast<const __half2*>(&gu.x)), g1 = __half22float2(*reinterpret_cast<const __half2*>(&gu.y));
      const float al = lg[i * H + h];
      acc[0] = fmaf(v0.x * g0.x, al, acc[0]); acc[1] = fmaf(v0.y * g0.y, al, acc[1]);
      acc[2] = fmaf(v1.x * g1.x, al, acc[2]); acc[3] = fmaf(v1.y * g1.y, al, acc[3]);
    }
    *reinterpret_cast<float4*>(a.hnode + (size_t)g * D + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---- coordinate update tail (models/mol_gnn.py:82-92): inv = mean([1, extra] * tanh(coord_mlp.2 output)),
// pos_r += sum_c (pos_r - pos_c) / max(|.|, 1e-8) * scale * inv.  One WARP per atom: lanes over its partner rows (a thread per
// atom walked ~n rows serially with 22 k threads in flight), fixed-order shuffle tree for the three sums.
__global__ void __launch_bounds__(256) k_wide_equi_out(const int* __restrict__ grp_row0, const int* __restrict__ grp_len,
                                                       const int* __restrict__ row_j, const float* __restrict__ c3, int ldc, int nslots,
                                                       const uint8_t* __restrict__ extra, const int* __restrict__ row_pair, int X,
                                                       float coord_scale, const float4* __restrict__ pos_in,
                                                       float4* __restrict__ pos_out, int Nn) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= Nn) return;
  const float4 p = pos_in[g];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  const int r0 = grp_row0[g], gl = grp_len[g];
  for (int i = lane; i < gl; i += 32) {
    const int R = r0 + i;
    const float4 q = pos_in[row_j[R]];
    const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
    const float nrm = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-8f);
    const float* c = c3 + (size_t)R * ldc;
    const uint8_t bits = extra[row_pair ? row_pair[R] : R];
    float cs[3] = {0.f, 0.f, 0.f};                  // coord_mlp.2 outputs: the partial sums of the GEMM's column slots, in order
    for (int s = 0; s < nslots; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(c + 4 * s);
      cs[0] += v.x; cs[1] += v.y; cs[2] += v.z;
    }
    float inv = tanhf(cs[0]);
    for (int x = 0; x < X; ++x) inv += ((bits >> x) & 1) ? tanhf(cs[1 + x]) : 0.f;
    const float f = coord_scale * inv / ((float)(1 + X) * nrm);
    sx = fmaf(dx, f, sx); sy = fmaf(dy, f, sy); sz = fmaf(dz, f, sz);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) pos_out[g] = make_float4(p.x + sx, p.y + sy, p.z + sz, 0.f);
}

// ---- last layer of edge_exist_mlp / edge_type_mlp (models/mol_gnn.py:574-578) + scatter to the dense grid.
// x[row] = [SiLU hidden of exist (hw) | SiLU hidden of type (hw)]; w4 [ch][hw]: row 0 reads the first half.
__global__ void k_wide_head_out(Plan p, const float* __restrict__ x, int ldx, int hw, const float* __restrict__ w4,
                                const float* __restrict__ b4, int ch, int both, float* __restrict__ out_dense) {
  // eight lanes per row (16-byte loads of the row, partial dots, three shuffles per output): one thread per row walked its
  // 2 hw floats with 32 sectors per load instruction (0.29 ms per call at GEOM nf = 384)
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, l8 = threadIdx.x & 7;
  if (row >= p.n_tiles * 128) return;
  const int g = p.row_g[row];
  if (g < 0) return;
  const float* xr = x + (size_t)row * ldx;
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = 0.f;
  for (int c = 4 * l8; c < 2 * hw; c += 32) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const bool first = c < hw;                              // output 0 reads the first half of the row, the others the second
    const int ci = first ? c : c - hw;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < ch && ((k == 0) == first)) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w4 + k * hw + ci));
        o[k] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, o[k]))));
      }
    }
  }
  const unsigned m = 0xffu << (threadIdx.x & 24);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < ch) {
      o[k] += __shfl_xor_sync(m, o[k], 4);
      o[k] += __shfl_xor_sync(m, o[k], 2);
      o[k] += __shfl_xor_sync(m, o[k], 1);
    }
  }
  const int N = p.N;
  const int dg = p.node_dense[g], dj = p.node_dense[p.row_j[row]];
  const int b = dg / N, ig = dg - b * N, ij = dj - b * N;
  float* dst = out_dense + (((size_t)b * N + ij) * N + ig) * ch;            // row (g, j) is the edge r = j -> c = g
  float* dst2 = out_dense + (((size_t)b * N + ig) * N + ij) * ch;           // pair plan: the same value is e_hat[b, i, j] too
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < ch && l8 == k) {
      const float v = o[k] + b4[k];
      dst[k] = v;
      if (both) dst2[k] = v;
    }
  }
}

}  // namespace

#define WIDE_OK() cudaGetLastError()

cudaError_t launch_wide_embed_in(const WideEmbedArgs& a, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(a.dist_flag, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const int R = a.p.n_tiles * 128;
  if (a.cond_x) k_wide_dist_flag<<<(R + 255) / 256, 256, 0, st>>>(a.p, a.cond_x, 3 + a.inn, a.dist_flag);
  k_wide_embed_in<<<R / 16, 256, 0, st>>>(a);
  return WIDE_OK();
}
cudaError_t launch_wide_put(const float* src, int ld, int M, int W, const int* valid, void* img1, int K1, int col1,
                            void* img2, int K2, int col2, void* img3, int K3, int col3, cudaStream_t st) {
  k_wide_put<<<(M + 15) / 16, 256, 0, st>>>(src, ld, M, W, valid, img1, K1, col1, img2, K2, col2, img3, K3, col3);
  return WIDE_OK();
}
cudaError_t launch_wide_dist(const Plan& p, const float* pos, const float* tab, int ld_tab, int off_gbf, const float* gbf,
                             int ld_gbf, int ed, void* img1, int K1, int col1, void* img2, int K2, int col2, cudaStream_t st) {
  k_wide_dist<<<p.n_tiles * 128 / 32, 256, 0, st>>>(p, reinterpret_cast<const float4*>(pos), tab, ld_tab, off_gbf, gbf,
                                                   ld_gbf, ed, img1, K1, col1, img2, K2, col2);
  return WIDE_OK();
}
cudaError_t launch_wide_ln(const WideLnArgs& a, cudaStream_t st) {
  const int rows_pad = (a.M + 127) / 128 * 128;
  const bool edge_img_only = a.out_img && !a.out32 && !a.y_img && !a.ybias && a.off_gate < 0;
  const bool narrow = a.Kimg <= 128;                 // two rows per warp
  const int g32 = rows_pad / 8, g16 = rows_pad / 16;
  if (edge_img_only && a.x_f16 && a.y && a.y2 && a.y_f16 && !narrow) {
    if (a.Kimg <= 384) k_wide_ln<1, 16, 3><<<g16, 256, 0, st>>>(a);      // two rows per warp, three pieces per lane
    else k_wide_ln<1, 32><<<g32, 256, 0, st>>>(a);
  }
  else if (a.out_img && a.out32 && !a.y_img && !a.x_f16 && a.y && a.y2 && !a.y_f16 && a.ybias && a.off_gate >= 0) {
    if (narrow) k_wide_ln<2, 16><<<g16, 256, 0, st>>>(a);
    else k_wide_ln<2, 32><<<g32, 256, 0, st>>>(a);
  } else if (edge_img_only && !a.x_f16 && !a.y) {
    if (narrow) k_wide_ln<3, 16><<<g16, 256, 0, st>>>(a);
    else k_wide_ln<3, 32><<<g32, 256, 0, st>>>(a);
  } else if (narrow) k_wide_ln<0, 16><<<g16, 256, 0, st>>>(a);
  else k_wide_ln<0, 32><<<g32, 256, 0, st>>>(a);
  return WIDE_OK();
}
cudaError_t launch_wide_attn(const WideAttnArgs& a, cudaStream_t st) {
  static const bool per_target = std::getenv("JODO_WIDE_ATTN_PER_TARGET") != nullptr;       // A/B switch
  if (!per_target && wide_attn_mol_ok(a)) return launch_wide_attn_mol(a, st);
  const int S = a.H - a.X, qkp = (S * a.sc + 31) & ~31;
  const size_t smem = (size_t)(5 * qkp + a.max_gl * a.H) * sizeof(float) + 2 * a.max_gl * sizeof(int);
  k_wide_attn<<<a.Nn, WA_THREADS, smem, st>>>(a);
  return WIDE_OK();
}
cudaError_t launch_wide_equi_out(const int* grp_row0, const int* grp_len, const int* row_j, const float* c3, int ldc, int nslots,
                                 const uint8_t* extra, const int* row_pair, int X, float coord_scale, const float* pos_in,
                                 float* pos_out, int Nn, cudaStream_t st) {
  k_wide_equi_out<<<(Nn + 7) / 8, 256, 0, st>>>(grp_row0, grp_len, row_j, c3, ldc, nslots, extra, row_pair, X, coord_scale,
                                                    reinterpret_cast<const float4*>(pos_in), reinterpret_cast<float4*>(pos_out), Nn);
  return WIDE_OK();
}
cudaError_t launch_wide_head_out(const Plan& p, const float* x, int ldx, int hw, const float* w4, const float* b4, int ch,
                                 int both, float* out_dense, cudaStream_t st) {
  const int R = p.n_tiles * 128;
  k_wide_head_out<<<(R * 8 + 255) / 256, 256, 0, st>>>(p, x, ldx, hw, w4, b4, ch, both, out_dense);
  return WIDE_OK();
}

}  // namespace jodo
