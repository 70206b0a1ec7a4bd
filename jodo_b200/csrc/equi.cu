// E(3)-equivariant coordinate update of one DGT block on edge tiles.
//
// reference MultiCondEquiUpdate.forward, models/mol_gnn.py:71-94 and CoorsNorm, models/layers.py:344-347:
//   u = cat[h[row], h[col], e, dist]; inv = LN(input_lin(u)) * (1 + scale) + shift;
//   inv = tanh(coord_mlp(inv)) -> [1 + 2]; inv = mean(inv * [1, adj2d, adjsp]);
//   pos[row] += sum_col (pos[row] - pos[col]) / max(|.|, 1e-8) * coord_scale * inv.
// The sum runs over the partners of `row`, so here the group atom g is ROW r and the partner j is COL c
// (same stored rows as the attention pass, roles swapped; edge features are symmetric).
// input_lin is hoisted: W[:, :D] h[g] + b and W[:, D:2D] h[j] come from the per-atom piece-major fp16 buffer AB (pre-added
// as half2 under the input_lin MMA), only the
// [e | dist] part (K = 128) runs per edge.
//
// fp16 operand images (same mantissa as tf32).  coord_mlp.0 (256x256, 128 KB, pre-scaled by 1/2 for the
// tanh form of SiLU) stays resident in shared memory; the input_lin image (64 KB) is bulk-copied per tile into the
// region that later holds the LayerNorm-modulated operand, and that copy as well as the next e tile are issued as
// soon as the MMA that read the region has completed, so they overlap the epilogues.
// 512 threads: warp w = tile rows 32*(w&3)..+31 (its TMEM lane quarter) x column quarter (w>>2) of the 256 hidden
// units, 16 columns at a time, so that four warps per scheduler cover each other's TMEM / gather latencies.
#include <cstdio>
#include "edge_common.cuh"

namespace jodo {

#ifdef JODO_PHASE_TIMING
__device__ long long g_equi_phase[16];
#define PHASE_MARK(i) do { if (t == 0 && blockIdx.x == 0) { long long c_ = clock64(); g_equi_phase[i] += c_ - ph_last; ph_last = c_; } } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

namespace {

constexpr int EQ_THREADS = 512;
constexpr int EQ_WC0 = 0;                        // 128 KB: coord_mlp.0 image (N = 256, K = 256: 4 chunks of 32 KB)
constexpr int EQ_X = 131072;                     // 64 KB: input_lin image (N = 256, K = 128), then LN-modulated A (K = 256)
constexpr int EQ_U = EQ_X + 65536;               // 32 KB: [e | GBF(d)] (K = 128); chunk 1 doubles as scratch
constexpr int EQ_MISC = EQ_U + 32768;            // barriers + tmem slot (128 B), then C3 [128] float4 and the group table
constexpr int EQ_C3 = EQ_MISC + 128;             // per-row coordinate contribution (dedicated: read across the tile end)
constexpr int EQ_GT = EQ_C3 + 128 * 16;          // group table: start | len << 8 [64], atom [64] (<= 64 groups per tile)
constexpr int EQ_SMEM = EQ_GT + 512;
static_assert(EQ_SMEM <= 232448, "shared memory budget");
// scratch inside U chunk 1 (free once the input_lin MMA has completed)
constexpr int EQ_SCR = EQ_U + 16384;
constexpr int EQ_LNS = EQ_SCR;                   // [128][4] float2
constexpr int EQ_W2 = EQ_LNS + 128 * 4 * 8;      // 8 KB: coord_mlp.2 image (N = 16, K = 256), bulk-copied per tile
static_assert(EQ_W2 % 1024 == 0 && EQ_W2 + 8192 <= EQ_MISC, "scratch overflows the U chunk");

// Uniform-conditioning fast path: when every molecule of the batch carries the same noise level (and context), the
// AdaLN rows are identical, and row 0's (shift[256] | scale[256] | gbf scale, shift) segment is copied into constant
// memory before the launch.  Per-column constants then reach the FMA pipe as constant-bank operands instead of one
// 16-byte load per (thread, 4 columns) -- the L1 write-back path is what bounds these kernels (profiles/).
__constant__ float c_eqmod[528];

#define EQ_DISPATCH(F, ...)                  \
  switch (cq) {                              \
    case 0: F<0>(__VA_ARGS__); break;        \
    case 1: F<1>(__VA_ARGS__); break;        \
    case 2: F<2>(__VA_ARGS__); break;        \
    default: F<3>(__VA_ARGS__); break;       \
  }

// distance features, columns [16 CQ, 16 CQ + 16) of the GBF chunk; constants are kernel-parameter operands
template <int CQ>
__device__ __forceinline__ void eq_gbf(const EquiArgs& a, float d, float scale, float shift, uint4 (&out)[2]) {
  const float x = fmaf(d, scale, shift);               // the table stores 1 + scale
  float df[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int col = 16 * CQ + i;
    if (col == 0) {
      df[i] = x;
    } else {
      const float w = (x - a.gbf4[4 * col]) * a.gbf4[4 * col + 1];
      df[i] = ex2_fast(-(w * w)) * a.gbf4[4 * col + 2];
    }
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    out[p].x = pack_h2(df[8 * p], df[8 * p + 1]); out[p].y = pack_h2(df[8 * p + 2], df[8 * p + 3]);
    out[p].z = pack_h2(df[8 * p + 4], df[8 * p + 5]); out[p].w = pack_h2(df[8 * p + 6], df[8 * p + 7]);
  }
}

// LN + modulate of hidden units [64 CQ, 64 CQ + 64) -> X chunk CQ
template <int CQ, bool UNI>
__device__ __forceinline__ void eq_pass2_t(uint32_t tm_x, uint8_t* X, int row, float mean, float rstd, const float* tr) {
  static_assert(UNI, "the general path is eq_pass2_gen");
  const float nmr = -mean * rstd;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float x[16];
    tmem_ld16(tmem_addr(tm_x, 64 * CQ + 16 * c), x);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int col = 64 * CQ + 16 * c + i;
      const float n = fmaf(x[i], rstd, nmr);
      x[i] = fmaf(n, c_eqmod[256 + col], c_eqmod[col]);
    }
    st_rowh<16>(X, row, CQ, 2 * c, x);
  }
}
template <int CQ> __device__ __forceinline__ void eq_pass2_uni(uint32_t tm_x, uint8_t* X, int row, float mean, float rstd, const float* tr) {
  eq_pass2_t<CQ, true>(tm_x, X, row, mean, rstd, tr);
}
// general path (per-molecule rows through loads): one code copy for all column quarters (instruction-cache footprint)
__device__ __noinline__ void eq_pass2_gen(uint32_t tm_x, uint8_t* X, int row, int cq, float mean, float rstd, const float* tr) {
  const float nmr = -mean * rstd;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float x[16];
    tmem_ld16(tmem_addr(tm_x, 64 * cq + 16 * c), x);
    const float* shift = tr + tab_equi(D_) + 64 * cq + 16 * c;
    const float* scale = shift + D_;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + i));
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + i));
      const float n0 = fmaf(x[i], rstd, nmr), n1 = fmaf(x[i + 1], rstd, nmr), n2 = fmaf(x[i + 2], rstd, nmr), n3 = fmaf(x[i + 3], rstd, nmr);
      x[i] = fmaf(n0, sc.x, sh.x);
      x[i + 1] = fmaf(n1, sc.y, sh.y);
      x[i + 2] = fmaf(n2, sc.z, sh.z);
      x[i + 3] = fmaf(n3, sc.w, sh.w);
    }
    st_rowh<16>(X, row, cq, 2 * c, x);
  }
}

// SiLU (h + h tanh h with h = x/2; the image is pre-scaled by 1/2) of hidden units [64 CQ, 64 CQ + 64), written back to
// tensor memory as packed fp16 pairs: the A operand of the coord_mlp.2 MMA (lane = row, column = two K elements).
// b0h[k] = coord_mlp.0 bias / 2 as kernel-parameter operands.
template <int CQ>
__device__ __forceinline__ void eq_silu_tm(const EquiArgs& a, uint32_t tm_c, uint32_t tm_s) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float x[16];
    tmem_ld16(tmem_addr(tm_c, 64 * CQ + 16 * c), x);
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float h0 = x[i] + a.b0h[64 * CQ + 16 * c + i], h1 = x[i + 1] + a.b0h[64 * CQ + 16 * c + i + 1];
      pk[i >> 1] = pack_h2(fmaf(h0, tanh_fast(h0), h0), fmaf(h1, tanh_fast(h1), h1));
    }
    tmem_st8(tmem_addr(tm_s, 32 * CQ + 8 * c), pk);
  }
  tmem_wait_st();
}

__global__ void __launch_bounds__(EQ_THREADS, 1) k_equi(const __grid_constant__ EquiArgs a) {
  if (a.skip_if_uniform && a.nonuni != nullptr && *a.nonuni == 0) return;     // uniform conditioning: equi_lin.cu does this block
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* X = smem + EQ_X;
  uint8_t* U = smem + EQ_U;
  uint8_t* misc = smem + EQ_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);   // 0: coord_mlp.0 image, 1: e tile, 2: input_lin image, 3/5: MMA in (N halves), 4/6: MMA c0 (N halves)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 96);    // bars: ... 7: MMA coord_mlp.2, 8: its weight image
  float2* LNS = reinterpret_cast<float2*>(smem + EQ_LNS);
  float4* C3 = reinterpret_cast<float4*>(smem + EQ_C3);
  uint32_t* gt_meta = reinterpret_cast<uint32_t*>(smem + EQ_GT);
  int* gt_node = reinterpret_cast<int*>(gt_meta + 64);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int rq = warp & 3, cq = warp >> 2;
  const int row = rq * 32 + lane;
  const int per = (a.p.n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, a.p.n_tiles);
  const uint8_t* win_img = static_cast<const uint8_t*>(a.win_img);   // (eight replicas to spread the per-tile re-fetch over L2: no gain)

  if (t == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    if (tile0 < tile1) {
      mbar_expect_tx(&bars[2], 65536);
      bulk_g2s(X, win_img, 65536, &bars[2]);
    }
    mbar_expect_tx(&bars[0], 131072);
    bulk_g2s(smem + EQ_WC0, a.wc0_img, 131072, &bars[0]);
  }
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_x = tmem, tm_c = tmem + 256;
  uint32_t par = 0;
  const float4* pos = reinterpret_cast<const float4*>(a.pos_in);
  float4* pos_out = reinterpret_cast<float4*>(a.pos_out);
  const int cb = 64 * cq;                         // first hidden column of this thread

  // Software pipeline across tiles: the row metadata, positions and distance features of tile i+1 are produced while
  // the coord_mlp.0 MMA of tile i runs; the hoisted per-atom parts of tile i are gathered and pre-added (half2) while
  // its input_lin MMA runs.
  const int tfirst = min(tile0, a.p.n_tiles - 1);
  RowInfo r = load_row(a.p, tfirst, row);
  int ng = a.p.tile_ngroups[tfirst];
  uint8_t ex = a.extra[r.pr];
  // the e chunk of a tile = its rows' pair rows, gathered from the pair-row store one tile ahead (edge_common.cuh):
  // the four warps that own a row quarter copy 8 of its rows each
  if (tile0 < tile1) gather_e16_warp<8>(U, a.e16, 32 * rq, 8 * cq, r.valid, r.pr, lane);
  float4 pg = pos[r.g], pj = pos[r.j];
  uint4 dfh[2];
  uint4 ab[8];                                    // A[g] + B[j] of this thread's 64 hidden units (fp16 pairs), one tile ahead
  RowInfo rn = r;
  int ngn = ng;
  uint8_t exn = ex;
  float4 pgn = pg, pjn = pj;
#ifdef JODO_PHASE_TIMING
  long long ph_last = clock64();
#endif
  // tile0 - 1 is a pipeline fill step: it only produces the distance features of tile0
  for (int tile = tile0 - 1; tile < tile1; ++tile) {
    if (tile >= tile0) {
    const float* tr = a.tab + (size_t)(uni ? 0 : r.mol) * a.ld_tab + a.tab_off;
    // distance features of this tile (computed one tile ahead) -> U chunk 1
#pragma unroll
    for (int p = 0; p < 2; ++p)
      *reinterpret_cast<uint4*>(U + img_piece(row, 1, 2 * cq + p, CHUNK_BYTES_A)) = dfh[p];
    // row metadata of the next tile
    const int nt_ = min(tile + 1, tile1 - 1);
    rn = load_row(a.p, nt_, row);
    ngn = a.p.tile_ngroups[nt_];
    cp_async_wait_all();                           // this thread's share of the gathered e chunk has landed
    fence_async_smem();
    sync_tc();
    PHASE_MARK(0);
    if (t == 0) {
      PHASE_MARK(1);
      mbar_wait(&bars[2], par);
      PHASE_MARK(2);
      tc_fence_after();
      // input_lin edge part, hidden units [0,128) then [128,256): the column quarters 0,1 start on the first half
      // while the tensor core works on the second
      const uint32_t idesc = umma_idesc_f16(128);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tm_x + 128 * hf, umma_desc_sw128(smem_u32(U) + (k >> 2) * CHUNK_BYTES_A + (k & 3) * 32),
                   umma_desc_sw128(smem_u32(X) + (k >> 2) * 32768 + hf * 16384 + (k & 3) * 32), idesc, k ? 1u : 0u);
        umma_commit(&bars[hf ? 5 : 3]);
      }
    }
    // the hoisted input_lin parts A[g] + B[j] of this tile were gathered (and pre-added as half2) one tile ahead
    mbar_wait(&bars[cq < 2 ? 3 : 5], par);
    PHASE_MARK(3);
    tc_fence_after();

    // ---- pass 1: x = acc + (A[g] + B[j]) over this thread's 64 hidden units, kept in TMEM; row statistics
    float mean, rstd;
    {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float x[16];
        tmem_ld16(tmem_addr(tm_x, cb + 16 * q), x);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t* u2 = reinterpret_cast<const uint32_t*>(&ab[2 * q + i]);
#pragma unroll
          for (int w = 0; w < 4; ++w) {              // fp32 accumulator + one half of the packed pair in one instruction (FHADD)
            const float v0 = fhadd_lo(u2[w], x[8 * i + 2 * w]), v1 = fhadd_hi(u2[w], x[8 * i + 2 * w + 1]);
            x[8 * i + 2 * w] = v0; x[8 * i + 2 * w + 1] = v1;
            s1 += v0; s2 = fmaf(v0, v0, s2);
            s1 += v1; s2 = fmaf(v1, v1, s2);
          }
        }
        tmem_st16(tmem_addr(tm_x, cb + 16 * q), x);
      }
      tmem_wait_st();
      if (cq < 2) mbar_wait(&bars[5], par);      // the scratch aliases the GBF chunk: the whole input_lin MMA must be done
      exn = a.extra[rn.pr];                      // (a dependent load behind row_pair: issued here, where rn has long arrived)
      if (tile + 1 < tile1)                      // U chunk 0 is consumed: gather the next tile's e rows (rn)
        gather_e16_warp<8>(U, a.e16, 32 * rq, 8 * cq, rn.valid, rn.pr, lane);
      if (t == 0) {
        mbar_expect_tx(&bars[8], 8192);          // the GBF chunk is consumed too: its tail takes the coord_mlp.2 image
        bulk_g2s(smem + EQ_W2, a.w2_img, 8192, &bars[8]);
      }
      if (cq == 0 && r.valid && row == r.gs) { gt_meta[r.gi] = (uint32_t)r.gs | ((uint32_t)r.gl << 8); gt_node[r.gi] = r.g; }
      LNS[row * 4 + cq] = make_float2(s1, s2);
      __syncthreads();
      PHASE_MARK(4);
      const float4 o01 = *reinterpret_cast<const float4*>(&LNS[row * 4]);
      const float4 o23 = *reinterpret_cast<const float4*>(&LNS[row * 4 + 2]);
      mean = (o01.x + o01.z + o23.x + o23.z) * (1.0f / 256.0f);
      rstd = rsqrtf(fmaxf((o01.y + o01.w + o23.y + o23.w) * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-6f);
    }
    // positions of the next tile's rows: in flight during pass 2
    pgn = pos[rn.g]; pjn = pos[rn.j];
    // ---- pass 2: LN + modulate -> X chunk cq (fp16, K = 256); the input_lin image there is no longer needed.
    // Padding rows carry finite garbage; they only feed their own (discarded) output rows.
    if (uni) { EQ_DISPATCH(eq_pass2_uni, tm_x, X, row, mean, rstd, tr); }
    else eq_pass2_gen(tm_x, X, row, cq, mean, rstd, tr);
    fence_async_smem();
    sync_tc();
    PHASE_MARK(5);
    if (t == 0) {
      if (tile == tile0) mbar_wait(&bars[0], 0);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_f16(128);                                // coord_mlp.0 (pre-scaled by 1/2), N halves
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
          umma_f16(tm_c + 128 * hf, umma_desc_sw128(smem_u32(X) + (k >> 2) * CHUNK_BYTES_A + (k & 3) * 32),
                   umma_desc_sw128(smem_u32(smem + EQ_WC0) + (k >> 2) * 32768 + hf * 16384 + (k & 3) * 32), idesc, k ? 1u : 0u);
        umma_commit(&bars[hf ? 6 : 4]);
      }
    }
    }   // tile >= tile0
    // under the MMA: distance features of the next tile
    {
      const float* trn = a.tab + (size_t)(uni ? 0 : rn.mol) * a.ld_tab + a.tab_off;
      const float gsc = uni ? c_eqmod[512] : trn[tab_gbf(D_)], gsh = uni ? c_eqmod[513] : trn[tab_gbf(D_) + 1];
      EQ_DISPATCH(eq_gbf, a, sq_dist(pgn, pjn), gsc, gsh, dfh);
    }
    // ... and its hoisted input_lin parts (piece-major fp16 rows, bias folded into the g part), pre-added as half2: the
    // gather (1 KB per edge from L2) runs under the coord_mlp.0 MMA instead of in front of pass 1
    {
      const uint4* pa = static_cast<const uint4*>(a.AB) + (size_t)(8 * cq) * a.ldab + rn.g;
      const uint4* pb = static_cast<const uint4*>(a.AB) + (size_t)(32 + 8 * cq) * a.ldab + rn.j;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 ua[4], ub[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ua[i] = __ldg(pa + (size_t)(4 * h + i) * a.ldab); ub[i] = __ldg(pb + (size_t)(4 * h + i) * a.ldab); }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half2* x2 = reinterpret_cast<__half2*>(&ua[i]);
          const __half2* y2 = reinterpret_cast<const __half2*>(&ub[i]);
#pragma unroll
          for (int k = 0; k < 4; ++k) x2[k] = __hadd2(x2[k], y2[k]);
          ab[4 * h + i] = ua[i];
        }
      }
    }
    if (tile >= tile0) {
    mbar_wait(&bars[cq < 2 ? 4 : 6], par);
    PHASE_MARK(6);
    tc_fence_after();

    // ---- SiLU -> fp16 A operand in tensor memory (the columns of x, dead since pass 2); coord_mlp.2 on the tensor core
    EQ_DISPATCH(eq_silu_tm, a, tm_c, tm_x);
    if (t == 0) {                                // X is consumed once both halves are done: fetch the input_lin image for the next tile
      mbar_wait(&bars[6], par);
      if (tile + 1 < tile1) {
        mbar_expect_tx(&bars[2], 65536);
        bulk_g2s(X, win_img, 65536, &bars[2]);
      }
    }
    sync_tc();
    PHASE_MARK(7);
    if (t == 0) {
      mbar_wait(&bars[8], par);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_f16(16);
#pragma unroll
      for (int k = 0; k < 16; ++k)               // K = 16 per step = 8 columns of packed pairs
        umma_f16_ts(tm_x + 128, tm_x + 8 * k, umma_desc_sw128(smem_u32(smem + EQ_W2) + (k >> 2) * 2048 + (k & 3) * 32), idesc, k ? 1u : 0u);
      umma_commit(&bars[7]);
    }
    if (cq == 0) {       // tanh, adjacency-weighted mean, coordinate contribution of this edge
      mbar_wait(&bars[7], par);
      tc_fence_after();
      float dd[16];
      tmem_ld16(tmem_addr(tm_x, 128), dd);
      const float w = (tanh_fast(dd[0]) + ((ex & 1) ? tanh_fast(dd[1]) : 0.f) + ((ex & 2) ? tanh_fast(dd[2]) : 0.f)) * (1.0f / 3.0f);
      const float dx = pg.x - pj.x, dy = pg.y - pj.y, dz = pg.z - pj.z;
      const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
      const float f = r.valid ? a.coord_scale * w / fmaxf(nrm, 1e-8f) : 0.f;
      C3[row] = make_float4(dx * f, dy * f, dz * f, 0.f);
    }
    tc_fence_before();
    __syncthreads();
    // per-atom sums of the coordinate contributions: one warp per group, lanes over its rows, shuffle tree
    for (int gi = warp; gi < ng; gi += EQ_THREADS / 32) {
      const int gs = gt_meta[gi] & 255u, gl = (gt_meta[gi] >> 8) & 255u;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int k = lane; k < gl; k += 32) { const float4 c = C3[gs + k]; sx += c.x; sy += c.y; sz += c.z; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
      }
      if (lane == 0) {
        const int node = gt_node[gi];
        const float4 p0 = pos[node];
        pos_out[node] = make_float4(p0.x + sx, p0.y + sy, p0.z + sz, 0.f);
      }
    }
    // no barrier here: C3 and the group table are dedicated buffers that are rewritten only after the next tile's
    // barriers; the scratch inside the GBF chunk (LNS, coord_mlp.2 image) was last read before the barrier above
    PHASE_MARK(8);
    par ^= 1;
    }   // tile >= tile0
    r = rn; ng = ngn; ex = exn; pg = pgn; pj = pjn;
  }
  if (tile0 >= tile1 && t == 0) mbar_wait(&bars[0], 0);   // never leave with bulk copies in flight
  sync_tc();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace

#ifdef JODO_PHASE_TIMING
extern "C" int jodo_debug_equi_phases(long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (out16) cudaMemcpyFromSymbol(out16, g_equi_phase, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_equi_phase, z, sizeof(z)); }
  return 0;
}
#endif

cudaError_t launch_equi(const EquiArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_equi, EQ_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  if ((e0 = const_tables_acquire(st)) != cudaSuccess) return e0;
  if (a.nonuni) {     // row 0 of the table feeds the uniform fast path (harmless when the batch is not uniform)
    cudaError_t e = cudaMemcpyToSymbolAsync(c_eqmod, a.tab + a.tab_off + tab_equi(D_), sizeof(float) * 528, 0,
                                            cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  }
  const int grid = a.p.n_tiles < num_sms ? a.p.n_tiles : num_sms;
  k_equi<<<grid, EQ_THREADS, EQ_SMEM, st>>>(a);
  if ((e0 = cudaGetLastError()) != cudaSuccess) return e0;
  return const_tables_release(st);
}

}  // namespace jodo
