// Coordinate update of one DGT block under UNIFORM conditioning, with coord_mlp.0 composed into input_lin.
//
// reference MultiCondEquiUpdate.forward, models/mol_gnn.py:71-94:
//   x = input_lin(u),  inv = LN(x) * (1 + scale) + shift,  c = coord_mlp.0(inv),  out = tanh(coord_mlp.2(SiLU(c)))
// LayerNorm without affine is  LN(x) = C x / sigma  with the centering matrix C = I - 11^T / D, and the modulation row
// (shift, scale) is the same for every edge when every molecule carries the same noise level (always in unconditional
// sampling: sampling.py:549 broadcasts one noise level).  Then
//   c = W0 (diag(1 + scale) C x / sigma + shift) + b0 = (M x) / sigma + d,   M = W0 diag(1 + scale) C,  d = W0 shift + b0,
// and M x is linear in u = [h_row | h_col | e | dist]:  M x = (M W_in) u + M b_in.  k_equi_compose forms M W_in once
// per call and block (0.17 GFLOP) from the table row of the step; the per-edge kernel then needs ONE K = 128 tensor-core
// product with N = 512 -- x (for the row statistics) and y = (M W_e)[e | dist] -- plus the hoisted per-atom parts of
// both (A[g] + B[j], YA[g] + YB[j]); the 256 x 256 product per edge, the LayerNorm operand image and the SiLU -> tensor
// memory -> coord_mlp.2 product of equi.cu disappear (coord_mlp.2's three dot products ride on the SiLU pass).
// Weights resident: input_lin edge part + its composed twin (2 x 64 KB); two [e | dist] tile buffers; tensor memory:
// x in columns [0, 256), y in [256, 512).  Per tile: MMA_x(i+1) runs under pass 2 of tile i, MMA_y(i+1) under the tail
// of tile i and pass 1 of tile i+1; three CTA barriers per tile.
// The general (per-molecule conditioning) case stays on equi.cu; both kernels test the device flag and one returns.
#include "edge_common.cuh"

namespace jodo {

// d' [256] | coord_mlp.2 rows 0..2 [3][256] | GBF (1 + scale, shift): uploaded from the compose kernel's staging row
__constant__ float c_eqlin[1040];

#ifdef JODO_PHASE_TIMING
__device__ long long g_equi_lin_phase[16];
#define EL_MARK(i) do { if (t == 0 && blockIdx.x == 0) { long long c_ = clock64(); g_equi_lin_phase[i] += c_ - ph_last; ph_last = c_; } } while (0)
#else
#define EL_MARK(i) do { } while (0)
#endif

namespace {

constexpr int EL_THREADS = 512;
constexpr int EL_WIN = 0;                        // 64 KB: input_lin edge part image (N = 256, K = 128: [e | dist])
constexpr int EL_WCE = 65536;                    // 64 KB: composed image (M W_e) / 2, same layout
constexpr int EL_U = 131072;                     // 2 x 32 KB: [e | GBF(d)] tiles (K = 128), double buffered
constexpr int EL_MISC = EL_U + 65536;            // barriers + tmem slot (128 B)
constexpr int EL_LNS = EL_MISC + 128;            // [128][4] float2: per column quarter (sum, sum of squares)
constexpr int EL_P3 = EL_LNS + 128 * 4 * 8;      // [128][4] float4: per column quarter partial coord_mlp.2 dots
constexpr int EL_C3 = EL_P3 + 128 * 4 * 16;      // [128] float4: per-row coordinate contribution
constexpr int EL_GT = EL_C3 + 128 * 16;          // group table: start | len << 8 [64], atom [64]
constexpr int EL_SMEM = EL_GT + 512;
static_assert(EL_SMEM <= 232448, "shared memory budget");

#define EL_DISPATCH(F, ...)                  \
  switch (cq) {                              \
    case 0: F<0>(__VA_ARGS__); break;        \
    case 1: F<1>(__VA_ARGS__); break;        \
    case 2: F<2>(__VA_ARGS__); break;        \
    default: F<3>(__VA_ARGS__); break;       \
  }

// distance features, columns [16 CQ, 16 CQ + 16) of the GBF chunk (constants are kernel-parameter operands)
template <int CQ>
__device__ __forceinline__ void el_gbf(const EquiLinArgs& a, float d, uint4 (&out)[2]) {
  const float x = fmaf(d, c_eqlin[1024], c_eqlin[1025]);               // the table stores 1 + scale
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

// pass 2 over hidden units [64 CQ, 64 CQ + 64): h = (y + yab) * rstd / 2 + d / 2 (the halves are folded into the composed
// image and d'), SiLU(2h) = h + h tanh(h), and the three coord_mlp.2 dot products over these columns.  The per-column
// constants are constant-bank operands (measured: the same loop fed by broadcast 16-byte shared-memory loads, one code
// copy for all column quarters, is 12 % slower).
template <int CQ>
__device__ __forceinline__ void el_pass2(uint32_t tm_y, const uint4 (&yab)[8], float rstd, float (&p)[3]) {
  p[0] = p[1] = p[2] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float y[16];
    tmem_ld16(tmem_addr(tm_y, 64 * CQ + 16 * q), y);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t* u2 = reinterpret_cast<const uint32_t*>(&yab[2 * q + i]);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int col = 64 * CQ + 16 * q + 8 * i + 2 * w + hh;
          const float v = hh ? fhadd_hi(u2[w], y[8 * i + 2 * w + 1]) : fhadd_lo(u2[w], y[8 * i + 2 * w]);
          const float h = fmaf(v, rstd, c_eqlin[col]);
          const float s = fmaf(h, tanh_fast(h), h);
          p[0] = fmaf(s, c_eqlin[256 + col], p[0]);
          p[1] = fmaf(s, c_eqlin[512 + col], p[1]);
          p[2] = fmaf(s, c_eqlin[768 + col], p[2]);
        }
      }
    }
  }
}

// Eight 16-byte pieces (this thread's 64 hidden units) of one per-atom part: piece0 = 0 A, 32 B, 64 YA, 96 YB; `v` = g or j.
// The partner parts (B[j], YB[j]: one L2 line per lane) are issued a phase ahead of their use, the group parts (A[g], YA[g]:
// a few lines per warp) closer to it, and the two are added as half2 only when both have had time to land.
__device__ __forceinline__ void el_load8(const EquiLinArgs& a, int piece0, int cq, int v, uint4 (&out)[8]) {
  const uint4* p = static_cast<const uint4*>(a.AB) + (size_t)(piece0 + 8 * cq) * a.ldab + v;
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = __ldg(p + (size_t)i * a.ldab);
}
__device__ __forceinline__ void el_add8(uint4 (&acc)[8], const uint4 (&b)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __half2* x2 = reinterpret_cast<__half2*>(&acc[i]);
    const __half2* y2 = reinterpret_cast<const __half2*>(&b[i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) x2[k] = __hadd2(x2[k], y2[k]);
  }
}

// per-row state kept across a tile: 6 registers instead of RowInfo's 8 + two float4 positions
struct ElRow { int g, j, pr; uint32_t meta; float dx, dy, dz; };       // meta: start | len << 8 | group << 16, bit 31 = valid row
__device__ __forceinline__ ElRow el_row(const Plan& p, int tile, int t) {
  ElRow r;
  const int R = tile * TILE_ROWS + t;
  const int g = p.row_g[R];
  const bool valid = g >= 0;
  r.g = valid ? g : 0;
  r.j = valid ? p.row_j[R] : 0;
  r.meta = (p.row_meta[R] & 0xFFFFFFu) | (valid ? 0x80000000u : 0u);
  r.pr = valid ? p.row_pair[R] : 0;
  r.dx = r.dy = r.dz = 0.f;
  return r;
}
__device__ __forceinline__ void el_delta(ElRow& r, const float4* __restrict__ pos) {
  const float4 pg = pos[r.g], pj = pos[r.j];
  r.dx = pg.x - pj.x; r.dy = pg.y - pj.y; r.dz = pg.z - pj.z;
}

// one K = 128 product of a [e | dist] tile with a resident N = 256 image, as two N = 128 halves
__device__ __forceinline__ void el_mma(uint32_t tm, const uint8_t* U, const uint8_t* W, uint64_t* bar) {
  const uint32_t idesc = umma_idesc_f16(128);
#pragma unroll
  for (int hf = 0; hf < 2; ++hf)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_f16(tm + 128 * hf, umma_desc_sw128(smem_u32(U) + (k >> 2) * CHUNK_BYTES_A + (k & 3) * 32),
               umma_desc_sw128(smem_u32(W) + (k >> 2) * 32768 + hf * 16384 + (k & 3) * 32), idesc, k ? 1u : 0u);
  umma_commit(bar);
}

__global__ void __launch_bounds__(EL_THREADS, 1) k_equi_lin(const __grid_constant__ EquiLinArgs a) {
  if (*a.nonuni != 0) return;                     // per-molecule conditioning: equi.cu does this block
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* misc = smem + EL_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);            // 0: weight images, 1: MMA x, 2: MMA y
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 96);
  float2* LNS = reinterpret_cast<float2*>(smem + EL_LNS);
  float4* P3 = reinterpret_cast<float4*>(smem + EL_P3);
  float4* C3 = reinterpret_cast<float4*>(smem + EL_C3);
  uint32_t* gt_meta = reinterpret_cast<uint32_t*>(smem + EL_GT);
  int* gt_node = reinterpret_cast<int*>(gt_meta + 64);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int rq = warp & 3, cq = warp >> 2;
  const int row = rq * 32 + lane;
  const int per = (a.p.n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, a.p.n_tiles);
  if (tile0 >= tile1) return;

  if (t == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 131072);
    bulk_g2s(smem + EL_WIN, a.win_img, 65536, &bars[0]);
    bulk_g2s(smem + EL_WCE, a.wce_img, 65536, &bars[0]);
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_x = tmem, tm_y = tmem + 256;
  const float4* pos = reinterpret_cast<const float4*>(a.pos_in);
  float4* pos_out = reinterpret_cast<float4*>(a.pos_out);
  const int cb = 64 * cq;

  // ---- prologue: tile0's operands, both of its products in flight
  ElRow r = el_row(a.p, tile0, row);
  int ng = a.p.tile_ngroups[tile0];
  uint8_t ex = a.extra[r.pr];
  el_delta(r, pos);
  uint4 ab[8];
  {
    uint8_t* U0 = smem + EL_U;
    gather_e16_warp<8>(U0, a.e16, 32 * rq, 8 * cq, (r.meta >> 31) != 0, r.pr, lane);
    uint4 dfh[2];
    EL_DISPATCH(el_gbf, a, r.dx * r.dx + r.dy * r.dy + r.dz * r.dz, dfh);
#pragma unroll
    for (int p = 0; p < 2; ++p) *reinterpret_cast<uint4*>(U0 + img_piece(row, 1, 2 * cq + p, CHUNK_BYTES_A)) = dfh[p];
    {
      uint4 ag[8];
      el_load8(a, 32, cq, r.j, ab);
      el_load8(a, 0, cq, r.g, ag);
      el_add8(ab, ag);
    }
    cp_async_wait_all();
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      el_mma(tm_x, U0, smem + EL_WIN, &bars[1]);
      el_mma(tm_y, U0, smem + EL_WCE, &bars[2]);
    }
  }
  ElRow rn = el_row(a.p, min(tile0 + 1, tile1 - 1), row);
  int ngn = a.p.tile_ngroups[min(tile0 + 1, tile1 - 1)];
  float4 pgn = pos[rn.g], pjn = pos[rn.j];         // positions of the next tile's rows, loaded a phase ahead of their use
  uint32_t par = 0;
#ifdef JODO_PHASE_TIMING
  long long ph_last = clock64();
#endif

  for (int tile = tile0; tile < tile1; ++tile) {
    const bool more = tile + 1 < tile1;
    uint8_t* Un = smem + EL_U + (((tile - tile0) & 1) ^ 1) * 32768;
    // next tile's adjacency bits: in flight during pass 1; the atoms of the tile after it (their positions are fetched
    // after pass 2 and used at the next tile's top: two dependent round trips, both off the critical path)
    const uint8_t exn = a.extra[rn.pr];
    const int R2 = min(tile + 2, tile1 - 1) * TILE_ROWS + row;
    const int g2 = max(a.p.row_g[R2], 0), j2 = max(a.p.row_j[R2], 0);

    // next tile's [e | dist] operand into the other buffer (its last reader, MMA y of tile - 1, completed before pass 2
    // of tile - 1): the pair-row copies are in flight during pass 1
    if (more) {
      rn.dx = pgn.x - pjn.x; rn.dy = pgn.y - pjn.y; rn.dz = pgn.z - pjn.z;
      gather_e16_warp<8>(Un, a.e16, 32 * rq, 8 * cq, (rn.meta >> 31) != 0, rn.pr, lane);
      uint4 dfh[2];
      EL_DISPATCH(el_gbf, a, rn.dx * rn.dx + rn.dy * rn.dy + rn.dz * rn.dz, dfh);
#pragma unroll
      for (int p = 0; p < 2; ++p) *reinterpret_cast<uint4*>(Un + img_piece(row, 1, 2 * cq + p, CHUNK_BYTES_A)) = dfh[p];
    }
    // the composed partner part YB[j] of this tile: in flight during pass 1
    uint4 yab[8];
    el_load8(a, 96, cq, r.j, yab);
    EL_MARK(0);
    // ---- pass 1: x = acc + (A[g] + B[j]) over this thread's 64 hidden units -> row statistics (x is not kept)
    mbar_wait(&bars[1], par);
    EL_MARK(1);
    tc_fence_after();
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
          for (int w = 0; w < 4; ++w) {
            const float v0 = fhadd_lo(u2[w], x[8 * i + 2 * w]), v1 = fhadd_hi(u2[w], x[8 * i + 2 * w + 1]);
            s1 += v0; s2 = fmaf(v0, v0, s2);
            s1 += v1; s2 = fmaf(v1, v1, s2);
          }
        }
      }
      LNS[row * 4 + cq] = make_float2(s1, s2);
    }
    EL_MARK(2);
    // the composed group part YA[g]: in flight across the barrier
    uint4 part[8];
    el_load8(a, 64, cq, r.g, part);
    if (more) {                                    // the next operand's copies and stores have landed
      cp_async_wait_all();
      fence_async_smem();
    }
    sync_tc();                                     // statistics visible; every x read done; next operand complete
    EL_MARK(3);
    if (t == 0 && more) el_mma(tm_x, Un, smem + EL_WIN, &bars[1]);        // x of the next tile, under pass 2
    // (behind the barrier: other warps may still have been summing the previous tile's groups)
    if (cq == 0 && (r.meta >> 31) && row == (int)(r.meta & 255u)) { gt_meta[(r.meta >> 16) & 255u] = r.meta & 0xFFFFu; gt_node[(r.meta >> 16) & 255u] = r.g; }

    // ---- pass 2: SiLU((y + yab) / sigma + d) and the three coord_mlp.2 dot products
    float rstd;
    {
      const float4 o01 = *reinterpret_cast<const float4*>(&LNS[row * 4]);
      const float4 o23 = *reinterpret_cast<const float4*>(&LNS[row * 4 + 2]);
      const float mean = (o01.x + o01.z + o23.x + o23.z) * (1.0f / 256.0f);
      rstd = rsqrtf(fmaxf((o01.y + o01.w + o23.y + o23.w) * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-6f);
    }
    el_add8(yab, part);
    EL_MARK(4);
    mbar_wait(&bars[2], par);
    EL_MARK(5);
    tc_fence_after();
    float p3[3];
    EL_DISPATCH(el_pass2, tm_y, yab, rstd, p3);
    P3[row * 4 + cq] = make_float4(p3[0], p3[1], p3[2], 0.f);
    EL_MARK(6);
    // the hoisted parts and row metadata of the following tiles: in flight across the tail
    if (more) el_load8(a, 32, cq, rn.j, ab);                                // B[j'] of the next tile
    const int t2 = min(tile + 2, tile1 - 1);
    const ElRow r2 = el_row(a.p, t2, row);
    pgn = pos[g2]; pjn = pos[j2];                  // (consumed by now: the next tile's deltas were formed at this tile's top)
    const int ng2 = a.p.tile_ngroups[t2];
    sync_tc();                                     // every y read done; partial dots visible
    EL_MARK(7);
    if (t == 0 && more) el_mma(tm_y, Un, smem + EL_WCE, &bars[2]);        // y of the next tile, under the tail and pass 1

    if (cq == 0) {       // tanh, adjacency-weighted mean, coordinate contribution of this edge
      const float4 q0 = P3[row * 4], q1 = P3[row * 4 + 1], q2 = P3[row * 4 + 2], q3 = P3[row * 4 + 3];
      const float d0 = (q0.x + q1.x) + (q2.x + q3.x), d1 = (q0.y + q1.y) + (q2.y + q3.y), d2 = (q0.z + q1.z) + (q2.z + q3.z);
      const float w = (tanh_fast(d0) + ((ex & 1) ? tanh_fast(d1) : 0.f) + ((ex & 2) ? tanh_fast(d2) : 0.f)) * (1.0f / 3.0f);
      const float dx = r.dx, dy = r.dy, dz = r.dz;
      const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
      const float f = (r.meta >> 31) ? a.coord_scale * w / fmaxf(nrm, 1e-8f) : 0.f;
      C3[row] = make_float4(dx * f, dy * f, dz * f, 0.f);
    }
    if (more) el_load8(a, 0, cq, rn.g, part);                               // A[g'] of the next tile, under the group sums
    EL_MARK(8);
    __syncthreads();
    EL_MARK(9);
    // per-atom sums of the coordinate contributions: one warp per group, lanes over its rows, shuffle tree
    for (int gi = warp; gi < ng; gi += EL_THREADS / 32) {
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
    // C3 / the group table are rewritten only after the next tile's first barrier; LNS after its pass 1 (behind this
    // tile's barriers); P3 after its second barrier
    EL_MARK(10);
    if (more) el_add8(ab, part);
    EL_MARK(11);
    par ^= 1;
    r = rn; ng = ngn; ex = exn;
    rn = r2; ngn = ng2;
  }
  sync_tc();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---- composition: per block,  M = W0 diag(s) C / 2  (s = 1 + scale of the step's table row, C = centering),
//   [YA | YB | Y_e] = M [W_A | W_B | W_e]  (input_lin [256, 640] in its own column order),  yb = M b_in,
//   d' = (W0 shift + b0) / 2.
// CTA = 64 x 64 output tile of one block: grid (L, 4 row tiles, 10 column tiles), 256 threads, 4 x 4 outputs each.
constexpr int CP_THREADS = 256;
__global__ void __launch_bounds__(CP_THREADS) k_equi_compose(const jodo_equi_compose_item* __restrict__ items,
                                                              const float* __restrict__ tab_row0, const int* __restrict__ nonuni) {
  if (*nonuni != 0) return;
  __shared__ float Ms[16][65];       // [m][n]: M tile, transposed
  __shared__ float Bs[16][64];       // [m][k]
  __shared__ float rm[64];           // row means of W0 diag(s)
  __shared__ float sv[256], shv[256];
  const jodo_equi_compose_item it = items[blockIdx.x];
  const int n0 = 64 * blockIdx.y, k0 = 64 * blockIdx.z;
  const int tid = threadIdx.x;
  const float* tr = tab_row0 + it.tab_off;
  for (int i = tid; i < 256; i += CP_THREADS) { shv[i] = tr[i]; sv[i] = tr[256 + i]; }     // equi (shift | 1 + scale)
  __syncthreads();
  {  // row means, d' and yb for this CTA's 64 rows: 4 threads per row
    const int rr = tid >> 2, part = tid & 3;
    const float* w0 = it.w0 + (size_t)(n0 + rr) * 256;
    float sm = 0.f, sd = 0.f;
    for (int m = part; m < 256; m += 4) { const float w = w0[m]; sm = fmaf(w, sv[m], sm); sd = fmaf(w, shv[m], sd); }
    sm += __shfl_xor_sync(0xffffffffu, sm, 1); sm += __shfl_xor_sync(0xffffffffu, sm, 2);
    sd += __shfl_xor_sync(0xffffffffu, sd, 1); sd += __shfl_xor_sync(0xffffffffu, sd, 2);
    const float mean = sm * (1.0f / 256.0f);
    if (part == 0) rm[rr] = mean;
    if (blockIdx.z == 0) {
      float yb = 0.f;
      for (int m = part; m < 256; m += 4) yb = fmaf(0.5f * (w0[m] * sv[m] - mean), it.bi[m], yb);
      yb += __shfl_xor_sync(0xffffffffu, yb, 1); yb += __shfl_xor_sync(0xffffffffu, yb, 2);
      if (part == 0) {
        it.ab_bias[512 + n0 + rr] = yb;
        it.consts[n0 + rr] = 0.5f * (sd + it.b0[n0 + rr]);
      }
    }
  }
  if (blockIdx.y == 0 && blockIdx.z == 0) {
    for (int i = tid; i < 768; i += CP_THREADS) it.consts[256 + i] = it.w2[i];
    if (tid < 2) it.consts[1024 + tid] = tab_row0[it.tab_off + 2 * 256 + tid];              // GBF (1 + scale, shift) follows the equi rows
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int m0 = 0; m0 < 256; m0 += 16) {
    for (int e = tid; e < 1024; e += CP_THREADS) {
      const int rr = e >> 4, mm = e & 15;                                          // M tile: coalesced over m
      Ms[mm][rr] = 0.5f * (it.w0[(size_t)(n0 + rr) * 256 + m0 + mm] * sv[m0 + mm] - rm[rr]);
      const int bm = e >> 6, bk = e & 63;                                          // input_lin tile: coalesced over k
      Bs[bm][bk] = it.wi[(size_t)(m0 + bm) * 640 + k0 + bk];
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = Ms[mm][4 * ty + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[mm][4 * tx + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // fp16 operand images: columns [0, 256) -> tile 2 (YA), [256, 512) -> tile 3 (YB) of the per-atom GEMM's weight image
  // ([N / 256][K / 64][256 rows][128 B]); [512, 640) -> the edge kernel's composed image ([K / 64][256 rows][128 B])
  const int k = k0 + 4 * tx;                      // 4 consecutive k = half a 16-byte piece
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * ty + i;
    const uint2 hv = make_uint2(pack_h2(acc[i][0], acc[i][1]), pack_h2(acc[i][2], acc[i][3]));
    uint8_t* base;
    int kk;
    if (k < 512) { base = static_cast<uint8_t*>(it.ab_img) + (size_t)(2 + (k >> 8)) * (4 * 256 * 128); kk = k & 255; }
    else { base = static_cast<uint8_t*>(it.wce_img); kk = k - 512; }
    const int chunk = kk >> 6, piece = (kk & 63) >> 3;
    *reinterpret_cast<uint2*>(base + (size_t)chunk * (256 * 128) + (size_t)n * 128 + ((piece ^ (n & 7)) << 4) + ((kk & 7) << 1)) = hv;
  }
}

}  // namespace

#ifdef JODO_PHASE_TIMING
extern "C" int jodo_debug_equi_lin_phases(long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (out16) cudaMemcpyFromSymbol(out16, g_equi_lin_phase, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_equi_lin_phase, z, sizeof(z)); }
  return 0;
}
#endif

cudaError_t launch_equi_compose(const jodo_equi_compose_item* items_dev, int L, const float* tab_row0, const int* nonuni,
                                cudaStream_t st) {
  k_equi_compose<<<dim3(L, 4, 10), CP_THREADS, 0, st>>>(items_dev, tab_row0, nonuni);
  return cudaGetLastError();
}

cudaError_t launch_equi_lin(const EquiLinArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_equi_lin, EL_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  if ((e0 = const_tables_acquire(st)) != cudaSuccess) return e0;
  cudaError_t e = cudaMemcpyToSymbolAsync(c_eqlin, a.consts, sizeof(float) * 1026, 0, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return e;
  const int grid = a.p.n_tiles < num_sms ? a.p.n_tiles : num_sms;
  k_equi_lin<<<grid, EL_THREADS, EL_SMEM, st>>>(a);
  if ((e0 = cudaGetLastError()) != cudaSuccess) return e0;
  return const_tables_release(st);
}

}  // namespace jodo
