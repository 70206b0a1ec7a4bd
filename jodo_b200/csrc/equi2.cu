// E(3)-equivariant coordinate update of one DGT block on edge tiles -- CTA-pair version (cta_group::2).
//
// Same arithmetic as equi.cu (reference MultiCondEquiUpdate.forward, models/mol_gnn.py:71-94; CoorsNorm,
// models/layers.py:344-347): per directed edge row (g = row r, j = col c)
//   x = input_lin([h_g | h_j | e | GBF(d)]) = A[g] + B[j] + W_e [e | GBF(d)]      (hoisted per-atom parts A, B)
//   y = LN(x) (1 + scale) + shift;  z = SiLU(coord_mlp.0 y);  w = mean(tanh(coord_mlp.2 z) * [1, adj2d, adjsp])
//   pos[g] += sum_j (pos[g] - pos[j]) / max(|.|, 1e-8) * coord_scale * w.
//
// What bounded equi.cu was not the tensor pipe (21 % busy) but one tile in flight per SM with every phase in lock-step:
// 224 KB of operands per tile (coord_mlp.0 resident, the input_lin image re-fetched per tile into the buffer that later
// holds the LayerNorm output) left no room for a second tile.  Here two CTAs on the two SMs of a TPC form a pair: every
// tcgen05.mma runs with cta_group::2 over the 256 rows of the pair's two tiles, each CTA holding HALF of every weight
// image (N / 2 rows), so that all weights are resident (100 KB per CTA) and the tile buffers are private to a stage.
// Each CTA then runs a two-stage pipeline over its tiles with two groups of 8 warps:
//   stage A (warps 0-7)   MMA1 = input_lin edge part -> x (TMEM cols 0-255); under it the next tile's metadata, positions
//                         and distance features; + A[g] + B[j], row statistics; LayerNorm + modulation -> X (fp16 image)
//   stage B (warps 8-15)  coord_mlp.0 MMA (A = X) -> TMEM cols 256-511; SiLU packed in place as fp16 pairs; coord_mlp.2
//                         MMA with A from tensor memory; tanh, adjacency mean, per-atom sums -> new positions
// so stage A of tile i + 1 overlaps stage B of tile i, and both tensor-core waits are covered by the other stage's math.
// Hand-offs: mbarriers. rdy1 / rdy2 / rdy3 live in the leader (cluster rank 0) and collect "operand ready" arrivals of
// both CTAs before its elected threads issue MMA1 / coord_mlp.0 / coord_mlp.2; m1 / c0 / c2 are completed in both CTAs by
// tcgen05.commit.multicast.  Every barrier completes exactly once per step, so its parity is step & 1.
#include <cstdlib>
#include "edge_common.cuh"

namespace jodo {

#ifdef JODO_PHASE_TIMING
__device__ long long g_equi2_phase[16];
// stage A marks are taken by thread 0, stage B marks by thread 256, of CTA 0 (the leader of pair 0)
#define E2_MARK(i) do { if (c.lt == 0 && blockIdx.x == 0) { long long c_ = clock64(); g_equi2_phase[i] += c_ - ph_last; ph_last = c_; } } while (0)
#define E2_MARK_INIT long long ph_last = clock64()
#else
#define E2_MARK(i) do { } while (0)
#define E2_MARK_INIT do { } while (0)
#endif

namespace {

constexpr int E2_THREADS = 512;
constexpr int E2_GROUP = 256;
constexpr int E2_WC0 = 0;                      // 64 KB: this CTA's 128 rows of coord_mlp.0 / 2 (K = 256: 4 chunks of 16 KB)
constexpr int E2_WIN = 65536;                  // 32 KB: this CTA's 128 rows of the input_lin edge part (K = 128)
constexpr int E2_X = E2_WIN + 32768;           // 64 KB: LayerNorm-modulated operand (K = 256)
constexpr int E2_U = E2_X + 65536;             // 32 KB: [e | GBF(d)] (K = 128)
constexpr int E2_W2 = E2_U + 32768;            // 8 KB: this CTA's 16 rows of coord_mlp.2 padded to N = 32 (K = 256: 4 chunks of 2 KB)
constexpr int E2_LNS = E2_W2 + 8192;           // 2 KB: LayerNorm partial sums [128][2] float2
constexpr int E2_C3 = E2_LNS + 2048;           // 2 x 2 KB: per-row coordinate contribution [128] float4, by step parity
constexpr int E2_GT = E2_C3 + 4096;            // 2 x 512 B: group table start | len << 8 [64], atom [64]
constexpr int E2_GP = E2_GT + 1024;            // 2 x 1 KB: position of every group atom [64] float4
constexpr int E2_BAR = E2_GP + 2048;           // mbarriers + tmem slot
constexpr int E2_SMEM = E2_BAR + 128;
static_assert(E2_SMEM <= 232448, "shared memory budget");
static_assert(E2_W2 % 1024 == 0 && E2_X % 1024 == 0 && E2_U % 1024 == 0, "operand images need 1024-byte alignment");

enum { B_W = 0, B_RDY1, B_M1, B_RDY2, B_C0, B_RDY3, B_C2, B_COUNT };

// row 0 of the table for the uniform-conditioning fast path: shift[256] | 1 + scale[256] | GBF (1 + scale, shift)
__constant__ float c_e2mod[528];
__constant__ int c_e2flags;            // debug switches (JODO_E2_FLAGS): 1 = spin on the pair barriers

__device__ __forceinline__ void ga_sync() { tc_fence_before(); named_bar_sync(1, E2_GROUP); tc_fence_after(); }
__device__ __forceinline__ void gb_sync() { tc_fence_before(); named_bar_sync(2, E2_GROUP); tc_fence_after(); }

struct E2Ctx {
  uint8_t* smem; uint64_t* bars; uint32_t tm_x, tm_c;
  int rank, lt, row, rq, lane, pt0, pt1, nsteps;
};
__device__ __forceinline__ int e2_tile(const E2Ctx& c, int s) {        // this CTA's tile of step s (clamped on the odd tail)
  const int t = c.pt0 + 2 * s + c.rank;
  return t < c.pt1 ? t : c.pt1 - 1;
}

// distance features, columns [32 CH, 32 CH + 32) of the GBF chunk -> four packed 16-byte pieces
template <int CH>
__device__ __forceinline__ void e2_gbf(const EquiArgs& a, float d, float scale, float shift, uint4 (&out)[4]) {
  const float x = fmaf(d, scale, shift);               // the table stores 1 + scale
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float df[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = 32 * CH + 8 * p + i;
      if (col == 0) {
        df[i] = x;
      } else {
        const float w = (x - a.gbf4[4 * col]) * a.gbf4[4 * col + 1];
        df[i] = ex2_fast(-(w * w)) * a.gbf4[4 * col + 2];
      }
    }
    out[p].x = pack_h2(df[0], df[1]); out[p].y = pack_h2(df[2], df[3]);
    out[p].z = pack_h2(df[4], df[5]); out[p].w = pack_h2(df[6], df[7]);
  }
}

// ---- stage A: input_lin MMA, + hoisted per-atom parts, LayerNorm + modulation -> X ---------------------------------------
template <int CH, bool UNI>
__device__ __forceinline__ void e2_stage_a(const EquiArgs& a, const E2Ctx& c) {
  uint8_t* X = c.smem + E2_X;
  uint8_t* U = c.smem + E2_U;
  float2* LNS = reinterpret_cast<float2*>(c.smem + E2_LNS);
  const int row = c.row, lane = c.lane, lt = c.lt, rq = c.rq;
  const float4* pos = reinterpret_cast<const float4*>(a.pos_in);
  const uint4* AB = static_cast<const uint4*>(a.AB);
  if (c.nsteps == 0) return;

  // prologue: the first tile's distance features and e rows; row metadata runs two steps ahead of its use so that no
  // dependent global load sits in front of an MMA issue
  RowInfo r = load_row(a.p, e2_tile(c, 0), row);
  RowInfo rn = c.nsteps > 1 ? load_row(a.p, e2_tile(c, 1), row) : r;
  {
    const float* tr = a.tab + (size_t)(UNI ? 0 : r.mol) * a.ld_tab + a.tab_off;
    const float gsc = UNI ? c_e2mod[512] : tr[tab_gbf(D_)], gsh = UNI ? c_e2mod[513] : tr[tab_gbf(D_) + 1];
    uint4 dfh[4];
    e2_gbf<CH>(a, sq_dist(pos[r.g], pos[r.j]), gsc, gsh, dfh);
#pragma unroll
    for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(U + img_piece(row, 1, 4 * CH + p, CHUNK_BYTES_A)) = dfh[p];
    gather_e16_warp<16>(U, a.e16, 32 * rq, 16 * CH, r.valid, r.pr, lane);
  }

  E2_MARK_INIT;
  for (int s = 0; s < c.nsteps; ++s) {
    const uint32_t ph = s & 1;
    const float* tr = a.tab + (size_t)(UNI ? 0 : r.mol) * a.ld_tab + a.tab_off;
    // loads whose addresses are known: the next tile's positions, the first chunk of this tile's per-atom parts, the
    // metadata of the tile after next
    const bool more = s + 1 < c.nsteps;
    const float4 pgn = pos[rn.g], pjn = pos[rn.j];
    const uint4* pa = AB + (size_t)(16 * CH) * a.ldab + r.g;            // A[g]: pieces [16 CH, 16 CH + 16) of the A half
    const uint4* pb = AB + (size_t)(32 + 16 * CH) * a.ldab + r.j;       // B[j]: the same pieces of the B half
    uint4 ua0 = __ldg(pa), ua1 = __ldg(pa + a.ldab), ub0 = __ldg(pb), ub1 = __ldg(pb + a.ldab);
    RowInfo rnn = rn;
    if (s + 2 < c.nsteps) rnn = load_row(a.p, e2_tile(c, s + 2), row);
    // ---- U of this step is complete in this CTA: tell the leader, which issues the input_lin MMA for the pair
    cp_async_wait_all();
    fence_async_smem();
    ga_sync();
    E2_MARK(0);
    if (lt == 0) {
      if (s == 0) mbar_wait(&c.bars[B_W], 0);            // this CTA's weight halves have landed
      mbar_arrive_cluster(&c.bars[B_RDY1], 0);
      if (c.rank == 0) {
        if (c_e2flags & 1) mbar_wait_cluster<true>(&c.bars[B_RDY1], ph); else mbar_wait_cluster<false>(&c.bars[B_RDY1], ph);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16_2sm(256);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_2sm(c.tm_x, umma_desc_sw128(smem_u32(U) + (k >> 2) * CHUNK_BYTES_A + (k & 3) * 32),
                       umma_desc_sw128(smem_u32(c.smem + E2_WIN) + (k >> 2) * CHUNK_BYTES_A + (k & 3) * 32), idesc, k ? 1u : 0u);
        umma_commit_2sm(&c.bars[B_M1], 3);
      }
    }
    E2_MARK(1);
    // ---- under the MMA: the next tile's distance features
    uint4 dfh[4];
    {
      const float* trn = a.tab + (size_t)(UNI ? 0 : rn.mol) * a.ld_tab + a.tab_off;
      const float gsc = UNI ? c_e2mod[512] : trn[tab_gbf(D_)], gsh = UNI ? c_e2mod[513] : trn[tab_gbf(D_) + 1];
      e2_gbf<CH>(a, sq_dist(pgn, pjn), gsc, gsh, dfh);
    }
    E2_MARK(2);
    mbar_wait(&c.bars[B_M1], ph);
    tc_fence_after();
    E2_MARK(3);
    if (more) {                                          // U is consumed: the next tile's operand rows go in
#pragma unroll
      for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(U + img_piece(row, 1, 4 * CH + p, CHUNK_BYTES_A)) = dfh[p];
      gather_e16_warp<16>(U, a.e16, 32 * rq, 16 * CH, rn.valid, rn.pr, lane);
    }
    // ---- pass 1: x = acc + (A[g] + B[j]) over this thread's 128 hidden units, kept in tensor memory; row statistics
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 na0, na1, nb0, nb1;
      if (q < 7) {
        na0 = __ldg(pa + (size_t)(2 * q + 2) * a.ldab); na1 = __ldg(pa + (size_t)(2 * q + 3) * a.ldab);
        nb0 = __ldg(pb + (size_t)(2 * q + 2) * a.ldab); nb1 = __ldg(pb + (size_t)(2 * q + 3) * a.ldab);
      }
      float x[16];
      tmem_ld16(tmem_addr(c.tm_x, 128 * CH + 16 * q), x);
      float fa[8], fb[8];
      unpack8(ua0, fa); unpack8(ub0, fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float v = x[e] + (fa[e] + fb[e]); x[e] = v; s1 += v; s2 = fmaf(v, v, s2); }
      unpack8(ua1, fa); unpack8(ub1, fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float v = x[8 + e] + (fa[e] + fb[e]); x[8 + e] = v; s1 += v; s2 = fmaf(v, v, s2); }
      tmem_st16(tmem_addr(c.tm_x, 128 * CH + 16 * q), x);
      if (q < 7) { ua0 = na0; ua1 = na1; ub0 = nb0; ub1 = nb1; }
    }
    tmem_wait_st();
    LNS[row * 2 + CH] = make_float2(s1, s2);
    ga_sync();
    const float2 o = LNS[row * 2 + (CH ^ 1)];
    const float mean = (s1 + o.x) * (1.0f / 256.0f);
    const float rstd = rsqrtf(fmaxf((s2 + o.y) * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-6f);
    const float nmr = -mean * rstd;
    E2_MARK(4);
    // ---- pass 2: LayerNorm + modulation -> X (fp16 operand image); X is free once coord_mlp.0 of the previous step is done
    if (s > 0) mbar_wait(&c.bars[B_C0], ph ^ 1u);
    E2_MARK(5);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float x[16];
      tmem_ld16(tmem_addr(c.tm_x, 128 * CH + 16 * q), x);
      if (UNI) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = 128 * CH + 16 * q + i;
          x[i] = fmaf(fmaf(x[i], rstd, nmr), c_e2mod[256 + col], c_e2mod[col]);
        }
      } else {
        const float* shift = tr + tab_equi(D_) + 128 * CH + 16 * q;
        const float* scale = shift + D_;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + i));
          const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + i));
          x[i] = fmaf(fmaf(x[i], rstd, nmr), sc.x, sh.x);
          x[i + 1] = fmaf(fmaf(x[i + 1], rstd, nmr), sc.y, sh.y);
          x[i + 2] = fmaf(fmaf(x[i + 2], rstd, nmr), sc.z, sh.z);
          x[i + 3] = fmaf(fmaf(x[i + 3], rstd, nmr), sc.w, sh.w);
        }
      }
      st_rowh<16>(X, row, 2 * CH + (q >> 2), 2 * (q & 3), x);
    }
    fence_async_smem();
    ga_sync();                                           // (also: every read of x is done before the next step's MMA1)
    if (lt == 0) mbar_arrive_cluster(&c.bars[B_RDY2], 0);
    E2_MARK(6);
    r = rn;
    rn = rnn;
  }
}

// ---- stage B: coord_mlp.0 MMA, SiLU in place, coord_mlp.2 MMA from tensor memory, per-atom coordinate sums ---------------
template <int CH>
__device__ __forceinline__ void e2_stage_b(const EquiArgs& a, const E2Ctx& c) {
  uint8_t* X = c.smem + E2_X;
  const int row = c.row, lane = c.lane, lt = c.lt;
  const int wb = lt >> 5;
  const float4* pos = reinterpret_cast<const float4*>(a.pos_in);
  float4* pos_out = reinterpret_cast<float4*>(a.pos_out);
  int ng_prev = 0;
  bool live_prev = false;
  // per-atom sums of a step run in the next step's coord_mlp.0 window: their inputs are double-buffered by step parity
  auto C3_of = [&](int s_) { return reinterpret_cast<float4*>(c.smem + E2_C3) + 128 * (s_ & 1); };
  auto gtm_of = [&](int s_) { return reinterpret_cast<uint32_t*>(c.smem + E2_GT) + 128 * (s_ & 1); };
  auto GP_of = [&](int s_) { return reinterpret_cast<float4*>(c.smem + E2_GP) + 64 * (s_ & 1); };
  auto group_sums = [&](int s_, int ng_, bool live_) {   // one warp per group, lanes over its rows, shuffle tree
    const float4* C3 = C3_of(s_);
    const uint32_t* gt_meta = gtm_of(s_);
    const int* gt_node = reinterpret_cast<const int*>(gt_meta + 64);
    const float4* GP = GP_of(s_);
    for (int gi = wb; gi < ng_; gi += E2_GROUP / 32) {
      const int gs = gt_meta[gi] & 255u, gl = (gt_meta[gi] >> 8) & 255u;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int k = lane; k < gl; k += 32) { const float4 v = C3[gs + k]; sx += v.x; sy += v.y; sz += v.z; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
      }
      if (lane == 0 && live_) {
        const float4 p0 = GP[gi];
        pos_out[gt_node[gi]] = make_float4(p0.x + sx, p0.y + sy, p0.z + sz, 0.f);
      }
    }
  };

  if (c.nsteps == 0) return;
  // row metadata, adjacency bits and positions of a step are fetched during the previous one
  RowInfo rn = load_row(a.p, e2_tile(c, 0), row);
  int ngn = a.p.tile_ngroups[e2_tile(c, 0)];
  uint8_t exn = a.extra[rn.pr];
  float4 pgn = pos[rn.g], pjn = pos[rn.j];
  E2_MARK_INIT;
  for (int s = 0; s < c.nsteps; ++s) {
    const uint32_t ph = s & 1;
    const bool live = c.pt0 + 2 * s + c.rank < c.pt1;    // the odd tail: the peer repeats the last tile and writes nothing
    const RowInfo r = rn;
    const int ng = ngn;
    const uint8_t ex = exn;
    const float4 pg = pgn, pj = pjn;
    if (CH == 0 && r.valid && row == r.gs) {
      uint32_t* gt_meta = gtm_of(s);
      gt_meta[r.gi] = (uint32_t)r.gs | ((uint32_t)r.gl << 8);
      reinterpret_cast<int*>(gt_meta + 64)[r.gi] = r.g;
      GP_of(s)[r.gi] = pg;
    }
    if (lt == 0) {
      if (s == 0) mbar_arrive_cluster(&c.bars[B_RDY2], 0);   // the accumulator columns start free
      if (c.rank == 0) {
        if (c_e2flags & 1) mbar_wait_cluster<true>(&c.bars[B_RDY2], ph); else mbar_wait_cluster<false>(&c.bars[B_RDY2], ph);              // both X operands written, both accumulators drained
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16_2sm(256);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          umma_f16_2sm(c.tm_c, umma_desc_sw128(smem_u32(X) + (k >> 2) * CHUNK_BYTES_A + (k & 3) * 32),
                       umma_desc_sw128(smem_u32(c.smem + E2_WC0) + (k >> 2) * CHUNK_BYTES_A + (k & 3) * 32), idesc, k ? 1u : 0u);
        umma_commit_2sm(&c.bars[B_C0], 3);
      }
    }
    E2_MARK(8);
    if (s + 1 < c.nsteps) {                              // next step's metadata: in flight under the MMA
      rn = load_row(a.p, e2_tile(c, s + 1), row);
      ngn = a.p.tile_ngroups[e2_tile(c, s + 1)];
    }
    if (s > 0) group_sums(s - 1, ng_prev, live_prev);    // the previous step's per-atom sums, under the MMA
    mbar_wait(&c.bars[B_C0], ph);
    tc_fence_after();
    E2_MARK(9);
    // SiLU (h + h tanh h, h = x / 2: the image and bias are pre-scaled) of hidden units [128 CH, 128 CH + 128), written back
    // in place as packed fp16 pairs: unit k of this half sits in column 128 CH + k / 2, the A operand of coord_mlp.2
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float x[16];
      tmem_ld16(tmem_addr(c.tm_c, 128 * CH + 16 * q), x);
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float h0 = x[i] + a.b0h[128 * CH + 16 * q + i], h1 = x[i + 1] + a.b0h[128 * CH + 16 * q + i + 1];
        pk[i >> 1] = pack_h2(fmaf(h0, tanh_fast(h0), h0), fmaf(h1, tanh_fast(h1), h1));
      }
      tmem_st8(tmem_addr(c.tm_c, 128 * CH + 8 * q), pk);
    }
    tmem_wait_st();
    if (s + 1 < c.nsteps) { exn = a.extra[rn.pr]; pgn = pos[rn.g]; pjn = pos[rn.j]; }      // in flight under coord_mlp.2 and the tail
    gb_sync();
    E2_MARK(10);
    if (lt == 0) {
      mbar_arrive_cluster(&c.bars[B_RDY3], 0);
      if (c.rank == 0) {
        if (c_e2flags & 1) mbar_wait_cluster<true>(&c.bars[B_RDY3], ph); else mbar_wait_cluster<false>(&c.bars[B_RDY3], ph);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16_2sm(32);   // A from tensor memory needs N >= 32 in pair mode: rows 3.. are zero
#pragma unroll
        for (int k = 0; k < 16; ++k)                     // K = 16 per step = 8 columns of packed pairs; D = columns 64..95
          umma_f16_ts_2sm(c.tm_c + 64, c.tm_c + 128 * (k >> 3) + 8 * (k & 7),
                          umma_desc_sw128(smem_u32(c.smem + E2_W2) + (k >> 2) * 2048 + (k & 3) * 32), idesc, k ? 1u : 0u);
        umma_commit_2sm(&c.bars[B_C2], 3);
      }
    }
    E2_MARK(11);
    if (CH == 0) {       // tanh, adjacency-weighted mean, coordinate contribution of this edge
      mbar_wait(&c.bars[B_C2], ph);
      tc_fence_after();
      float dd[16];
      tmem_ld16(tmem_addr(c.tm_c, 64), dd);
      const float w = (tanh_fast(dd[0]) + ((ex & 1) ? tanh_fast(dd[1]) : 0.f) + ((ex & 2) ? tanh_fast(dd[2]) : 0.f)) * (1.0f / 3.0f);
      const float dx = pg.x - pj.x, dy = pg.y - pj.y, dz = pg.z - pj.z;
      const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
      const float f = r.valid ? a.coord_scale * w / fmaxf(nrm, 1e-8f) : 0.f;
      C3_of(s)[row] = make_float4(dx * f, dy * f, dz * f, 0.f);
    }
    E2_MARK(12);
    gb_sync();                                           // C3 and the group table are complete; the accumulator is drained
    E2_MARK(13);
    if (lt == 0 && s + 1 < c.nsteps) mbar_arrive_cluster(&c.bars[B_RDY2], 0);
    ng_prev = ng;
    live_prev = live;
    E2_MARK(14);
  }
  gb_sync();
  group_sums(c.nsteps - 1, ng_prev, live_prev);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(E2_THREADS, 1) k_equi2(const __grid_constant__ EquiArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + E2_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + E2_BAR + 96);
  const int t = threadIdx.x, warp = t >> 5;
  const int rank = (int)cluster_ctarank();
  const int pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  int per = (a.p.n_tiles + pairs - 1) / pairs;
  per = (per + 1) & ~1;                                  // whole steps of two tiles
  const int pt0 = pair * per;
  const int pt1 = min(pt0 + per, a.p.n_tiles);

  if (t == 0) {
    mbar_init(&bars[B_W], 1);
    mbar_init(&bars[B_RDY1], 2);                         // stage A of both CTAs
    mbar_init(&bars[B_M1], 1);
    mbar_init(&bars[B_RDY2], 4);                         // stage A (X written) and stage B (accumulator drained) of both CTAs
    mbar_init(&bars[B_C0], 1);
    mbar_init(&bars[B_RDY3], 2);                         // stage B of both CTAs
    mbar_init(&bars[B_C2], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[B_W], 65536 + 32768 + 8192);
    const uint8_t* wc0 = static_cast<const uint8_t*>(a.wc0_img);
    const uint8_t* win = static_cast<const uint8_t*>(a.win_img);
    const uint8_t* w2 = static_cast<const uint8_t*>(a.w2_img32);
    for (int kc = 0; kc < 4; ++kc) bulk_g2s(smem + E2_WC0 + kc * 16384, wc0 + kc * 32768 + rank * 16384, 16384, &bars[B_W]);
    for (int kc = 0; kc < 2; ++kc) bulk_g2s(smem + E2_WIN + kc * 16384, win + kc * 32768 + rank * 16384, 16384, &bars[B_W]);
    for (int kc = 0; kc < 4; ++kc) bulk_g2s(smem + E2_W2 + kc * 2048, w2 + kc * 4096 + rank * 2048, 2048, &bars[B_W]);
  }
  if (warp == 0) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                    // barriers initialised and tensor memory allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;

  E2Ctx c;
  c.smem = smem; c.bars = bars; c.tm_x = tmem; c.tm_c = tmem + 256;
  c.rank = rank; c.lt = t & 255; c.rq = warp & 3; c.lane = t & 31; c.row = (warp & 3) * 32 + (t & 31);
  c.pt0 = pt0; c.pt1 = pt1; c.nsteps = pt1 > pt0 ? (pt1 - pt0 + 1) >> 1 : 0;
  const int ch = (warp >> 2) & 1;
  if (t < E2_GROUP) {
    if (ch == 0) { if (uni) e2_stage_a<0, true>(a, c); else e2_stage_a<0, false>(a, c); }
    else         { if (uni) e2_stage_a<1, true>(a, c); else e2_stage_a<1, false>(a, c); }
  } else {
    if (ch == 0) e2_stage_b<0>(a, c); else e2_stage_b<1>(a, c);
  }
  if (t == 0) mbar_wait(&bars[B_W], 0);                  // never leave with bulk copies in flight
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                    // the peer's shared memory is an MMA operand until the very end
  if (warp == 0) tmem_dealloc_2sm<512>(tmem);
}

}  // namespace

#ifdef JODO_PHASE_TIMING
extern "C" int jodo_debug_equi2_phases(long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (out16) cudaMemcpyFromSymbol(out16, g_equi2_phase, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_equi2_phase, z, sizeof(z)); }
  return 0;
}
#endif

cudaError_t launch_equi2(const EquiArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_equi2, E2_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  if ((e0 = const_tables_acquire(st)) != cudaSuccess) return e0;
  if (a.nonuni) {     // row 0 of the table feeds the uniform fast path (harmless when the batch is not uniform)
    e0 = cudaMemcpyToSymbolAsync(c_e2mod, a.tab + a.tab_off + tab_equi(D_), sizeof(float) * 528, 0, cudaMemcpyDeviceToDevice, st);
    if (e0 != cudaSuccess) return e0;
  }
  static int flags = -1;
  if (flags < 0) {
    const char* f = std::getenv("JODO_E2_FLAGS");
    flags = f ? std::atoi(f) : 0;
    cudaMemcpyToSymbol(c_e2flags, &flags, sizeof(int));
  }
  const int max_pairs = num_sms / 2;
  const int want = (a.p.n_tiles + 1) / 2;
  const int pairs = want < max_pairs ? want : max_pairs;
  k_equi2<<<2 * pairs, E2_THREADS, E2_SMEM, st>>>(a);
  if ((e0 = cudaGetLastError()) != cudaSuccess) return e0;
  return const_tables_release(st);
}

}  // namespace jodo
