// Edge-gated attention of one DGT block on edge tiles (persistent, weights resident in shared memory).
//
// reference: EquivariantMixBlock.forward, models/mol_gnn.py:284-287,293-297 (distance GBF, block edge_emb,
// norm1_edge + modulate) and TransMixLayer.message, models/layers.py:159-186 (lin_edge0/tanh trilinear
// logits with 1/sqrt(out_channels), two adjacency heads first, PyG softmax over the sources of a target,
// lin_edge1/tanh gated values, sum onto the target).
//
// A tile holds complete groups: group atom g = attention TARGET c, partner j = SOURCE r.  The block-input edge features
// and the adjacency bits are stored once per unordered pair (they are symmetric); every directed row fetches its pair's
// row through the plan's row_pair map.
// Per tile:  A0 = [GBF(d) | e]  --MMA1--> e1 --LN/modulate--> en  --MMA2--> g0 --(logits)--  --MMA3--> g1 --(messages)
//   (fp16 operand images, fp32 accumulation; e1, g0 and g1 reuse the same 256 TMEM columns one after the other, the
//   g1 MMA runs under the softmax)
//   logits (k[j] . q[g] . tanh(g0)), softmax per group through shared memory,
//   msg = v[j] * tanh(g1) * alpha, summed per group, written to hnode[g].
//
// 512 threads = two independent groups of 8 warps, each walking its own tiles (group-local named barriers, private
// operand / staging buffers, 256 TMEM columns, its own mbarriers) and sharing the three resident weight images, so
// that one group's tensor-core round trips, barriers and gather latencies are covered by the other group's math.
// Inside a group warp w works on tile rows 32*(w&3)..+31 (its TMEM lane quarter) and on column half (w>>2) of every
// row vector (32 of 64 edge features, 128 of 256 q/k/g0 columns = 7 of 14 heads, 128 of 256 value columns = 8 of 16
// heads).  q | k | v rows are piece-major fp16 (edge_common.cuh) so that the partner gathers read whole lines;
// per-column constants are kernel-parameter / constant-memory operands (uniform-conditioning fast path, see equi.cu).
#include "edge_common.cuh"

namespace jodo {

__constant__ float c_atmod[130];        // row 0 of the table: edge shift_msa[64], scale_msa[64], then GBF scale, shift

#ifdef JODO_PHASE_TIMING
__device__ long long g_attn_phase[16];
#define AT_MARK(i) do { if (c.lt == 0 && c.grp == 0 && HALF == 0 && blockIdx.x == 0) { long long c_ = clock64(); g_attn_phase[i] += c_ - ph_last; ph_last = c_; } } while (0)
#else
#define AT_MARK(i) do { } while (0)
#endif

namespace {

constexpr int AT_THREADS = 512;
constexpr int AT_GROUP = 256;
constexpr int AT_WE = 0;                          // 16 KB: block edge_emb image (N=64, K=128)
constexpr int AT_W0 = AT_WE + 16384;              // 32 KB: lin_edge0 image (N=256 split-head, K=64)
constexpr int AT_W1 = AT_W0 + 32768;              // 32 KB: lin_edge1 image
constexpr int AT_GRP = AT_W1 + 32768;             // per group:
constexpr int G_A0 = 0;                           //   32 KB fp16: chunk 0 = GBF(d), then en, then message image of half 0;
                                                  //                chunk 1 = e (bulk-copied one tile ahead)
constexpr int G_MIB = 32768;                      //   17 KB: logits / exp values [128][17] fp32 + 1/(sum + 1e-16) [128][16] during
constexpr int G_LG = G_MIB;                       //          the softmax, then the message image of half 1 (16 KB)
constexpr int G_GI = G_LG + 128 * 17 * 4;
constexpr int G_IND = G_MIB + 17408;              //   16 KB: row -> group indicator, fp16 K-major image [64 slots][128 rows]
constexpr int G_LN = G_IND + 16384;               //   LayerNorm partial sums [128][2] float2
constexpr int G_GT = G_LN + 128 * 2 * 8;          //   group table: start | len << 8 [128], atom [128]
constexpr int G_BYTES = ((G_GT + 1024 + 1023) / 1024) * 1024;
constexpr int AT_MISC = AT_GRP + 2 * G_BYTES;     // barriers, tmem slot
constexpr int AT_SMEM = AT_MISC + 128;
static_assert(AT_SMEM <= 232448, "shared memory budget");

constexpr int SC = 18;        // sub_channels = 256 // 14   (models/layers.py:112)
constexpr int HQ = 126;       // q/k/g0 columns of one half = 7 heads x 18

// Shared-memory descriptor of an MN-major SWIZZLE_128B operand (cute/arch/mma_sm100_desc.hpp, make_umma_desc<Major::MN>):
// 64 consecutive MN elements (128 B) per K row, 8-row swizzle atoms of 1024 B (SBO), MN blocks of 64 `lbo_bytes` apart.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor with an MN-major A operand (bit 15), M = 128
__device__ __host__ constexpr uint32_t umma_idesc_f16_amn(int n) { return umma_idesc_f16(n) | (1u << 15); }

__device__ __forceinline__ void at_group_sync(int grp) {
  tc_fence_before();
  named_bar_sync(1 + grp, AT_GROUP);
  tc_fence_after();
}

struct AtCtx {
  uint8_t* gs;                 // group-private shared memory
  const uint8_t* WE; const uint8_t* W0; const uint8_t* W1;
  uint64_t* bar_w; uint64_t* bar_e; uint64_t* bar_m;
  uint32_t tm;
  int grp, lt, row, rq, lane, tile0, tile1;
};

template <int HALF, bool UNI>
__device__ __forceinline__ void at_group_loop(const AttnArgs& a, const AtCtx& c) {
  uint8_t* A0 = c.gs + G_A0;
  uint8_t* MI = HALF == 0 ? A0 : c.gs + G_MIB;       // this half's 64-column message image (MN-major A operand chunk)
  uint8_t* IND = c.gs + G_IND;
  float* LG = reinterpret_cast<float*>(c.gs + G_LG);
  float* GI = reinterpret_cast<float*>(c.gs + G_GI);
  float2* LNS = reinterpret_cast<float2*>(c.gs + G_LN);
  uint32_t* gt_meta = reinterpret_cast<uint32_t*>(c.gs + G_GT);
  int* gt_node = reinterpret_cast<int*>(gt_meta + 128);
  const int row = c.row, lane = c.lane, lt = c.lt, rq = c.rq;
  const uint32_t tm = c.tm;
  uint32_t ind_prev = 0xFFFFFFFFu;                  // byte offset of this row's 1.0 in the indicator image
  const float4* pos = reinterpret_cast<const float4*>(a.pos);
  uint32_t par_m = 0;


  // row metadata of a tile is fetched one tile ahead
  const int tfirst = min(c.tile0, a.p.n_tiles - 1);
  RowInfo rn = load_row(a.p, tfirst, row);
  int ngn = a.p.tile_ngroups[tfirst];
  uint8_t exn = a.extra[rn.pr];
  // the e chunk of a tile = its rows' pair rows, gathered from the pair-row store one tile ahead (edge_common.cuh):
  // the two warps that own a row quarter copy 16 of its rows each
  if (c.tile0 < c.tile1) gather_e16_warp<16>(A0 + CHUNK_BYTES_A, a.e16, 32 * rq, 16 * HALF, rn.valid, rn.pr, lane);
#ifdef JODO_PHASE_TIMING
  long long ph_last = clock64();
#endif
  for (int tile = c.tile0; tile < c.tile1; tile += 2) {
    const RowInfo r = rn;
    const int ng = ngn;
    const uint8_t ex = exn;
    if (HALF == 0 && r.valid && row == r.gs) { gt_meta[r.gi] = (uint32_t)r.gs | ((uint32_t)r.gl << 8); gt_node[r.gi] = r.g; }
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off;

    // ---- distance features -> A0 chunk 0, columns [32*HALF, 32*HALF+32)
    {
      const float gsc = UNI ? c_atmod[128] : tr[tab_gbf(D_)], gsh = UNI ? c_atmod[129] : tr[tab_gbf(D_) + 1];
      const float d = sq_dist(pos[r.j], pos[r.g]);
      const float x = fmaf(d, gsc, gsh);                  // the table stores 1 + scale
      float df[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = 32 * HALF + i;
        if (col == 0) {
          df[i] = x;
        } else {
          const float w = (x - a.gbf4[4 * col]) * a.gbf4[4 * col + 1];
          df[i] = ex2_fast(-(w * w)) * a.gbf4[4 * col + 2];
        }
      }
      st_rowh<32>(A0, row, 0, 4 * HALF, df);       // padding rows: finite garbage, masked by alpha = 0 below
    }
    AT_MARK(0);
    cp_async_wait_all();                                 // this thread's share of the gathered e chunk has landed
    fence_async_smem();
    at_group_sync(c.grp);
    AT_MARK(1);
    if (lt == 0) {
      if (tile == c.tile0) mbar_wait(c.bar_w, 0);
      tc_fence_after();
      mma_tile_h(tm, smem_u32(A0), smem_u32(c.WE), 64, 2, false);                   // e1 = edge_emb([dist | e])
      umma_commit(c.bar_m);
    }
    mbar_wait(c.bar_m, par_m);
    AT_MARK(2);
    par_m ^= 1;
    tc_fence_after();
    // row -> group indicator for the aggregation MMA: B[slot][row] = 1 (K-major image, two chunks of 64 rows)
    if (HALF == 0) {
      if (ind_prev != 0xFFFFFFFFu) *reinterpret_cast<uint16_t*>(IND + ind_prev) = 0;
      if (r.valid) {
        ind_prev = (uint32_t)(row >> 6) * 8192u + (uint32_t)r.gi * 128u + (((uint32_t)((row & 63) >> 3) ^ ((uint32_t)r.gi & 7u)) << 4) + ((uint32_t)(row & 7) << 1);
        *reinterpret_cast<uint16_t*>(IND + ind_prev) = 0x3C00;      // fp16 1.0
      } else {
        ind_prev = 0xFFFFFFFFu;
      }
    }

    // ---- en = LN(e1) * (1 + scale_msa) + shift_msa  -> A0 chunk 0
    {
      float x[32];
      tmem_ld32(tmem_addr(tm, 32 * HALF), x);
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { x[i] += a.b_emb[32 * HALF + i]; s += x[i]; q = fmaf(x[i], x[i], q); }
      LNS[row * 2 + HALF] = make_float2(s, q);
      at_group_sync(c.grp);                              // (also: every e1 read is done before MMA2 overwrites the columns)
      const float2 o = LNS[row * 2 + (HALF ^ 1)];
      const float mean = (s + o.x) * (1.0f / 64.0f);
      const float var = fmaxf((q + o.y) * (1.0f / 64.0f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-6f);
      const float nmr = -mean * rstd;
      float scv[32], shv[32];                              // general path: this molecule's row, 16-byte loads
      if (!UNI) { ldg_row<32>(tr + tab_edge(D_) + ED_ + 32 * HALF, scv); ldg_row<32>(tr + tab_edge(D_) + 32 * HALF, shv); }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = 32 * HALF + i;
        const float n = fmaf(x[i], rstd, nmr);
        const float sc = UNI ? c_atmod[64 + col] : scv[i];
        const float sh = UNI ? c_atmod[col] : shv[i];
        x[i] = fmaf(n, sc, sh);
      }
      st_rowh<32>(A0, row, 0, 4 * HALF, x);
    }
    fence_async_smem();
    AT_MARK(3);
    at_group_sync(c.grp);
    AT_MARK(4);
    if (lt == 0) {
      mma_tile_h(tm, smem_u32(A0), smem_u32(c.W0), 256, 1, false);                  // g0 pre-activation
      umma_commit(c.bar_m);
    }
    // ---- logits of this half's 7 heads: a[s] = sum_ch q[g,s,ch] k[j,s,ch] tanh(g0[s,ch]) / sqrt(16)
    // q / k pieces of a 16-column chunk are loaded one chunk ahead of their use
    H16 qc, kc;
    {
      const uint4* qb = static_cast<const uint4*>(a.qkv) + (size_t)(16 * HALF) * a.ldq + r.g;
      const uint4* kb = static_cast<const uint4*>(a.qkv) + (size_t)(32 + 16 * HALF) * a.ldq + r.j;
      qc.u[0] = __ldg(qb); qc.u[1] = __ldg(qb + a.ldq);
      kc.u[0] = __ldg(kb); kc.u[1] = __ldg(kb + a.ldq);
      mbar_wait(c.bar_m, par_m);
      AT_MARK(5);
      par_m ^= 1;
      tc_fence_after();
      float lg[7];
#pragma unroll
      for (int s = 0; s < 7; ++s) lg[s] = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        H16 qn, kn;
        if (ch < 7) {
          qn.u[0] = __ldg(qb + (size_t)(2 * ch + 2) * a.ldq); qn.u[1] = __ldg(qb + (size_t)(2 * ch + 3) * a.ldq);
          kn.u[0] = __ldg(kb + (size_t)(2 * ch + 2) * a.ldq); kn.u[1] = __ldg(kb + (size_t)(2 * ch + 3) * a.ldq);
        }
        float acc[16];
        tmem_ld16(tmem_addr(tm, 128 * HALF + 16 * ch), acc);
        // q k straight from the packed fp16 pairs (FHFMA: exact fp32 product of two halves), no unpacking
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t* q2 = reinterpret_cast<const uint32_t*>(&qc.u[i]);
          const uint32_t* k2 = reinterpret_cast<const uint32_t*>(&kc.u[i]);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int col = 16 * ch + 8 * i + 2 * w;          // SC and HQ are even: a pair never straddles two heads
            if (col < HQ) {
              lg[col / SC] = fmaf(fhfma_lo(q2[w], k2[w], 0.f), tanh_fast(acc[8 * i + 2 * w]), lg[col / SC]);
              lg[col / SC] = fmaf(fhfma_hi(q2[w], k2[w], 0.f), tanh_fast(acc[8 * i + 2 * w + 1]), lg[col / SC]);
            }
          }
        }
        if (ch < 7) { qc = qn; kc = kn; }
      }
      // head order of the reference: the two adjacency heads first (1 where adjacent, -1e10 otherwise,
      // models/layers.py:170-174), then the 14 computed heads
      LG[row * 17 + HALF] = ((ex >> HALF) & 1) ? 1.0f : -1e10f;
#pragma unroll
      for (int s = 0; s < 7; ++s) LG[row * 17 + 2 + 7 * HALF + s] = lg[s] * 0.25f;
    }
    // row metadata of this group's next tile: loaded here (after the register-hungry logit loop, so that the prefetch is
    // not spilled) and used below for the gather of the next e chunk, which MMA1 of this tile has long consumed
    {
      const int nt_ = tile + 2 < c.tile1 ? tile + 2 : tile;
      rn = load_row(a.p, nt_, row);
      ngn = a.p.tile_ngroups[nt_];
    }
    AT_MARK(6);
    at_group_sync(c.grp);                                  // logits visible; every g0 read done
    AT_MARK(7);
    if (lt == 0) {
      mma_tile_h(tm, smem_u32(A0), smem_u32(c.W1), 256, 1, false);                  // g1 pre-activation, under the softmax
      umma_commit(c.bar_m);
    }
    // first value chunk in flight under the softmax
    const uint4* vb = static_cast<const uint4*>(a.qkv) + (size_t)(64 + 16 * HALF) * a.ldq + r.j;
    H16 vc;
    vc.u[0] = __ldg(vb); vc.u[1] = __ldg(vb + a.ldq);
    exn = a.extra[rn.pr];
    if (tile + 2 < c.tile1)
      gather_e16_warp<16>(A0 + CHUNK_BYTES_A, a.e16, 32 * rq, 16 * HALF, rn.valid, rn.pr, lane);
    // (group, head): max, exp in place, 1 / (sum + 1e-16)      (PyG softmax, models/layers.py:178).
    {
      // L lanes share a (group, head): lane `sub` takes the rows sub, sub + L, ... of the group.  Every group sums its
      // exponentials in FOUR interleaved partial sums combined as (s0 + s1) + (s2 + s3), whatever L is (zeros are added
      // where a lane holds no part), so the arithmetic depends on the group alone and not on what else shares the tile.
      const int shl = ng <= 4 ? 2 : (ng <= 8 ? 1 : 0);       // ng * 16 << shl is a whole number of warps when shl > 0
      const int L = 1 << shl;
      for (int it = lt; it < ((ng * 16) << shl); it += AT_GROUP) {
        const int sub = it & (L - 1), item = it >> shl, gi = item >> 4, h = item & 15;
        const int gs = gt_meta[gi] & 255u, gl = (gt_meta[gi] >> 8) & 255u;
        float m = -INFINITY;
        for (int rr = gs + sub; rr < gs + gl; rr += L) m = fmaxf(m, LG[rr * 17 + h]);
        if (shl >= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        if (shl == 2) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int rr = gs + sub; rr < gs + gl; rr += L) {
          const float e = ex2_fast((LG[rr * 17 + h] - m) * 1.4426950408889634f);
          LG[rr * 17 + h] = e;
          const int k = (rr - gs) & 3;
          s4[0] += k == 0 ? e : 0.f; s4[1] += k == 1 ? e : 0.f; s4[2] += k == 2 ? e : 0.f; s4[3] += k == 3 ? e : 0.f;
        }
        float s01 = s4[0] + s4[1], s23 = s4[2] + s4[3];
        if (shl >= 1) { s01 += __shfl_xor_sync(0xffffffffu, s01, 1); s23 += __shfl_xor_sync(0xffffffffu, s23, 1); }
        if (shl == 2) { s01 += __shfl_xor_sync(0xffffffffu, s01, 2); s23 += __shfl_xor_sync(0xffffffffu, s23, 2); }
        if (sub == 0) GI[gi * 16 + h] = 1.0f / ((s01 + s23) + 1e-16f);
      }
    }
    AT_MARK(8);
    named_bar_sync(1 + c.grp, AT_GROUP);
    AT_MARK(9);
    float alpha[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) alpha[h] = r.valid ? LG[row * 17 + 8 * HALF + h] * GI[r.gi * 16 + 8 * HALF + h] : 0.f;
    named_bar_sync(1 + c.grp, AT_GROUP);                 // the logits buffer becomes the message image of half 1

    // ---- messages -> fp16 image; group sums on the tensor core:  D^T[col][slot] = sum_row msg[row][col] * ind[row][slot]
    // (A = message image, MN-major: M = 128 value columns of this pass, K = 128 rows; B = indicator, N = 64 slots).
    // Two passes of 64 columns per half; D^T(p) lands in TMEM columns [64p, 64p+64), whose g1 values are consumed by then.
    mbar_wait(c.bar_m, par_m);
    AT_MARK(10);
    par_m ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int ch = 4 * p + cc;                      // value head of this half = 16 columns
        H16 vn;
        if (ch < 7) { vn.u[0] = __ldg(vb + (size_t)(2 * ch + 2) * a.ldq); vn.u[1] = __ldg(vb + (size_t)(2 * ch + 3) * a.ldq); }
        float acc[16];
        tmem_ld16(tmem_addr(tm, 128 * HALF + 16 * ch), acc);
        const float al = p == 0 ? alpha[cc] : alpha[4 + cc];
        // msg = (v alpha) tanh(g1) on packed pairs: v is fp16 already, the message image is fp16 anyway
        const uint32_t al2 = pack_h2(al, al);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t* v2 = reinterpret_cast<const uint32_t*>(&vc.u[i]);
          uint4 o;
          uint32_t* o2 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int w = 0; w < 4; ++w)
            o2[w] = mul_h2(mul_h2(v2[w], al2), pack_h2(tanh_fast(acc[8 * i + 2 * w]), tanh_fast(acc[8 * i + 2 * w + 1])));
          *reinterpret_cast<uint4*>(MI + img_piece(row, 0, 2 * cc + i, CHUNK_BYTES_A)) = o;
        }
        if (ch < 7) vc = vn;
      }
      fence_async_smem();
      AT_MARK(11);
      at_group_sync(c.grp);
      AT_MARK(12);
      if (lt == 0) {
        const uint32_t idesc = umma_idesc_f16_amn(64);
        const uint32_t sa = smem_u32(A0), lbo = (uint32_t)G_MIB, sb = smem_u32(IND);
#pragma unroll
        for (int k = 0; k < 8; ++k)                     // 16 rows per step
          umma_f16(tm + 64 * p, umma_desc_sw128_mn(sa + k * 2048, lbo), umma_desc_sw128(sb + (k >> 2) * 8192 + (k & 3) * 32), idesc, k ? 1u : 0u);
        umma_commit(c.bar_m);
      }
      mbar_wait(c.bar_m, par_m);
      AT_MARK(13);
      par_m ^= 1;
      tc_fence_after();
      // D^T lanes = value columns (lanes 0..63: half 0, 64..127: half 1); this warp drains slots [32 HALF, 32 HALF + 32)
      if (32 * HALF < ng) {
        float dsum[32];
        tmem_ld32(tmem_addr(tm, 64 * p + 32 * HALF), dsum);
        float* dst = a.hnode + 128 * (rq >> 1) + 64 * p + 32 * (rq & 1) + lane;
#pragma unroll
        for (int sl = 0; sl < 32; ++sl)
          if (32 * HALF + sl < ng) dst[(size_t)gt_node[32 * HALF + sl] * D_] = dsum[sl];
      }
    }
    tc_fence_before();                                       // g1 reads are ordered before the next tile's MMA1 by its group_sync
    AT_MARK(14);
  }
}

__global__ void __launch_bounds__(AT_THREADS, 1) k_attn(const __grid_constant__ AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_MISC);   // 0: weights, 1,2: e tile of group 0,1, 3,4: MMA of group 0,1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + AT_MISC + 64);
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t >> 8, lt = t & 255, lw = lt >> 5;
  const int per = (a.p.n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, a.p.n_tiles);

  if (t == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 16384 + 32768 + 32768);
    bulk_g2s(smem + AT_WE, a.w_emb_img, 16384, &bars[0]);
    bulk_g2s(smem + AT_W0, a.w0_img, 32768, &bars[0]);
    bulk_g2s(smem + AT_W1, a.w1_img, 32768, &bars[0]);
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  for (int i = lt; i < 16384 / 16; i += AT_GROUP)       // indicator images start empty
    reinterpret_cast<uint4*>(smem + AT_GRP + grp * G_BYTES + G_IND)[i] = make_uint4(0u, 0u, 0u, 0u);
  sync_tc();
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;

  AtCtx c;
  c.gs = smem + AT_GRP + grp * G_BYTES;
  c.WE = smem + AT_WE; c.W0 = smem + AT_W0; c.W1 = smem + AT_W1;
  c.bar_w = &bars[0]; c.bar_e = &bars[1 + grp]; c.bar_m = &bars[3 + grp];
  c.tm = *tmem_slot + 256u * grp;
  c.grp = grp; c.lt = lt; c.rq = lw & 3; c.lane = t & 31; c.row = (lw & 3) * 32 + (t & 31);
  c.tile0 = tile0 + grp; c.tile1 = tile1;
  if ((lw >> 2) == 0) {
    if (uni) at_group_loop<0, true>(a, c); else at_group_loop<0, false>(a, c);
  } else {
    if (uni) at_group_loop<1, true>(a, c); else at_group_loop<1, false>(a, c);
  }
  if (t == 0) mbar_wait(&bars[0], 0);                    // never leave with bulk copies in flight
  sync_tc();
  if (warp == 0) tmem_dealloc<512>(*tmem_slot);
}

}  // namespace

#ifdef JODO_PHASE_TIMING
extern "C" int jodo_debug_attn_phases(long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (out16) cudaMemcpyFromSymbol(out16, g_attn_phase, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_attn_phase, z, sizeof(z)); }
  return 0;
}
#endif

cudaError_t launch_attn(const AttnArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_attn, AT_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  if ((e0 = const_tables_acquire(st)) != cudaSuccess) return e0;
  if (a.nonuni) {     // row 0 of the table feeds the uniform fast path: edge (shift, scale)_msa and the GBF pair
    cudaError_t e = cudaMemcpyToSymbolAsync(c_atmod, a.tab + a.tab_off + tab_edge(D_), sizeof(float) * 128, 0,
                                            cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(c_atmod, a.tab + a.tab_off + tab_gbf(D_), sizeof(float) * 2, sizeof(float) * 128,
                                cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  }
  const int grid = a.p.n_tiles < 2 * num_sms ? (a.p.n_tiles + 1) / 2 : num_sms;
  k_attn<<<grid, AT_THREADS, AT_SMEM, st>>>(a);
  if ((e0 = cudaGetLastError()) != cudaSuccess) return e0;
  return const_tables_release(st);
}

}  // namespace jodo
