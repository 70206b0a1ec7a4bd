// Edge-gated attention of one DGT block on edge tiles (persistent, weights resident in shared memory).
//
// reference: EquivariantMixBlock.forward, models/mol_gnn.py:284-287,293-297 (distance GBF, block edge_emb,
// norm1_edge + modulate) and TransMixLayer.message, models/layers.py:159-186 (lin_edge0/tanh trilinear
// logits with 1/sqrt(out_channels), two adjacency heads first, PyG softmax over the sources of a target,
// lin_edge1/tanh gated values, sum onto the target).
//
// A tile holds complete groups: group atom g = attention TARGET c, partner j = SOURCE r.
// Per tile:  A0 = [GBF(d) | e]  --MMA1--> e1 --LN/modulate--> en  --MMA2--> g0 (TMEM 0..255)
//                                                               --MMA3--> g1 (TMEM 256..511)
//   (fp16 operand images, fp32 accumulation: same mantissa as tf32 at half the shared memory and twice the rate)
//   logits (k[j] . q[g] . tanh(g0)), softmax per group through shared memory,
//   msg = v[j] * tanh(g1) * alpha, summed per group, written to hnode[g].
//
// 256 threads: warp w works on tile rows 32*(w&3)..+31 (its TMEM lane quarter) and on column half (w>>2) of every
// row vector (32 of 64 edge features, 128 of 256 q/k/g0 columns = 7 of 14 heads, 128 of 256 value columns = 8 of 16
// heads).  Each CTA walks a contiguous range of tiles so that the q/k/v rows of a molecule stay hot in L1/L2.
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int AT_THREADS = 256;
constexpr int AT_A0 = 0;                          // 32 KB fp16: chunk 0 = GBF(d) then en; chunk 1 = e (bulk-copied)
constexpr int AT_WE = 32768;                      // 16 KB: block edge_emb image (N=64, K=128)
constexpr int AT_W0 = AT_WE + 16384;              // 32 KB: lin_edge0 image (N=256, K=64)
constexpr int AT_W1 = AT_W0 + 32768;              // 32 KB: lin_edge1 image
constexpr int AT_S = AT_W1 + 32768;               // 32 KB: message staging, per column half [128 rows][32] fp32, xor-swizzled
constexpr int AT_LG = AT_S + 32768;               // logits / exp values [128][17] fp32
constexpr int AT_GI = AT_LG + 128 * 17 * 4;       // 1 / (sum + 1e-16) per (group, head)  [128][16]
constexpr int AT_LN = AT_GI + 128 * 16 * 4;       // LayerNorm partial sums [128][2] float2
constexpr int AT_MISC = AT_LN + 128 * 2 * 8;      // barriers, tmem slot, GBF constants, bias, group table
constexpr int AT_SMEM = AT_MISC + 128 + 768 + 256 + 512 + 512;
static_assert(AT_SMEM <= 232448, "shared memory budget");

constexpr int SC = 18;        // sub_channels = 256 // 14   (models/layers.py:112)
constexpr int HQ = 126;       // q/k/g0 columns of one half = 7 heads x 18

__global__ void __launch_bounds__(AT_THREADS, 1) k_attn(AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* A0 = smem + AT_A0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_MISC);   // 0: weights, 1: e tile, 2..4: MMA1..3
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + AT_MISC + 64);
  float* gbf = reinterpret_cast<float*>(smem + AT_MISC + 128);
  float* bemb = gbf + 192;
  uint32_t* gt_meta = reinterpret_cast<uint32_t*>(bemb + 64);     // [128] group start | len << 8
  int* gt_node = reinterpret_cast<int*>(gt_meta + 128);           // [128] group atom
  float* LG = reinterpret_cast<float*>(smem + AT_LG);
  float* GI = reinterpret_cast<float*>(smem + AT_GI);
  float2* LNS = reinterpret_cast<float2*>(smem + AT_LN);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int rq = warp & 3, half = warp >> 2;
  const int row = rq * 32 + lane;
  float* S = reinterpret_cast<float*>(smem + AT_S) + half * (128 * 32);

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
    if (tile0 < tile1) {
      mbar_expect_tx(&bars[1], CHUNK_BYTES_A);
      bulk_g2s(A0 + CHUNK_BYTES_A, reinterpret_cast<const uint8_t*>(a.e16) + (size_t)tile0 * CHUNK_BYTES_A, CHUNK_BYTES_A, &bars[1]);
    }
  }
  for (int i = t; i < 192; i += AT_THREADS) gbf[i] = a.gbf[i];
  if (t < 64) bemb[t] = a.b_emb[t];
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  uint32_t par = 0;
  const float4* pos = reinterpret_cast<const float4*>(a.pos);

  // row metadata of a tile is fetched one tile ahead
  RowInfo rn = load_row(a.p, min(tile0, a.p.n_tiles - 1), row);
  int ngn = a.p.tile_ngroups[min(tile0, a.p.n_tiles - 1)];
  uint8_t exn = a.extra[(size_t)min(tile0, a.p.n_tiles - 1) * TILE_ROWS + row];
  for (int tile = tile0; tile < tile1; ++tile) {
    const RowInfo r = rn;
    const int ng = ngn;
    const uint8_t ex = exn;
    if (half == 0 && r.valid && row == r.gs) { gt_meta[r.gi] = (uint32_t)r.gs | ((uint32_t)r.gl << 8); gt_node[r.gi] = r.g; }
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off;

    // ---- distance features -> A0 chunk 0, columns [32*half, 32*half+32)
    {
      float df[32];
      if (r.valid) {
        const float d = sq_dist(pos[r.j], pos[r.g]);
        gbf_eval_half(d, tr[tab_gbf(D_)], tr[tab_gbf(D_) + 1], gbf, half, df);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) df[i] = 0.f;
      }
      st_rowh<32>(A0, row, 0, 4 * half, df);
    }
    {
      const int nt_ = min(tile + 1, tile1 - 1);
      rn = load_row(a.p, nt_, row);
      ngn = a.p.tile_ngroups[nt_];
      exn = a.extra[(size_t)nt_ * TILE_ROWS + row];
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      if (tile == tile0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], par);
      tc_fence_after();
      mma_tile_h(tmem + 256, smem_u32(A0), smem_u32(smem + AT_WE), 64, 2, false);     // e1 = edge_emb([dist | e])
      umma_commit(&bars[2]);
    }
    mbar_wait(&bars[2], par);
    tc_fence_after();
    if (t == 0 && tile + 1 < tile1) {                  // the e chunk is consumed: prefetch the next tile's
      mbar_expect_tx(&bars[1], CHUNK_BYTES_A);
      bulk_g2s(A0 + CHUNK_BYTES_A, reinterpret_cast<const uint8_t*>(a.e16) + (size_t)(tile + 1) * CHUNK_BYTES_A, CHUNK_BYTES_A, &bars[1]);
    }

    // ---- en = LN(e1) * (1 + scale_msa) + shift_msa  -> A0 chunk 0
    {
      float x[32];
      tmem_ld32(tmem_addr(tmem, 256 + 32 * half), x);
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { x[i] += bemb[32 * half + i]; s += x[i]; q = fmaf(x[i], x[i], q); }
      LNS[row * 2 + half] = make_float2(s, q);
      __syncthreads();
      const float2 o = LNS[row * 2 + (half ^ 1)];
      const float mean = (s + o.x) * (1.0f / 64.0f);
      const float var = fmaxf((q + o.y) * (1.0f / 64.0f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-6f);
      const float* shift = tr + tab_edge(D_) + 32 * half;
      const float* scale = shift + ED_;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 sh = *reinterpret_cast<const float4*>(shift + i);
        const float4 sc = *reinterpret_cast<const float4*>(scale + i);
        x[i] = r.valid ? fmaf((x[i] - mean) * rstd, 1.0f + sc.x, sh.x) : 0.f;
        x[i + 1] = r.valid ? fmaf((x[i + 1] - mean) * rstd, 1.0f + sc.y, sh.y) : 0.f;
        x[i + 2] = r.valid ? fmaf((x[i + 2] - mean) * rstd, 1.0f + sc.z, sh.z) : 0.f;
        x[i + 3] = r.valid ? fmaf((x[i + 3] - mean) * rstd, 1.0f + sc.w, sh.w) : 0.f;
      }
      st_rowh<32>(A0, row, 0, 4 * half, x);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mma_tile_h(tmem, smem_u32(A0), smem_u32(smem + AT_W0), 256, 1, false);        // g0 pre-activation
      umma_commit(&bars[3]);
      mma_tile_h(tmem + 256, smem_u32(A0), smem_u32(smem + AT_W1), 256, 1, false);  // g1 pre-activation
      umma_commit(&bars[4]);
    }
    // ---- logits of this half's 7 heads: a[s] = sum_ch q[g,s,ch] k[j,s,ch] tanh(g0[s,ch]) / sqrt(16)
    // q / k rows are fp16; the loads of a 32-column chunk are issued one chunk ahead of its use
    const uint16_t* qkv16 = static_cast<const uint16_t*>(a.qkv);
    const uint16_t* qrow = qkv16 + (size_t)r.g * a.ldq + 128 * half;
    const uint16_t* krow = qkv16 + (size_t)r.j * a.ldq + D_ + 128 * half;
    const uint16_t* vrow = qkv16 + (size_t)r.j * a.ldq + 2 * D_ + 128 * half;
    H32 qc = ldg_h32(qrow), kc = ldg_h32(krow);
    mbar_wait(&bars[3], par);
    tc_fence_after();
    H32 vc;
    {
      float lg[7];
#pragma unroll
      for (int s = 0; s < 7; ++s) lg[s] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        H32 qn, kn;
        if (c < 3) { qn = ldg_h32(qrow + (c + 1) * 32); kn = ldg_h32(krow + (c + 1) * 32); }
        else vc = ldg_h32(vrow);
        float acc[32];
        tmem_ld32(tmem_addr(tmem, 128 * half + c * 32), acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float qf[8], kf[8];
          unpack8(qc.u[i], qf);
          unpack8(kc.u[i], kf);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = c * 32 + 8 * i + e;
            if (col < HQ) lg[col / SC] = fmaf(qf[e] * kf[e], tanh_fast(acc[8 * i + e]), lg[col / SC]);
          }
        }
        if (c < 3) { qc = qn; kc = kn; }
      }
      // head order of the reference: the two adjacency heads first (1 where adjacent, -1e10 otherwise,
      // models/layers.py:170-174), then the 14 computed heads
      LG[row * 17 + half] = ((ex >> half) & 1) ? 1.0f : -1e10f;
#pragma unroll
      for (int s = 0; s < 7; ++s) LG[row * 17 + 2 + 7 * half + s] = lg[s] * 0.25f;
    }
    __syncthreads();
    // (group, head): max, exp in place, 1 / (sum + 1e-16)      (PyG softmax, models/layers.py:178)
    for (int it = t; it < ng * 16; it += AT_THREADS) {
      const int gi = it >> 4, h = it & 15;
      const int gs = gt_meta[gi] & 255u, gl = (gt_meta[gi] >> 8) & 255u;
      float m = -INFINITY;
      for (int rr = gs; rr < gs + gl; ++rr) m = fmaxf(m, LG[rr * 17 + h]);
      float s = 0.f;
      for (int rr = gs; rr < gs + gl; ++rr) {
        const float e = ex2_fast((LG[rr * 17 + h] - m) * 1.4426950408889634f);
        LG[rr * 17 + h] = e;
        s += e;
      }
      GI[gi * 16 + h] = 1.0f / (s + 1e-16f);
    }
    __syncthreads();
    float alpha[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) alpha[h] = r.valid ? LG[row * 17 + 8 * half + h] * GI[r.gi * 16 + 8 * half + h] : 0.f;
    // groups whose first row lies in this warp's 32 rows are summed by this warp
    const uint32_t starts = __ballot_sync(0xffffffffu, r.valid && row == r.gs);

    // ---- messages and per-group sums
    mbar_wait(&bars[4], par);
    tc_fence_after();
    {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        H32 vn;
        if (c < 3) vn = ldg_h32(vrow + (c + 1) * 32);
        float acc[32];
        tmem_ld32(tmem_addr(tmem, 256 + 128 * half + c * 32), acc);
        float* srow = S + row * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float vf[8];
          unpack8(vc.u[i], vf);
          const float al = alpha[2 * c + (i >> 1)];
#pragma unroll
          for (int e = 0; e < 8; ++e) srow[(8 * i + e) ^ lane] = vf[e] * tanh_fast(acc[8 * i + e]) * al;
        }
        if (c < 3) vc = vn;
        named_bar_sync(1 + half, 128);
        uint32_t m = starts;
        while (m) {
          const int r0 = __ffs(m) - 1;
          m &= m - 1;
          const int gl = __shfl_sync(0xffffffffu, r.gl, r0);
          const int node = __shfl_sync(0xffffffffu, r.g, r0);
          const int R0 = rq * 32 + r0;
          float sum = 0.f;
          for (int rr = R0; rr < R0 + gl; ++rr) sum += S[rr * 32 + (lane ^ (rr & 31))];
          a.hnode[(size_t)node * D_ + 128 * half + c * 32 + lane] = sum;
        }
        named_bar_sync(1 + half, 128);
      }
    }
    sync_tc();
    par ^= 1;
  }
  if (tile0 >= tile1 && t == 0) mbar_wait(&bars[0], 0);   // never leave with bulk copies in flight
  sync_tc();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace

cudaError_t launch_attn(const AttnArgs& a, int num_sms, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int grid = a.p.n_tiles < num_sms ? a.p.n_tiles : num_sms;
  k_attn<<<grid, AT_THREADS, AT_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
