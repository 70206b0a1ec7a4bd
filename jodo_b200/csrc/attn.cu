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
//   logits (thread per row, k[j] . q[g] . tanh(g0)), softmax per group through shared memory,
//   msg = v[j] * tanh(g1) * alpha, summed per group, written to hnode[g].
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int AT_A0 = 0;                          // 64 KB: chunks 0,1 = GBF(d) then en; chunks 2,3 = e, later scratch
constexpr int AT_WE = 65536;                      // 32 KB: block edge_emb image (N=64, K=128)
constexpr int AT_W0 = AT_WE + 32768;              // 64 KB: lin_edge0 image (N=256, K=64)
constexpr int AT_W1 = AT_W0 + 65536;              // 64 KB: lin_edge1 image
constexpr int AT_MISC = AT_W1 + 65536;            // barriers, tmem slot, GBF constants, bias, group table
constexpr int AT_SMEM = AT_MISC + 128 + 768 + 256 + 512 + 512;
// scratch inside A0 chunks 2,3 (free once MMA1 has completed)
constexpr int AT_LG = 32768;                      // logits [128][17] fp32
constexpr int AT_GM = AT_LG + 128 * 17 * 4;       // group max  [128][16]
constexpr int AT_GS = AT_GM + 128 * 16 * 4;       // group sum  [128][16]
constexpr int AT_S = 32768;                       // message staging [128][33] fp32 (aliases LG/GM/GS)
static_assert(AT_GS + 128 * 16 * 4 <= 65536, "softmax scratch overflows A0");
static_assert(AT_S + 128 * 33 * 4 <= 65536, "staging overflows A0");
static_assert(AT_SMEM <= 232448, "shared memory budget");

constexpr int SC = 18;        // sub_channels = 256 // 14   (models/layers.py:112)
constexpr int QK = 252;       // 14 * 18

__global__ void __launch_bounds__(ET, 1) k_attn(AttnArgs a) {
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
  float* GM = reinterpret_cast<float*>(smem + AT_GM);
  float* GS = reinterpret_cast<float*>(smem + AT_GS);
  float* S = reinterpret_cast<float*>(smem + AT_S);

  const int t = threadIdx.x;
  if (t == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 32768 + 65536 + 65536);
    bulk_g2s(smem + AT_WE, a.w_emb_img, 32768, &bars[0]);
    bulk_g2s(smem + AT_W0, a.w0_img, 65536, &bars[0]);
    bulk_g2s(smem + AT_W1, a.w1_img, 65536, &bars[0]);
  }
  for (int i = t; i < 192; i += ET) gbf[i] = a.gbf[i];
  if (t < 64) bemb[t] = a.b_emb[t];
  if (t < 32) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  uint32_t par_e = 0, par1 = 0, par2 = 0, par3 = 0;
  bool first = true;
  const float4* pos = reinterpret_cast<const float4*>(a.pos);

  for (int tile = blockIdx.x; tile < a.p.n_tiles; tile += gridDim.x) {
    const RowInfo r = load_row(a.p, tile, t);
    const int ng = a.p.tile_ngroups[tile];
    if (r.valid && t == r.gs) { gt_meta[r.gi] = (uint32_t)r.gs | ((uint32_t)r.gl << 8); gt_node[r.gi] = r.g; }
    if (t == 0) {
      mbar_expect_tx(&bars[1], E_TILE_BYTES);
      bulk_g2s(A0 + 32768, reinterpret_cast<const uint8_t*>(a.e_in) + (size_t)tile * a.e_tile_bytes, E_TILE_BYTES, &bars[1]);
    }
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off;
    const uint8_t ex = a.extra[(size_t)tile * TILE_ROWS + t];

    // ---- distance features -> A0 chunks 0,1
    {
      float df[64];
      if (r.valid) {
        const float d = sq_dist(pos[r.j], pos[r.g]);
        gbf_eval(d, tr[tab_gbf(D_)], tr[tab_gbf(D_) + 1], gbf, df);
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) df[i] = 0.f;
      }
      st_row64<true>(A0, t, 0, df);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      if (first) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], par_e);
      tc_fence_after();
      mma_tile(tmem, smem_u32(A0), smem_u32(smem + AT_WE), 64, 4, false);     // e1 = edge_emb([dist | e])
      umma_commit(&bars[2]);
    }
    first = false;
    par_e ^= 1;
    mbar_wait(&bars[2], par1);
    par1 ^= 1;
    tc_fence_after();

    // ---- en = LN(e1) * (1 + scale_msa) + shift_msa  -> A0 chunks 0,1
    {
      float x[64];
      float h0[32], h1[32];
      tmem_ld32(tmem_addr(tmem, 0), h0);
      tmem_ld32(tmem_addr(tmem, 32), h1);
#pragma unroll
      for (int i = 0; i < 32; ++i) { x[i] = h0[i] + bemb[i]; x[32 + i] = h1[i] + bemb[32 + i]; }
      ln_mod64(x, tr + tab_edge(D_), tr + tab_edge(D_) + ED_);
      if (!r.valid) {
#pragma unroll
        for (int i = 0; i < 64; ++i) x[i] = 0.f;
      }
      st_row64<true>(A0, t, 0, x);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mma_tile(tmem, smem_u32(A0), smem_u32(smem + AT_W0), 256, 2, false);        // g0 pre-activation
      umma_commit(&bars[3]);
      mma_tile(tmem + 256, smem_u32(A0), smem_u32(smem + AT_W1), 256, 2, false);  // g1 pre-activation
      umma_commit(&bars[4]);
    }
    mbar_wait(&bars[3], par2);
    par2 ^= 1;
    tc_fence_after();

    // ---- logits: a[s] = sum_ch q[g,s,ch] k[j,s,ch] tanh(g0[s,ch]) / sqrt(16)
    float lg[14];
#pragma unroll
    for (int s = 0; s < 14; ++s) lg[s] = 0.f;
    {
      const float* qrow = a.qkv + (size_t)r.g * a.ldq;
      const float* krow = a.qkv + (size_t)r.j * a.ldq + D_;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float acc[32];
        tmem_ld32(tmem_addr(tmem, c * 32), acc);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int col = c * 32 + i;
          if (col < QK) {
            const float4 q4 = *reinterpret_cast<const float4*>(qrow + col);
            const float4 k4 = *reinterpret_cast<const float4*>(krow + col);
            lg[col / SC] += q4.x * k4.x * tanh_f(acc[i]);
            if (col + 1 < QK) lg[(col + 1) / SC] += q4.y * k4.y * tanh_f(acc[i + 1]);
            if (col + 2 < QK) lg[(col + 2) / SC] += q4.z * k4.z * tanh_f(acc[i + 2]);
            if (col + 3 < QK) lg[(col + 3) / SC] += q4.w * k4.w * tanh_f(acc[i + 3]);
          }
        }
      }
    }
    // extra heads first: 1 where adjacent, -1e10 otherwise (models/layers.py:170-174)
    LG[t * 17 + 0] = (ex & 1) ? 1.0f : -1e10f;
    LG[t * 17 + 1] = (ex & 2) ? 1.0f : -1e10f;
#pragma unroll
    for (int s = 0; s < 14; ++s) LG[t * 17 + 2 + s] = lg[s] * 0.25f;
    __syncthreads();
    for (int it = t; it < ng * 16; it += ET) {       // (group, head) -> max and sum of exp over the group's rows
      const int gi = it >> 4, h = it & 15;
      const int gs = gt_meta[gi] & 255u, gl = (gt_meta[gi] >> 8) & 255u;
      float m = -INFINITY;
      for (int rr = gs; rr < gs + gl; ++rr) m = fmaxf(m, LG[rr * 17 + h]);
      float s = 0.f;
      for (int rr = gs; rr < gs + gl; ++rr) s += __expf(LG[rr * 17 + h] - m);
      GM[gi * 16 + h] = m;
      GS[gi * 16 + h] = s;
    }
    __syncthreads();
    float alpha[16];
#pragma unroll
    for (int h = 0; h < 16; ++h)
      alpha[h] = r.valid ? __expf(LG[t * 17 + h] - GM[r.gi * 16 + h]) / (GS[r.gi * 16 + h] + 1e-16f) : 0.f;
    __syncthreads();                                  // S aliases LG/GM/GS

    // ---- messages and per-group sums
    mbar_wait(&bars[4], par3);
    par3 ^= 1;
    tc_fence_after();
    {
      const float* vrow = a.qkv + (size_t)r.j * a.ldq + 2 * D_;
      const int ch = t & 31, slot = t >> 5;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float acc[32];
        tmem_ld32(tmem_addr(tmem, 256 + c * 32), acc);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(vrow + c * 32 + i);
          const float al = alpha[(c * 32 + i) / 16];
          S[t * 33 + i] = v4.x * tanh_f(acc[i]) * al;
          S[t * 33 + i + 1] = v4.y * tanh_f(acc[i + 1]) * al;
          S[t * 33 + i + 2] = v4.z * tanh_f(acc[i + 2]) * al;
          S[t * 33 + i + 3] = v4.w * tanh_f(acc[i + 3]) * al;
        }
        __syncthreads();
        for (int gi = slot; gi < ng; gi += 4) {
          const int gs = gt_meta[gi] & 255u, gl = (gt_meta[gi] >> 8) & 255u;
          float sum = 0.f;
          for (int rr = gs; rr < gs + gl; ++rr) sum += S[rr * 33 + ch];
          a.hnode[(size_t)gt_node[gi] * D_ + c * 32 + ch] = sum;
        }
        __syncthreads();
      }
    }
    fence_async_smem();        // scratch written above is overwritten by the next tile's bulk copy
    sync_tc();
  }
  if (t < 32) tmem_dealloc<512>(tmem);
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
  k_attn<<<grid, ET, AT_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
