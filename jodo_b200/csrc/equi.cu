// E(3)-equivariant coordinate update of one DGT block on edge tiles.
//
// reference MultiCondEquiUpdate.forward, models/mol_gnn.py:71-94 and CoorsNorm, models/layers.py:344-347:
//   u = cat[h[row], h[col], e, dist]; inv = LN(input_lin(u)) * (1 + scale) + shift;
//   inv = tanh(coord_mlp(inv)) -> [1 + 2]; inv = mean(inv * [1, adj2d, adjsp]);
//   pos[row] += sum_col (pos[row] - pos[col]) / max(|.|, 1e-8) * coord_scale * inv.
// The sum runs over the partners of `row`, so here the group atom g is ROW r and the partner j is COL c
// (same stored rows as the attention pass, roles swapped; edge features are symmetric).
// input_lin is hoisted: W[:, :D] h[g] + W[:, D:2D] h[j] come from the per-atom buffer AB, only the
// [e | dist] part (K = 128) runs per edge.  coord_mlp.0 (256x256) is streamed in 8 K-chunks of 32 KB.
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int EQ_AIN = 0;                        // 64 KB: [e | GBF(d)] (K = 128); later 2 x 32 KB coord_mlp.0 chunk ring
constexpr int EQ_A3 = 65536;                     // 128 KB: input_lin image landing zone, then LN-modulated A (K = 256)
constexpr int EQ_MISC = EQ_A3 + 131072;
constexpr int EQ_SMEM = EQ_MISC + 128 + (192 + 256 + 256 + 768 + 128 * 3) * 4 + 512 + 512;
static_assert(EQ_SMEM <= 232448, "shared memory budget");
constexpr int WC_CHUNK = 256 * 128;              // one K-chunk of coord_mlp.0 (N = 256) = 32 KB

__global__ void __launch_bounds__(ET, 1) k_equi(EquiArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* AIN = smem + EQ_AIN;
  uint8_t* A3 = smem + EQ_A3;
  uint8_t* misc = smem + EQ_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);   // 0: e tile, 1: input_lin image, 2: MMA, 3,4: chunk landed, 5,6: chunk consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* gbf = reinterpret_cast<float*>(misc + 128);    // [192]
  float* b_in = gbf + 192;                              // [256]
  float* b_c0 = b_in + 256;                             // [256]
  float* wc2 = b_c0 + 256;                              // [3][256]
  float* C3 = wc2 + 768;                                // [128][3] per-row coordinate contributions
  uint32_t* gt_meta = reinterpret_cast<uint32_t*>(C3 + 384);
  int* gt_node = reinterpret_cast<int*>(gt_meta + 128);

  const int t = threadIdx.x;
  if (t == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  for (int i = t; i < 192; i += ET) gbf[i] = a.gbf[i];
  for (int i = t; i < 256; i += ET) { b_in[i] = a.b_in[i]; b_c0[i] = a.b_c0[i]; }
  for (int i = t; i < 768; i += ET) wc2[i] = a.wc2[i];
  if (t < 32) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_x = tmem, tm_c = tmem + 256;
  uint32_t par_e = 0, par_w = 0, par_m = 0;
  uint32_t par_land[2] = {0, 0}, par_free[2] = {0, 0};          // thread 0 only
  const float4* pos = reinterpret_cast<const float4*>(a.pos_in);
  float4* pos_out = reinterpret_cast<float4*>(a.pos_out);

  for (int tile = blockIdx.x; tile < a.p.n_tiles; tile += gridDim.x) {
    const RowInfo r = load_row(a.p, tile, t);
    const int ng = a.p.tile_ngroups[tile];
    if (r.valid && t == r.gs) { gt_meta[r.gi] = (uint32_t)r.gs | ((uint32_t)r.gl << 8); gt_node[r.gi] = r.g; }
    if (t == 0) {
      mbar_expect_tx(&bars[0], E_TILE_BYTES);
      bulk_g2s(AIN, reinterpret_cast<const uint8_t*>(a.e) + (size_t)tile * a.e_tile_bytes, E_TILE_BYTES, &bars[0]);
      mbar_expect_tx(&bars[1], 131072);
      bulk_g2s(A3, a.win_img, 131072, &bars[1]);
    }
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off;
    const uint8_t ex = a.extra[(size_t)tile * TILE_ROWS + t];
    const float4 pg = pos[r.g], pj = pos[r.j];
    {
      float df[64];
      if (r.valid) {
        gbf_eval(sq_dist(pg, pj), tr[tab_gbf(D_)], tr[tab_gbf(D_) + 1], gbf, df);
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) df[i] = 0.f;
      }
      st_row64<true>(AIN, t, 2, df);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mbar_wait(&bars[0], par_e);
      mbar_wait(&bars[1], par_w);
      tc_fence_after();
      mma_tile(tm_x, smem_u32(AIN), smem_u32(A3), 256, 4, false);       // input_lin edge part
      umma_commit(&bars[2]);
    }
    par_e ^= 1; par_w ^= 1;
    mbar_wait(&bars[2], par_m);
    par_m ^= 1;
    tc_fence_after();
    if (t == 0) {      // AIN is free now: start streaming coord_mlp.0 chunks 0 and 1
      for (int s = 0; s < 2; ++s) {
        mbar_expect_tx(&bars[3 + s], WC_CHUNK);
        bulk_g2s(AIN + s * WC_CHUNK, reinterpret_cast<const uint8_t*>(a.wc0_img) + (size_t)s * WC_CHUNK, WC_CHUNK, &bars[3 + s]);
      }
    }

    // ---- pass 1: x = acc + A[g] + B[j] + b, kept in TMEM; row statistics (shifted by the first element)
    float mean, rstd;
    {
      const float* ag = a.AB + (size_t)r.g * a.ldab;
      const float* bj = a.AB + (size_t)r.j * a.ldab + D_;
      float sh0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float x[32];
        tmem_ld32(tmem_addr(tm_x, c * 32), x);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 u = *reinterpret_cast<const float4*>(ag + c * 32 + i);
          const float4 v = *reinterpret_cast<const float4*>(bj + c * 32 + i);
          x[i] += u.x + v.x + b_in[c * 32 + i];
          x[i + 1] += u.y + v.y + b_in[c * 32 + i + 1];
          x[i + 2] += u.z + v.z + b_in[c * 32 + i + 2];
          x[i + 3] += u.w + v.w + b_in[c * 32 + i + 3];
        }
        if (c == 0) sh0 = x[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = x[i] - sh0; s1 += d; s2 += d * d; }
        tmem_st32(tmem_addr(tm_x, c * 32), x);
      }
      tmem_wait_st();
      const float m1 = s1 * (1.0f / 256.0f);
      mean = sh0 + m1;
      rstd = rsqrtf(fmaxf(s2 * (1.0f / 256.0f) - m1 * m1, 0.f) + 1e-6f);
    }
    // ---- pass 2: LN + modulate -> A3 (K = 256); the input_lin image there is no longer needed
    {
      const float* shift = tr + tab_equi(D_);
      const float* scale = shift + D_;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float x[32];
        tmem_ld32(tmem_addr(tm_x, c * 32), x);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 sh = *reinterpret_cast<const float4*>(shift + c * 32 + i);
          const float4 sc = *reinterpret_cast<const float4*>(scale + c * 32 + i);
          x[i] = r.valid ? (x[i] - mean) * rstd * (1.0f + sc.x) + sh.x : 0.f;
          x[i + 1] = r.valid ? (x[i + 1] - mean) * rstd * (1.0f + sc.y) + sh.y : 0.f;
          x[i + 2] = r.valid ? (x[i + 2] - mean) * rstd * (1.0f + sc.z) + sh.z : 0.f;
          x[i + 3] = r.valid ? (x[i + 3] - mean) * rstd * (1.0f + sc.w) + sh.w : 0.f;
        }
        st_row32<true>(A3, t, c, x);
      }
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {      // coord_mlp.0: 8 K-chunks through a 2-deep ring in AIN
      for (int kc = 0; kc < 8; ++kc) {
        const int s = kc & 1;
        mbar_wait(&bars[3 + s], par_land[s]);
        par_land[s] ^= 1;
        tc_fence_after();
        mma_tile(tm_c, smem_u32(A3 + kc * CHUNK_BYTES_A), smem_u32(AIN + s * WC_CHUNK), 256, 1, kc > 0);
        if (kc + 2 < 8) {
          umma_commit(&bars[5 + s]);
          mbar_wait(&bars[5 + s], par_free[s]);
          par_free[s] ^= 1;
          mbar_expect_tx(&bars[3 + s], WC_CHUNK);
          bulk_g2s(AIN + s * WC_CHUNK, reinterpret_cast<const uint8_t*>(a.wc0_img) + (size_t)(kc + 2) * WC_CHUNK, WC_CHUNK,
                   &bars[3 + s]);
        }
      }
      umma_commit(&bars[2]);
    }
    mbar_wait(&bars[2], par_m);
    par_m ^= 1;
    tc_fence_after();

    // ---- coord_mlp.2 on CUDA cores, tanh, adjacency-weighted mean, coordinate contribution
    {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float x[32];
        tmem_ld32(tmem_addr(tm_c, c * 32), x);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = silu_f(x[i] + b_c0[c * 32 + i]);
          o0 += s * wc2[c * 32 + i];
          o1 += s * wc2[256 + c * 32 + i];
          o2 += s * wc2[512 + c * 32 + i];
        }
      }
      const float w = (tanh_f(o0) + ((ex & 1) ? tanh_f(o1) : 0.f) + ((ex & 2) ? tanh_f(o2) : 0.f)) * (1.0f / 3.0f);
      const float dx = pg.x - pj.x, dy = pg.y - pj.y, dz = pg.z - pj.z;
      const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
      const float f = r.valid ? a.coord_scale * w / fmaxf(nrm, 1e-8f) : 0.f;
      C3[t * 3 + 0] = dx * f; C3[t * 3 + 1] = dy * f; C3[t * 3 + 2] = dz * f;
    }
    __syncthreads();
    if (t < ng) {
      const int gs = gt_meta[t] & 255u, gl = (gt_meta[t] >> 8) & 255u;
      const int node = gt_node[t];
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int rr = gs; rr < gs + gl; ++rr) { sx += C3[rr * 3]; sy += C3[rr * 3 + 1]; sz += C3[rr * 3 + 2]; }
      const float4 p0 = pos[node];
      pos_out[node] = make_float4(p0.x + sx, p0.y + sy, p0.z + sz, 0.f);
    }
    fence_async_smem();
    sync_tc();
  }
  if (t < 32) tmem_dealloc<512>(tmem);
}

}  // namespace

cudaError_t launch_equi(const EquiArgs& a, int num_sms, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_equi, cudaFuncAttributeMaxDynamicSharedMemorySize, EQ_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int grid = a.p.n_tiles < num_sms ? a.p.n_tiles : num_sms;
  k_equi<<<grid, ET, EQ_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
