// Edge residual + edge FFN + edge_l projection of one DGT block on edge tiles (persistent, weights resident).
//
// reference models/mol_gnn.py:304-305 (h_edge = node2edge_lin(h_node[row] + h_node[col]), hoisted to
// P[g] + P[j] + bias with P = W h_node per atom), :313-317 (gated residual from the block INPUT e, norm2_edge,
// modulate, FFN with SiLU, gated residual on the post-norm value) and :568 (edge_i projection that feeds the
// concatenated edge hiddens).  Everything is symmetric in (g, j), so the row orientation does not matter here.
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int EU_EA = 0;                       // 32 KB: e tile -> e2 -> e_out (A operand, K = 64)
constexpr int EU_A2 = 32768;                   // 32 KB: SiLU(hidden chunk) (A operand, K = 64)
constexpr int EU_W = 65536;                    // weights: r x 16 KB (ff3 chunks) | r x 16 KB (ff4 chunks) | 4 KB (edge_l)
constexpr int EU_W_IMG = 64 * 2 * 128;         // one (N=64, K=64) image = 16 KB
constexpr int EU_WL_IMG = 16 * 2 * 128;        // (N=16, K=64) image = 4 KB
__host__ __device__ constexpr int eu_misc(int r) { return EU_W + 2 * r * EU_W_IMG + EU_WL_IMG; }
__host__ __device__ constexpr int eu_smem(int r) { return eu_misc(r) + 128 + (64 * r + 64 + 64 + 16) * 4; }

__global__ void __launch_bounds__(ET, 1) k_edge_update(EdgeUpdateArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  const int r_ = a.r;
  uint8_t* EA = smem + EU_EA;
  uint8_t* A2 = smem + EU_A2;
  uint8_t* W3 = smem + EU_W;
  uint8_t* W4 = W3 + r_ * EU_W_IMG;
  uint8_t* WL = W4 + r_ * EU_W_IMG;
  uint8_t* misc = smem + eu_misc(r_);
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);     // 0: weights, 1: e tile, 2: MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* b3 = reinterpret_cast<float*>(misc + 128);       // [64 r]
  float* b4 = b3 + 64 * r_;                               // [64]
  float* bn = b4 + 64;                                    // [64] node2edge_lin bias
  float* bl = bn + 64;                                    // [16]

  const int t = threadIdx.x;
  if (t == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 2 * r_ * EU_W_IMG + EU_WL_IMG);
    bulk_g2s(W3, a.w3_img, r_ * EU_W_IMG, &bars[0]);
    bulk_g2s(W4, a.w4_img, r_ * EU_W_IMG, &bars[0]);
    bulk_g2s(WL, a.wl_img, EU_WL_IMG, &bars[0]);
  }
  for (int i = t; i < 64 * r_; i += ET) b3[i] = a.b3[i];
  if (t < 64) { b4[t] = a.b4[t]; bn[t] = a.b_n2e[t]; }
  if (t < 16) bl[t] = a.bl[t];
  if (t < 32) tmem_alloc<256>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_y = tmem, tm_f = tmem + 64, tm_l = tmem + 128;
  uint32_t par_e = 0, par_m = 0;
  bool first = true;

  for (int tile = blockIdx.x; tile < a.p.n_tiles; tile += gridDim.x) {
    const RowInfo r = load_row(a.p, tile, t);
    if (t == 0) {
      mbar_expect_tx(&bars[1], E_TILE_BYTES);
      bulk_g2s(EA, reinterpret_cast<const uint8_t*>(a.e_in) + (size_t)tile * a.e_tile_bytes, E_TILE_BYTES, &bars[1]);
    }
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off + tab_edge(D_);
    // h_edge = P[g] + P[j] + b
    float x[64];
    {
      const float* pg = a.P + (size_t)r.g * a.ldp;
      const float* pj = a.P + (size_t)r.j * a.ldp;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 u = *reinterpret_cast<const float4*>(pg + i);
        const float4 v = *reinterpret_cast<const float4*>(pj + i);
        x[i] = u.x + v.x + bn[i]; x[i + 1] = u.y + v.y + bn[i + 1];
        x[i + 2] = u.z + v.z + bn[i + 2]; x[i + 3] = u.w + v.w + bn[i + 3];
      }
    }
    mbar_wait(&bars[1], par_e);
    par_e ^= 1;
    float e2[64];
    {
      float e[64];
      ld_row64(EA, t, 0, e);
#pragma unroll
      for (int i = 0; i < 64; ++i) e2[i] = e[i] + tr[2 * ED_ + i] * x[i];        // e + gate_msa * h_edge
      ln_mod64(e2, tr + 3 * ED_, tr + 4 * ED_);                                    // shift_mlp, scale_mlp
      if (!r.valid) {
#pragma unroll
        for (int i = 0; i < 64; ++i) e2[i] = 0.f;
      }
      st_row64<true>(EA, t, 0, e2);
    }
    fence_async_smem();
    sync_tc();
    for (int hc = 0; hc < r_; ++hc) {
      if (t == 0) {
        if (first) { mbar_wait(&bars[0], 0); first = false; }
        mma_tile(tm_f, smem_u32(EA), smem_u32(W3 + hc * EU_W_IMG), 64, 2, false);
        umma_commit(&bars[2]);
      }
      mbar_wait(&bars[2], par_m);
      par_m ^= 1;
      tc_fence_after();
      {
        float h0[32], h1[32], s[64];
        tmem_ld32(tmem_addr(tm_f, 0), h0);
        tmem_ld32(tmem_addr(tm_f, 32), h1);
#pragma unroll
        for (int i = 0; i < 32; ++i) { s[i] = silu_f(h0[i] + b3[hc * 64 + i]); s[32 + i] = silu_f(h1[i] + b3[hc * 64 + 32 + i]); }
        st_row64<true>(A2, t, 0, s);
      }
      fence_async_smem();
      sync_tc();
      if (t == 0) {
        mma_tile(tm_y, smem_u32(A2), smem_u32(W4 + hc * EU_W_IMG), 64, 2, hc > 0);
        umma_commit(&bars[2]);
      }
      mbar_wait(&bars[2], par_m);      // A2 and tm_f are reused by the next chunk
      par_m ^= 1;
      tc_fence_after();
    }
    first = false;
    // e_out = e2 + gate_mlp * (y + b4)
    {
      float h0[32], h1[32];
      tmem_ld32(tmem_addr(tm_y, 0), h0);
      tmem_ld32(tmem_addr(tm_y, 32), h1);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        e2[i] = r.valid ? e2[i] + tr[5 * ED_ + i] * (h0[i] + b4[i]) : 0.f;
        e2[32 + i] = r.valid ? e2[32 + i] + tr[5 * ED_ + 32 + i] * (h1[i] + b4[32 + i]) : 0.f;
      }
      uint8_t* dst = reinterpret_cast<uint8_t*>(a.e_out) + (size_t)tile * E_TILE_BYTES;
#pragma unroll
      for (int p = 0; p < 16; ++p)
        *reinterpret_cast<float4*>(dst + img_piece(t, p >> 3, p & 7, CHUNK_BYTES_A)) =
            make_float4(e2[4 * p], e2[4 * p + 1], e2[4 * p + 2], e2[4 * p + 3]);
      st_row64<true>(EA, t, 0, e2);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mma_tile(tm_l, smem_u32(EA), smem_u32(WL), 16, 2, false);
      umma_commit(&bars[2]);
    }
    mbar_wait(&bars[2], par_m);
    par_m ^= 1;
    tc_fence_after();
    {
      float h[16];
      tmem_ld16(tmem_addr(tm_l, 0), h);
      uint8_t* dst = reinterpret_cast<uint8_t*>(a.eh_img) + (size_t)tile * a.eh_tile_bytes;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        if (i < a.ce) {
          const int col = a.eh_col + i;
          const float4 o = r.valid ? make_float4(h[i] + bl[i], h[i + 1] + bl[i + 1], h[i + 2] + bl[i + 2], h[i + 3] + bl[i + 3])
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(dst + img_piece(t, col >> 5, (col & 31) >> 2, CHUNK_BYTES_A)) = o;
        }
      }
    }
    fence_async_smem();
    sync_tc();
  }
  if (t < 32) tmem_dealloc<256>(tmem);
}

}  // namespace

cudaError_t launch_edge_update(const EdgeUpdateArgs& a, int num_sms, cudaStream_t st) {
  if (a.r < 1 || a.r > 4 || a.ce % 4 || a.ce > 16 || a.eh_col % 4) return cudaErrorInvalidValue;
  static int attr_bytes = 0;
  const int bytes = eu_smem(a.r);
  if (attr_bytes < bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_update, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    attr_bytes = bytes;
  }
  const int per_sm = bytes <= 110 * 1024 ? 2 : 1;
  const int cap = per_sm * num_sms;
  const int grid = a.p.n_tiles < cap ? a.p.n_tiles : cap;
  k_edge_update<<<grid, ET, bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
