// Edge residual + edge FFN + edge_l projection of one DGT block on edge tiles (persistent, weights resident).
//
// reference models/mol_gnn.py:304-305 (h_edge = node2edge_lin(h_node[row] + h_node[col]), hoisted to
// P[g] + P[j] + bias with P = W h_node per atom), :313-317 (gated residual from the block INPUT e, norm2_edge,
// modulate, FFN with SiLU, gated residual on the post-norm value) and :568 (edge_i projection that feeds the
// concatenated edge hiddens).  Everything is symmetric in (g, j), so the row orientation does not matter here.
//
// 256 threads: warp w works on tile rows 32*(w&3)..+31 and on column half (w>>2): 32 of the 64 edge features and
// 32r of the 64r hidden units.  The fp32 edge state is streamed through a 2-deep bulk-copy ring and rewritten in
// place; the fp16 operand copy and the edge-hidden slice are written next to it.
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int EU_THREADS = 256;
constexpr int EU_EA = 0;                       // 2 x 32 KB: fp32 e tiles (bulk-copy ring)
constexpr int EU_A = 65536;                    // 16 KB fp16: e2, then e_out (A operand, K = 64)
constexpr int EU_A2 = EU_A + 16384;            // r x 16 KB fp16: SiLU(hidden) (A operand, K = 64 r)
__host__ __device__ constexpr int eu_w3(int r) { return EU_A2 + r * 16384; }       // (N = 64 r, K = 64): 8 r KB
__host__ __device__ constexpr int eu_w4(int r) { return eu_w3(r) + r * 8192; }     // (N = 64, K = 64 r): 8 r KB
__host__ __device__ constexpr int eu_wl(int r) { return eu_w4(r) + r * 8192; }     // (N = 16, K = 64): 2 KB
__host__ __device__ constexpr int eu_misc(int r) { return eu_wl(r) + 2048; }
__host__ __device__ constexpr int eu_smem(int r) { return eu_misc(r) + 128 + (64 * r + 64 + 64 + 16) * 4 + 128 * 2 * 8; }

__global__ void __launch_bounds__(EU_THREADS, 1) k_edge_update(EdgeUpdateArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  const int r_ = a.r;
  uint8_t* A = smem + EU_A;
  uint8_t* A2 = smem + EU_A2;
  uint8_t* W3 = smem + eu_w3(r_);
  uint8_t* W4 = smem + eu_w4(r_);
  uint8_t* WL = smem + eu_wl(r_);
  uint8_t* misc = smem + eu_misc(r_);
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);     // 0: weights, 1,2: e tile ring, 3: MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* b3 = reinterpret_cast<float*>(misc + 128);       // [64 r]
  float* b4 = b3 + 64 * r_;                               // [64]
  float* bn = b4 + 64;                                    // [64] node2edge_lin bias
  float* bl = bn + 64;                                    // [16]
  float2* LNS = reinterpret_cast<float2*>(bl + 16);       // [128][2]

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int rq = warp & 3, half = warp >> 2;
  const int row = rq * 32 + lane;
  const int per = (a.p.n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, a.p.n_tiles);

  if (t == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 2 * r_ * 8192 + 2048);
    bulk_g2s(W3, a.w3_img, r_ * 8192, &bars[0]);
    bulk_g2s(W4, a.w4_img, r_ * 8192, &bars[0]);
    bulk_g2s(WL, a.wl_img, 2048, &bars[0]);
    if (tile0 < tile1) {
      mbar_expect_tx(&bars[1], E_TILE_BYTES);
      bulk_g2s(smem + EU_EA, reinterpret_cast<const uint8_t*>(a.e32) + (size_t)tile0 * E_TILE_BYTES, E_TILE_BYTES, &bars[1]);
    }
  }
  for (int i = t; i < 64 * r_; i += EU_THREADS) b3[i] = a.b3[i];
  if (t < 64) { b4[t] = a.b4[t]; bn[t] = a.b_n2e[t]; }
  if (t < 16) bl[t] = a.bl[t];
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_f = tmem, tm_y = tmem + 256, tm_l = tmem + 320;
  uint32_t par_e[2] = {0, 0}, par_m = 0;

  RowInfo rn = load_row(a.p, min(tile0, a.p.n_tiles - 1), row);      // row metadata is fetched one tile ahead
  const uint16_t* p16 = static_cast<const uint16_t*>(a.P);
  for (int tile = tile0; tile < tile1; ++tile) {
    const int buf = (tile - tile0) & 1;
    uint8_t* EA = smem + EU_EA + buf * E_TILE_BYTES;
    const RowInfo r = rn;
    if (t == 0 && tile + 1 < tile1) {          // the other ring slot was fully consumed in the previous iteration
      mbar_expect_tx(&bars[1 + (buf ^ 1)], E_TILE_BYTES);
      bulk_g2s(smem + EU_EA + (buf ^ 1) * E_TILE_BYTES, reinterpret_cast<const uint8_t*>(a.e32) + (size_t)(tile + 1) * E_TILE_BYTES,
               E_TILE_BYTES, &bars[1 + (buf ^ 1)]);
    }
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off + tab_edge(D_);
    const int c0 = 32 * half;
    // h_edge = P[g] + P[j] + b   (own 32 columns)
    float x[32];
    {
      const H32 u = ldg_h32(p16 + (size_t)r.g * a.ldp + c0);
      const H32 v = ldg_h32(p16 + (size_t)r.j * a.ldp + c0);
      rn = load_row(a.p, min(tile + 1, tile1 - 1), row);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float uf[8], vf[8];
        unpack8(u.u[i], uf);
        unpack8(v.u[i], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) x[8 * i + e] = uf[e] + vf[e] + bn[c0 + 8 * i + e];
      }
    }
    mbar_wait(&bars[1 + buf], par_e[buf]);
    par_e[buf] ^= 1;
    float e2[32];
    {
      float e[32];
      ld_row32(EA, row, half, e);
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        e2[i] = fmaf(tr[2 * ED_ + c0 + i], x[i], e[i]);        // e + gate_msa * h_edge
        s += e2[i];
        q = fmaf(e2[i], e2[i], q);
      }
      LNS[row * 2 + half] = make_float2(s, q);
      __syncthreads();
      const float2 o = LNS[row * 2 + (half ^ 1)];
      const float mean = (s + o.x) * (1.0f / 64.0f);
      const float rstd = rsqrtf(fmaxf((q + o.y) * (1.0f / 64.0f) - mean * mean, 0.f) + 1e-6f);
      const float* shift = tr + 3 * ED_ + c0;                   // shift_mlp, scale_mlp
      const float* scale = tr + 4 * ED_ + c0;
#pragma unroll
      for (int i = 0; i < 32; ++i) e2[i] = r.valid ? fmaf((e2[i] - mean) * rstd, 1.0f + scale[i], shift[i]) : 0.f;
      st_rowh<32>(A, row, 0, 4 * half, e2);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      if (tile == tile0) mbar_wait(&bars[0], 0);
      tc_fence_after();
      mma_tile_h(tm_f, smem_u32(A), smem_u32(W3), 64 * r_, 1, false);
      umma_commit(&bars[3]);
    }
    mbar_wait(&bars[3], par_m);
    par_m ^= 1;
    tc_fence_after();
    for (int q = 0; q < r_; ++q) {             // SiLU(hidden) -> A2, 32 columns at a time
      const int blk = r_ * half + q;
      float h[32];
      tmem_ld32(tmem_addr(tm_f, 32 * blk), h);
#pragma unroll
      for (int i = 0; i < 32; ++i) h[i] = silu_fast(h[i] + b3[32 * blk + i]);
      st_rowh<32>(A2, row, blk >> 1, 4 * (blk & 1), h);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mma_tile_h(tm_y, smem_u32(A2), smem_u32(W4), 64, r_, false);
      umma_commit(&bars[3]);
    }
    mbar_wait(&bars[3], par_m);
    par_m ^= 1;
    tc_fence_after();
    // e_out = e2 + gate_mlp * (y + b4)
    {
      float y[32];
      tmem_ld32(tmem_addr(tm_y, c0), y);
#pragma unroll
      for (int i = 0; i < 32; ++i) e2[i] = r.valid ? fmaf(tr[5 * ED_ + c0 + i], y[i] + b4[c0 + i], e2[i]) : 0.f;
      uint8_t* dst = reinterpret_cast<uint8_t*>(a.e32) + (size_t)tile * E_TILE_BYTES;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        *reinterpret_cast<float4*>(dst + img_piece(row, half, p, CHUNK_BYTES_A)) =
            make_float4(e2[4 * p], e2[4 * p + 1], e2[4 * p + 2], e2[4 * p + 3]);
      st_rowh<32>(reinterpret_cast<uint8_t*>(a.e16) + (size_t)tile * CHUNK_BYTES_A, row, 0, 4 * half, e2);
      st_rowh<32>(A, row, 0, 4 * half, e2);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mma_tile_h(tm_l, smem_u32(A), smem_u32(WL), 16, 1, false);
      umma_commit(&bars[3]);
    }
    mbar_wait(&bars[3], par_m);
    par_m ^= 1;
    tc_fence_after();
    if (half == 0) {
      float h[16];
      tmem_ld16(tmem_addr(tm_l, 0), h);
      uint8_t* dst = reinterpret_cast<uint8_t*>(a.eh) + (size_t)tile * a.eh_tile_bytes;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < a.ce) {
          const int col = a.eh_col + i;
          const uint32_t off = img_piece(row, col >> 6, (col & 63) >> 3, CHUNK_BYTES_A) + ((col & 7) << 1);
          const uint32_t hv = pack_h2(r.valid ? h[i] + bl[i] : 0.f, 0.f);
          *reinterpret_cast<uint16_t*>(dst + off) = (uint16_t)(hv & 0xFFFFu);
        }
      }
    }
    sync_tc();
  }
  if (tile0 >= tile1 && t == 0) mbar_wait(&bars[0], 0);   // never leave with bulk copies in flight
  sync_tc();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace

cudaError_t launch_edge_update(const EdgeUpdateArgs& a, int num_sms, cudaStream_t st) {
  if (a.r < 1 || a.r > 4 || a.ce < 1 || a.ce > 16) return cudaErrorInvalidValue;
  static int attr_bytes = 0;
  const int bytes = eu_smem(a.r);
  if (attr_bytes < bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_update, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    attr_bytes = bytes;
  }
  const int grid = a.p.n_tiles < num_sms ? a.p.n_tiles : num_sms;
  k_edge_update<<<grid, EU_THREADS, bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
