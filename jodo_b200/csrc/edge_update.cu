// Edge residual + edge FFN + edge_l projection of one DGT block on edge tiles (persistent, weights resident).
//
// reference models/mol_gnn.py:304-305 (h_edge = node2edge_lin(h_node[row] + h_node[col]), hoisted to
// P[g] + P[j] + bias with P = W h_node per atom), :313-317 (gated residual from the block INPUT e, norm2_edge,
// modulate, FFN with SiLU, gated residual on the post-norm value) and :568 (edge_i projection that feeds the
// concatenated edge hiddens).  Everything is symmetric in (g, j), so the row orientation does not matter here.
//
// 512 threads = two independent groups of 8 warps, each walking its own tiles (group-local named barriers, its own
// operand buffers, 256 TMEM columns and mbarrier), sharing the resident weight images: while one group waits for a
// tensor-core round trip the other one computes.  Inside a group warp w works on tile rows 32*(w&3)..+31 and on column
// half (w>>2): 32 of the 64 edge features and 64 of every 128 hidden units.
// Data movement: the fp32 edge state lives in HBM in a piece-major tile layout ([16 pieces][128 rows][16 B], private to
// jodo_edge_embed / jodo_edge_update) so that the row-per-thread loads and stores are fully coalesced; the fp16 operand
// copy for the other kernels is written to shared memory once and leaves through one bulk store; per-column constants
// are kernel-parameter / constant-memory operands (uniform-conditioning fast path, see equi.cu).
#include "edge_common.cuh"

namespace jodo {

// rows of the fp32 edge state whose fp16 operand copy saturated (see jodo_saturation_count)
__device__ unsigned int g_sat_edge_update;

__constant__ float c_eumod[384];       // row 0 of the edge AdaLN table: (shift, scale, gate)_msa, (shift, scale, gate)_mlp

namespace {

constexpr int EU_THREADS = 512;
constexpr int EU_GROUP = 256;
// shared memory: [W3 r*16 KB][W4 r*16 KB][WL 2 KB][group 0: A 16 KB, A2 32 KB][group 1: ...][misc]
__host__ __device__ constexpr int eu_w3() { return 0; }
__host__ __device__ constexpr int eu_w4(int r) { return r * 8192; }
__host__ __device__ constexpr int eu_wl(int r) { return 2 * r * 8192; }
// each group also owns a 32 KB buffer into which its next fp32 edge tile is bulk-copied one tile ahead
__host__ __device__ constexpr int eu_gbytes(int r) { return 49152 + 32768; }
__host__ __device__ constexpr int eu_grp(int r, int g) { return 2 * r * 8192 + 2048 + g * eu_gbytes(r); }
__host__ __device__ constexpr int eu_misc(int r) { return eu_grp(r, 2); }
__host__ __device__ constexpr int eu_smem(int r) { return eu_misc(r) + 128; }     // r = 4: 226 KB + 128 B

__device__ __forceinline__ void group_sync(int grp) {        // group-local barrier that also orders tcgen05 traffic
  tc_fence_before();
  named_bar_sync(1 + grp, EU_GROUP);
  tc_fence_after();
}
// shared::cta -> global bulk store (TMA engine)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct GroupCtx {
  uint8_t* A; uint8_t* A2; const uint8_t* W3; const uint8_t* W4; const uint8_t* WL;
  uint64_t* bar_w; uint64_t* bar_m; uint64_t* bar_e; float2* LNS; uint8_t* EA;
  uint32_t tm_f, tm_y, tm_l;
  int grp, lt, row, tile0, tile1;
};

template <int HALF, bool UNI, int R>
__device__ __forceinline__ void eu_group_loop(const EdgeUpdateArgs& a, const GroupCtx& c) {
  constexpr int C0 = 32 * HALF;
  constexpr int NCH = R / 2;                      // hidden chunks of 128 units
  const int row = c.row, lt = c.lt;
  uint32_t par_m = 0, par_e = 0;
  uint8_t* e32 = reinterpret_cast<uint8_t*>(a.e32);
  if (lt == 0 && c.tile0 < c.tile1) {
    mbar_expect_tx(c.bar_e, E_TILE_BYTES);
    bulk_g2s(c.EA, e32 + (size_t)c.tile0 * E_TILE_BYTES, E_TILE_BYTES, c.bar_e);
  }
  RowInfo rn = load_row(a.p, min(c.tile0, a.p.n_tiles - 1), row);      // row metadata is fetched one tile ahead
  for (int tile = c.tile0; tile < c.tile1; tile += 2) {
    const RowInfo r = rn;
    // ---- loads of this tile: fp32 e (own 32 columns, piece-major tile), P[g], P[j] (piece-major fp16)
    float4 ev[8];
    {                                              // staged tile: [16 pieces][128 rows][16 B], conflict-free row reads
      mbar_wait(c.bar_e, par_e);
      par_e ^= 1;
      const float4* src = reinterpret_cast<const float4*>(c.EA) + (8 * HALF) * 128 + row;
#pragma unroll
      for (int p = 0; p < 8; ++p) ev[p] = src[p * 128];
    }
    const H32 pu = ldg_pm32(a.P, a.ldp, r.g, 4 * HALF);
    const H32 pv = ldg_pm32(a.P, a.ldp, r.j, 4 * HALF);
    rn = load_row(a.p, tile + 2 < c.tile1 ? tile + 2 : tile, row);
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off + tab_edge(D_);
    // ---- e2 = LN(e + gate_msa * (P[g] + P[j] + b)) * (1 + scale_mlp) + shift_mlp
    float e2[32];
    {
      float gv[32];                                  // general path: this molecule's row, 16-byte loads
      if (!UNI) ldg_row<32>(tr + 2 * ED_ + C0, gv);
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float uf[8], vf[8];
        unpack8(pu.u[i], uf);
        unpack8(pv.u[i], vf);
        const float ee[8] = {ev[2 * i].x, ev[2 * i].y, ev[2 * i].z, ev[2 * i].w, ev[2 * i + 1].x, ev[2 * i + 1].y, ev[2 * i + 1].z, ev[2 * i + 1].w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int col = C0 + 8 * i + k;
          const float g = UNI ? c_eumod[2 * ED_ + col] : gv[8 * i + k];
          const float v = fmaf(g, (uf[k] + vf[k]) + a.b_n2e[col], ee[k]);
          e2[8 * i + k] = v;
          s += v;
          q = fmaf(v, v, q);
        }
      }
      c.LNS[row * 2 + HALF] = make_float2(s, q);
      if (lt == 0) bulk_wait_read();               // the previous tile's bulk store has finished reading A
      named_bar_sync(1 + c.grp, EU_GROUP);
      if (lt == 0 && tile + 2 < c.tile1) {       // the staged tile is consumed: fetch this group's next one
        mbar_expect_tx(c.bar_e, E_TILE_BYTES);
        bulk_g2s(c.EA, e32 + (size_t)(tile + 2) * E_TILE_BYTES, E_TILE_BYTES, c.bar_e);
      }
      const float2 o = c.LNS[row * 2 + (HALF ^ 1)];
      const float mean = (s + o.x) * (1.0f / 64.0f);
      const float rstd = rsqrtf(fmaxf((q + o.y) * (1.0f / 64.0f) - mean * mean, 0.f) + 1e-6f);
      const float nmr = -mean * rstd;
      float scv[32], shv[32];
      if (!UNI) { ldg_row<32>(tr + 4 * ED_ + C0, scv); ldg_row<32>(tr + 3 * ED_ + C0, shv); }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = C0 + i;
        const float n = fmaf(e2[i], rstd, nmr);
        const float sc = UNI ? c_eumod[4 * ED_ + col] : scv[i];
        const float sh = UNI ? c_eumod[3 * ED_ + col] : shv[i];
        e2[i] = fmaf(n, sc, sh);               // padding rows carry finite garbage until the final select
      }
      st_rowh<32>(c.A, row, 0, 4 * HALF, e2);
    }
    fence_async_smem();
    group_sync(c.grp);
    if (lt == 0) {
      if (tile == c.tile0) mbar_wait(c.bar_w, 0);
      tc_fence_after();
      mma_tile_h(c.tm_f, smem_u32(c.A), smem_u32(c.W3), 128, 1, false);          // hidden chunk 0
      umma_commit(c.bar_m);
    }
#pragma unroll
    for (int hc = 0; hc < NCH; ++hc) {
      mbar_wait(c.bar_m, par_m);                   // MMA1(hc) done (and, for hc > 0, MMA2(hc-1): A2 is free again)
      par_m ^= 1;
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 2; ++q) {                // SiLU(hidden) -> A2: own 64 of the 128 units, 32 at a time
        float h[32];
        tmem_ld32(tmem_addr(c.tm_f, 64 * HALF + 32 * q), h);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = h[i] + a.b3[128 * hc + 64 * HALF + 32 * q + i];     // image and bias pre-scaled by 1/2
          h[i] = fmaf(x, tanh_fast(x), x);
        }
        st_rowh<32>(c.A2, row, HALF, 4 * q, h);
      }
      fence_async_smem();
      group_sync(c.grp);
      if (lt == 0) {
        mma_tile_h(c.tm_y, smem_u32(c.A2), smem_u32(c.W4 + hc * 16384), 64, 2, hc > 0);
        if (hc + 1 < NCH) mma_tile_h(c.tm_f, smem_u32(c.A), smem_u32(c.W3 + (hc + 1) * 16384), 128, 1, false);
        umma_commit(c.bar_m);
      }
    }
    mbar_wait(c.bar_m, par_m);
    par_m ^= 1;
    tc_fence_after();
    // ---- e_out = e2 + gate_mlp * (y + b4): fp32 state (coalesced piece-major store), fp16 operand copy via A
    {
      float y[32];
      tmem_ld32(tmem_addr(c.tm_y, C0), y);
      float gv[32];
      if (!UNI) ldg_row<32>(tr + 5 * ED_ + C0, gv);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = C0 + i;
        const float g = UNI ? c_eumod[5 * ED_ + col] : gv[i];
        e2[i] = r.valid ? fmaf(g, y[i] + a.b4[col], e2[i]) : 0.f;
      }
      {                                            // the fp16 operand copy of the edge state clamps at +-65504: count it
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fabsf(e2[i]));
        if (mx > 65504.f) atomicAdd(&g_sat_edge_update, 1u);
      }
      float4* dst = reinterpret_cast<float4*>(e32 + (size_t)tile * E_TILE_BYTES) + (8 * HALF) * 128 + row;
#pragma unroll
      for (int p = 0; p < 8; ++p) dst[p * 128] = make_float4(e2[4 * p], e2[4 * p + 1], e2[4 * p + 2], e2[4 * p + 3]);
      st_rowh<32>(c.A, row, 0, 4 * HALF, e2);
    }
    fence_async_smem();
    group_sync(c.grp);
    if (lt == 0) {
      mma_tile_h(c.tm_l, smem_u32(c.A), smem_u32(c.WL), 16, 1, false);
      umma_commit(c.bar_m);
      bulk_s2g(reinterpret_cast<uint8_t*>(a.e16) + (size_t)tile * CHUNK_BYTES_A, c.A, CHUNK_BYTES_A);
    }
    mbar_wait(c.bar_m, par_m);
    par_m ^= 1;
    tc_fence_after();
    if (HALF == 0) {
      float h[16];
      tmem_ld16(tmem_addr(c.tm_l, 0), h);
      uint8_t* dst = reinterpret_cast<uint8_t*>(a.eh) + (size_t)tile * a.eh_tile_bytes;
      if (a.ce == 16 && (a.eh_col & 7) == 0) {     // 16 columns = two whole 16-byte pieces of the row
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = r.valid ? h[i] + a.bl[i] : 0.f;
        const int col = a.eh_col;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          uint4 o;
          o.x = pack_h2(h[8 * p], h[8 * p + 1]); o.y = pack_h2(h[8 * p + 2], h[8 * p + 3]);
          o.z = pack_h2(h[8 * p + 4], h[8 * p + 5]); o.w = pack_h2(h[8 * p + 6], h[8 * p + 7]);
          *reinterpret_cast<uint4*>(dst + img_piece(row, col >> 6, ((col & 63) >> 3) + p, CHUNK_BYTES_A)) = o;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i < a.ce) {
            const int col = a.eh_col + i;
            const uint32_t off = img_piece(row, col >> 6, (col & 63) >> 3, CHUNK_BYTES_A) + ((col & 7) << 1);
            const uint32_t hv = pack_h2(r.valid ? h[i] + a.bl[i] : 0.f, 0.f);
            *reinterpret_cast<uint16_t*>(dst + off) = (uint16_t)(hv & 0xFFFFu);
          }
        }
      }
    }
    tc_fence_before();                              // tm_l / tm_y reads are ordered before the next tile's MMAs by the next group_sync
  }
  if (lt == 0) bulk_wait_read();
}

template <int R>
__global__ void __launch_bounds__(EU_THREADS, 1) k_edge_update(const __grid_constant__ EdgeUpdateArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* misc = smem + eu_misc(R);
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);     // 0: weights, 1,2: MMA of group 0,1, 3,4: staged e tile of group 0,1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t >> 8, lt = t & 255, lw = lt >> 5;
  const int per = (a.p.n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, a.p.n_tiles);

  if (t == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 2 * R * 8192 + 2048);
    bulk_g2s(smem + eu_w3(), a.w3_img, R * 8192, &bars[0]);
    bulk_g2s(smem + eu_w4(R), a.w4_img, R * 8192, &bars[0]);
    bulk_g2s(smem + eu_wl(R), a.wl_img, 2048, &bars[0]);
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot + 256u * grp;
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;

  GroupCtx c;
  c.A = smem + eu_grp(R, grp);
  c.A2 = c.A + 16384;
  c.W3 = smem + eu_w3(); c.W4 = smem + eu_w4(R); c.WL = smem + eu_wl(R);
  c.bar_w = &bars[0]; c.bar_m = &bars[1 + grp]; c.bar_e = &bars[3 + grp];
  c.EA = c.A + 49152;
  c.LNS = reinterpret_cast<float2*>(c.A2);           // LayerNorm partial sums live in A2, which is rewritten only later (SiLU)
  c.tm_f = tmem; c.tm_y = tmem + 128; c.tm_l = tmem + 192;
  c.grp = grp; c.lt = lt; c.row = (lw & 3) * 32 + (t & 31);
  c.tile0 = tile0 + grp; c.tile1 = tile1;
  if ((lw >> 2) == 0) {
    if (uni) eu_group_loop<0, true, R>(a, c); else eu_group_loop<0, false, R>(a, c);
  } else {
    if (uni) eu_group_loop<1, true, R>(a, c); else eu_group_loop<1, false, R>(a, c);
  }
  if (t == 0) mbar_wait(&bars[0], 0);                // never leave with the weight copies in flight
  sync_tc();
  if (warp == 0) tmem_dealloc<512>(*tmem_slot);
}

}  // namespace

cudaError_t sat_count_edge_update(unsigned int* out, bool reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_sat_edge_update, sizeof(unsigned int));
  if (e == cudaSuccess && reset) { const unsigned int z = 0; e = cudaMemcpyToSymbol(g_sat_edge_update, &z, sizeof(z)); }
  return e;
}

cudaError_t launch_edge_update(const EdgeUpdateArgs& a, int num_sms, cudaStream_t st) {
  if ((a.r != 2 && a.r != 4) || a.ce < 1 || a.ce > 16) return cudaErrorInvalidValue;
  static DevAttr attr2 = {}, attr4 = {};
  const int bytes = eu_smem(a.r);
  cudaError_t e0 = a.r == 2 ? ensure_dyn_smem(k_edge_update<2>, bytes, attr2) : ensure_dyn_smem(k_edge_update<4>, bytes, attr4);
  if (e0 != cudaSuccess) return e0;
  if ((e0 = const_tables_acquire(st)) != cudaSuccess) return e0;
  if (a.nonuni) {     // row 0 of the edge AdaLN table feeds the uniform fast path
    cudaError_t e = cudaMemcpyToSymbolAsync(c_eumod, a.tab + a.tab_off + tab_edge(D_), sizeof(float) * 384, 0,
                                            cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  }
  const int grid = a.p.n_tiles < 2 * num_sms ? (a.p.n_tiles + 1) / 2 : num_sms;
  if (a.r == 2) k_edge_update<2><<<grid, EU_THREADS, bytes, st>>>(a);
  else k_edge_update<4><<<grid, EU_THREADS, bytes, st>>>(a);
  if ((e0 = cudaGetLastError()) != cudaSuccess) return e0;
  return const_tables_release(st);
}

}  // namespace jodo
