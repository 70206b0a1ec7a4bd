// E(3)-equivariant coordinate update of one DGT block on edge tiles.
//
// reference MultiCondEquiUpdate.forward, models/mol_gnn.py:71-94 and CoorsNorm, models/layers.py:344-347:
//   u = cat[h[row], h[col], e, dist]; inv = LN(input_lin(u)) * (1 + scale) + shift;
//   inv = tanh(coord_mlp(inv)) -> [1 + 2]; inv = mean(inv * [1, adj2d, adjsp]);
//   pos[row] += sum_col (pos[row] - pos[col]) / max(|.|, 1e-8) * coord_scale * inv.
// The sum runs over the partners of `row`, so here the group atom g is ROW r and the partner j is COL c
// (same stored rows as the attention pass, roles swapped; edge features are symmetric).
// input_lin is hoisted: W[:, :D] h[g] + W[:, D:2D] h[j] come from the per-atom buffer AB, only the
// [e | dist] part (K = 128) runs per edge.
//
// fp16 operand images (same mantissa as tf32).  coord_mlp.0 (256x256, 128 KB) stays resident in shared memory; the
// input_lin image (64 KB) is bulk-copied per tile into the region that later holds the LayerNorm-modulated operand,
// and that copy as well as the next e tile are issued as soon as the MMA that read the region has completed, so they
// overlap the epilogues.  256 threads: warp w = tile rows 32*(w&3)..+31 x column half (w>>2) of the 256 hidden units.
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int EQ_THREADS = 256;
constexpr int EQ_WC0 = 0;                        // 128 KB: coord_mlp.0 image (N = 256, K = 256: 4 chunks of 32 KB)
constexpr int EQ_X = 131072;                     // 64 KB: input_lin image (N = 256, K = 128), then LN-modulated A (K = 256)
constexpr int EQ_U = EQ_X + 65536;               // 32 KB: [e | GBF(d)] (K = 128); chunk 1 doubles as scratch
constexpr int EQ_MISC = EQ_U + 32768;
constexpr int EQ_SMEM = EQ_MISC + 128 + 768;
static_assert(EQ_SMEM <= 232448, "shared memory budget");
// scratch inside U chunk 1 (free once the input_lin MMA has completed)
constexpr int EQ_SCR = EQ_U + 16384;
constexpr int EQ_LNS = EQ_SCR;                   // [128][2] float2
constexpr int EQ_DOT = EQ_LNS + 128 * 2 * 8;     // [128][2][4] floats
constexpr int EQ_C3 = EQ_DOT + 128 * 2 * 16;     // [128][4] floats
constexpr int EQ_GT = EQ_C3 + 128 * 16;          // group tables 2 x [128] ints
static_assert(EQ_GT + 1024 <= EQ_MISC, "scratch overflows the U chunk");

__global__ void __launch_bounds__(EQ_THREADS, 1) k_equi(EquiArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* X = smem + EQ_X;
  uint8_t* U = smem + EQ_U;
  uint8_t* misc = smem + EQ_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);   // 0: coord_mlp.0 image, 1: e tile, 2: input_lin image, 3: MMA in, 4: MMA c0
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* gbf = reinterpret_cast<float*>(misc + 128);    // [192]
  float2* LNS = reinterpret_cast<float2*>(smem + EQ_LNS);
  float4* DOT = reinterpret_cast<float4*>(smem + EQ_DOT);
  float4* C3 = reinterpret_cast<float4*>(smem + EQ_C3);
  uint32_t* gt_meta = reinterpret_cast<uint32_t*>(smem + EQ_GT);
  int* gt_node = reinterpret_cast<int*>(gt_meta + 128);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int rq = warp & 3, half = warp >> 2;
  const int row = rq * 32 + lane;
  const int per = (a.p.n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, a.p.n_tiles);

  if (t == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    if (tile0 < tile1) {
      mbar_expect_tx(&bars[1], CHUNK_BYTES_A);
      bulk_g2s(U, reinterpret_cast<const uint8_t*>(a.e16) + (size_t)tile0 * CHUNK_BYTES_A, CHUNK_BYTES_A, &bars[1]);
      mbar_expect_tx(&bars[2], 65536);
      bulk_g2s(X, a.win_img, 65536, &bars[2]);
    }
    mbar_expect_tx(&bars[0], 131072);
    bulk_g2s(smem + EQ_WC0, a.wc0_img, 131072, &bars[0]);
  }
  for (int i = t; i < 192; i += EQ_THREADS) gbf[i] = a.gbf[i];
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_x = tmem, tm_c = tmem + 256;
  uint32_t par = 0;
  const float4* pos = reinterpret_cast<const float4*>(a.pos_in);
  float4* pos_out = reinterpret_cast<float4*>(a.pos_out);
  const int cb = 128 * half;                      // first hidden column of this thread

  // row metadata of a tile is fetched one tile ahead
  RowInfo rn = load_row(a.p, min(tile0, a.p.n_tiles - 1), row);
  int ngn = a.p.tile_ngroups[min(tile0, a.p.n_tiles - 1)];
  uint8_t exn = a.extra[(size_t)min(tile0, a.p.n_tiles - 1) * TILE_ROWS + row];
  const uint16_t* ab16 = static_cast<const uint16_t*>(a.AB);
  for (int tile = tile0; tile < tile1; ++tile) {
    const RowInfo r = rn;
    const int ng = ngn;
    const uint8_t ex = exn;
    const float* tr = a.tab + (size_t)r.mol * a.ld_tab + a.tab_off;
    const float4 pg = pos[r.g], pj = pos[r.j];
    // hoisted input_lin parts (fp16 rows): first 32-column chunk in flight before the MMA wait
    const uint16_t* ag = ab16 + (size_t)r.g * a.ldab + cb;
    const uint16_t* bj = ab16 + (size_t)r.j * a.ldab + D_ + cb;
    H32 uc = ldg_h32(ag), vc = ldg_h32(bj);
    {
      float df[32];
      if (r.valid) {
        gbf_eval_half(sq_dist(pg, pj), tr[tab_gbf(D_)], tr[tab_gbf(D_) + 1], gbf, half, df);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) df[i] = 0.f;
      }
      st_rowh<32>(U, row, 1, 4 * half, df);
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
      mbar_wait(&bars[1], par);
      mbar_wait(&bars[2], par);
      tc_fence_after();
      mma_tile_h(tm_x, smem_u32(U), smem_u32(X), 256, 2, false);       // input_lin edge part
      umma_commit(&bars[3]);
    }
    mbar_wait(&bars[3], par);
    tc_fence_after();
    if (t == 0 && tile + 1 < tile1) {            // U chunk 0 is consumed: prefetch the next e tile
      mbar_expect_tx(&bars[1], CHUNK_BYTES_A);
      bulk_g2s(U, reinterpret_cast<const uint8_t*>(a.e16) + (size_t)(tile + 1) * CHUNK_BYTES_A, CHUNK_BYTES_A, &bars[1]);
    }
    if (half == 0 && r.valid && row == r.gs) { gt_meta[r.gi] = (uint32_t)r.gs | ((uint32_t)r.gl << 8); gt_node[r.gi] = r.g; }

    // ---- pass 1: x = acc + A[g] + B[j] + b over this thread's 128 hidden units, kept in TMEM; row statistics
    float mean, rstd;
    {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        H32 un, vn;
        if (c < 3) { un = ldg_h32(ag + (c + 1) * 32); vn = ldg_h32(bj + (c + 1) * 32); }
        float x[32];
        tmem_ld32(tmem_addr(tm_x, cb + c * 32), x);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float uf[8], vf[8];
          unpack8(uc.u[i], uf);
          unpack8(vc.u[i], vf);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.b_in + cb + c * 32 + 8 * i));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.b_in + cb + c * 32 + 8 * i + 4));
          x[8 * i] += uf[0] + vf[0] + b0.x; x[8 * i + 1] += uf[1] + vf[1] + b0.y;
          x[8 * i + 2] += uf[2] + vf[2] + b0.z; x[8 * i + 3] += uf[3] + vf[3] + b0.w;
          x[8 * i + 4] += uf[4] + vf[4] + b1.x; x[8 * i + 5] += uf[5] + vf[5] + b1.y;
          x[8 * i + 6] += uf[6] + vf[6] + b1.z; x[8 * i + 7] += uf[7] + vf[7] + b1.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { s1 += x[i]; s2 = fmaf(x[i], x[i], s2); }
        tmem_st32(tmem_addr(tm_x, cb + c * 32), x);
        if (c < 3) { uc = un; vc = vn; }
      }
      tmem_wait_st();
      LNS[row * 2 + half] = make_float2(s1, s2);
      __syncthreads();
      const float2 o = LNS[row * 2 + (half ^ 1)];
      mean = (s1 + o.x) * (1.0f / 256.0f);
      rstd = rsqrtf(fmaxf((s2 + o.y) * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-6f);
    }
    // ---- pass 2: LN + modulate -> X (fp16, K = 256); the input_lin image there is no longer needed
    {
      const float* shift = tr + tab_equi(D_) + cb;
      const float* scale = shift + D_;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float x[32];
        tmem_ld32(tmem_addr(tm_x, cb + c * 32), x);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c * 32 + i));
          const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c * 32 + i));
          x[i] = r.valid ? fmaf((x[i] - mean) * rstd, 1.0f + sc.x, sh.x) : 0.f;
          x[i + 1] = r.valid ? fmaf((x[i + 1] - mean) * rstd, 1.0f + sc.y, sh.y) : 0.f;
          x[i + 2] = r.valid ? fmaf((x[i + 2] - mean) * rstd, 1.0f + sc.z, sh.z) : 0.f;
          x[i + 3] = r.valid ? fmaf((x[i + 3] - mean) * rstd, 1.0f + sc.w, sh.w) : 0.f;
        }
        st_rowh<32>(X, row, 2 * half + (c >> 1), 4 * (c & 1), x);
      }
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      if (tile == tile0) mbar_wait(&bars[0], 0);
      tc_fence_after();
      mma_tile_h(tm_c, smem_u32(X), smem_u32(smem + EQ_WC0), 256, 4, false);     // coord_mlp.0
      umma_commit(&bars[4]);
    }
    mbar_wait(&bars[4], par);
    tc_fence_after();
    if (t == 0 && tile + 1 < tile1) {            // X is consumed: fetch the input_lin image for the next tile
      mbar_expect_tx(&bars[2], 65536);
      bulk_g2s(X, a.win_img, 65536, &bars[2]);
    }

    // ---- SiLU, coord_mlp.2 on CUDA cores (partial dots over this thread's 128 hidden units)
    {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float x[32];
        tmem_ld32(tmem_addr(tm_c, cb + c * 32), x);
        const float* bc = a.b_c0 + cb + c * 32;
        const float* w = a.wc2 + cb + c * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bc + i));
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + i));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + 256 + i));
          const float4 w2 = __ldg(reinterpret_cast<const float4*>(w + 512 + i));
          const float s0 = silu_fast(x[i] + b4.x), s1 = silu_fast(x[i + 1] + b4.y);
          const float s2 = silu_fast(x[i + 2] + b4.z), s3 = silu_fast(x[i + 3] + b4.w);
          o0 = fmaf(s0, w0.x, o0); o0 = fmaf(s1, w0.y, o0); o0 = fmaf(s2, w0.z, o0); o0 = fmaf(s3, w0.w, o0);
          o1 = fmaf(s0, w1.x, o1); o1 = fmaf(s1, w1.y, o1); o1 = fmaf(s2, w1.z, o1); o1 = fmaf(s3, w1.w, o1);
          o2 = fmaf(s0, w2.x, o2); o2 = fmaf(s1, w2.y, o2); o2 = fmaf(s2, w2.z, o2); o2 = fmaf(s3, w2.w, o2);
        }
      }
      DOT[row * 2 + half] = make_float4(o0, o1, o2, 0.f);
    }
    __syncthreads();
    if (half == 0) {     // tanh, adjacency-weighted mean, coordinate contribution of this edge
      const float4 p = DOT[row * 2], q = DOT[row * 2 + 1];
      const float w = (tanh_fast(p.x + q.x) + ((ex & 1) ? tanh_fast(p.y + q.y) : 0.f) + ((ex & 2) ? tanh_fast(p.z + q.z) : 0.f)) *
                      (1.0f / 3.0f);
      const float dx = pg.x - pj.x, dy = pg.y - pj.y, dz = pg.z - pj.z;
      const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
      const float f = r.valid ? a.coord_scale * w / fmaxf(nrm, 1e-8f) : 0.f;
      C3[row] = make_float4(dx * f, dy * f, dz * f, 0.f);
    }
    __syncthreads();
    if (t < ng) {
      const int gs = gt_meta[t] & 255u, gl = (gt_meta[t] >> 8) & 255u;
      const int node = gt_node[t];
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int rr = gs; rr < gs + gl; ++rr) { const float4 c = C3[rr]; sx += c.x; sy += c.y; sz += c.z; }
      const float4 p0 = pos[node];
      pos_out[node] = make_float4(p0.x + sx, p0.y + sy, p0.z + sz, 0.f);
    }
    fence_async_smem();        // the scratch is overwritten by the next tile's GBF rows
    sync_tc();
    par ^= 1;
  }
  if (tile0 >= tile1 && t == 0) mbar_wait(&bars[0], 0);   // never leave with bulk copies in flight
  sync_tc();
  if (warp == 0) tmem_dealloc<512>(tmem);
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
  k_equi<<<grid, EQ_THREADS, EQ_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
