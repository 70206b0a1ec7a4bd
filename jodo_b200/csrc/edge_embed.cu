// Model-level edge embedding on edge tiles (prologue of the DGT forward).
//
// reference models/mol_gnn.py:517-557: self-conditioning features, cond_adj_2d / cond_adj_spatial,
// dist0 = zeros if every cond distance is 0 else GBF(d0), e0 = edge_emb(cat[edge_x, cond_edge_x, dist0]).
// Rows are the unordered pairs (g, j), g < j, of the pair plan (the inputs are symmetric: edge_x[b, j, g] is read); inputs
// are gathered from the dense padded batch.
#include "edge_common.cuh"

namespace jodo {

namespace {

// any squared cond distance over real ordered pairs != 0  (batch-global branch, models/mol_gnn.py:544)
__global__ void k_dist_flag(Plan p, const float* __restrict__ cond_x, int w, int* __restrict__ flag) {
  const int R = blockIdx.x * blockDim.x + threadIdx.x;
  if (R >= p.n_tiles * TILE_ROWS) return;
  const int g = p.row_g[R];
  if (g < 0) return;
  const float* a = cond_x + (size_t)p.node_dense[g] * w;
  const float* b = cond_x + (size_t)p.node_dense[p.row_j[R]] * w;
  const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  if (dx * dx + dy * dy + dz * dz != 0.0f) atomicOr(flag, 1);
}

constexpr int EE_A = 0;                       // A operand: 3 chunks (K = 96) = 48 KB
constexpr int EE_W = 3 * CHUNK_BYTES_A;       // weight image: 3 chunks x 64 rows x 128 B = 24 KB
constexpr int EE_OUT = EE_W + 3 * 64 * 128;   // 16 KB: fp16 operand image of the embedded tile, bulk-stored to e16 and eh
constexpr int EE_MISC = EE_OUT + CHUNK_BYTES_A;
constexpr int EE_SMEM = EE_MISC + 64 + 64 + 768 + 256;

__global__ void __launch_bounds__(ET, 1) k_edge_embed(EdgeEmbedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* A = smem + EE_A;
  uint8_t* W = smem + EE_W;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + EE_MISC);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + EE_MISC + 64);
  float* gbf = reinterpret_cast<float*>(smem + EE_MISC + 128);
  float* bias = gbf + 192;

  const int t = threadIdx.x;
  if (t == 0) {
    mbar_init(bar_w, 1); mbar_init(bar_m, 1);
    fence_barrier_init();
    mbar_expect_tx(bar_w, 3 * 64 * 128);
    bulk_g2s(W, a.w_img, 3 * 64 * 128, bar_w);
  }
  for (int i = t; i < 192; i += ET) gbf[i] = a.gbf[i];
  if (t < 64) bias[t] = a.bias[t];
  if (t < 32) tmem_alloc<64>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  const int use_gbf = a.cond_x ? *a.dist_flag : 0;
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;
  const int N = a.p.N, w = 3 + a.inn, ch = a.ch;
  uint32_t par_m = 0;
  bool w_ready = false;

  for (int tile = blockIdx.x; tile < a.p.n_tiles; tile += gridDim.x) {
    const RowInfo r = load_row(a.p, tile, t);
    const int dg = a.p.node_dense[r.g], dj = a.p.node_dense[r.j];
    const int b = dg / N, ig = dg - b * N, ij = dj - b * N;
    const size_t eoff = (((size_t)b * N + ij) * N + ig) * ch;     // edge_x[b, r=j, c=g]
    float d0 = 0.f;
    bool bad = false;                                  // a non-finite input of this row (NaN isolation, jodo_b200.h)
    if (a.cond_x && r.valid) {
      const float* cg = a.cond_x + (size_t)dg * w;
      const float* cj = a.cond_x + (size_t)dj * w;
      const float dx = cj[0] - cg[0], dy = cj[1] - cg[1], dz = cj[2] - cg[2];
      d0 = dx * dx + dy * dy + dz * dz;
      if (!isfinite(d0)) { d0 = 0.f; bad = true; }
    }
    float v[64];
    if (use_gbf && r.valid) {
      const float* tr = a.tab + (size_t)(uni ? 0 : r.mol) * a.ld_tab;
      gbf_eval(d0, tr[0], tr[1], gbf, v);
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = 0.f;
    }
    st_row64<true>(A, t, 0, v);
    float s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = 0.f;
    float c0 = 0.f;
    if (r.valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {          // ch <= 8
        if (i < ch) {
          float xv = a.edge_x[eoff + i];
          float cv = a.cond_edge_x ? a.cond_edge_x[eoff + i] : 0.f;
          if (!isfinite(xv)) { xv = 0.f; bad = true; }
          if (!isfinite(cv)) { cv = 0.f; bad = true; }
          s[i] = xv;
          if (i == 0) c0 = cv;
          // second block of ch columns; indices are compile-time after unrolling both loops
#pragma unroll
          for (int k = 0; k < 16; ++k) if (k == ch + i) s[k] = cv;
        }
      }
    }
    st_row32<true>(A, t, 2, s);
    // adjacency heads: cond_adj_2d (models/mol_gnn.py:520-525), cond_adj_spatial (models/utils.py:111-119)
    uint8_t bits = 0;
    if (r.valid) {
      const bool a2d = a.cond_edge_x ? (c0 >= a.edge_th) : true;
      const bool asp = d0 <= a.spatial_cut;
      bits = (a2d ? 1 : 0) | (asp ? 2 : 0);
    }
    a.extra[(size_t)tile * TILE_ROWS + t] = bits;
    if (bad && a.mol_bad) atomicOr(a.mol_bad + r.mol, 1);

    fence_async_smem();
    sync_tc();
    if (t == 0) {
      if (!w_ready) mbar_wait(bar_w, 0);
      mma_tile(tmem, smem_u32(A), smem_u32(W), 64, 3, false);
      umma_commit(bar_m);
    }
    w_ready = true;
    mbar_wait(bar_m, par_m);
    par_m ^= 1;
    tc_fence_after();
    float e[64];
    {
      float h0[32], h1[32];
      tmem_ld32(tmem_addr(tmem, 0), h0);
      tmem_ld32(tmem_addr(tmem, 32), h1);
#pragma unroll
      for (int i = 0; i < 32; ++i) { e[i] = r.valid ? h0[i] + bias[i] : 0.f; e[32 + i] = r.valid ? h1[i] + bias[32 + i] : 0.f; }
    }
    // fp32 master copy (piece-major tile: [16 pieces][128 rows][16 B], coalesced) + fp16 operand copies (edge state and
    // edge-hidden chunk 0)
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.e32) + (size_t)tile * E_TILE_BYTES) + t;
#pragma unroll
    for (int p = 0; p < 16; ++p) dst[p * 128] = make_float4(e[4 * p], e[4 * p + 1], e[4 * p + 2], e[4 * p + 3]);
    // fp16 copies through shared memory: one image, two bulk stores (row-per-thread global stores touch 32 lines each)
    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // previous tile's stores have read OUT
    __syncthreads();
    st_rowh<64>(smem + EE_OUT, t, 0, 0, e);
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint8_t*>(a.e16) + (size_t)tile * CHUNK_BYTES_A),
                   "r"(smem_u32(smem + EE_OUT)), "r"((uint32_t)CHUNK_BYTES_A) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint8_t*>(a.eh) + (size_t)tile * a.eh_tile_bytes),
                   "r"(smem_u32(smem + EE_OUT)), "r"((uint32_t)CHUNK_BYTES_A) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (t < 32) tmem_dealloc<64>(tmem);
}

}  // namespace

cudaError_t launch_dist_flag(const EdgeEmbedArgs& a, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(a.dist_flag, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  if (!a.cond_x) return cudaSuccess;
  const int R = a.p.n_tiles * TILE_ROWS;
  k_dist_flag<<<(R + 255) / 256, 256, 0, st>>>(a.p, a.cond_x, 3 + a.inn, a.dist_flag);
  return cudaGetLastError();
}

cudaError_t launch_edge_embed(const EdgeEmbedArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_edge_embed, EE_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  const int grid = a.p.n_tiles < 2 * num_sms ? a.p.n_tiles : 2 * num_sms;
  k_edge_embed<<<grid, ET, EE_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
