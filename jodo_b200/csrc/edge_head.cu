// Edge prediction heads on edge tiles: edge_exist_mlp and edge_type_mlp over the concatenated edge hiddens.
//
// reference models/mol_gnn.py:466-479 (two 3-layer MLPs with SiLU), :571-578 (cat of edge hiddens, cat of the
// two outputs, scatter to the dense [B,N,N,ch] tensor).  Both MLPs share their input, so layer 0 is one
// N=128 GEMM ([exist.0 ; type.0]), layer 2 one block-diagonal N=64 GEMM, layer 4 (32 -> 1 and 32 -> ch-1)
// runs on the CUDA cores.  Rows are the unordered pairs (g, j) of the pair plan; each is written to out[b, i_j, i_g, :]
// and out[b, i_g, i_j, :].
// fp16 operand images; all weights resident; the next tile is bulk-copied while the current one is in the MLP.
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int EH_KEH = 192;                       // concatenated width (64 + ce*L padded to a multiple of 64)
constexpr int EH_THREADS = 256;                   // two independent groups of 4 warps, each walking its own tiles
constexpr int EH_W0 = 0;                          // 48 KB: layer-0 image, N=128, K=192 (shared)
constexpr int EH_W2 = EH_W0 + 3 * 128 * 128;      // 16 KB: layer-2 image, N=64, K=128 (shared)
constexpr int EH_GRP = EH_W2 + 2 * 64 * 128;      // per group: 48 KB tile image (3 chunks of 64 columns) + 32 KB SiLU(layer 0)
constexpr int EH_GBYTES = 5 * CHUNK_BYTES_A;
constexpr int EH_MISC = EH_GRP + 2 * EH_GBYTES;
constexpr int EH_SMEM = EH_MISC + 128 + (128 + 64 + 8 * 32 + 8) * 4;
static_assert(EH_SMEM <= 232448, "shared memory budget");

__global__ void __launch_bounds__(EH_THREADS, 1) k_edge_head(EdgeHeadArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  const int grp = threadIdx.x >> 7;
  uint8_t* IN = smem + EH_GRP + grp * EH_GBYTES;
  uint8_t* A2 = IN + 3 * CHUNK_BYTES_A;
  uint8_t* misc = smem + EH_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);   // 0: weights, 1,2: tile of group 0,1, 3,4: MMA of group 0,1
  uint64_t* bar_t = &bars[1 + grp];
  uint64_t* bar_m = &bars[3 + grp];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* b0 = reinterpret_cast<float*>(misc + 128);     // [128]
  float* b2 = b0 + 128;                                 // [64]
  float* w4 = b2 + 64;                                  // [ch][32]
  float* b4 = w4 + 8 * 32;                              // [ch]

  const int t = threadIdx.x & 127;                     // thread inside the group = tile row
  const int ch = a.ch;
  const int per = (a.p.n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per + grp;
  const int tile1 = min(blockIdx.x * per + per, a.p.n_tiles);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 3 * 128 * 128 + 2 * 64 * 128);
    bulk_g2s(smem + EH_W0, a.w0_img, 3 * 128 * 128, &bars[0]);
    bulk_g2s(smem + EH_W2, a.w2_img, 2 * 64 * 128, &bars[0]);
  }
  if (threadIdx.x < 128) b0[threadIdx.x] = a.b0[threadIdx.x];
  if (threadIdx.x < 64) b2[threadIdx.x] = a.b2[threadIdx.x];
  for (int i = threadIdx.x; i < ch * 32; i += EH_THREADS) w4[i] = a.w4[i];
  if (threadIdx.x < ch) b4[threadIdx.x] = a.b4[threadIdx.x];
  if (threadIdx.x < 32) tmem_alloc<512>(tmem_slot);
  sync_tc();
  if (t == 0 && tile0 < tile1) {
    mbar_expect_tx(bar_t, 3 * CHUNK_BYTES_A);
    bulk_g2s(IN, reinterpret_cast<const uint8_t*>(a.eh) + (size_t)tile0 * a.eh_tile_bytes, 3 * CHUNK_BYTES_A, bar_t);
  }
  const uint32_t tmem = *tmem_slot + 256u * grp;
  uint32_t par_t = 0, par_m = 0;
  const int N = a.p.N;
  auto group_sync = [&]() { tc_fence_before(); named_bar_sync(1 + grp, 128); tc_fence_after(); };

  for (int tile = tile0; tile < tile1; tile += 2) {
    const RowInfo r = load_row(a.p, tile, t);
    if (t == 0) {
      if (tile == tile0) mbar_wait(&bars[0], 0);
      mbar_wait(bar_t, par_t);
      tc_fence_after();
      mma_tile_h(tmem, smem_u32(IN), smem_u32(smem + EH_W0), 128, 3, false);
      umma_commit(bar_m);
    }
    par_t ^= 1;
    mbar_wait(bar_m, par_m);
    par_m ^= 1;
    tc_fence_after();
    if (t == 0 && tile + 2 < tile1) {      // the tile region is free: fetch this group's next one
      mbar_expect_tx(bar_t, 3 * CHUNK_BYTES_A);
      bulk_g2s(IN, reinterpret_cast<const uint8_t*>(a.eh) + (size_t)(tile + 2) * a.eh_tile_bytes, 3 * CHUNK_BYTES_A, bar_t);
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[32];
      tmem_ld32(tmem_addr(tmem, c * 32), x);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = silu_fast(x[i] + b0[c * 32 + i]);
      st_rowh<32>(A2, t, c >> 1, 4 * (c & 1), x);
    }
    fence_async_smem();
    group_sync();
    if (t == 0) {
      mma_tile_h(tmem + 128, smem_u32(A2), smem_u32(smem + EH_W2), 64, 2, false);
      umma_commit(bar_m);
    }
    mbar_wait(bar_m, par_m);
    par_m ^= 1;
    tc_fence_after();
    {
      float h0[32], h1[32];
      tmem_ld32(tmem_addr(tmem, 128), h0);       // exist hidden (32)
      tmem_ld32(tmem_addr(tmem, 160), h1);       // type hidden (32)
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float s0 = silu_fast(h0[i] + b2[i]);
        const float s1 = silu_fast(h1[i] + b2[32 + i]);
        o[0] += s0 * w4[i];
#pragma unroll
        for (int k = 1; k < 8; ++k) if (k < ch) o[k] += s1 * w4[k * 32 + i];
      }
      if (r.valid) {
        const int dg = a.p.node_dense[r.g], dj = a.p.node_dense[r.j];
        const int b = dg / N, ig = dg - b * N, ij = dj - b * N;
        // a pair row serves both orientations: 0.5 (e[i,j] + e[j,i]) of the reference (models/mol_gnn.py:579) averages two
        // values computed from identical (symmetric) features, i.e. is that value
        float* dst = a.out_dense + (((size_t)b * N + ij) * N + ig) * ch;
        float* dst2 = a.out_dense + (((size_t)b * N + ig) * N + ij) * ch;
        const bool bad = a.mol_bad != nullptr && a.mol_bad[r.mol] != 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < ch) { const float v = bad ? __int_as_float(0x7fc00000) : o[k] + b4[k]; dst[k] = v; dst2[k] = v; }
      }
    }
    group_sync();
  }
  if (threadIdx.x == 0) mbar_wait(&bars[0], 0);          // never leave with the weight copies in flight
  sync_tc();
  if (threadIdx.x < 32) tmem_dealloc<512>(*tmem_slot);
}

}  // namespace

cudaError_t launch_edge_head(const EdgeHeadArgs& a, int num_sms, cudaStream_t st) {
  if (a.keh != EH_KEH || a.ch < 1 || a.ch > 8) return cudaErrorInvalidValue;
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_edge_head, EH_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  const int grid = a.p.n_tiles < 2 * num_sms ? (a.p.n_tiles + 1) / 2 : num_sms;
  k_edge_head<<<grid, EH_THREADS, EH_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
