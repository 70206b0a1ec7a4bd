// Edge prediction heads on edge tiles: edge_exist_mlp and edge_type_mlp over the concatenated edge hiddens.
//
// reference models/mol_gnn.py:466-479 (two 3-layer MLPs with SiLU), :571-578 (cat of edge hiddens, cat of the
// two outputs, scatter to the dense [B,N,N,ch] tensor).  Both MLPs share their input, so layer 0 is one
// N=128 GEMM ([exist.0 ; type.0]), layer 2 one block-diagonal N=64 GEMM, layer 4 (32 -> 1 and 32 -> ch-1)
// runs on the CUDA cores.  Row (g, j) is the directed edge r=j -> c=g and is written to out[b, i_j, i_g, :].
#include "edge_common.cuh"

namespace jodo {

namespace {

constexpr int EH_KEH = 192;                       // concatenated width (64 + ce*L padded to a multiple of 32)
constexpr int EH_IN = 0;                          // 96 KB: tile image (6 chunks); later A2 (64 KB) + layer-2 image (32 KB)
constexpr int EH_W0 = 6 * CHUNK_BYTES_A;          // 96 KB: layer-0 image, N=128, K=192 (6 chunks x 16 KB)
constexpr int EH_MISC = EH_W0 + 6 * 128 * 128;
constexpr int EH_SMEM = EH_MISC + 128 + (128 + 64 + 8 * 32 + 8) * 4;
static_assert(EH_SMEM <= 232448, "shared memory budget");
constexpr int EH_W2_BYTES = 4 * 64 * 128;         // N=64, K=128 -> 32 KB

__global__ void __launch_bounds__(ET, 1) k_edge_head(EdgeHeadArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_smem_alignment(smem);
  uint8_t* IN = smem + EH_IN;
  uint8_t* W0 = smem + EH_W0;
  uint8_t* W2 = IN + 4 * CHUNK_BYTES_A;           // 64 KB into the tile region
  uint8_t* misc = smem + EH_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);   // 0: layer-0 image, 1: tile, 2: MMA, 3: layer-2 image
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* b0 = reinterpret_cast<float*>(misc + 128);     // [128]
  float* b2 = b0 + 128;                                 // [64]
  float* w4 = b2 + 64;                                  // [ch][32]
  float* b4 = w4 + 8 * 32;                              // [ch]

  const int t = threadIdx.x;
  const int ch = a.ch;
  if (t == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], 6 * 128 * 128);
    bulk_g2s(W0, a.w0_img, 6 * 128 * 128, &bars[0]);
  }
  b0[t] = a.b0[t];
  if (t < 64) b2[t] = a.b2[t];
  for (int i = t; i < ch * 32; i += ET) w4[i] = a.w4[i];
  if (t < ch) b4[t] = a.b4[t];
  if (t < 32) tmem_alloc<256>(tmem_slot);
  sync_tc();
  const uint32_t tmem = *tmem_slot;
  uint32_t par_t = 0, par_m = 0, par_w2 = 0;
  bool first = true;
  const int N = a.p.N;

  for (int tile = blockIdx.x; tile < a.p.n_tiles; tile += gridDim.x) {
    const RowInfo r = load_row(a.p, tile, t);
    if (t == 0) {
      mbar_expect_tx(&bars[1], 6 * CHUNK_BYTES_A);
      bulk_g2s(IN, reinterpret_cast<const uint8_t*>(a.eh_img) + (size_t)tile * a.eh_tile_bytes, 6 * CHUNK_BYTES_A, &bars[1]);
      if (first) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], par_t);
      tc_fence_after();
      mma_tile(tmem, smem_u32(IN), smem_u32(W0), 128, 6, false);
      umma_commit(&bars[2]);
    }
    first = false;
    par_t ^= 1;
    mbar_wait(&bars[2], par_m);
    par_m ^= 1;
    tc_fence_after();
    if (t == 0) {      // the tile region is free: fetch the layer-2 image behind the A2 area
      mbar_expect_tx(&bars[3], EH_W2_BYTES);
      bulk_g2s(W2, a.w2_img, EH_W2_BYTES, &bars[3]);
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float x[32];
      tmem_ld32(tmem_addr(tmem, c * 32), x);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = silu_f(x[i] + b0[c * 32 + i]);
      st_row32<true>(IN, t, c, x);
    }
    fence_async_smem();
    sync_tc();
    if (t == 0) {
      mbar_wait(&bars[3], par_w2);
      tc_fence_after();
      mma_tile(tmem + 128, smem_u32(IN), smem_u32(W2), 64, 4, false);
      umma_commit(&bars[2]);
    }
    par_w2 ^= 1;
    mbar_wait(&bars[2], par_m);
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
        const float s0 = silu_f(h0[i] + b2[i]);
        const float s1 = silu_f(h1[i] + b2[32 + i]);
        o[0] += s0 * w4[i];
#pragma unroll
        for (int k = 1; k < 8; ++k) if (k < ch) o[k] += s1 * w4[k * 32 + i];
      }
      if (r.valid) {
        const int dg = a.p.node_dense[r.g], dj = a.p.node_dense[r.j];
        const int b = dg / N, ig = dg - b * N, ij = dj - b * N;
        float* dst = a.out_dense + (((size_t)b * N + ij) * N + ig) * ch;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < ch) dst[k] = o[k] + b4[k];
      }
    }
    fence_async_smem();
    sync_tc();
  }
  if (t < 32) tmem_dealloc<256>(tmem);
}

}  // namespace

cudaError_t launch_edge_head(const EdgeHeadArgs& a, int num_sms, cudaStream_t st) {
  if (a.keh != EH_KEH || a.ch < 1 || a.ch > 8) return cudaErrorInvalidValue;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_head, cudaFuncAttributeMaxDynamicSharedMemorySize, EH_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int grid = a.p.n_tiles < num_sms ? a.p.n_tiles : num_sms;
  k_edge_head<<<grid, ET, EH_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
