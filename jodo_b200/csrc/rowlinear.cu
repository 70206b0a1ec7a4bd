// Row-wise linear map on the tensor cores:  C[M, N] = epi( act_in(A[M, K]) * W[N, K]^T + bias ).
//
// Used for everything that is one dense row per atom or per molecule: the noise-level / context
// embedding MLPs and the per-molecule AdaLN tables (reference models/mol_gnn.py:534, 291-294, 78,
// models/layers.py:330), node_emb, q/k/v, the node FFN, the hoisted node parts of node2edge_lin and
// input_lin, node_l and the node head (SURVEY.md §8a "Hoists").
//
// One CTA = one 128-row x NT-column output tile.  K is streamed in 32-column chunks through a 2-stage
// shared-memory ring: the weight chunk is a pre-swizzled image copied by the TMA engine
// (cp.async.bulk), the activation chunk is loaded coalesced by the threads, transformed, rounded to
// tf32 and stored in the same swizzled layout; one thread issues tcgen05.mma (kind::tf32) into a
// TMEM accumulator; the epilogue reads TMEM with one thread per output row.
#include "common.cuh"
#include "kernels.h"

namespace jodo {

namespace {

constexpr int RL_THREADS = 128;
constexpr int RL_STAGES = 2;
constexpr int RL_A_STAGE = TILE_ROWS * 128;        // 16 KB
constexpr int RL_W_STAGE = 256 * 128;              // 32 KB (NT <= 256)
constexpr int RL_SMEM = 1024 + RL_STAGES * (RL_A_STAGE + RL_W_STAGE) + 256;

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_SILU) return silu_f(x);
  if (act == ACT_GELU) return gelu_f(x);
  return x;
}

__global__ void __launch_bounds__(RL_THREADS) k_rowlinear(RowLinearArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_buf = smem;                                   // RL_STAGES x 16 KB
  uint8_t* w_buf = smem + RL_STAGES * RL_A_STAGE;          // RL_STAGES x 32 KB
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(w_buf + RL_STAGES * RL_W_STAGE);   // [2] weights landed
  uint64_t* bar_m = bar_w + RL_STAGES;                                             // [2] MMAs of a stage done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_m + RL_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TILE_ROWS;
  const int nt = a.NT;
  const int n0 = blockIdx.y * nt;
  const int nk = a.K / 32;

  if (tid == 0) {
    for (int s = 0; s < RL_STAGES; ++s) { mbar_init(&bar_w[s], 1); mbar_init(&bar_m[s], 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // weight image of this N tile: [K/32][NT][128 B]
  const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.Wimg) + (size_t)blockIdx.y * nk * nt * 128;
  const uint32_t w_chunk_bytes = (uint32_t)nt * 128u;

  uint32_t par_w[RL_STAGES] = {0, 0}, par_m[RL_STAGES] = {0, 0};
  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc & 1;
    if (kc >= RL_STAGES) {                       // stage s is free once the MMAs of chunk kc-2 completed
      mbar_wait(&bar_m[s], par_m[s]);
      par_m[s] ^= 1;
    }
    if (tid == 0) {
      mbar_expect_tx(&bar_w[s], w_chunk_bytes);
      bulk_g2s(w_buf + s * RL_W_STAGE, wimg + (size_t)kc * w_chunk_bytes, w_chunk_bytes, &bar_w[s]);
    }
    // activation chunk: 8 lanes cover one 128-byte row segment, a warp covers 4 rows per pass
    {
      uint8_t* dst = a_buf + s * RL_A_STAGE;
      const int piece = lane & 7;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = warp * 32 + it * 4 + (lane >> 3);
        const int gr = m0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < a.M) v = *reinterpret_cast<const float4*>(a.A + (size_t)gr * a.lda + kc * 32 + piece * 4);
        v.x = to_tf32(apply_act(v.x, a.act_in));
        v.y = to_tf32(apply_act(v.y, a.act_in));
        v.z = to_tf32(apply_act(v.z, a.act_in));
        v.w = to_tf32(apply_act(v.w, a.act_in));
        *reinterpret_cast<float4*>(dst + img_piece(r, 0, piece, CHUNK_BYTES_A)) = v;
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&bar_w[s], par_w[s]);
      par_w[s] ^= 1;
      tc_fence_after();
      mma_tile(tmem, smem_u32(a_buf + s * RL_A_STAGE), smem_u32(w_buf + s * RL_W_STAGE), nt, 1, kc > 0);
      umma_commit(&bar_m[s]);
    }
  }
  // the last commit covers every earlier MMA
  {
    const int s = (nk - 1) & 1;
    mbar_wait(&bar_m[s], par_m[s]);
  }
  tc_fence_after();

  // epilogue: thread t owns output row m0 + t
  const int gr = m0 + tid;
  const bool live = gr < a.M;
  const int mol = (live && a.row_mol) ? a.row_mol[gr] : 0;
  for (int c0 = 0; c0 < nt; c0 += 32) {
    float v[32];
    if (nt - c0 >= 32) {
      tmem_ld32(tmem_addr(tmem, c0), v);
    } else {
      float h[16];
      tmem_ld16(tmem_addr(tmem, c0), h);
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i] = h[i]; v[i + 16] = 0.f; }
    }
    const int ncols = (nt - c0 >= 32) ? 32 : 16;
    if (live) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (i < ncols) {
          const int col = n0 + c0 + i;
          float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          if (a.bias) {
            const float4 b = *reinterpret_cast<const float4*>(a.bias + col);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (a.epi == EPI_ACT) {
            o.x = apply_act(o.x, a.act_out); o.y = apply_act(o.y, a.act_out);
            o.z = apply_act(o.z, a.act_out); o.w = apply_act(o.w, a.act_out);
          } else if (a.epi == EPI_ADD) {
            const float4 x = *reinterpret_cast<const float4*>(a.aux + (size_t)gr * a.ld_aux + col);
            o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
          } else if (a.epi == EPI_GATED_RES) {   // out = res + gate[mol] * (acc + bias)
            const float4 x = *reinterpret_cast<const float4*>(a.aux + (size_t)gr * a.ld_aux + col);
            const float4 g = *reinterpret_cast<const float4*>(a.gate + (size_t)mol * a.ld_gate + col);
            o.x = x.x + g.x * o.x; o.y = x.y + g.y * o.y; o.z = x.z + g.z * o.z; o.w = x.w + g.w * o.w;
          }
          *reinterpret_cast<float4*>(a.C + (size_t)gr * a.ldc + col) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace

const char* check_rowlinear(const RowLinearArgs& a) {
  if (a.M <= 0) return "rowlinear: M <= 0";
  if (a.K <= 0 || a.K % 32) return "rowlinear: K must be a positive multiple of 32";
  if (!(a.NT == 16 || a.NT == 32 || a.NT == 64 || a.NT == 128 || a.NT == 256)) return "rowlinear: NT must be 16/32/64/128/256";
  if (a.N <= 0 || a.N % a.NT) return "rowlinear: N must be a multiple of NT";
  if (a.lda % 4 || a.ldc % 4) return "rowlinear: lda/ldc must be multiples of 4";
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.C) | reinterpret_cast<uintptr_t>(a.Wimg)) & 15)
    return "rowlinear: pointers must be 16-byte aligned";
  if ((a.epi == EPI_ADD || a.epi == EPI_GATED_RES) && (!a.aux || a.ld_aux % 4)) return "rowlinear: aux missing";
  if (a.epi == EPI_GATED_RES && (!a.gate || !a.row_mol || a.ld_gate % 4)) return "rowlinear: gate/row_mol missing";
  return nullptr;
}

cudaError_t launch_rowlinear(const RowLinearArgs& a, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_rowlinear, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((a.M + TILE_ROWS - 1) / TILE_ROWS, a.N / a.NT);
  k_rowlinear<<<grid, RL_THREADS, RL_SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
