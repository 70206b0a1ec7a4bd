// Row-wise linear map on the tensor cores:  C[M, N] = epi( act_in(A[M, K]) * W[N, K]^T + bias ).
//
// Used for everything that is one dense row per atom or per molecule: the noise-level / context
// embedding MLPs and the per-molecule AdaLN tables (reference models/mol_gnn.py:534, 291-294, 78,
// models/layers.py:330), node_emb, q/k/v, the node FFN, the hoisted node parts of node2edge_lin and
// input_lin, node_l and the node head (SURVEY.md §8a "Hoists").
//
// One CTA = one 128-row x NT-column output tile, two CTAs per SM (each owns <= 256 TMEM columns and <= 96 KB of
// shared memory) so that one CTA's epilogue overlaps the other's main loop.  K is streamed in 64-column chunks
// through a 2-stage ring: the weight chunk is a pre-swizzled fp16 image copied by the TMA engine (cp.async.bulk); the
// fp32 activation chunk is loaded coalesced into registers one chunk ahead, transformed, converted to fp16
// (same mantissa as tf32, saturating) and stored in the same swizzled layout; producers signal an mbarrier instead of
// a block-wide barrier; one thread issues tcgen05.mma (kind::f16, fp32 accumulation in TMEM) and its commit frees
// the stage.  Epilogue: warp w reads TMEM lane quarter (w&3), column half (w>>2).
#include "common.cuh"
#include "kernels.h"

namespace jodo {

namespace {

constexpr int RL_THREADS = 256;
constexpr int RL_STAGES = 2;
constexpr int RL_A_STAGE = TILE_ROWS * 128;        // 16 KB: [128 rows][64 fp16]
constexpr int RL_W_STAGE = 256 * 128;              // 32 KB (NT <= 256)
constexpr int RL_SMEM = 1024 + RL_STAGES * (RL_A_STAGE + RL_W_STAGE) + 256;

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_SILU) return silu_f(x);
  if (act == ACT_GELU) return gelu_f(x);
  return x;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(RL_THREADS, 2) k_rowlinear(RowLinearArgs a) {
  if (a.skip_if_zero && *a.skip_if_zero == 0) return;   // uniform conditioning: jodo_row0_linear computes row 0, which is every row
  // declared with its alignment (not aligned by pointer arithmetic): the compiler must see a shared-memory address, or every
  // staging access below becomes a generic LD / ST instead of LDS / STS
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* a_buf = smem;                                   // RL_STAGES x 16 KB
  uint8_t* w_buf = smem + RL_STAGES * RL_A_STAGE;          // RL_STAGES x 32 KB
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(w_buf + RL_STAGES * RL_W_STAGE);   // [2] weights landed
  uint64_t* bar_a = bar_w + RL_STAGES;                                             // [2] activations stored (256 arrivals)
  uint64_t* bar_m = bar_a + RL_STAGES;                                             // [2] MMAs of a stage done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_m + RL_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * TILE_ROWS;
  const int nt = a.NT;
  const int n0 = blockIdx.y * nt;
  const int nk = a.K / 64;

  if (tid == 0) {
    for (int s = 0; s < RL_STAGES; ++s) { mbar_init(&bar_w[s], 1); mbar_init(&bar_a[s], RL_THREADS); mbar_init(&bar_m[s], 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // weight image of this N tile: [K/64][NT][128 B]
  const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.Wimg) + (size_t)blockIdx.y * nk * nt * 128;
  const uint32_t w_chunk_bytes = (uint32_t)nt * 128u;

  // activation chunk: 16 threads cover the 64 columns of one row, a pass covers 16 rows, 8 passes.  Two register sets: the
  // chunk after next is requested as soon as a set has been stored, so a chunk's global loads have a whole loop iteration
  // (one tensor-core round trip) to land instead of the few instructions between two iterations.
  const int c4 = tid & 15, r0 = tid >> 4;
  float4 va[8], vb[8];
  auto load_chunk = [&](int kc, float4 (&v)[8]) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int gr = m0 + p * 16 + r0;
      v[p] = gr < a.M ? __ldg(reinterpret_cast<const float4*>(a.A + (size_t)gr * a.lda + kc * 64) + c4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_chunk(0, va);
  if (nk > 1) load_chunk(1, vb);

  uint32_t par_m[RL_STAGES] = {0, 0}, par_w[RL_STAGES] = {0, 0}, par_a[RL_STAGES] = {0, 0};
  auto step = [&](int kc, int s, float4 (&v)[8]) {      // s = kc & 1 as a literal: the parities stay in registers
    if (kc >= RL_STAGES) {                       // stage s is free once the MMAs of chunk kc-2 completed
      mbar_wait(&bar_m[s], par_m[s]);
      par_m[s] ^= 1;
    }
    if (tid == 0) {
      mbar_expect_tx(&bar_w[s], w_chunk_bytes);
      bulk_g2s(w_buf + s * RL_W_STAGE, wimg + (size_t)kc * w_chunk_bytes, w_chunk_bytes, &bar_w[s]);
    }
    {
      uint8_t* dst = a_buf + s * RL_A_STAGE;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int r = p * 16 + r0;
        uint2 o;
        o.x = pack_h2(apply_act(v[p].x, a.act_in), apply_act(v[p].y, a.act_in));
        o.y = pack_h2(apply_act(v[p].z, a.act_in), apply_act(v[p].w, a.act_in));
        *reinterpret_cast<uint2*>(dst + img_piece(r, 0, c4 >> 1, CHUNK_BYTES_A) + ((c4 & 1) << 3)) = o;
      }
    }
    fence_async_smem();
    mbar_arrive(&bar_a[s]);
    if (kc + 2 < nk) load_chunk(kc + 2, v);
    if (tid == 0) {
      mbar_wait(&bar_a[s], par_a[s]);
      mbar_wait(&bar_w[s], par_w[s]);
      tc_fence_after();
      mma_tile_h(tmem, smem_u32(a_buf + s * RL_A_STAGE), smem_u32(w_buf + s * RL_W_STAGE), nt, 1, kc > 0);
      umma_commit(&bar_m[s]);
    }
    par_a[s] ^= 1;
    par_w[s] ^= 1;
  };
  for (int kc = 0; kc < nk; kc += 2) {
    step(kc, 0, va);
    if (kc + 1 < nk) step(kc + 1, 1, vb);
  }
  // the last commit covers every earlier MMA
  {
    const int s = (nk - 1) & 1;
    mbar_wait(&bar_m[s], par_m[s]);
  }
  tc_fence_after();

  // epilogue: warp w = TMEM lane quarter (w & 3); column half (w >> 2) when the tile is at least 64 wide
  const int half = warp >> 2;
  const int gr = m0 + (warp & 3) * 32 + (tid & 31);
  const bool live = gr < a.M;
  const int mol = (live && a.row_mol) ? a.row_mol[gr] : 0;
  const int cw = nt >= 64 ? nt / 2 : nt;                    // columns this thread handles
  const int cbeg = nt >= 64 ? half * cw : 0;
  if (nt >= 64 || half == 0) {
    for (int c0 = cbeg; c0 < cbeg + cw; c0 += 32) {
      float x[32];
      const int ncols = (cbeg + cw - c0 >= 32) ? 32 : 16;
      if (ncols == 32) {
        tmem_ld32(tmem_addr(tmem, c0), x);
      } else {
        float h[16];
        tmem_ld16(tmem_addr(tmem, c0), h);
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[i] = h[i]; x[i + 16] = 0.f; }
      }
      if (live) {
        float* C32 = static_cast<float*>(a.C);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (i < ncols) {
            const int col = n0 + c0 + i;
            float4 o = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
            if (a.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + col));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (a.epi == EPI_ACT) {
              o.x = apply_act(o.x, a.act_out); o.y = apply_act(o.y, a.act_out);
              o.z = apply_act(o.z, a.act_out); o.w = apply_act(o.w, a.act_out);
            } else if (a.epi == EPI_ADD) {
              const float4 y = *reinterpret_cast<const float4*>(a.aux + (size_t)gr * a.ld_aux + col);
              o.x += y.x; o.y += y.y; o.z += y.z; o.w += y.w;
            } else if (a.epi == EPI_GATED_RES) {   // out = res + gate[mol] * (acc + bias)
              const float4 y = *reinterpret_cast<const float4*>(a.aux + (size_t)gr * a.ld_aux + col);
              const float4 g = __ldg(reinterpret_cast<const float4*>(a.gate + (size_t)mol * a.ld_gate + col));
              o.x = fmaf(g.x, o.x, y.x); o.y = fmaf(g.y, o.y, y.y); o.z = fmaf(g.z, o.z, y.z); o.w = fmaf(g.w, o.w, y.w);
            }
            if (!a.out_f16) *reinterpret_cast<float4*>(C32 + (size_t)gr * a.ldc + col) = o;
            x[i] = o.x; x[i + 1] = o.y; x[i + 2] = o.z; x[i + 3] = o.w;
          }
        }
        if (a.out_f16) {
          uint16_t* C16 = static_cast<uint16_t*>(a.C) + (size_t)gr * a.ldc + n0 + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            if (i < ncols) {
              uint4 o;
              o.x = pack_h2(x[i], x[i + 1]); o.y = pack_h2(x[i + 2], x[i + 3]);
              o.z = pack_h2(x[i + 4], x[i + 5]); o.w = pack_h2(x[i + 6], x[i + 7]);
              *reinterpret_cast<uint4*>(C16 + i) = o;
            }
          }
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
  if (a.K <= 0 || a.K % 64) return "rowlinear: K must be a positive multiple of 64";
  if (!(a.NT == 16 || a.NT == 32 || a.NT == 64 || a.NT == 128 || a.NT == 256)) return "rowlinear: NT must be 16/32/64/128/256";
  if (a.N <= 0 || a.N % a.NT) return "rowlinear: N must be a multiple of NT";
  if (a.lda % 4 || a.ldc % (a.out_f16 ? 8 : 4)) return "rowlinear: lda/ldc must be multiples of 4 (8 for fp16 output)";
  if (a.out_f16 && a.epi != EPI_STORE) return "rowlinear: fp16 output supports EPI_STORE only";
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.C) | reinterpret_cast<uintptr_t>(a.Wimg)) & 15)
    return "rowlinear: pointers must be 16-byte aligned";
  if ((a.epi == EPI_ADD || a.epi == EPI_GATED_RES) && (!a.aux || a.ld_aux % 4)) return "rowlinear: aux missing";
  if (a.epi == EPI_GATED_RES && (!a.gate || !a.row_mol || a.ld_gate % 4)) return "rowlinear: gate/row_mol missing";
  return nullptr;
}

cudaError_t launch_rowlinear(const RowLinearArgs& a, cudaStream_t stream) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_rowlinear, RL_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  dim3 grid((a.M + TILE_ROWS - 1) / TILE_ROWS, a.N / a.NT);
  k_rowlinear<<<grid, RL_THREADS, RL_SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace jodo
