// Wide path, fused coordinate branch of one DGT block (reference MultiCondEquiUpdate.forward, models/mol_gnn.py:71-94):
//
//   x = input_lin edge part (per PAIR, a GEMM upstream) + A[row] + B[col]     (hoisted per-atom parts of input_lin)
//   inv = LN(x) (1 + scale) + shift                                            -> fp16 operand tile in SHARED memory
//   c = coord_mlp.0(inv) on the tensor core, SiLU, coord_mlp.2 as three fp32 row dots -> c3[row, 4 slot + k]
//
// The unfused sequence (k_wide_ln -> u_img -> k_imglinear) writes the LayerNorm's operand image to HBM (768 B per edge row at
// nf = 384) and reads it back once per 128-column tile of the GEMM; its LayerNorm kernel was the top kernel of the wide
// path.  Here one persistent CTA owns a 128-row tile for ALL N columns: eight LayerNorm warps gather the three fp16 rows
// of every edge row (its pair's input_lin part through row_pair, A[g], B[j]), reduce the row statistics and leave the
// modulated rows as the K-major SWIZZLE_128B A operand in shared memory; one warp streams the weight chunks through a ring
// with the TMA engine, one thread issues tcgen05.mma into N <= 384 TMEM columns, eight epilogue warps drain them.  The
// LayerNorm of tile i+1 runs under the epilogue of tile i.  The row sums over the partners of an atom stay in
// k_wide_equi_out (loose plans let a group straddle tiles).
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

namespace jodo {
namespace {

constexpr int WE_THREADS = 576;                      // warp 0 producer, 1 MMA, 2..9 epilogue, 10..17 LayerNorm
constexpr int WE_CHUNK = TILE_ROWS * 128;            // 16 KB: [128 rows][64 fp16], A chunk or weight chunk (NT = 128)
constexpr int WE_STAGES = 3;
constexpr int WE_STG_ROW = 144;
constexpr int WE_STG_BUF = TILE_ROWS * WE_STG_ROW;   // 18 KB
constexpr int WE_STG_BYTES = 2 * 2 * WE_STG_BUF;     // 2 column-half teams x 2 buffers

__device__ __forceinline__ void we_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void we_unpack8(const uint4 u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 we_pack8(const float* v) {
  uint4 o;
  o.x = pack_h2(v[0], v[1]); o.y = pack_h2(v[2], v[3]); o.z = pack_h2(v[4], v[5]); o.w = pack_h2(v[6], v[7]);
  return o;
}

// KC = D / 64 K chunks, NTN = D / 128 column tiles (D = 256 or 384)
template <int KC>
__global__ void __launch_bounds__(WE_THREADS, 1) k_wide_equi(WideEquiArgs a) {
  constexpr int D = 64 * KC, NTN = D / 128, PH = D / 16;       // PH = 16-byte pieces per half row
  constexpr int BT = PH % 6 == 0 ? 6 : 4;                      // pieces per load batch of the LayerNorm threads
  // declared with its alignment (not aligned by pointer arithmetic): the compiler must see a shared-memory address, or every
  // staging access below becomes a generic LD / ST instead of LDS / STS
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* A = smem;                                            // [KC][128][128 B]
  uint8_t* ring = A + KC * WE_CHUNK;                            // WE_STAGES weight chunks
  uint8_t* stg_base = ring + WE_STAGES * WE_CHUNK;
  uint64_t* bar_wfull = reinterpret_cast<uint64_t*>(stg_base + WE_STG_BYTES);
  uint64_t* bar_wempty = bar_wfull + WE_STAGES;
  uint64_t* bar_afull = bar_wempty + WE_STAGES;                 // 256 arrivals: the LayerNorm threads
  uint64_t* bar_aempty = bar_afull + 1;                         // the tile's MMAs have read A
  uint64_t* bar_tfull = bar_aempty + 1;                         // accumulators complete
  uint64_t* bar_tempty = bar_tfull + 1;                         // 8 arrivals: the epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 1);
  float* mods = reinterpret_cast<float*>(stg_base + WE_STG_BYTES + 256);     // [2 D]: shift | 1 + scale of the tile's first molecule

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = (a.M + TILE_ROWS - 1) / TILE_ROWS;
  if (tid == 0) {
    for (int s = 0; s < WE_STAGES; ++s) { mbar_init(&bar_wfull[s], 1); mbar_init(&bar_wempty[s], 1); }
    mbar_init(bar_afull, 256); mbar_init(bar_aempty, 1); mbar_init(bar_tfull, 1); mbar_init(bar_tempty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer: NTN x KC chunks per tile
    if (lane == 0) {
      const uint8_t* W = static_cast<const uint8_t*>(a.Wimg);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
        for (int c = 0; c < NTN * KC; ++c, ++it) {
          const int s = it % WE_STAGES;
          mbar_wait(&bar_wempty[s], ((it / WE_STAGES) & 1u) ^ 1u);
          mbar_expect_tx(&bar_wfull[s], WE_CHUNK);
          bulk_g2s(ring + s * WE_CHUNK, W + (size_t)c * WE_CHUNK, WE_CHUNK, &bar_wfull[s]);
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
        mbar_wait(bar_tempty, (ti & 1u) ^ 1u);                 // the previous tile's accumulators are drained
        mbar_wait(bar_afull, ti & 1u);                         // this tile's operand rows are in shared memory
        tc_fence_after();
        for (int nt = 0; nt < NTN; ++nt)
          for (int kc = 0; kc < KC; ++kc, ++it) {
            const int s = it % WE_STAGES;
            mbar_wait(&bar_wfull[s], (it / WE_STAGES) & 1u);
            tc_fence_after();
            const uint32_t sa = smem_u32(A + kc * WE_CHUNK), sw = smem_u32(ring + s * WE_CHUNK);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem + 128u * nt, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sw + kk * 32), idesc, (kc | kk) ? 1u : 0u);
            umma_commit(&bar_wempty[s]);
          }
        umma_commit(bar_aempty);
        umma_commit(bar_tfull);
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue: bias, SiLU, three row dots per row
    const int ew = warp - 2;
    const int rq = warp & 3;
    const int team = ew >> 2;
    const int wt = ((rq - 2) & 3);
    const int row = rq * 32 + lane;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    uint8_t* stg = stg_base + team * 2 * WE_STG_BUF;
    const int r0 = wt * 4 + rsub;
    uint32_t ti = 0, sb = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      float dacc[8][3];
#pragma unroll
      for (int it = 0; it < 8; ++it) dacc[it][0] = dacc[it][1] = dacc[it][2] = 0.f;
      mbar_wait(bar_tfull, ti & 1u);
      tc_fence_after();
      for (int nt = 0; nt < NTN; ++nt)
        for (int c0 = team * 64; c0 < (team + 1) * 64; c0 += 32, sb ^= 1u) {
          uint8_t* buf = stg + sb * WE_STG_BUF;
          const int col = nt * 128 + c0 + c4;
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + col));
          const float4 dw0 = __ldg(reinterpret_cast<const float4*>(a.dot_w + col));
          const float4 dw1 = __ldg(reinterpret_cast<const float4*>(a.dot_w + D + col));
          const float4 dw2 = __ldg(reinterpret_cast<const float4*>(a.dot_w + 2 * D + col));
          {
            float x[32];
            tmem_ld32(tmem + ((uint32_t)rq << 21) + (uint32_t)(nt * 128 + c0), x);
#pragma unroll
            for (int p = 0; p < 8; ++p)
              *reinterpret_cast<float4*>(buf + row * WE_STG_ROW + p * 16) = make_float4(x[4 * p], x[4 * p + 1], x[4 * p + 2], x[4 * p + 3]);
          }
          named_bar_sync(1 + team, 128);
          const uint8_t* src = buf + r0 * WE_STG_ROW + c4 * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float4 o = *reinterpret_cast<const float4*>(src + it * 16 * WE_STG_ROW);
            o.x = silu_fast(o.x + b.x); o.y = silu_fast(o.y + b.y); o.z = silu_fast(o.z + b.z); o.w = silu_fast(o.w + b.w);
            dacc[it][0] = fmaf(o.x, dw0.x, fmaf(o.y, dw0.y, fmaf(o.z, dw0.z, fmaf(o.w, dw0.w, dacc[it][0]))));
            dacc[it][1] = fmaf(o.x, dw1.x, fmaf(o.y, dw1.y, fmaf(o.z, dw1.z, fmaf(o.w, dw1.w, dacc[it][1]))));
            dacc[it][2] = fmaf(o.x, dw2.x, fmaf(o.y, dw2.y, fmaf(o.z, dw2.z, fmaf(o.w, dw2.w, dacc[it][2]))));
          }
        }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) we_arrive(bar_tempty);
      // halving butterfly over the 8 lanes of a row group (see imglinear.cu): lane s ends with the three sums of row s
      float w12[12], w6[6], w3[3];
      {
        const bool hi = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const float lo_v = dacc[i / 3][i % 3], hi_v = dacc[4 + i / 3][i % 3];
          w12[i] = (hi ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, hi ? lo_v : hi_v, 4);
        }
      }
      {
        const bool hi = (lane & 2) != 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) w6[i] = (hi ? w12[6 + i] : w12[i]) + __shfl_xor_sync(0xffffffffu, hi ? w12[i] : w12[6 + i], 2);
      }
      {
        const bool hi = (lane & 1) != 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) w3[i] = (hi ? w6[3 + i] : w6[i]) + __shfl_xor_sync(0xffffffffu, hi ? w6[i] : w6[3 + i], 1);
      }
      const int gr = tile * TILE_ROWS + r0 + 16 * (lane & 7);
      if (gr < a.M) *reinterpret_cast<float4*>(a.out + (size_t)gr * a.ld_out + 4 * team) = make_float4(w3[0], w3[1], w3[2], 0.f);
    }
  } else {
    // ------------------------------------------------------------------ LayerNorm producers: two threads per row
    const int tl = tid - 320;
    const int row = tl >> 1, half = tl & 1;
    const uint16_t* U = static_cast<const uint16_t*>(a.U);
    const uint16_t* AB = static_cast<const uint16_t*>(a.AB);
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      const int R = tile * TILE_ROWS + row;
      const int g = R < a.M ? __ldg(a.row_g + R) : -1;
      const bool valid = g >= 0;
      const int j = valid ? __ldg(a.row_j + R) : 0;
      const int pr = valid ? (a.xi ? __ldg(a.xi + R) : R) : 0;
      const int mol = valid ? __ldg(a.row_mol + R) : 0;
      const uint4* pu = reinterpret_cast<const uint4*>(U + (size_t)pr * a.ldu) + half * PH;
      const uint4* pa = reinterpret_cast<const uint4*>(AB + (size_t)(valid ? g : 0) * a.ldab) + half * PH;
      const uint4* pb = reinterpret_cast<const uint4*>(AB + (size_t)j * a.ldab + D) + half * PH;
      // the modulation rows (shift | 1 + scale) of the tile's first molecule, staged once per tile: most rows of a tile belong
      // to it (a molecule of n atoms owns n (n - 1) consecutive rows); rows of other molecules read the table in HBM / L2
      const int R0 = min(tile * TILE_ROWS, a.M - 1);
      int m0 = 0;
      {
        const int g0 = __ldg(a.row_g + R0);
        m0 = g0 >= 0 ? __ldg(a.row_mol + R0) : 0;
      }
      named_bar_sync(3, 256);                                  // the previous tile's pass 2 has read the staged rows
      {
        const float* t0 = a.tab + (size_t)m0 * a.ld_tab;
        for (int c = tl; c < 2 * D; c += 256) mods[c] = c < D ? __ldg(t0 + a.off_shift + c) : __ldg(t0 + a.off_scale + c - D);
      }
      mbar_wait(bar_aempty, (ti & 1u) ^ 1u);                   // the previous tile's MMAs have read the operand
      // pass 1: v = u + a + b -> fp16 rows in the operand tile, row statistics in fp32 (3 BT 16-byte loads in flight per thread)
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int i0 = 0; i0 < PH; i0 += BT) {
        uint4 xu[BT], xa[BT], xb[BT];
#pragma unroll
        for (int i = 0; i < BT; ++i) { xu[i] = __ldg(pu + i0 + i); xa[i] = __ldg(pa + i0 + i); xb[i] = __ldg(pb + i0 + i); }
#pragma unroll
        for (int i = 0; i < BT; ++i) {
          float u[8], va[8], vb[8];
          we_unpack8(xu[i], u); we_unpack8(xa[i], va); we_unpack8(xb[i], vb);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            u[e] = valid ? u[e] + (va[e] + vb[e]) : 0.f;
            s1 += u[e];
            s2 = fmaf(u[e], u[e], s2);
          }
          const int gp = half * PH + i0 + i;                  // global 16-byte piece of the row
          *reinterpret_cast<uint4*>(A + (gp >> 3) * WE_CHUNK + row * 128 + (((gp & 7) ^ (row & 7)) << 4)) = we_pack8(u);
        }
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      const float mean = s1 * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(s2 * (1.0f / D) - mean * mean, 0.f) + 1e-6f);
      named_bar_sync(3, 256);                                  // staged modulation rows visible
      // pass 2: normalise + modulate in place (the table stores 1 + scale)
      const bool own = mol == m0;
      const float* t = a.tab + (size_t)mol * a.ld_tab;
#pragma unroll 2
      for (int i = 0; i < PH; ++i) {
        const int gp = half * PH + i;
        uint4* p = reinterpret_cast<uint4*>(A + (gp >> 3) * WE_CHUNK + row * 128 + (((gp & 7) ^ (row & 7)) << 4));
        float v[8];
        we_unpack8(*p, v);
        float4 sh0, sh1, sc0, sc1;
        if (own) {
          sh0 = *reinterpret_cast<const float4*>(mods + 8 * gp); sh1 = *reinterpret_cast<const float4*>(mods + 8 * gp + 4);
          sc0 = *reinterpret_cast<const float4*>(mods + D + 8 * gp); sc1 = *reinterpret_cast<const float4*>(mods + D + 8 * gp + 4);
        } else {
          sh0 = __ldg(reinterpret_cast<const float4*>(t + a.off_shift + 8 * gp)); sh1 = __ldg(reinterpret_cast<const float4*>(t + a.off_shift + 8 * gp + 4));
          sc0 = __ldg(reinterpret_cast<const float4*>(t + a.off_scale + 8 * gp)); sc1 = __ldg(reinterpret_cast<const float4*>(t + a.off_scale + 8 * gp + 4));
        }
        const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = valid ? fmaf((v[e] - mean) * rstd, sc[e], sh[e]) : 0.f;
        *p = we_pack8(v);
      }
      fence_async_smem();                                      // the tensor core reads these rows through the async proxy
      we_arrive(bar_afull);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

constexpr int we_smem(int KC) { return KC * WE_CHUNK + WE_STAGES * WE_CHUNK + WE_STG_BYTES + 256 + 2 * 64 * KC * 4; }
static_assert(we_smem(6) <= 232448, "shared memory budget");

template <int KC>
cudaError_t launch_we(const WideEquiArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e = ensure_dyn_smem(k_wide_equi<KC>, we_smem(KC), attr);
  if (e != cudaSuccess) return e;
  const int tiles = (a.M + TILE_ROWS - 1) / TILE_ROWS;
  k_wide_equi<KC><<<tiles < num_sms ? tiles : num_sms, WE_THREADS, we_smem(KC), st>>>(a);
  return cudaGetLastError();
}

}  // namespace

const char* check_wide_equi(const WideEquiArgs& a) {
  if (a.M <= 0) return "wide_equi: M <= 0";
  if (a.D != 256 && a.D != 384) return "wide_equi: built for D = 256 and D = 384";
  if (!a.U || !a.AB || !a.row_g || !a.row_j || !a.row_mol || !a.tab || !a.Wimg || !a.bias || !a.dot_w || !a.out) return "wide_equi: null buffer";
  if ((a.ldu % 8) || a.ldu < a.D || (a.ldab % 8) || a.ldab < 2 * a.D || (a.ld_tab % 4) || (a.off_shift % 4) || (a.off_scale % 4) || (a.ld_out % 4) || a.ld_out < 8)
    return "wide_equi: bad strides";
  if ((reinterpret_cast<uintptr_t>(a.U) | reinterpret_cast<uintptr_t>(a.AB) | reinterpret_cast<uintptr_t>(a.tab) | reinterpret_cast<uintptr_t>(a.bias) |
       reinterpret_cast<uintptr_t>(a.dot_w) | reinterpret_cast<uintptr_t>(a.out)) & 15)
    return "wide_equi: pointers must be 16-byte aligned";
  if (reinterpret_cast<uintptr_t>(a.Wimg) & 127) return "wide_equi: weight image must be 128-byte aligned";
  return nullptr;
}

cudaError_t launch_wide_equi(const WideEquiArgs& a, int num_sms, cudaStream_t st) {
  return a.D == 384 ? launch_we<6>(a, num_sms, st) : launch_we<4>(a, num_sms, st);
}

}  // namespace jodo
