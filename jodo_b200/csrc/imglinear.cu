// Persistent tensor-core row-linear on fp16 operand images:  C[M, N] = epi( A[M, K] * W[N, K]^T + bias ).
//
// This is the per-atom GEMM of every DGT block (reference models/layers.py:147-149 lin_query/key/value,
// models/mol_gnn.py:262-264,309-311 node FFN, :304-305 hoisted node2edge_lin, :73,79 hoisted input_lin parts,
// :567 node_i): the activations already live in HBM as fp16 operand images (written by jodo_ln_mod_img or by the
// epilogue of the previous GEMM), so both operands reach shared memory through the TMA engine with no register
// staging, and the kernel is a warp-specialised pipeline:
//
//   warp 0      producer: cp.async.bulk of the (A chunk, W chunk) pair of every K step into a ring of stages
//   warp 1      MMA issuer: tcgen05.mma kind::f16 into one of two TMEM accumulators; tcgen05.commit frees the stage
//   warps 2..9  epilogue: TMEM -> registers -> bias / activation / gated residual -> fp32 rows, fp16 rows and/or
//               the fp16 operand image of the next GEMM; then the accumulator is handed back to the MMA warp
//
// A CTA walks work units (128-row tile, NT-column tile) with the column tile fastest, so the CTAs that share an A
// tile run at the same time and re-read it from L2.  The epilogue of unit i overlaps the main loop of unit i+1.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace jodo {

// fp16 operands written by this kernel that exceeded +-65504 and were clamped (per-atom operands the edge kernels gather:
// q | k | v, the hoisted input_lin / node2edge_lin parts, activation images).  Read through jodo_saturation_count.
__device__ unsigned int g_sat_imglinear;

namespace {

// Epilogue warps: 8 (two column-half teams, double-buffered staging) or 16 (four column-quarter teams, one staging
// buffer each -- the same shared memory; NT >= 128).  The GEMMs here are short in K, so a 128 x NT tile's epilogue
// (activation, packing, stores) takes longer than its MMAs: with 16 warps it keeps up (k_imglinear<MODE, ACT, 16>).
constexpr int il_threads(int ew) { return 64 + 32 * ew; }
constexpr int IL_A_STAGE = TILE_ROWS * 128;          // 16 KB: [128 rows][64 fp16]
constexpr int IL_RING_BYTES = 147456;                // 144 KB of stages (3 x 48 KB at NT = 256)
constexpr int IL_MAX_STAGES = 6;
constexpr int IL_STG_ROW = 144;                      // staging row: 32 fp32 + 16 bytes of padding (conflict-free both ways)
constexpr int IL_STG_BUF = TILE_ROWS * IL_STG_ROW;   // 18 KB
constexpr int IL_STG_BYTES = 2 * 2 * IL_STG_BUF;     // 2 column-half teams x 2 buffers, or 4 column-quarter teams x 1
constexpr int IL_SMEM = 1024 + IL_RING_BYTES + IL_STG_BYTES + 512;
static_assert(IL_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == ACT_SILU) return silu_fast(x);
  if (act == ACT_GELU) return gelu_f(x);
  if (act == ACT_TANH) return tanh_fast(x);
  return x;
}

// MODE < 0: every epilogue option is a run-time flag.  MODE >= 0 fixes them at compile time (the four combinations
// a DGT block uses), which removes the predicated paths from the epilogue's dependent instruction stream:
//   bits [0,2) epilogue | bit 2 fp32 rows | bit 3 fp16 rows | bit 4 fp16 image | bit 5 second image | bit 6 fused row dots
constexpr int il_mode(int epi, bool c32, bool c16, bool cimg, bool cimg2 = false, bool dot = false) {
  return epi | (c32 ? 4 : 0) | (c16 ? 8 : 0) | (cimg ? 16 : 0) | (cimg2 ? 32 : 0) | (dot ? 64 : 0);
}
// A mode of its own: JODO_EPI_STORE with ONLY the piece-major fp16 output (q | k | v and the hoisted parts the edge kernels
// gather).  In that layout the 16 bytes of (piece, row) are contiguous over rows, which is the accumulator's own layout
// (lane = row): every thread packs its 32 columns into four pieces and stores them directly -- consecutive lanes write
// consecutive 16-byte slots, no staging through shared memory, no team barrier.  (The row-major pass wrote 16 sectors of
// 16 useful bytes per store instruction here.)
constexpr int IL_MODE_PM16 = 128;
// JODO_EPI_LN_MOD: LayerNorm + modulation of the accumulator row, written as the fp16 operand image of the next GEMM (16
// epilogue warps, N = NT = 128: warp = lane quarter x 32-column quarter, one chunk per thread; the row statistics cross
// the four column quarters through shared memory and one barrier of the 16 warps).
constexpr int IL_MODE_LN = 256;

template <int MODE, int ACT = ACT_SILU, int EW = 8>
__global__ void __launch_bounds__(il_threads(EW), 1) k_imglinear(ImgLinearArgs a) {
  constexpr int NTEAM = EW / 4, NBUF = EW == 8 ? 2 : 1;
  if (a.skip_if_zero && *a.skip_if_zero == 0) return;     // uniform conditioning: nothing to do (warp-uniform, before any setup)
  // declared with its alignment (not aligned by pointer arithmetic): the compiler must see a shared-memory address, or every
  // staging access below becomes a generic LD / ST instead of LDS / STS
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int nt = a.NT;
  const int stage_bytes = IL_A_STAGE + nt * 128;
  int stages = IL_RING_BYTES / stage_bytes;
  if (stages > IL_MAX_STAGES) stages = IL_MAX_STAGES;
  uint8_t* stg_base = smem + IL_RING_BYTES;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + IL_RING_BYTES + IL_STG_BYTES);   // operands landed
  uint64_t* bar_empty = bar_full + IL_MAX_STAGES;                              // MMAs of the stage done
  uint64_t* bar_tfull = bar_empty + IL_MAX_STAGES;                             // [2] accumulator complete
  uint64_t* bar_tempty = bar_tfull + 2;                                        // [2] accumulator drained (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = a.K / 64;
  const int ntn = a.N / nt;
  const int mt = (a.M + TILE_ROWS - 1) / TILE_ROWS;
  const int units = mt * ntn;

  if (tid == 0) {
    for (int s = 0; s < IL_MAX_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_tfull[s], 1); mbar_init(&bar_tempty[s], EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const uint8_t* Aimg = static_cast<const uint8_t*>(a.Aimg);
      const uint8_t* Wimg = static_cast<const uint8_t*>(a.Wimg);
      const uint32_t wbytes = (uint32_t)nt * 128u;
      uint32_t it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int m = u / ntn, n = u - m * ntn;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1u;
          mbar_wait(&bar_empty[s], ph ^ 1u);
          uint8_t* dst = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&bar_full[s], IL_A_STAGE + wbytes);
          bulk_g2s(dst, Aimg + ((size_t)m * nk + kc) * IL_A_STAGE, IL_A_STAGE, &bar_full[s]);
          bulk_g2s(dst + IL_A_STAGE, Wimg + ((size_t)n * nk + kc) * wbytes, wbytes, &bar_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(nt);
      uint32_t it = 0, ai = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ai) {
        const uint32_t ab = ai & 1u, aph = (ai >> 1) & 1u;
        mbar_wait(&bar_tempty[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t td = tmem + ab * 256u;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1u;
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sw = sa + IL_A_STAGE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(td, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sw + kk * 32), idesc, (kc | kk) ? 1u : 0u);
          umma_commit(&bar_empty[s]);
        }
        umma_commit(&bar_tfull[ab]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps = 2 column-half teams)
    // Each team drains its half of the accumulator 32 columns at a time: TMEM -> registers (thread = row) -> padded
    // fp32 staging rows in shared memory -> team barrier -> "row-major" pass in which 8 lanes cover the 128 bytes of
    // one row, so that bias / residual / gate loads and every store are coalesced.
    const int ew = warp - 2;
    const int rq = warp & 3;                       // TMEM lane quarter this warp may read
    const int team = ew >> 2;                      // column half (quarter with 16 warps)
    const int wt = ((rq - 2) & 3);                 // warp index inside the team (0..3), any bijection works
    const int row = rq * 32 + lane;
    const int cw = nt / NTEAM;                     // columns per team (32, 64 or 128)
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const int epi = MODE < 0 ? a.epi : ((MODE == IL_MODE_PM16 || MODE == IL_MODE_LN) ? EPI_STORE : (MODE & 3));
    const bool has32 = MODE < 0 ? a.C32 != nullptr : (MODE & 4) != 0;
    const bool has16 = MODE < 0 ? a.C16 != nullptr : (MODE & 8) != 0;
    const bool hasimg = MODE < 0 ? a.Cimg != nullptr : (MODE & 16) != 0;
    const bool hasimg2 = MODE < 0 ? a.Cimg2 != nullptr : (MODE & 32) != 0;
    const bool hasdot = MODE < 0 ? a.dot_out != nullptr : (MODE & 64) != 0;
    const int act = a.act_out;
    const float* __restrict__ bias = a.bias;
    const float* __restrict__ aux = a.aux;
    const float* __restrict__ gate = a.gate;
    float* __restrict__ C32 = a.C32;
    uint16_t* __restrict__ C16 = static_cast<uint16_t*>(a.C16);
    uint8_t* __restrict__ Cimg = static_cast<uint8_t*>(a.Cimg);
    uint8_t* __restrict__ Cimg2 = static_cast<uint8_t*>(a.Cimg2);
    const int ld_aux = a.ld_aux, ld_gate = a.ld_gate, ldc32 = a.ldc32, ldc16 = a.ldc16, Mrows = a.M;
    // placement of the image outputs (normalised by the launcher: k = destination columns, col0 = first column, ncols = limit)
    const int nco1 = a.cimg_k / 64, c01 = a.cimg_col0, nc1 = a.cimg_ncols;
    const int nco2 = a.cimg2_k / 64, c02 = a.cimg2_col0, nc2 = a.cimg2_ncols;
    const float* __restrict__ dot_w = a.dot_w;
    float* __restrict__ dot_out = a.dot_out;
    const bool c16_pm = a.c16_piece_major != 0;
    const bool gate_row0 = a.nonuni != nullptr && *a.nonuni == 0;       // uniform conditioning: every molecule's gate row is row 0
    uint8_t* stg = stg_base + team * NBUF * IL_STG_BUF;
    const int r0 = wt * 4 + rsub;                  // this thread's rows in the row-major pass: r0 + 16 * it
    uint32_t ai = 0, sb = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ai) {
      const int m = u / ntn, n = u - m * ntn;
      const uint32_t ab = ai & 1u, aph = (ai >> 1) & 1u;
      const int gr0 = m * TILE_ROWS + r0;
      int mol[8];
      if (epi == EPI_GATED_RES) {
#pragma unroll
        for (int it = 0; it < 8; ++it) mol[it] = (!gate_row0 && (gr0 + 16 * it) < Mrows) ? __ldg(a.row_mol + gr0 + 16 * it) : 0;
      }
      float dacc[8][3];                              // fused row dots: this thread's 8 rows x 3 outputs over the team's columns
      if (hasdot) {
#pragma unroll
        for (int it = 0; it < 8; ++it) dacc[it][0] = dacc[it][1] = dacc[it][2] = 0.f;
      }
      mbar_wait(&bar_tfull[ab], aph);
      tc_fence_after();
      if (MODE == IL_MODE_LN) {
        const int gr = m * TILE_ROWS + row;
        const int c0 = team * 32, W = a.ln_cols;
        float x[32];
        tmem_ld32(tmem + ab * 256u + ((uint32_t)rq << 21) + (uint32_t)c0, x);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&bar_tempty[ab]);     // the accumulator is in registers: the next unit's MMAs may start
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + c0) + p);
          x[4 * p] += bv.x; x[4 * p + 1] += bv.y; x[4 * p + 2] += bv.z; x[4 * p + 3] += bv.w;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (c0 + 4 * p + i < W) { s += x[4 * p + i]; q = fmaf(x[4 * p + i], x[4 * p + i], q); }
          }
        }
        // partial sums of the four column quarters; two buffers by unit parity (a warp may be one unit ahead of another)
        float2* lns = reinterpret_cast<float2*>(stg_base) + (ai & 1u) * (TILE_ROWS * 4);
        lns[row * 4 + team] = make_float2(s, q);
        named_bar_sync(1, 32 * EW);
        const float4 p01 = *reinterpret_cast<const float4*>(&lns[row * 4]);
        const float4 p23 = *reinterpret_cast<const float4*>(&lns[row * 4 + 2]);
        const float inv_w = 1.0f / (float)W;
        const float mean = (p01.x + p01.z + p23.x + p23.z) * inv_w;
        const float rstd = rsqrtf(fmaxf((p01.y + p01.w + p23.y + p23.w) * inv_w - mean * mean, 0.f) + 1e-6f);
        const float nmr = -mean * rstd;
        const bool live = gr < Mrows && (!a.ln_valid || __ldg(a.ln_valid + gr) >= 0);
        const int mol_ = (live && !gate_row0) ? __ldg(a.row_mol + gr) : 0;
        const float* t = gate + (size_t)mol_ * ld_gate;
        uint8_t* dst = Cimg + (size_t)m * 2 * IL_A_STAGE + (size_t)(c0 >> 6) * IL_A_STAGE;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int col = c0 + 8 * p;
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (live && col < W) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(t + a.ln_off_scale + col));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(t + a.ln_off_scale + col + 4));
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(t + a.ln_off_shift + col));
            const float4 h1 = __ldg(reinterpret_cast<const float4*>(t + a.ln_off_shift + col + 4));
            const float* v = &x[8 * p];
            o.x = pack_h2(fmaf(fmaf(v[0], rstd, nmr), s0.x, h0.x), fmaf(fmaf(v[1], rstd, nmr), s0.y, h0.y));
            o.y = pack_h2(fmaf(fmaf(v[2], rstd, nmr), s0.z, h0.z), fmaf(fmaf(v[3], rstd, nmr), s0.w, h0.w));
            o.z = pack_h2(fmaf(fmaf(v[4], rstd, nmr), s1.x, h1.x), fmaf(fmaf(v[5], rstd, nmr), s1.y, h1.y));
            o.w = pack_h2(fmaf(fmaf(v[6], rstd, nmr), s1.z, h1.z), fmaf(fmaf(v[7], rstd, nmr), s1.w, h1.w));
          }
          *reinterpret_cast<uint4*>(dst + img_piece(row, 0, ((col & 63) >> 3), IL_A_STAGE)) = o;
        }
        continue;
      }
      if (MODE == IL_MODE_PM16) {
        const int gr = m * TILE_ROWS + row;
        for (int c0 = team * cw; c0 < (team + 1) * cw; c0 += 32) {
          const int colb = n * nt + c0;
          float x[32];
          tmem_ld32(tmem + ab * 256u + ((uint32_t)rq << 21) + (uint32_t)c0, x);
          if (bias) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + colb) + p);
              x[4 * p] += b.x; x[4 * p + 1] += b.y; x[4 * p + 2] += b.z; x[4 * p + 3] += b.w;
            }
          }
          float mx = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fabsf(x[i]));
          if (mx > 65504.f) atomicAdd(&g_sat_imglinear, 1u);
          if (gr < Mrows) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              uint4 o;
              o.x = pack_h2(x[8 * p], x[8 * p + 1]); o.y = pack_h2(x[8 * p + 2], x[8 * p + 3]);
              o.z = pack_h2(x[8 * p + 4], x[8 * p + 5]); o.w = pack_h2(x[8 * p + 6], x[8 * p + 7]);
              *reinterpret_cast<uint4*>(C16 + ((size_t)((colb >> 3) + p) * ldc16 + gr) * 8) = o;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&bar_tempty[ab]);
        continue;
      }
      for (int c0 = team * cw; c0 < (team + 1) * cw; c0 += 32, sb = (NBUF == 2 ? sb ^ 1u : 0u)) {
        uint8_t* buf = stg + sb * IL_STG_BUF;
        const int col = n * nt + c0 + c4;
        // loads that do not depend on the accumulator go first
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + col));
        float4 dw0, dw1, dw2;
        if (hasdot) {
          dw0 = __ldg(reinterpret_cast<const float4*>(dot_w + col));
          dw1 = __ldg(reinterpret_cast<const float4*>(dot_w + a.N + col));
          dw2 = __ldg(reinterpret_cast<const float4*>(dot_w + 2 * a.N + col));
        }
        // residual / gate rows: loaded ahead of the accumulator with 8 warps; with 16 warps (96 registers) inside the row loop
        constexpr bool PREFETCH = EW == 8;
        float4 y[PREFETCH ? 8 : 1], g[PREFETCH ? 8 : 1];
        if (PREFETCH && epi == EPI_GATED_RES) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int gr = gr0 + 16 * it;
            if (gr < Mrows) {
              y[it] = *reinterpret_cast<const float4*>(aux + (size_t)gr * ld_aux + col);
              g[it] = __ldg(reinterpret_cast<const float4*>(gate + (size_t)mol[it] * ld_gate + col));
            } else {
              y[it] = g[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
        {
          float x[32];
          tmem_ld32(tmem + ab * 256u + ((uint32_t)rq << 21) + (uint32_t)c0, x);
          if (NBUF == 1) named_bar_sync(1 + team, 128);   // one buffer per team: the row-major pass of the previous chunk is done
#pragma unroll
          for (int p = 0; p < 8; ++p)
            *reinterpret_cast<float4*>(buf + row * IL_STG_ROW + p * 16) = make_float4(x[4 * p], x[4 * p + 1], x[4 * p + 2], x[4 * p + 3]);
        }
        named_bar_sync(1 + team, 128);             // staging buffer sb complete (the other buffer may still be read)
        const uint8_t* src = buf + r0 * IL_STG_ROW + c4 * 4;
        const int colA = col + c01, colB = col + c02;
        uint8_t* img = (hasimg && col < nc1) ? Cimg + ((size_t)m * nco1 + (colA >> 6)) * IL_A_STAGE + ((colA & 4) << 1) : nullptr;
        uint8_t* img2 = (hasimg2 && col < nc2) ? Cimg2 + ((size_t)m * nco2 + (colB >> 6)) * IL_A_STAGE + ((colB & 4) << 1) : nullptr;
        const int piece = (colA & 63) >> 3, piece2 = (colB & 63) >> 3;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = r0 + 16 * it;
          const int gr = gr0 + 16 * it;
          const bool live = gr < Mrows;
          float4 o = *reinterpret_cast<const float4*>(src + it * 16 * IL_STG_ROW);
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          if (epi == EPI_ACT) {
            if (MODE >= 0 && ACT == ACT_TANH) {     // compiled: tanh (lin_edge0 | lin_edge1 of the wide path)
              o.x = tanh_fast(o.x); o.y = tanh_fast(o.y); o.z = tanh_fast(o.z); o.w = tanh_fast(o.w);
            } else if (MODE >= 0 || act == ACT_SILU) {     // the other compiled modes use SiLU (FFNs)
              o.x = silu_fast(o.x); o.y = silu_fast(o.y); o.z = silu_fast(o.z); o.w = silu_fast(o.w);
            } else {
              o.x = act_apply(o.x, act); o.y = act_apply(o.y, act); o.z = act_apply(o.z, act); o.w = act_apply(o.w, act);
            }
          } else if (epi == EPI_GATED_RES) {        // out = res + gate[mol] * (acc + bias)
            float4 yy, gg;
            if (PREFETCH) { yy = y[it]; gg = g[it]; }
            else if (live) {
              yy = *reinterpret_cast<const float4*>(aux + (size_t)gr * ld_aux + col);
              gg = __ldg(reinterpret_cast<const float4*>(gate + (size_t)mol[it] * ld_gate + col));
            } else { yy = gg = make_float4(0.f, 0.f, 0.f, 0.f); }
            o.x = fmaf(gg.x, o.x, yy.x); o.y = fmaf(gg.y, o.y, yy.y);
            o.z = fmaf(gg.z, o.z, yy.z); o.w = fmaf(gg.w, o.w, yy.w);
          }
          if (!live) o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (hasdot) {
            dacc[it][0] = fmaf(o.x, dw0.x, fmaf(o.y, dw0.y, fmaf(o.z, dw0.z, fmaf(o.w, dw0.w, dacc[it][0]))));
            dacc[it][1] = fmaf(o.x, dw1.x, fmaf(o.y, dw1.y, fmaf(o.z, dw1.z, fmaf(o.w, dw1.w, dacc[it][1]))));
            dacc[it][2] = fmaf(o.x, dw2.x, fmaf(o.y, dw2.y, fmaf(o.z, dw2.z, fmaf(o.w, dw2.w, dacc[it][2]))));
          }
          if (has32 && live) *reinterpret_cast<float4*>(C32 + (size_t)gr * ldc32 + col) = o;
          const uint2 hh = make_uint2(pack_h2(o.x, o.y), pack_h2(o.z, o.w));
          if ((has16 || hasimg || hasimg2) && fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))) > 65504.f)
            atomicAdd(&g_sat_imglinear, 1u);       // an fp16 operand saturated (cvt.satfinite clamps silently): count it
          if (has16 && live) {
            if (c16_pm) *reinterpret_cast<uint2*>(C16 + ((size_t)(col >> 3) * ldc16 + gr) * 8 + (col & 7)) = hh;
            else *reinterpret_cast<uint2*>(C16 + (size_t)gr * ldc16 + col) = hh;
          }
          if (hasimg && img) *reinterpret_cast<uint2*>(img + img_piece(r, 0, piece, IL_A_STAGE)) = hh;
          if (hasimg2 && img2) *reinterpret_cast<uint2*>(img2 + img_piece(r, 0, piece2, IL_A_STAGE)) = hh;
        }
      }
      if (hasdot) {
        // The 8 lanes of a row group hold partial sums of the same 8 rows x 3 outputs.  Butterfly with halving: after the
        // exchanges over lane bits 2, 1, 0 (12 + 6 + 3 shuffles) lane s of the group holds the three complete sums of row s.
        const int slot = n * NTEAM + team;
        float w12[12], w6[6], w3[3];
        {
          const bool hi = (lane & 4) != 0;
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const float lo_v = dacc[i / 3][i % 3], hi_v = dacc[4 + i / 3][i % 3];
            const float keep = hi ? hi_v : lo_v, send = hi ? lo_v : hi_v;
            w12[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        {
          const bool hi = (lane & 2) != 0;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float keep = hi ? w12[6 + i] : w12[i], send = hi ? w12[i] : w12[6 + i];
            w6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
        }
        {
          const bool hi = (lane & 1) != 0;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float keep = hi ? w6[3 + i] : w6[i], send = hi ? w6[i] : w6[3 + i];
            w3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
        }
        const int gr = gr0 + 16 * (lane & 7);
        if (gr < Mrows) *reinterpret_cast<float4*>(dot_out + (size_t)gr * a.ld_dot + 4 * slot) = make_float4(w3[0], w3[1], w3[2], 0.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&bar_tempty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}


// ---------------------------------------------------------------------------------------------------------------------
// Fused-row-dots GEMM of the wide path's coordinate branch (coord_mlp.0 -> SiLU -> coord_mlp.2 dots; reference
// models/mol_gnn.py:84-88), rows = directed edges.  With K = N = D the streaming kernel above moves (A tile + W tile) per
// 128 x NT output tile through L2 -- 64 FLOP per byte at NT = 128, which is the L2 bandwidth cap of the chip (~6300 B/clk),
// not the tensor pipe.  Here a work unit is TWO 128-row tiles x one NT-column tile (NT = 192 at D = 384, 256 at D = 256 / 512):
// every W chunk is fetched once for both row tiles (two MMAs per chunk into two accumulators), and the epilogue has no
// HBM output except the dots, so it needs no row-major staging: 16 epilogue warps read their rows straight from tensor
// memory (lane = row; warp = lane quarter x {row tile, column half}), with the per-column constants (bias, the three dot
// weights) broadcast from a shared-memory table.
constexpr int ID_EPI_WARPS = 16;
constexpr int ID_THREADS = 64 + 32 * ID_EPI_WARPS;
constexpr int ID_RING_BYTES = 196608;                // 3 stages at NT = 192 / 256, 4 at NT = 128
constexpr int ID_CTAB_BYTES = 8192;                  // float4 per output column, N <= 512
constexpr int ID_SMEM = ID_RING_BYTES + ID_CTAB_BYTES + 512;
static_assert(ID_SMEM <= 232448, "shared memory budget");

__global__ void __launch_bounds__(ID_THREADS, 1) k_imglinear_dot2(ImgLinearArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int nt = a.NT;
  const int stage_bytes = 2 * IL_A_STAGE + nt * 128;
  int stages = ID_RING_BYTES / stage_bytes;
  if (stages > IL_MAX_STAGES) stages = IL_MAX_STAGES;
  float4* ctab = reinterpret_cast<float4*>(smem + ID_RING_BYTES);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + ID_RING_BYTES + ID_CTAB_BYTES);
  uint64_t* bar_empty = bar_full + IL_MAX_STAGES;
  uint64_t* bar_tfull = bar_empty + IL_MAX_STAGES;            // [2]
  uint64_t* bar_tempty = bar_tfull + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = a.K / 64;
  const int ntn = a.N / nt;
  const int mt = (a.M + TILE_ROWS - 1) / TILE_ROWS;
  const int units = ((mt + 1) / 2) * ntn;
  // accumulators: NT = 128 -- two sets of (tile 0, tile 1) at columns 0 | 128 and 256 | 384, so the epilogue of one unit runs
  // under the MMAs of the next; wider tiles -- one set at columns 0 | 256
  const uint32_t nbuf = nt <= 128 ? 2u : 1u;
  const uint32_t tile_cols = nt <= 128 ? 128u : 256u;

  if (tid == 0) {
    for (int s = 0; s < IL_MAX_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_tfull[s], 1); mbar_init(&bar_tempty[s], ID_EPI_WARPS); }
    fence_barrier_init();
  }
  for (int c = tid; c < a.N; c += ID_THREADS)
    ctab[c] = make_float4(a.bias ? __ldg(a.bias + c) : 0.f, __ldg(a.dot_w + c), __ldg(a.dot_w + a.N + c), __ldg(a.dot_w + 2 * a.N + c));
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* Aimg = static_cast<const uint8_t*>(a.Aimg);
      const uint8_t* Wimg = static_cast<const uint8_t*>(a.Wimg);
      const uint32_t wbytes = (uint32_t)nt * 128u;
      uint32_t it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int pr = u / ntn, n = u - pr * ntn;
        const int m0 = 2 * pr, m1 = min(2 * pr + 1, mt - 1);    // odd tail: the second accumulator repeats the first tile (not stored)
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1u;
          mbar_wait(&bar_empty[s], ph ^ 1u);
          uint8_t* dst = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&bar_full[s], 2 * IL_A_STAGE + wbytes);
          bulk_g2s(dst, Aimg + ((size_t)m0 * nk + kc) * IL_A_STAGE, IL_A_STAGE, &bar_full[s]);
          bulk_g2s(dst + IL_A_STAGE, Aimg + ((size_t)m1 * nk + kc) * IL_A_STAGE, IL_A_STAGE, &bar_full[s]);
          bulk_g2s(dst + 2 * IL_A_STAGE, Wimg + ((size_t)n * nk + kc) * wbytes, wbytes, &bar_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(nt);
      uint32_t it = 0, ai = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ai) {
        const uint32_t ab = ai % nbuf, aph = (ai / nbuf) & 1u;
        const uint32_t td = tmem + ab * 256u;
        mbar_wait(&bar_tempty[ab], aph ^ 1u);
        tc_fence_after();
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1u;
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sw = sa + 2 * IL_A_STAGE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t wd = umma_desc_sw128(sw + kk * 32);
            umma_f16(td, umma_desc_sw128(sa + kk * 32), wd, idesc, (kc | kk) ? 1u : 0u);
            umma_f16(td + tile_cols, umma_desc_sw128(sa + IL_A_STAGE + kk * 32), wd, idesc, (kc | kk) ? 1u : 0u);
          }
          umma_commit(&bar_empty[s]);
        }
        umma_commit(&bar_tfull[ab]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int rq = warp & 3;                       // TMEM lane quarter this warp may read
    const int q = ew >> 2;                         // (row tile, column half): warps 4q+2 .. 4q+5 cover the four lane quarters
    const int tile = q >> 1, half = q & 1;
    const int cw = nt / 2;                         // 64, 96 or 128 columns per warp: chunks of 32
    const int row = rq * 32 + lane;
    const uint32_t tbase0 = tmem + (uint32_t)tile * tile_cols + ((uint32_t)rq << 21) + (uint32_t)(half * cw);
    uint32_t ai = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ai) {
      const int pr = u / ntn, n = u - pr * ntn;
      const int m = 2 * pr + tile;
      const float4* ct = ctab + n * nt + half * cw;
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      const uint32_t ab = ai % nbuf, aph = (ai / nbuf) & 1u;
      const uint32_t tbase = tbase0 + ab * 256u;
      mbar_wait(&bar_tfull[ab], aph);
      tc_fence_after();
      for (int c0 = 0; c0 < cw; c0 += 32) {
        float x[32];
        tmem_ld32(tbase + (uint32_t)c0, x);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 k = ct[c0 + i];             // (bias, w0, w1, w2) of this column: one broadcast LDS.128
          const float s = silu_fast(x[i] + k.x);
          d0 = fmaf(s, k.y, d0); d1 = fmaf(s, k.z, d1); d2 = fmaf(s, k.w, d2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&bar_tempty[ab]);  // this accumulator set is drained
      const int gr = m * TILE_ROWS + row;
      if (m < mt && gr < a.M)
        *reinterpret_cast<float4*>(a.dot_out + (size_t)gr * a.ld_dot + 4 * (n * 2 + half)) = make_float4(d0, d1, d2, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace

cudaError_t sat_count_imglinear(unsigned int* out, bool reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_sat_imglinear, sizeof(unsigned int));
  if (e == cudaSuccess && reset) { const unsigned int z = 0; e = cudaMemcpyToSymbol(g_sat_imglinear, &z, sizeof(z)); }
  return e;
}

const char* check_imglinear(const ImgLinearArgs& a) {
  if (a.M <= 0) return "imglinear: M <= 0";
  if (a.K <= 0 || a.K % 64) return "imglinear: K must be a positive multiple of 64";
  const bool dot_only = a.dot_out && !a.C32 && !a.C16 && !a.Cimg;
  if (!(a.NT == 64 || a.NT == 128 || a.NT == 256 || (a.NT == 192 && dot_only))) return "imglinear: NT must be 64/128/256 (192: fused row dots only)";
  if (a.N <= 0 || a.N % a.NT) return "imglinear: N must be a multiple of NT";
  if (!a.Aimg || !a.Wimg) return "imglinear: operand image missing";
  if ((reinterpret_cast<uintptr_t>(a.Aimg) | reinterpret_cast<uintptr_t>(a.Wimg)) & 127) return "imglinear: images must be 128-byte aligned";
  if (!a.C32 && !a.C16 && !a.Cimg && !a.dot_out) return "imglinear: no output";
  if (a.epi == EPI_LN_MOD) {
    if (a.N != 128 || a.NT != 128 || !a.Cimg || a.C32 || a.C16 || a.Cimg2 || a.dot_out || a.cimg_k)
      return "imglinear: JODO_EPI_LN_MOD needs N = NT = 128 and the image output alone (default placement)";
    if (a.ln_cols <= 0 || a.ln_cols > 128 || (a.ln_cols % 8) || !a.gate || !a.row_mol || (a.ld_gate % 4) || (a.ln_off_shift % 4) ||
        (a.ln_off_scale % 4) || (reinterpret_cast<uintptr_t>(a.gate) & 15))
      return "imglinear: JODO_EPI_LN_MOD needs ln_cols % 8 == 0 (<= 128), the table (gate, ld_gate, row_mol) and 16-byte aligned offsets";
    if (reinterpret_cast<uintptr_t>(a.Cimg) & 127) return "imglinear: image output must be 128-byte aligned";
    return nullptr;
  }
  if (a.dot_out && (!a.dot_w || a.epi != EPI_ACT || a.ld_dot < 4 * 2 * (a.N / a.NT) || (a.ld_dot % 4) || ((reinterpret_cast<uintptr_t>(a.dot_w) | reinterpret_cast<uintptr_t>(a.dot_out)) & 15)))
    return "imglinear: fused row dots need dot_w, JODO_EPI_ACT and ld_dot >= 8 N / NT";
  if (a.Cimg2 && !a.Cimg) return "imglinear: Cimg2 needs Cimg";
  if (a.Cimg && a.cimg_k && ((a.cimg_k % 64) || (a.cimg_col0 % 8) || (a.cimg_ncols % 4) || a.cimg_col0 + a.cimg_ncols > a.cimg_k || a.cimg_ncols > a.N))
    return "imglinear: bad placement of the image output";
  if (a.Cimg2 && ((reinterpret_cast<uintptr_t>(a.Cimg2) & 127) || a.cimg2_k <= 0 || (a.cimg2_k % 64) || (a.cimg2_col0 % 8) || (a.cimg2_ncols % 4) ||
                  a.cimg2_col0 + a.cimg2_ncols > a.cimg2_k || a.cimg2_ncols > a.N))
    return "imglinear: bad placement of the second image output";
  if (a.C32 && ((a.ldc32 % 4) || (reinterpret_cast<uintptr_t>(a.C32) & 15))) return "imglinear: fp32 output must be 16-byte aligned rows";
  if (a.C16 && !a.c16_piece_major && ((a.ldc16 % 8) || (reinterpret_cast<uintptr_t>(a.C16) & 15))) return "imglinear: fp16 output must be 16-byte aligned rows";
  if (a.C16 && a.c16_piece_major && (a.ldc16 < a.M || (reinterpret_cast<uintptr_t>(a.C16) & 15))) return "imglinear: piece-major fp16 output needs ldc16 >= M rows";
  if (a.Cimg && ((a.N % 64) || (reinterpret_cast<uintptr_t>(a.Cimg) & 127))) return "imglinear: image output needs N % 64 == 0";
  if (a.epi != EPI_STORE && a.epi != EPI_ACT && a.epi != EPI_GATED_RES) return "imglinear: unsupported epilogue";
  if (a.epi == EPI_GATED_RES && (!a.aux || !a.gate || !a.row_mol || (a.ld_aux % 4) || (a.ld_gate % 4))) return "imglinear: aux/gate/row_mol missing";
  return nullptr;
}

namespace {
template <int MODE, int ACT, int EW>
cudaError_t launch_mode_ew(const ImgLinearArgs& a, int grid, cudaStream_t stream) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_imglinear<MODE, ACT, EW>, IL_SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  k_imglinear<MODE, ACT, EW><<<grid, il_threads(EW), IL_SMEM, stream>>>(a);
  return cudaGetLastError();
}
// WIDE_OK: this compiled mode also exists with 16 epilogue warps (taken when NT >= 128; JODO_IL_EPI8=1: always 8).  Measured on
// B200 (GEOM nf = 384): tanh(lin_edge0 | lin_edge1) 0.299 -> 0.253 ms, ff_linear3 0.160 -> 0.132, input_lin edge part 0.145 -> 0.134.
// The gated-residual modes stay on 8 warps: without the registers to prefetch their residual / gate rows they got slower
// (ff_linear2 0.041 -> 0.050 ms, ff_linear4 0.210 -> 0.232).
template <int MODE, int ACT = ACT_SILU, bool WIDE_OK = false>
cudaError_t launch_mode(const ImgLinearArgs& a, int grid, cudaStream_t stream) {
  if (WIDE_OK) {
    static const bool epi8 = [] { const char* e = getenv("JODO_IL_EPI8"); return e && e[0] == '1'; }();
    if (a.NT >= 128 && !epi8) return launch_mode_ew<MODE, ACT, WIDE_OK ? 16 : 8>(a, grid, stream);
  }
  return launch_mode_ew<MODE, ACT, 8>(a, grid, stream);
}
}  // namespace

cudaError_t launch_imglinear(const ImgLinearArgs& a_in, int num_sms, cudaStream_t stream) {
  ImgLinearArgs a = a_in;
  if (a.cimg_k == 0) { a.cimg_k = a.N; a.cimg_col0 = 0; a.cimg_ncols = a.N; }       // default placement: the image is the output
  const int units = ((a.M + TILE_ROWS - 1) / TILE_ROWS) * (a.N / a.NT);
  const int grid = units < num_sms ? units : num_sms;
  const int mode = il_mode(a.epi, a.C32 != nullptr, a.C16 != nullptr, a.Cimg != nullptr, a.Cimg2 != nullptr, a.dot_out != nullptr);
  const bool silu_or_none = a.epi != EPI_ACT || a.act_out == ACT_SILU;
  if (a.epi == EPI_LN_MOD) return launch_mode_ew<IL_MODE_LN, ACT_SILU, 16>(a, grid, stream);
  if (mode == il_mode(EPI_ACT, false, false, false, false, true) && a.act_out == ACT_SILU && a.NT >= 128 && a.N <= 512) {
    // wide: coord_mlp.0 + coord_mlp.2 dots -- two row tiles per W chunk, 16 epilogue warps (JODO_DOT_LEGACY=1: the streaming kernel)
    static const bool legacy = [] { const char* e = getenv("JODO_DOT_LEGACY"); return e && e[0] == '1'; }();
    if (!legacy || a.NT == 192) {
      static DevAttr attr = {};
      cudaError_t e0 = ensure_dyn_smem(k_imglinear_dot2, ID_SMEM, attr);
      if (e0 != cudaSuccess) return e0;
      const int u2 = ((a.M + 2 * TILE_ROWS - 1) / (2 * TILE_ROWS)) * (a.N / a.NT);
      k_imglinear_dot2<<<u2 < num_sms ? u2 : num_sms, ID_THREADS, ID_SMEM, stream>>>(a);
      return cudaGetLastError();
    }
  }
  if (a.epi == EPI_STORE && a.C16 && a.c16_piece_major && !a.C32 && !a.Cimg && !a.dot_out)
    return launch_mode<IL_MODE_PM16, ACT_SILU, true>(a, grid, stream);                                                                              // q|k|v, hoisted parts (fused path)
  if (silu_or_none) {
    switch (mode) {
      case il_mode(EPI_STORE, false, true, false): return launch_mode<il_mode(EPI_STORE, false, true, false), ACT_SILU, true>(a, grid, stream);      // q|k|v, hoisted parts
      case il_mode(EPI_STORE, true, false, false): return launch_mode<il_mode(EPI_STORE, true, false, false), ACT_SILU, true>(a, grid, stream);      // node_i
      case il_mode(EPI_ACT, false, false, true): return launch_mode<il_mode(EPI_ACT, false, false, true), ACT_SILU, true>(a, grid, stream);          // ff_linear1
      case il_mode(EPI_GATED_RES, true, false, true): return launch_mode<il_mode(EPI_GATED_RES, true, false, true)>(a, grid, stream);  // ff_linear2
      case il_mode(EPI_GATED_RES, true, false, false): return launch_mode<il_mode(EPI_GATED_RES, true, false, false)>(a, grid, stream);  // wide: ff_linear4, head accumulation
      case il_mode(EPI_GATED_RES, true, false, true, true): return launch_mode<il_mode(EPI_GATED_RES, true, false, true, true)>(a, grid, stream);  // wide: ff_linear4 -> e32 + [e | dist] + heads' operand
      case il_mode(EPI_ACT, false, false, false, false, true): return launch_mode<il_mode(EPI_ACT, false, false, false, false, true)>(a, grid, stream);  // wide: coord_mlp.0 + coord_mlp.2 dots
      // (the run-time-flag variant below is 3x slower on these shapes: 0.79 - 0.91 vs 0.27 ms on [490 k x 128] x [128 x 768])
      case il_mode(EPI_ACT, true, false, false): return launch_mode<il_mode(EPI_ACT, true, false, false), ACT_SILU, true>(a, grid, stream);          // wide: second layer of the edge heads
      case il_mode(EPI_STORE, true, false, true, true): return launch_mode<il_mode(EPI_STORE, true, false, true, true), ACT_SILU, true>(a, grid, stream);  // wide: model-level edge_emb -> e32 + two operand copies
      default: break;
    }
  } else if (a.act_out == ACT_TANH && mode == il_mode(EPI_ACT, false, true, false)) {
    return launch_mode<il_mode(EPI_ACT, false, true, false), ACT_TANH, true>(a, grid, stream);                                            // wide: tanh(lin_edge0 | lin_edge1)
  }
  return launch_mode<-1>(a, grid, stream);
}

}  // namespace jodo
