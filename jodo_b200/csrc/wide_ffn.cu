// Wide path: the edge FFN of one DGT block as ONE kernel on pair tiles (weights resident, edge state read and written once).
//
// reference models/mol_gnn.py:304-305 (h_edge = node2edge_lin(h_node[row] + h_node[col]), hoisted to P[i] + P[j] + bias with
// P = W h_node per atom) and :313-317:
//   e2 = norm2_edge(e + gate_msa * h_edge) * (1 + scale_mlp) + shift_mlp;  e_out = e2 + gate_mlp * ff_linear4(SiLU(ff_linear3(e2)))
// Everything is symmetric in (i, j), so the rows are the plan's unordered pairs.
//
// The unfused wide path ran this as a LayerNorm row kernel + two streaming GEMMs (jodo_wide_ln:e2, jodo_imglinear:ff3 / ff4),
// moving e2 (fp32 + fp16 image) and the hidden image through HBM: 0.52 ms per block at GEOM nf = 384 (490 k pairs).  Here the
// two weight images (ed = 96, r = 2: 48 + 36 KB) stay in shared memory and a tile's e2 / hidden rows never leave the SM:
// HBM traffic is e32 in + out and the two fp16 copies of the new state (the e columns of the [e | dist] operand of the
// next GEMMs, the block's slot of the edge heads' operand).
//
// 512 threads = two independent groups of 8 warps walking alternate tiles (group-local named barriers, own operand buffer,
// 256 TMEM columns and mbarrier): while one group waits on a tensor-core round trip or its gathers the other computes.
// Inside a group warp w = tile rows 32 (w & 3) .. +31 (its TMEM lane quarter) x column half (w >> 2).
#include "common.cuh"
#include "kernels.h"

namespace jodo {

// rows whose fp16 operand copy of the new edge state saturated (see jodo_saturation_count)
__device__ unsigned int g_sat_wide_ffn;

namespace {

constexpr int WF_THREADS = 512;
constexpr int WF_GROUP = 256;

template <int ED, int R>
struct WfCfg {
  static constexpr int H = ED * R;                 // hidden width
  static constexpr int CPT = ED / 2;               // edge columns per thread
  static constexpr int HPT = H / 2;                // hidden columns per thread
  static constexpr int KC1 = (ED + 63) / 64;       // K chunks of ff_linear3's operand
  static constexpr int KC2 = H / 64;               // K chunks of ff_linear4's operand
  static constexpr int W3_BYTES = KC1 * H * 128;   // [KC1][H rows][128 B]
  static constexpr int W4_BYTES = KC2 * ED * 128;  // [KC2][ED rows][128 B]
  static constexpr int A_BYTES = (KC2 > KC1 ? KC2 : KC1) * CHUNK_BYTES_A;     // e2 image, then the hidden image (same buffer)
  static constexpr int OFF_W4 = W3_BYTES;
  static constexpr int OFF_G = W3_BYTES + W4_BYTES;
  static constexpr int OFF_LNS = OFF_G + 2 * A_BYTES;          // [2 groups][128 rows][2 halves] float2
  static constexpr int OFF_CONST = OFF_LNS + 2 * 128 * 2 * 8;  // b3 / 2 [H] | b4 [ED] | node2edge bias [ED]
  static constexpr int OFF_MISC = OFF_CONST + (H + 2 * ED) * 4;
  static constexpr int SMEM = OFF_MISC + 128;
  static_assert(ED % 32 == 0 && ED <= 128 && H % 64 == 0 && H <= 256, "sizes");
  static_assert(W3_BYTES % 1024 == 0 && W4_BYTES % 1024 == 0, "operand images must stay 1024-byte aligned");
  static_assert(OFF_MISC % 16 == 0 && SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ void wf_group_sync(int grp) {     // group-local barrier that also orders tcgen05 traffic
  tc_fence_before();
  named_bar_sync(1 + grp, WF_GROUP);
  tc_fence_after();
}
__device__ __forceinline__ void wf_ld8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void wf_ldg8(const float* __restrict__ p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// 8 consecutive floats of a shared-memory table (16-byte aligned): two LDS.128 instead of eight scalar loads
__device__ __forceinline__ void wf_lds8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ uint4 wf_pack8(const float* v) {
  uint4 o;
  o.x = pack_h2(v[0], v[1]); o.y = pack_h2(v[2], v[3]); o.z = pack_h2(v[4], v[5]); o.w = pack_h2(v[6], v[7]);
  return o;
}
// address of the 16-byte piece holding columns [col8, col8 + 8) of `row` in an HBM operand image with K columns
__device__ __forceinline__ uint4* wf_img(void* img, int row, int col8, int K) {
  const int tile = row >> 7, r = row & 127, chunk = col8 >> 6, piece = (col8 & 63) >> 3;
  return reinterpret_cast<uint4*>(static_cast<uint8_t*>(img) + ((size_t)tile * (K >> 6) + chunk) * CHUNK_BYTES_A + (size_t)r * 128 +
                                  ((piece ^ (r & 7)) << 4));
}

template <int ED, int R, int HALF>
__device__ __forceinline__ void wf_group_loop(const WideFfnArgs& a, uint8_t* smem, int grp, int lt, int row, uint32_t tm,
                                              uint64_t* bar_w, uint64_t* bar_m, int tile0, int tile1) {
  using C = WfCfg<ED, R>;
  constexpr int CPT = C::CPT, C0 = CPT * HALF, NP = CPT / 8, HPT = C::HPT, H = C::H;
  uint8_t* A = smem + C::OFF_G + grp * C::A_BYTES;
  float2* LNS = reinterpret_cast<float2*>(smem + C::OFF_LNS) + grp * 256;
  const float* b3h = reinterpret_cast<const float*>(smem + C::OFF_CONST);
  const float* b4 = b3h + H;
  const float* nb = b4 + ED;
  const uint32_t w3 = smem_u32(smem), w4 = smem_u32(smem + C::OFF_W4), sa = smem_u32(A);
  uint32_t par = 0;
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;     // uniform conditioning: every molecule's table row is row 0
  // row metadata is fetched one tile ahead
  int in_ = -1, jn_ = 0, mn_ = 0;
  if (tile0 < tile1) {
    const int gr = tile0 * 128 + row;
    in_ = __ldg(a.pair_i + gr); jn_ = __ldg(a.pair_j + gr); mn_ = __ldg(a.pair_mol + gr);
  }
  for (int tile = tile0; tile < tile1; tile += 2) {
    const int gr = tile * 128 + row;
    const int pi = in_, pj = jn_, mol = mn_;
    if (tile + 2 < tile1) {
      const int g2 = gr + 256;
      in_ = __ldg(a.pair_i + g2); jn_ = __ldg(a.pair_j + g2); mn_ = __ldg(a.pair_mol + g2);
    }
    const bool valid = pi >= 0;
    const float* t = a.tab + (size_t)((valid && !uni) ? mol : 0) * a.ld_tab;
    float* er = a.e32 + (size_t)gr * a.lde + C0;
    const float* p_i = a.P + (size_t)(valid ? pi : 0) * a.ldp + C0;
    const float* p_j = a.P + (size_t)(valid ? pj : 0) * a.ldp + C0;
    // ---- e2 = LN(e + gate_msa * (P[i] + P[j] + b)) * (1 + scale_mlp) + shift_mlp   (the table stores 1 + scale)
    float e2[CPT];
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float x[8], yi[8], yj[8], g[8], bb[8];
      wf_ld8(er + 8 * p, x);
      wf_ldg8(p_i + 8 * p, yi);
      wf_ldg8(p_j + 8 * p, yj);
      wf_ldg8(t + a.off_gate + C0 + 8 * p, g);
      wf_lds8(nb + C0 + 8 * p, bb);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = fmaf(g[k], (yi[k] + yj[k]) + bb[k], x[k]);
        e2[8 * p + k] = v;
        s += v;
        q = fmaf(v, v, q);
      }
    }
    LNS[row * 2 + HALF] = make_float2(s, q);
    named_bar_sync(1 + grp, WF_GROUP);
    {
      const float2 o = LNS[row * 2 + (HALF ^ 1)];
      const float mean = (s + o.x) * (1.0f / ED);
      const float rstd = rsqrtf(fmaxf((q + o.y) * (1.0f / ED) - mean * mean, 0.f) + 1e-6f);
      const float nmr = -mean * rstd;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float sc[8], sh[8];
        wf_ldg8(t + a.off_scale + C0 + 8 * p, sc);
        wf_ldg8(t + a.off_shift + C0 + 8 * p, sh);
#pragma unroll
        for (int k = 0; k < 8; ++k) e2[8 * p + k] = fmaf(fmaf(e2[8 * p + k], rstd, nmr), sc[k], sh[k]);
        const int col = C0 + 8 * p;
        *reinterpret_cast<uint4*>(A + img_piece(row, col >> 6, (col & 63) >> 3, CHUNK_BYTES_A)) = wf_pack8(&e2[8 * p]);
      }
    }
    fence_async_smem();
    wf_group_sync(grp);
    if (lt == 0) {
      if (tile == tile0) mbar_wait(bar_w, 0);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_f16(H);                       // hidden = e2 * W3^T  (image and bias pre-scaled by 1/2)
#pragma unroll
      for (int kk = 0; kk < ED / 16; ++kk)
        umma_f16(tm, umma_desc_sw128(sa + (kk >> 2) * CHUNK_BYTES_A + (kk & 3) * 32),
                 umma_desc_sw128(w3 + (kk >> 2) * (H * 128) + (kk & 3) * 32), idesc, kk ? 1u : 0u);
      umma_commit(bar_m);
    }
    mbar_wait(bar_m, par);
    par ^= 1;
    tc_fence_after();
    // ---- SiLU(hidden) -> fp16 operand of ff_linear4, into the buffer the e2 image no longer needs (its MMA has completed)
#pragma unroll
    for (int c = 0; c < HPT / 16; ++c) {
      float h[16];
      const int h0 = HPT * HALF + 16 * c;
      tmem_ld16(tmem_addr(tm, h0), h);
#pragma unroll
      for (int i8 = 0; i8 < 2; ++i8) {
        float bb[8];
        wf_lds8(b3h + h0 + 8 * i8, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = h[8 * i8 + i] + bb[i];
          h[8 * i8 + i] = fmaf(x, tanh_fast(x), x);
        }
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int col = h0 + 8 * p;
        *reinterpret_cast<uint4*>(A + img_piece(row, col >> 6, (col & 63) >> 3, CHUNK_BYTES_A)) = wf_pack8(&h[8 * p]);
      }
    }
    fence_async_smem();
    wf_group_sync(grp);                               // every thread has read its hidden columns: the accumulator is reused
    if (lt == 0) {
      const uint32_t idesc = umma_idesc_f16(ED);
#pragma unroll
      for (int kk = 0; kk < H / 16; ++kk)
        umma_f16(tm, umma_desc_sw128(sa + (kk >> 2) * CHUNK_BYTES_A + (kk & 3) * 32),
                 umma_desc_sw128(w4 + (kk >> 2) * (ED * 128) + (kk & 3) * 32), idesc, kk ? 1u : 0u);
      umma_commit(bar_m);
    }
    mbar_wait(bar_m, par);
    par ^= 1;
    tc_fence_after();
    // ---- e_out = e2 + gate_mlp * (y + b4): fp32 state in place, fp16 copies into the operand images
    float mx = 0.f;
#pragma unroll
    for (int c = 0; c < CPT / 16; ++c) {
      float y[16];
      tmem_ld16(tmem_addr(tm, C0 + 16 * c), y);
      float g[16], bb[16];
      wf_ldg8(t + a.off_gate2 + C0 + 16 * c, g);
      wf_ldg8(t + a.off_gate2 + C0 + 16 * c + 8, g + 8);
      wf_lds8(b4 + C0 + 16 * c, bb);
      wf_lds8(b4 + C0 + 16 * c + 8, bb + 8);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = valid ? fmaf(g[i], y[i] + bb[i], e2[16 * c + i]) : 0.f;
        e2[16 * c + i] = v;
        mx = fmaxf(mx, fabsf(v));
      }
    }
    if (mx > 65504.f) atomicAdd(&g_sat_wide_ffn, 1u);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      *reinterpret_cast<float4*>(er + 8 * p) = make_float4(e2[8 * p], e2[8 * p + 1], e2[8 * p + 2], e2[8 * p + 3]);
      *reinterpret_cast<float4*>(er + 8 * p + 4) = make_float4(e2[8 * p + 4], e2[8 * p + 5], e2[8 * p + 6], e2[8 * p + 7]);
      const uint4 o = wf_pack8(&e2[8 * p]);
      if (a.img1) *wf_img(a.img1, gr, a.col1 + C0 + 8 * p, a.k1) = o;
      if (a.img2) *wf_img(a.img2, gr, a.col2 + C0 + 8 * p, a.k2) = o;
    }
    tc_fence_before();          // the accumulator reads are ordered before the next tile's MMA by its group barrier
  }
}

template <int ED, int R>
__global__ void __launch_bounds__(WF_THREADS, 1) k_wide_ffn(const __grid_constant__ WideFfnArgs a) {
  using C = WfCfg<ED, R>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* misc = smem + C::OFF_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);     // 0: weights, 1, 2: MMA of group 0, 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t >> 8, lt = t & 255, lw = lt >> 5;
  const int n_tiles = a.M >> 7;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, n_tiles);

  if (t == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    mbar_expect_tx(&bars[0], C::W3_BYTES + C::W4_BYTES);
    bulk_g2s(smem, a.w3_img, C::W3_BYTES, &bars[0]);
    bulk_g2s(smem + C::OFF_W4, a.w4_img, C::W4_BYTES, &bars[0]);
  }
  {
    float* cst = reinterpret_cast<float*>(smem + C::OFF_CONST);
    for (int i = t; i < C::H; i += WF_THREADS) cst[i] = __ldg(a.b3 + i);
    for (int i = t; i < ED; i += WF_THREADS) { cst[C::H + i] = __ldg(a.b4 + i); cst[C::H + ED + i] = __ldg(a.n2e_bias + i); }
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot + 256u * grp;
  const int row = (lw & 3) * 32 + (t & 31);
  if ((lw >> 2) == 0) wf_group_loop<ED, R, 0>(a, smem, grp, lt, row, tm, &bars[0], &bars[1 + grp], tile0 + grp, tile1);
  else wf_group_loop<ED, R, 1>(a, smem, grp, lt, row, tm, &bars[0], &bars[1 + grp], tile0 + grp, tile1);
  if (t == 0) mbar_wait(&bars[0], 0);                // never leave with the weight copies in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(*tmem_slot);
}

template <int ED, int R>
cudaError_t wf_launch(const WideFfnArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_wide_ffn<ED, R>, WfCfg<ED, R>::SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  const int n_tiles = a.M >> 7;
  const int grid = n_tiles < 2 * num_sms ? (n_tiles + 1) / 2 : num_sms;
  k_wide_ffn<ED, R><<<grid, WF_THREADS, WfCfg<ED, R>::SMEM, st>>>(a);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------------
// The same stage for hidden widths beyond 256 (GEOM-Drugs nf = 384 with mlp_ratio = 4: ed = 96, r ed = 384), where the two
// weight images (96 + 72 KB) leave no room for two tiles' operands.  Here nothing is resident: a CTA of 8 warps (tile rows x
// column half) walks the hidden units in chunks of 64 -- MMA1(hc) -> SiLU -> MMA2(hc) accumulating into y -- with the matching
// ff_linear3 rows (16 KB) and ff_linear4 K chunk (12 KB) streamed from L2 through two stages each, and two hidden accumulators
// so that MMA1(hc + 1) runs under the SiLU pass of chunk hc.  108 KB of shared memory, 224 TMEM columns: TWO CTAs per SM, so one
// tile's gathers and tensor-core round trips hide behind the other's arithmetic (a first version with ff_linear3 resident and
// one 16-warp CTA per SM took 0.47 ms per launch at GEOM nf = 384 -- 18 us of exposed latency per tile).
constexpr int WS_THREADS = 256;
constexpr int WS_HC = 64;

template <int ED, int R>
struct WsCfg {
  static constexpr int H = ED * R;
  static constexpr int NCH = H / WS_HC;            // hidden chunks
  static constexpr int CPT = ED / 2;               // edge columns per thread
  static constexpr int KC1 = (ED + 63) / 64;
  static constexpr int W3S_BYTES = KC1 * WS_HC * 128;   // rows [64 hc, 64 hc + 64) of every K chunk of [KC1][H][128 B]
  static constexpr int W4S_BYTES = ED * 128;            // K chunk hc of [H / 64][ED][128 B]
  static constexpr int OFF_W4S = 2 * W3S_BYTES;
  static constexpr int OFF_A = OFF_W4S + 2 * W4S_BYTES;
  static constexpr int OFF_A2 = OFF_A + KC1 * CHUNK_BYTES_A;
  static constexpr int OFF_LNS = OFF_A2 + CHUNK_BYTES_A;              // [128 rows][2 halves] float2
  static constexpr int OFF_CONST = OFF_LNS + 128 * 2 * 8;
  static constexpr int OFF_MISC = OFF_CONST + (H + 2 * ED) * 4;
  static constexpr int SMEM = (OFF_MISC + 128 + 1023) / 1024 * 1024;      // whole KB: the second CTA's window stays 1024-byte aligned
  static_assert(ED % 32 == 0 && ED <= 128 && H % WS_HC == 0 && 2 * WS_HC + ED <= 256, "sizes");
  static_assert(W3S_BYTES % 1024 == 0 && W4S_BYTES % 1024 == 0, "operand images must stay 1024-byte aligned");
  static_assert(OFF_MISC % 16 == 0 && 2 * (SMEM + 1024) <= 233472, "two CTAs per SM");
};

template <int ED, int R>
__global__ void __launch_bounds__(WS_THREADS, 2) k_wide_ffn_stream(const __grid_constant__ WideFfnArgs a) {
  using C = WsCfg<ED, R>;
  constexpr int H = C::H, NCH = C::NCH, CPT = C::CPT, NP = CPT / 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* A = smem + C::OFF_A;
  uint8_t* A2 = smem + C::OFF_A2;
  float2* LNS = reinterpret_cast<float2*>(smem + C::OFF_LNS);
  float* cst = reinterpret_cast<float*>(smem + C::OFF_CONST);
  const float* b3h = cst;
  const float* b4 = cst + H;
  const float* nb = b4 + ED;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_MISC);   // 0, 1: ff_linear3 slices, 2, 3: ff_linear4 slices, 4, 5: MMA1 (by step parity), 6: MMA2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_MISC + 64);
  // MMA1 of step it + 1 is issued as soon as step it's has been seen complete by ONE thread, and takes well under a microsecond:
  // on a single barrier its phase could complete before a slower warp has polled the phase of step it, and that warp would
  // then wait for the phase after next (a hang).  Two barriers by step parity: the next phase of either one is only started
  // behind a CTA barrier that every waiter of the current one has passed.  MMA2's commit needs no such care (the next MMA2
  // is issued behind the step's CTA barrier).
  uint64_t* bar_f = &bars[4];
  uint64_t* bar_y = &bars[6];
  const int t = threadIdx.x, warp = t >> 5;
  const int half = warp >> 2;
  const int row = (warp & 3) * 32 + (t & 31);
  const int n_tiles = a.M >> 7;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per;
  const int tile1 = min(tile0 + per, n_tiles);
  const int n_it = max(tile1 - tile0, 0) * NCH;              // (tile, hidden chunk) steps of this CTA
  const uint8_t* w3g = static_cast<const uint8_t*>(a.w3_img);
  const uint8_t* w4g = static_cast<const uint8_t*>(a.w4_img);

  // ff_linear3 rows of hidden chunk (k % NCH) -> slot k & 1 (one bulk copy per K chunk)
  auto load_w3 = [&](int k) {
    const int st = k & 1, hc = k % NCH;
    mbar_expect_tx(&bars[st], C::W3S_BYTES);
    for (int c = 0; c < C::KC1; ++c)
      bulk_g2s(smem + st * C::W3S_BYTES + c * (WS_HC * 128), w3g + ((size_t)c * H + (size_t)hc * WS_HC) * 128, WS_HC * 128, &bars[st]);
  };
  auto load_w4 = [&](int k) {
    const int st = k & 1, hc = k % NCH;
    mbar_expect_tx(&bars[2 + st], C::W4S_BYTES);
    bulk_g2s(smem + C::OFF_W4S + st * C::W4S_BYTES, w4g + (size_t)hc * C::W4S_BYTES, C::W4S_BYTES, &bars[2 + st]);
  };
  if (t == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    for (int k = 0; k < 2 && k < n_it; ++k) { load_w3(k); load_w4(k); }
  }
  for (int i = t; i < H; i += WS_THREADS) cst[i] = __ldg(a.b3 + i);
  for (int i = t; i < ED; i += WS_THREADS) { cst[H + i] = __ldg(a.b4 + i); cst[H + ED + i] = __ldg(a.n2e_bias + i); }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_y = tm + 2 * WS_HC;                      // hidden accumulators at columns 0 and 64, y behind them
  const uint32_t sa = smem_u32(A), sa2 = smem_u32(A2);
  const int C0 = CPT * half;
  uint32_t par_y = 0;
  const bool uni = a.nonuni != nullptr && *a.nonuni == 0;     // uniform conditioning: every molecule's table row is row 0
  int it = 0;                                                // running (tile, hidden chunk) index
  // MMA1 of step k: hidden chunk = e2 * W3[rows of the chunk]^T (image and bias pre-scaled by 1/2); one thread
  auto mma1 = [&](int k) {
    mbar_wait(&bars[k & 1], (uint32_t)(k >> 1) & 1u);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(WS_HC);
    const uint32_t w3s = smem_u32(smem + (k & 1) * C::W3S_BYTES);
#pragma unroll
    for (int kk = 0; kk < ED / 16; ++kk)
      umma_f16(tm + (uint32_t)((k & 1) * WS_HC), umma_desc_sw128(sa + (kk >> 2) * CHUNK_BYTES_A + (kk & 3) * 32),
               umma_desc_sw128(w3s + (kk >> 2) * (WS_HC * 128) + (kk & 3) * 32), idesc, kk ? 1u : 0u);
    umma_commit(&bar_f[k & 1]);
  };
  int in_ = -1, jn_ = 0, mn_ = 0;
  if (tile0 < tile1) {
    const int gr = tile0 * 128 + row;
    in_ = __ldg(a.pair_i + gr); jn_ = __ldg(a.pair_j + gr); mn_ = __ldg(a.pair_mol + gr);
  }
  for (int tile = tile0; tile < tile1; ++tile) {
    const int gr = tile * 128 + row;
    const int pi = in_, pj = jn_, mol = mn_;
    if (tile + 1 < tile1) {
      const int g2 = gr + 128;
      in_ = __ldg(a.pair_i + g2); jn_ = __ldg(a.pair_j + g2); mn_ = __ldg(a.pair_mol + g2);
    }
    const bool valid = pi >= 0;
    const float* tr = a.tab + (size_t)((valid && !uni) ? mol : 0) * a.ld_tab;
    float* er = a.e32 + (size_t)gr * a.lde + C0;
    const float* p_i = a.P + (size_t)(valid ? pi : 0) * a.ldp + C0;
    const float* p_j = a.P + (size_t)(valid ? pj : 0) * a.ldp + C0;
    // ---- e2 = LN(e + gate_msa * (P[i] + P[j] + b)) * (1 + scale_mlp) + shift_mlp   (the table stores 1 + scale)
    float e2[CPT];
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float x[8], yi[8], yj[8], g[8], bb[8];
      wf_ld8(er + 8 * p, x);
      wf_ldg8(p_i + 8 * p, yi);
      wf_ldg8(p_j + 8 * p, yj);
      wf_ldg8(tr + a.off_gate + C0 + 8 * p, g);
      wf_lds8(nb + C0 + 8 * p, bb);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = fmaf(g[k], (yi[k] + yj[k]) + bb[k], x[k]);
        e2[8 * p + k] = v;
        s += v;
        q = fmaf(v, v, q);
      }
    }
    LNS[row * 2 + half] = make_float2(s, q);
    __syncthreads();
    {
      const float2 o = LNS[row * 2 + (half ^ 1)];
      const float mean = (s + o.x) * (1.0f / ED);
      const float rstd = rsqrtf(fmaxf((q + o.y) * (1.0f / ED) - mean * mean, 0.f) + 1e-6f);
      const float nmr = -mean * rstd;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float sc[8], sh[8];
        wf_ldg8(tr + a.off_scale + C0 + 8 * p, sc);
        wf_ldg8(tr + a.off_shift + C0 + 8 * p, sh);
#pragma unroll
        for (int k = 0; k < 8; ++k) e2[8 * p + k] = fmaf(fmaf(e2[8 * p + k], rstd, nmr), sc[k], sh[k]);
        const int col = C0 + 8 * p;
        *reinterpret_cast<uint4*>(A + img_piece(row, col >> 6, (col & 63) >> 3, CHUNK_BYTES_A)) = wf_pack8(&e2[8 * p]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (t == 0) mma1(it);
#pragma unroll 1
    for (int hc = 0; hc < NCH; ++hc, ++it) {
      mbar_wait(&bar_f[it & 1], (uint32_t)(it >> 1) & 1u);    // MMA1 of this step has completed: its ff_linear3 slot is free
      tc_fence_after();
      if (t == 0) {
        if (it + 2 < n_it) load_w3(it + 2);
        if (hc + 1 < NCH) mma1(it + 1);                       // into the other hidden accumulator, under this SiLU pass
      }
      float h[32];
      tmem_ld32(tmem_addr(tm, (it & 1) * WS_HC + 32 * half), h);
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {
        float bb[8];
        wf_lds8(b3h + hc * WS_HC + 32 * half + 8 * i8, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = h[8 * i8 + i] + bb[i];
          h[8 * i8 + i] = fmaf(x, tanh_fast(x), x);
        }
      }
      if (hc > 0) {                                           // MMA2 of the previous step has consumed the hidden image and its slice
        mbar_wait(bar_y, par_y);
        par_y ^= 1;
        if (t == 0 && it + 1 < n_it) load_w4(it + 1);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
        *reinterpret_cast<uint4*>(A2 + img_piece(row, 0, 4 * half + p, CHUNK_BYTES_A)) = wf_pack8(&h[8 * p]);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (t == 0) {
        mbar_wait(&bars[2 + (it & 1)], (uint32_t)(it >> 1) & 1u);     // this chunk's ff_linear4 slice has landed
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(ED);
        const uint32_t w4s = smem_u32(smem + C::OFF_W4S + (it & 1) * C::W4S_BYTES);
#pragma unroll
        for (int kk = 0; kk < WS_HC / 16; ++kk)
          umma_f16(tm_y, umma_desc_sw128(sa2 + kk * 32), umma_desc_sw128(w4s + kk * 32), idesc, (hc | kk) ? 1u : 0u);
        umma_commit(bar_y);
      }
    }
    mbar_wait(bar_y, par_y);                                  // MMA2 of the last hidden chunk
    par_y ^= 1;
    tc_fence_after();
    // The ff_linear4 slice of step k is requested once MMA2(k - 2) has been seen complete.  `it` now names the next tile's
    // first step, which waits for no MMA2 inside the loop, so the slot freed just now is refilled here for step it + 1.
    if (t == 0 && it + 1 < n_it) load_w4(it + 1);
    // ---- e_out = e2 + gate_mlp * (y + b4): fp32 state in place, fp16 copies into the operand images
    float mx = 0.f;
#pragma unroll
    for (int c = 0; c < CPT / 16; ++c) {
      float y[16], g[16], bb[16];
      tmem_ld16(tmem_addr(tm_y, C0 + 16 * c), y);
      wf_ldg8(tr + a.off_gate2 + C0 + 16 * c, g);
      wf_ldg8(tr + a.off_gate2 + C0 + 16 * c + 8, g + 8);
      wf_lds8(b4 + C0 + 16 * c, bb);
      wf_lds8(b4 + C0 + 16 * c + 8, bb + 8);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = valid ? fmaf(g[i], y[i] + bb[i], e2[16 * c + i]) : 0.f;
        e2[16 * c + i] = v;
        mx = fmaxf(mx, fabsf(v));
      }
    }
    if (mx > 65504.f) atomicAdd(&g_sat_wide_ffn, 1u);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      *reinterpret_cast<float4*>(er + 8 * p) = make_float4(e2[8 * p], e2[8 * p + 1], e2[8 * p + 2], e2[8 * p + 3]);
      *reinterpret_cast<float4*>(er + 8 * p + 4) = make_float4(e2[8 * p + 4], e2[8 * p + 5], e2[8 * p + 6], e2[8 * p + 7]);
      const uint4 o = wf_pack8(&e2[8 * p]);
      if (a.img1) *wf_img(a.img1, gr, a.col1 + C0 + 8 * p, a.k1) = o;
      if (a.img2) *wf_img(a.img2, gr, a.col2 + C0 + 8 * p, a.k2) = o;
    }
    tc_fence_before();          // the accumulator reads are ordered before the next tile's MMAs by its barriers
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tm);
}

template <int ED, int R>
cudaError_t ws_launch(const WideFfnArgs& a, int num_sms, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e0 = ensure_dyn_smem(k_wide_ffn_stream<ED, R>, WsCfg<ED, R>::SMEM, attr);
  if (e0 != cudaSuccess) return e0;
  const int n_tiles = a.M >> 7;
  k_wide_ffn_stream<ED, R><<<n_tiles < 2 * num_sms ? n_tiles : 2 * num_sms, WS_THREADS, WsCfg<ED, R>::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t sat_count_wide_ffn(unsigned int* out, bool reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_sat_wide_ffn, sizeof(unsigned int));
  if (e == cudaSuccess && reset) { const unsigned int z = 0; e = cudaMemcpyToSymbol(g_sat_wide_ffn, &z, sizeof(z)); }
  return e;
}

const char* check_wide_ffn(const WideFfnArgs& a) {
  if (a.M <= 0 || (a.M % 128)) return "jodo_wide_edge_ffn: M must be a positive multiple of 128 (whole pair tiles)";
  const bool sizes = (a.H == 2 * a.ed && (a.ed == 32 || a.ed == 64 || a.ed == 96)) || (a.H == 4 * a.ed && (a.ed == 32 || a.ed == 64 || a.ed == 96));
  if (!sizes) return "jodo_wide_edge_ffn: built for ed = 32 / 64 / 96 with H = 2 ed or 4 ed";
  if (!a.e32 || !a.P || !a.pair_i || !a.pair_j || !a.pair_mol || !a.n2e_bias || !a.tab || !a.w3_img || !a.b3 || !a.w4_img || !a.b4)
    return "jodo_wide_edge_ffn: null input";
  if ((a.lde % 4) || (a.ldp % 4) || (a.ld_tab % 4) || (a.off_gate % 4) || (a.off_shift % 4) || (a.off_scale % 4) || (a.off_gate2 % 4) ||
      a.lde < a.ed || a.ldp < a.ed)
    return "jodo_wide_edge_ffn: rows and table segments must be 16-byte aligned";
  if (((reinterpret_cast<uintptr_t>(a.e32) | reinterpret_cast<uintptr_t>(a.P) | reinterpret_cast<uintptr_t>(a.tab) |
        reinterpret_cast<uintptr_t>(a.w3_img) | reinterpret_cast<uintptr_t>(a.w4_img)) & 15))
    return "jodo_wide_edge_ffn: pointers must be 16-byte aligned";
  if (a.img1 && ((a.k1 % 64) || (a.col1 % 8) || a.col1 + a.ed > a.k1 || (reinterpret_cast<uintptr_t>(a.img1) & 127)))
    return "jodo_wide_edge_ffn: bad placement of the first image output";
  if (a.img2 && ((a.k2 % 64) || (a.col2 % 8) || a.col2 + a.ed > a.k2 || (reinterpret_cast<uintptr_t>(a.img2) & 127)))
    return "jodo_wide_edge_ffn: bad placement of the second image output";
  return nullptr;
}

cudaError_t launch_wide_ffn(const WideFfnArgs& a, int num_sms, cudaStream_t st) {
  const int r = a.H / a.ed;
  if (r == 2) {
    if (a.ed == 32) return wf_launch<32, 2>(a, num_sms, st);
    if (a.ed == 64) return wf_launch<64, 2>(a, num_sms, st);
    return wf_launch<96, 2>(a, num_sms, st);
  }
  if (a.ed == 32) return wf_launch<32, 4>(a, num_sms, st);
  if (a.ed == 64) return wf_launch<64, 4>(a, num_sms, st);
  return ws_launch<96, 4>(a, num_sms, st);                   // hidden width 384: streamed weights, two CTAs per SM
}

}  // namespace jodo
