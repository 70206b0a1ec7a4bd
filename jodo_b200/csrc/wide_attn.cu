// Wide path attention, molecule-staged variant (TransMixLayer.message + aggregation, reference models/layers.py:157-186).
//
// k_wide_attn (wide.cu) runs one CTA per target atom and fetches k[j] | v[j] of every source from L2 again for every
// target: ~3 KB per edge row through L2, which is what bounded it (GEOM nf = 384: 2.9 GB per launch).  Here a CTA owns up
// to 32 target atoms of ONE molecule and first stages k | v of all the molecule's atoms in shared memory (fp16, n x 1.5 KB);
// a warp then walks the sources of one target with only the per-pair rows tanh(lin_edge0) | tanh(lin_edge1) streaming in
// (1.5 KB per edge row), two rows in flight.  Lane l owns q/k columns [EQ l, EQ l + EQ) and value columns [EV l, EV l + EV):
// at most two learned heads per lane on the q/k side (sc >= EQ), exactly one value head (C % EV == 0).  Softmax over the
// sources is online (running max / sum per lane for its value head, accumulators rescaled when the max moves): PyG's
// exp(a - max) / (sum + 1e-16) with the final max, one pass over the rows.
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

namespace jodo {
namespace {

constexpr int WM_THREADS = 512;
constexpr int WM_TG = 32;              // target atoms per CTA (two per warp)

template <int N>
__device__ __forceinline__ void ld_halves(const uint16_t* p, uint32_t (&out)[N / 2]) {     // N halves, 8-byte aligned
  static_assert(N % 4 == 0, "whole 8-byte pieces");
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    const uint2 u = *reinterpret_cast<const uint2*>(p + 4 * i);
    out[2 * i] = u.x; out[2 * i + 1] = u.y;
  }
}

template <int EQ, int EV>
__global__ void __launch_bounds__(WM_THREADS) k_wide_attn_mol(WideAttnArgs a) {
  extern __shared__ __align__(16) uint8_t wm_smem[];
  // the target chunks of one molecule are adjacent in launch order: a pair's row is read by the CTA of either end, and the
  // second read then hits L2 (with the molecule index fastest the re-read came from HBM: 1.23 GB per launch for 0.75 GB of rows)
  const int b = blockIdx.y;
  const int a0 = a.mol_start[b], n = a.mol_start[b + 1] - a0;
  const int t0 = blockIdx.x * WM_TG;
  if (t0 >= n) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int KQ = 32 * EQ, KV = 32 * EV;                 // staged row widths (halves): padded q/k part, value part (= D)
  uint16_t* ks = reinterpret_cast<uint16_t*>(wm_smem);      // [n][KQ]
  uint16_t* vs = ks + (size_t)n * KQ;                       // [n][KV]
  const int H = a.H, X = a.X, S = H - X, sc = a.sc, qk = S * sc, C = a.D / H;
  // ---- stage k | v of the molecule (16-byte pieces; k beyond the real width reads the zero padding of the rows)
  for (int i = tid; i < n * (KQ / 8); i += WM_THREADS) {
    const int r = i / (KQ / 8), p = i - r * (KQ / 8);
    reinterpret_cast<uint4*>(ks)[i] = *reinterpret_cast<const uint4*>(a.qkv + (size_t)(a0 + r) * a.ldq + a.k_off + 8 * p);
  }
  for (int i = tid; i < n * (KV / 8); i += WM_THREADS) {
    const int r = i / (KV / 8), p = i - r * (KV / 8);
    reinterpret_cast<uint4*>(vs)[i] = *reinterpret_cast<const uint4*>(a.qkv + (size_t)(a0 + r) * a.ldq + a.v_off + 8 * p);
  }
  __syncthreads();
  // ---- static lane geometry: q/k elements [EQ lane, EQ lane + EQ) belong to learned heads hA (first `cut`) and hA + 1
  const int e0 = EQ * lane;
  const int hA = e0 / sc;
  const int cut = min(EQ, (hA + 1) * sc - e0);              // elements of this lane in head hA
  const int hv = (EV * lane) / C;                           // the value head of this lane's columns
  // learned head s (owner lane s < S) sums partials of lanes [lo, lo + 4): its elements start in lane lo
  const int lo = (lane * sc) / EQ;
  const float inv = rsqrtf((float)C);

  for (int tt = warp; tt < WM_TG; tt += WM_THREADS / 32) {
    const int tl = t0 + tt;
    if (tl >= n) break;
    const int g = a0 + tl;
    const int gl = a.grp_len[g], row0 = a.grp_row0[g];
    float* dst = a.hnode + (size_t)g * a.D + EV * lane;
    float acc[EV];
#pragma unroll
    for (int i = 0; i < EV; ++i) acc[i] = 0.f;
    if (gl > 0) {
      // q of the target, pre-scaled by 1 / sqrt(C)  (models/layers.py:167: sqrt(out_channels))
      float q[EQ];
      {
        uint32_t qh[EQ / 2];
        ld_halves<EQ>(a.qkv + (size_t)g * a.ldq + e0, qh);
#pragma unroll
        for (int i = 0; i < EQ / 2; ++i) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&qh[i]));
          q[2 * i] = (e0 + 2 * i < qk) ? f.x * inv : 0.f;
          q[2 * i + 1] = (e0 + 2 * i + 1 < qk) ? f.y * inv : 0.f;
        }
      }
      float m = -INFINITY, l = 0.f;
      // row k's per-pair rows are loaded one iteration ahead
      uint32_t g0n[EQ / 2], g1n[EV / 2];
      int jn, pn;
      {
        const int R = row0;
        jn = a.row_j[R] - a0;
        pn = a.row_pair ? a.row_pair[R] : R;
        ld_halves<EQ>(a.G + (size_t)pn * a.ldg + e0, g0n);
        ld_halves<EV>(a.G + (size_t)pn * a.ldg + a.g1_off + EV * lane, g1n);
      }
      uint8_t bits_n = a.extra[pn];
      for (int k = 0; k < gl; ++k) {
        uint32_t g0[EQ / 2], g1[EV / 2];
#pragma unroll
        for (int i = 0; i < EQ / 2; ++i) g0[i] = g0n[i];
#pragma unroll
        for (int i = 0; i < EV / 2; ++i) g1[i] = g1n[i];
        const int j = jn;
        const uint8_t bits = bits_n;
        if (k + 1 < gl) {
          const int R = row0 + k + 1;
          jn = a.row_j[R] - a0;
          pn = a.row_pair ? a.row_pair[R] : R;
          ld_halves<EQ>(a.G + (size_t)pn * a.ldg + e0, g0n);
          ld_halves<EV>(a.G + (size_t)pn * a.ldg + a.g1_off + EV * lane, g1n);
          bits_n = a.extra[pn];
        }
        // ---- partial logits: sum over this lane's elements of q k tanh(g0), split at the head boundary
        uint32_t kh[EQ / 2];
        ld_halves<EQ>(ks + (size_t)j * KQ + e0, kh);
        float pa = 0.f, pb = 0.f;
#pragma unroll
        for (int i = 0; i < EQ / 2; ++i) {
          const float t0_ = fhfma_lo(kh[i], g0[i], 0.f) * q[2 * i], t1_ = fhfma_hi(kh[i], g0[i], 0.f) * q[2 * i + 1];
          if (2 * i < cut) pa += t0_; else pb += t0_;
          if (2 * i + 1 < cut) pa += t1_; else pb += t1_;
        }
        // learned head s = lane: partials of the lanes its elements live in (at most 4: sc <= 3 EQ + 1)
        float lg = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int src = min(lo + d, 31);
          const float xa = __shfl_sync(0xffffffffu, pa, src), xb = __shfl_sync(0xffffffffu, pb, src);
          const int hs = (EQ * src) / sc;                    // first head of lane src
          if (lo + d < 32) lg += (hs == lane ? xa : 0.f) + (hs + 1 == lane ? xb : 0.f);
        }
        // the logit of this lane's value head: adjacency heads first (models/layers.py:170-174), then the learned ones
        const float lgl = __shfl_sync(0xffffffffu, lg, max(hv - X, 0));
        const float aa = hv < X ? (((bits >> hv) & 1) ? 1.0f : -1e10f) : lgl;
        const float mn = fmaxf(m, aa);
        const float corr = __expf(m - mn), pe = __expf(aa - mn);         // first row: m = -inf -> corr = 0
        l = fmaf(l, corr, pe);
        m = mn;
        uint32_t vh[EV / 2];
        ld_halves<EV>(vs + (size_t)j * KV + EV * lane, vh);
#pragma unroll
        for (int i = 0; i < EV / 2; ++i) {
          acc[2 * i] = fmaf(acc[2 * i], corr, fhfma_lo(vh[i], g1[i], 0.f) * pe);
          acc[2 * i + 1] = fmaf(acc[2 * i + 1], corr, fhfma_hi(vh[i], g1[i], 0.f) * pe);
        }
      }
      const float rs = 1.0f / (l + 1e-16f);                  // PyG softmax: exp(a - max) / (sum + 1e-16)
#pragma unroll
      for (int i = 0; i < EV; ++i) acc[i] *= rs;
    }
#pragma unroll
    for (int i = 0; i < EV; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
  }
}

template <int EQ, int EV>
cudaError_t launch_mol(const WideAttnArgs& a, size_t smem, cudaStream_t st) {
  static DevAttr attr = {};
  cudaError_t e = ensure_dyn_smem(k_wide_attn_mol<EQ, EV>, (int)smem, attr);
  if (e != cudaSuccess) return e;
  k_wide_attn_mol<EQ, EV><<<dim3((a.n_max + WM_TG - 1) / WM_TG, a.B), WM_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

// true when the molecule-staged kernel covers these sizes (the launcher of wide.cu falls back to the per-target kernel)
bool wide_attn_mol_ok(const WideAttnArgs& a) {
  if (!a.mol_start || a.B <= 0 || a.n_max <= 0) return false;
  const int S = a.H - a.X, qk = S * a.sc, C = a.D / a.H;
  const int EQ = ((qk + 31) / 32 + 3) & ~3, EV = a.D / 32;
  if (!((EQ == 12 && EV == 12) || (EQ == 8 && EV == 8))) return false;
  if (a.D % 32 || a.sc < EQ || C % EV || S > 32 || a.X > 8) return false;
  if (a.k_off < 32 * EQ || a.v_off - a.k_off < 32 * EQ || a.g1_off < 32 * EQ) return false;    // padded parts readable
  if ((a.ldq % 8) || (a.k_off % 8) || (a.v_off % 8) || (a.ldg % 4) || (a.g1_off % 4)) return false;
  return (size_t)a.n_max * (32 * EQ + 32 * EV) * 2 <= 200 * 1024;
}

cudaError_t launch_wide_attn_mol(const WideAttnArgs& a, cudaStream_t st) {
  const int S = a.H - a.X, qk = S * a.sc;
  const int EQ = ((qk + 31) / 32 + 3) & ~3;
  const size_t smem = (size_t)a.n_max * (32 * EQ + a.D) * 2;
  if (EQ == 12) return launch_mol<12, 12>(a, smem, st);
  return launch_mol<8, 8>(a, smem, st);
}

}  // namespace jodo
