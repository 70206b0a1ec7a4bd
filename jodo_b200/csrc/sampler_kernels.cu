// Fused reverse-SDE update of the ancestral sampler (the caller of the denoiser).
//
// reference sampling.py:569-589: x_mean = c_x x + c_p pred;  x = x_mean + sigma * z_node;  the same for the dense
// bond tensor with symmetric noise.  The noise construction of reference models/utils.py:67-99 is fused in:
//   z_node = [ (raw_pos * mask) - mean over the molecule's atoms , raw_feat * mask ]
//   z_edge[b,i,j,c] = raw_edge[b,c,max(i,j),min(i,j)] for i != j (tril(-1) + transpose), times the edge mask.
// The raw standard-normal draws come from the caller (torch.randn in the reference's call order), so the random
// stream is the reference's.  Products and sums are rounded separately (no FMA contraction), like the torch ops.
#include "common.cuh"
#include "kernels.h"

namespace jodo {

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- in-kernel noise: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.  A different random stream than the reference's
// torch.randn by construction (own parity chain: oracle/philox_ref.py, tests/test_philox.py).  Counter = (index lo, index hi,
// step, stream), key = the sampler's 64-bit seed; every draw is a pure function of (seed, step, element), so a replayed
// CUDA graph only needs the step number from device memory, and the symmetric edge noise needs no transposition pass.
struct PhiloxKey { uint32_t lo, hi; };
enum { PHILOX_POS = 0, PHILOX_FEAT = 1, PHILOX_EDGE = 2 };
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, PhiloxKey k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.lo, lo1, hi0 ^ c.w ^ k.hi, lo0);
    k.lo += 0x9E3779B9u; k.hi += 0xBB67AE85u;
  }
  return c;
}
// four standard normals of counter (idx, step, stream): u = r 2^-32 + 2^-33 in (0, 1], z = sqrt(-2 ln u1) (cos, sin)(2 pi u2)
__device__ __forceinline__ float4 philox_normal4(unsigned long long idx, uint32_t step, uint32_t stream, PhiloxKey k) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), step, stream), k);
  const float s = 2.3283064365386963e-10f, h = 1.1641532182693481e-10f;      // 2^-32, 2^-33
  const float u0 = fmaf((float)r.x, s, h), u1 = fmaf((float)r.y, s, h), u2 = fmaf((float)r.z, s, h), u3 = fmaf((float)r.w, s, h);
  const float ra = sqrtf(-2.0f * logf(fminf(u0, 1.0f))), rb = sqrtf(-2.0f * logf(fminf(u2, 1.0f)));
  float sa, ca, sb, cb;
  sincospif(2.0f * u1, &sa, &ca);
  sincospif(2.0f * u3, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}
__device__ __forceinline__ float f4_get(const float4 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
struct PhiloxArgs { PhiloxKey key; uint32_t step; int on; };
// raw draws of one atom / one bond entry: from the caller's buffers, or generated
__device__ __forceinline__ void raw_pos3(const float* raw_pos, size_t atom, const PhiloxArgs& ph, float (&r)[3]) {
  if (ph.on) { const float4 z = philox_normal4(atom, ph.step, PHILOX_POS, ph.key); r[0] = z.x; r[1] = z.y; r[2] = z.z; }
  else { r[0] = raw_pos[atom * 3]; r[1] = raw_pos[atom * 3 + 1]; r[2] = raw_pos[atom * 3 + 2]; }
}
__global__ void k_philox_normal(unsigned long long n4, uint32_t step, uint32_t stream, PhiloxKey key, float4* out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) out[i] = philox_normal4(i, step, stream, key);
}

// one warp per molecule
__global__ void k_ancestral_nodes(const float* __restrict__ x, const float* __restrict__ pred, const float* __restrict__ raw_pos,
                                  const float* __restrict__ raw_feat, const float* __restrict__ node_mask, int B, int N, int F,
                                  float c_x, float c_p, float sigma, const float* __restrict__ coef, PhiloxArgs ph,
                                  float* __restrict__ x_new, float* __restrict__ x_mean) {
  if (coef) { c_x = coef[0]; c_p = coef[1]; sigma = coef[2]; if (ph.on) ph.step = (uint32_t)coef[4]; }     // per-step values of a replayed CUDA graph
  const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* m = node_mask + (size_t)b * N;
  float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
  for (int i = lane; i < N; i += 32) {
    const float mk = m[i];
    float r[3];
    raw_pos3(raw_pos, (size_t)b * N + i, ph, r);
    sx += r[0] * mk; sy += r[1] * mk; sz += r[2] * mk;
    cnt += mk;
  }
  sx = warp_sum_f(sx); sy = warp_sum_f(sy); sz = warp_sum_f(sz); cnt = warp_sum_f(cnt);
  const float mx = sx / cnt, my = sy / cnt, mz = sz / cnt;
  const int nf = F - 3, q4 = (nf + 3) >> 2;
  for (int i = lane; i < N; i += 32) {
    const float mk = m[i];
    const size_t o = ((size_t)b * N + i) * F;
    float r[3];
    raw_pos3(raw_pos, (size_t)b * N + i, ph, r);
    const float zc[3] = {__fsub_rn(__fmul_rn(r[0], mk), __fmul_rn(mx, mk)), __fsub_rn(__fmul_rn(r[1], mk), __fmul_rn(my, mk)),
                         __fsub_rn(__fmul_rn(r[2], mk), __fmul_rn(mz, mk))};
    float4 zf = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < F; ++k) {
      float rf = 0.f;
      if (k >= 3) {
        if (ph.on) {
          if (((k - 3) & 3) == 0) zf = philox_normal4(((unsigned long long)b * N + i) * q4 + ((k - 3) >> 2), ph.step, PHILOX_FEAT, ph.key);
          rf = f4_get(zf, (k - 3) & 3);
        } else {
          rf = raw_feat[((size_t)b * N + i) * nf + (k - 3)];
        }
      }
      const float z = k < 3 ? zc[k] : __fmul_rn(rf, mk);
      const float mean = __fadd_rn(__fmul_rn(c_x, x[o + k]), __fmul_rn(c_p, pred[o + k]));
      x_mean[o + k] = mean;
      x_new[o + k] = __fadd_rn(mean, __fmul_rn(sigma, z));
    }
  }
}

__global__ void k_ancestral_edges(const float* __restrict__ ex, const float* __restrict__ epred, const float* __restrict__ raw,
                                  const float* __restrict__ edge_mask, int B, int N, int ch, float c_x, float c_p, float sigma,
                                  const float* __restrict__ coef, PhiloxArgs ph, float* __restrict__ e_new, float* __restrict__ e_mean) {
  if (coef) { c_x = coef[0]; c_p = coef[1]; sigma = coef[2]; if (ph.on) ph.step = (uint32_t)coef[4]; }
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [B, N, N]
  const long long total = (long long)B * N * N;
  if (idx >= total) return;
  const int j = (int)(idx % N);
  const long long t = idx / N;
  const int i = (int)(t % N);
  const long long b = t / N;
  const float mk = edge_mask[idx];
  const int hi = i > j ? i : j, lo = i > j ? j : i;
  for (int c = 0; c < ch; ++c) {
    float rz = 0.f;
    if (i != j) {
      const unsigned long long lin = (unsigned long long)(((b * ch + c) * N + hi) * N + lo);
      rz = ph.on ? f4_get(philox_normal4(lin >> 2, ph.step, PHILOX_EDGE, ph.key), (int)(lin & 3)) : raw[lin];
    }
    const float z = __fmul_rn(rz, mk);
    const size_t o = (size_t)idx * ch + c;
    const float mean = __fadd_rn(__fmul_rn(c_x, ex[o]), __fmul_rn(c_p, epred[o]));
    e_mean[o] = mean;
    e_new[o] = __fadd_rn(mean, __fmul_rn(sigma, z));
  }
}

// ---- DPM-Solver++ update (reference mix_dpm_solver.py:44-59 positions, :61-91 first order, :93-150 second order) ----------
// One kernel form covers both updates of the order-2 singlestep:
//   atoms / bonds:  out = a * start - b * P0 - c * (P1 - P0)         (c = 0 and P1 unused in the intermediate update)
//   positions:      out = cx * pos_in + cp * pos_pred + sigma * z,   z = CoM-free masked normal (sigma = 0 on the last step)
// coef = device array {a, b, c, cx, cp, sigma}: a captured CUDA graph of an outer step is replayed with per-step values.
// Operation order and rounding follow the torch expressions of the reference (products and sums rounded separately).
__global__ void k_dpm_nodes(const float* __restrict__ x_start, const float* __restrict__ pos_in, int ld_pos,
                            const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ raw_pos,
                            const float* __restrict__ node_mask, int B, int N, int F, const float* __restrict__ coef,
                            float* __restrict__ out) {
  const float a = coef[0], bq = coef[1], cq = coef[2], cx = coef[3], cp = coef[4], sigma = coef[5];
  const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* m = node_mask + (size_t)b * N;
  float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
  for (int i = lane; i < N; i += 32) {
    const float mk = m[i];
    const float* r = raw_pos + ((size_t)b * N + i) * 3;
    sx += r[0] * mk; sy += r[1] * mk; sz += r[2] * mk;
    cnt += mk;
  }
  sx = warp_sum_f(sx); sy = warp_sum_f(sy); sz = warp_sum_f(sz); cnt = warp_sum_f(cnt);
  const float mean3[3] = {sx / cnt, sy / cnt, sz / cnt};
  const float* pp = p1 ? p1 : p0;                        // the prediction that moves the positions
  for (int i = lane; i < N; i += 32) {
    const float mk = m[i];
    const size_t o = ((size_t)b * N + i) * F;
    const float* r = raw_pos + ((size_t)b * N + i) * 3;
    for (int k = 0; k < 3; ++k) {
      const float z = __fsub_rn(__fmul_rn(r[k], mk), __fmul_rn(mean3[k], mk));
      const float mean = __fadd_rn(__fmul_rn(cx, pos_in[((size_t)b * N + i) * ld_pos + k]), __fmul_rn(cp, pp[o + k]));
      out[o + k] = __fadd_rn(mean, __fmul_rn(sigma, z));
    }
    for (int k = 3; k < F; ++k) {
      float v = __fsub_rn(__fmul_rn(a, x_start[o + k]), __fmul_rn(bq, p0[o + k]));
      if (p1) v = __fsub_rn(v, __fmul_rn(cq, __fsub_rn(p1[o + k], p0[o + k])));
      out[o + k] = v;
    }
  }
}

__global__ void k_dpm_edges(const float* __restrict__ e_start, const float* __restrict__ e0, const float* __restrict__ e1,
                            long long total, const float* __restrict__ coef, float* __restrict__ out) {
  const float a = coef[0], bq = coef[1], cq = coef[2];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = __fsub_rn(__fmul_rn(a, e_start[i]), __fmul_rn(bq, e0[i]));
  if (e1) v = __fsub_rn(v, __fmul_rn(cq, __fsub_rn(e1[i], e0[i])));
  out[i] = v;
}

}  // namespace

cudaError_t launch_dpm_update(const float* x_start, const float* pos_in, int ld_pos, const float* p0, const float* p1,
                              const float* raw_pos, const float* node_mask, const float* e_start, const float* e0,
                              const float* e1, int B, int N, int F, int ch, const float* coef, float* x_out, float* e_out,
                              cudaStream_t st) {
  k_dpm_nodes<<<(B + 7) / 8, 256, 0, st>>>(x_start, pos_in, ld_pos, p0, p1, raw_pos, node_mask, B, N, F, coef, x_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long total = (long long)B * N * N * ch;
  k_dpm_edges<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e_start, e0, e1, total, coef, e_out);
  return cudaGetLastError();
}

cudaError_t launch_ancestral_update(const float* x, const float* pred, const float* raw_pos, const float* raw_feat,
                                    const float* node_mask, const float* ex, const float* epred, const float* raw_edge,
                                    const float* edge_mask, int B, int N, int F, int ch, float c_x, float c_p, float sigma,
                                    const float* coef, int philox, unsigned long long seed, unsigned int step, float* x_new,
                                    float* x_mean, float* e_new, float* e_mean, cudaStream_t st) {
  PhiloxArgs ph;
  ph.key.lo = (uint32_t)seed; ph.key.hi = (uint32_t)(seed >> 32); ph.step = step; ph.on = philox;
  k_ancestral_nodes<<<(B + 7) / 8, 256, 0, st>>>(x, pred, raw_pos, raw_feat, node_mask, B, N, F, c_x, c_p, sigma, coef, ph, x_new, x_mean);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long total = (long long)B * N * N;
  k_ancestral_edges<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ex, epred, raw_edge, edge_mask, B, N, ch, c_x, c_p, sigma,
                                                                    coef, ph, e_new, e_mean);
  return cudaGetLastError();
}

cudaError_t launch_philox_normal(unsigned long long n4, unsigned long long seed, unsigned int step, unsigned int stream_id,
                                 float* out, cudaStream_t st) {
  PhiloxKey key;
  key.lo = (uint32_t)seed; key.hi = (uint32_t)(seed >> 32);
  k_philox_normal<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(n4, step, stream_id, key, reinterpret_cast<float4*>(out));
  return cudaGetLastError();
}

}  // namespace jodo
