// Shared device primitives for the jodo_b200 kernels (sm_100a only).
//
//  * tcgen05 tensor-core MMA (kind::tf32, M=128, cta_group::1) with accumulators in TMEM
//  * TMEM allocation / tcgen05.ld / tcgen05.st
//  * mbarrier + cp.async.bulk (TMA engine, 1-D bulk copies of pre-swizzled operand images)
//  * the "operand image" layout shared by HBM and shared memory
//
// Operand image (K-major, SWIZZLE_128B, 4-byte elements): a [rows x K] operand is stored as
// K/32 chunks; chunk kc holds columns [32kc, 32kc+32) of every row, one 128-byte line per row,
// line r at byte r*128, and inside a line the 16-byte piece p = (col%32)/4 sits at piece slot
// p ^ (r & 7).  With a 1024-byte aligned base this is exactly the canonical UMMA K-major
// SWIZZLE_128B layout (8-row x 128-byte atoms, SBO = 1024 B), so an image can be copied from HBM
// to shared memory with one linear bulk copy and handed to tcgen05.mma through a descriptor.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jodo {

constexpr int TILE_ROWS = 128;           // rows per edge tile == UMMA M
constexpr int CHUNK_COLS = 32;           // fp32/tf32 columns per 128-byte swizzle line
constexpr int CHUNK_BYTES_A = TILE_ROWS * 128;   // one K-chunk of a 128-row A operand (16 KB)

// ---------------------------------------------------------------- small helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float to_tf32(float x) {       // round-to-nearest tf32, returned as fp32 bits
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// Hardware tanh (MUFU.TANH, max relative error 2^-11 -- the same grade as one tf32 operand rounding; measured on the
// parity fixtures: no change of the end-to-end error against the fp64 reference) and SiLU built on it:
// x*sigmoid(x) = h + h*tanh(h), h = x/2.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_fast(h), h);
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float tanh_f(float x) {         // accurate to ~1e-6 abs (2 MUFU)
  float e = __expf(-2.0f * fabsf(x));
  float t = (1.0f - e) / (1.0f + e);
  return copysignf(t, x);
}

// byte offset of element (row, col) inside an operand image whose chunks are `chunk_bytes` apart
__device__ __host__ __forceinline__ uint32_t img_off(int row, int col, uint32_t chunk_bytes) {
  return (uint32_t)(col >> 5) * chunk_bytes + (uint32_t)row * 128u +
         ((((uint32_t)(col & 31) >> 2) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)col & 3u) << 2);
}
// byte offset of the 16-byte piece `p` (0..7) of row `row` in chunk `kc`
__device__ __forceinline__ uint32_t img_piece(int row, int kc, int p, uint32_t chunk_bytes) {
  return (uint32_t)kc * chunk_bytes + (uint32_t)row * 128u + ((((uint32_t)p) ^ ((uint32_t)row & 7u)) << 4);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// A wait that has lasted ~2 s (4e9 SM cycles) is a protocol bug: trap instead of hanging the device.
__device__ __forceinline__ bool mbar_timed_out(long long& t0) {
  const long long now = clock64();
  if (t0 == 0) { t0 = now; return false; }
  return now - t0 > 4000000000ll;
}
// Spin on the phase with parity `parity`.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  // try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires)
  // instead of burning issue slots that the CTA's other warp group could use
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (!done && (spin & 255u) == 255u && mbar_timed_out(t0)) __trap();   // a barrier that never completes is a bug: do not hang the device
  }
}

// ---------------------------------------------------------------- bulk async copy (TMA engine, 1-D)
// global -> shared::cta, completion counted in bytes on `bar`.  size % 16 == 0, 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// make generic-proxy writes to shared memory visible to the async proxy (TMA engine / tensor core)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {   // one full warp
  static_assert(NCOLS == 32 || NCOLS == 64 || NCOLS == 128 || NCOLS == 256 || NCOLS == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive columns -> 32 registers; thread i of warp w reads TMEM lane 32*(w%4)+i
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors + issue
// Shared-memory matrix descriptor, K-major SWIZZLE_128B (see cute/arch/mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64: 1024 B)
//   | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, M=128, N=n:
//   [4,6) c_format=1(F32) | [7,10) a_format=2(TF32) | [10,13) b_format=2 | [17,23) N>>3 | [24,29) M>>4
__device__ __host__ constexpr uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// A operand from tensor memory (lane = row, each 32-bit column = two consecutive fp16 K elements), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[128 x N] (TMEM, fp32) (+)= A[128 x 32*nchunks] * W[N x 32*nchunks]^T; both operands are images in
// shared memory (A chunks CHUNK_BYTES_A apart, W chunks N*128 bytes apart).  One thread calls this.
__device__ __forceinline__ void mma_tile(uint32_t tmem_d, uint32_t a_saddr, uint32_t w_saddr, int n, int nchunks,
                                         bool accumulate) {
  const uint32_t idesc = umma_idesc_tf32(n);
  uint32_t acc = accumulate ? 1u : 0u;
  for (int kc = 0; kc < nchunks; ++kc) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 4 x (K=8 tf32 = 32 bytes) per 128-byte swizzle line
      uint64_t ad = umma_desc_sw128(a_saddr + kc * CHUNK_BYTES_A + kk * 32);
      uint64_t bd = umma_desc_sw128(w_saddr + kc * (n * 128) + kk * 32);
      umma_tf32(tmem_d, ad, bd, idesc, acc);
      acc = 1u;
    }
  }
}

// ---------------------------------------------------------------- fp16 operands (kind::f16)
// fp16 has the same 10-bit mantissa as tf32, so an fp16 operand rounds exactly like a tf32 one but takes half the
// shared memory and runs the tensor core at twice the rate; values beyond +-65504 saturate (cvt .satfinite).
// fp16 operand image: same K-major SWIZZLE_128B layout, a chunk is 64 columns (128 bytes per row).
//   [4,6) c_format=1(F32) | [7,10) a_format=0(F16) | [10,13) b_format=0 | [17,23) N>>3 | [24,29) M>>4
__device__ __host__ constexpr uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[128 x N] (TMEM, fp32) (+)= A[128 x 64*nchunks] * W[N x 64*nchunks]^T, fp16 operand images in shared memory
// (A chunks CHUNK_BYTES_A apart, W chunks N*128 bytes apart).  One thread calls this.
__device__ __forceinline__ void mma_tile_h(uint32_t tmem_d, uint32_t a_saddr, uint32_t w_saddr, int n, int nchunks,
                                           bool accumulate) {
  const uint32_t idesc = umma_idesc_f16(n);
  uint32_t acc = accumulate ? 1u : 0u;
  for (int kc = 0; kc < nchunks; ++kc) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 4 x (K=16 fp16 = 32 bytes) per 128-byte swizzle line
      uint64_t ad = umma_desc_sw128(a_saddr + kc * CHUNK_BYTES_A + kk * 32);
      uint64_t bd = umma_desc_sw128(w_saddr + kc * (n * 128) + kk * 32);
      umma_f16(tmem_d, ad, bd, idesc, acc);
      acc = 1u;
    }
  }
}
// two fp32 -> packed fp16x2 (lo in bits [0,16)), round to nearest, saturating
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------- packed-half CUDA-core math
// The epilogues that feed fp16 operand images anyway do their elementwise math on fp16 pairs: one MUFU / FMA-pipe
// instruction per two elements, and fp16 x fp16 + fp32 accumulation in one instruction (FHFMA, sm_100: fma.rn.f32.f16)
// instead of two conversions and an FFMA.
__device__ __forceinline__ uint32_t tanh_h2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t mul_h2(uint32_t a, uint32_t b) {
  uint32_t y;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b));
  return y;
}
__device__ __forceinline__ uint32_t fma_h2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t y;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(y) : "r"(a), "r"(b), "r"(c));
  return y;
}
// c + a.lo * b.lo  /  c + a.hi * b.hi, products and sums in fp32
__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t b, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, al, bl, %3;\n\t}"
      : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t b, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, ah, bh, %3;\n\t}"
      : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}
// c + a.lo / c + a.hi (fp32 sum of an fp32 and one half of a packed pair: FHADD)
__device__ __forceinline__ float fhadd_lo(uint32_t a, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah;\n\tmov.b32 {al, ah}, %1;\n\tadd.rn.f32.f16 %0, al, %2;\n\t}" : "=f"(d) : "r"(a), "f"(c));
  return d;
}
__device__ __forceinline__ float fhadd_hi(uint32_t a, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah;\n\tmov.b32 {al, ah}, %1;\n\tadd.rn.f32.f16 %0, ah, %2;\n\t}" : "=f"(d) : "r"(a), "f"(c));
  return d;
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2) and clusters
// A cluster of two CTAs on the two SMs of a TPC runs one tcgen05.mma over M = 256 rows: each CTA supplies the A operand of
// its own 128 rows and HALF of the B operand (N / 2 rows of W at the same shared-memory offset in both CTAs); the
// accumulator of a CTA's 128 rows (all N columns) lands in its own tensor memory.  The leader (cluster rank 0) issues the
// instruction; tcgen05.commit with .multicast arrives on the mbarrier at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (count 1) on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait with cluster-scope acquire (the arrivals may come from the peer CTA).  SPIN: poll with test_wait instead of suspending
// in try_wait (the wake-up after a remote arrival is on the critical path of the one thread that then issues an MMA).
template <bool SPIN = false>
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    if (SPIN) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}"
          : "=r"(done)
          : "r"(addr), "r"(parity)
          : "memory");
    } else {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}"
          : "=r"(done)
          : "r"(addr), "r"(parity), "r"(0x989680u)
          : "memory");
    }
    if (!done && (spin & 255u) == 255u && mbar_timed_out(t0)) __trap();
  }
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot) {   // the same warp index in both CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// kind::f16 instruction descriptor for the pair: M = 256 over both CTAs, N = n (all of it)
__device__ __host__ constexpr uint32_t umma_idesc_f16_2sm(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// A operand from tensor memory (each CTA's own lanes), B halves from both CTAs' shared memory
__device__ __forceinline__ void umma_f16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// all previously issued tcgen05.mma of this thread arrive (count 1) on `bar` in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// TMEM address of (lane base of this warp, column)
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int col) {
  return base + (((uint32_t)(threadIdx.x >> 5) & 3u) << 21) + (uint32_t)col;   // (warp%4)*32 lanes << 16
}

}  // namespace jodo
