// Weight packing on the device: reference parameter tensors -> the operand images / tables the kernels consume.
//
// One launch handles the whole model.  The host (jodo_b200/pack.py) records what the reference's parameters contribute to
// every packed piece as a table of ITEMS; an item copies a source sub-matrix (fp32, row-major) into a destination piece
// at a (row, column) offset with dst = src * scale + add, in one of three destination formats:
//   JODO_PACK_F32      plain fp32 row-major matrix / vector (biases, per-molecule table biases, Fourier weights, ...)
//   JODO_PACK_IMG_F16  fp16 K-major SWIZZLE_128B operand image [N/nt][K/64][nt rows][128 B] (csrc/common.cuh), saturating
//   JODO_PACK_IMG_TF32 tf32-rounded fp32 image [N/nt][K/32][nt rows][128 B]
// Zero padding comes from a zero-filled destination buffer.  Replaces ~1000 small ATen launches per (re)pack.
#include "kernels.h"

namespace jodo {

namespace {

constexpr int PK_BLOCK = 256;
constexpr int PK_PER_BLOCK = 2048;      // source elements per thread block

__global__ void __launch_bounds__(PK_BLOCK) k_pack_items(const jodo_pack_item* __restrict__ items, const int* __restrict__ blk_item,
                                                         const int* __restrict__ blk_first) {
  const int it = blk_item[blockIdx.x];
  const jodo_pack_item I = items[it];
  const long long total = (long long)I.rows * I.cols;
  const long long e0 = (long long)(blockIdx.x - blk_first[it]) * PK_PER_BLOCK;
  for (long long e = e0 + threadIdx.x; e < total && e < e0 + PK_PER_BLOCK; e += PK_BLOCK) {
    const int r = (int)(e / I.cols), c = (int)(e - (long long)r * I.cols);
    const float v = fmaf(I.src[(size_t)r * I.src_ld + c], I.scale, I.add);
    const int R = I.row0 + r, C = I.col0 + c;
    if (I.kind == JODO_PACK_F32) {
      static_cast<float*>(I.dst)[(size_t)R * I.dst_ld + C] = v;
    } else if (I.kind == JODO_PACK_IMG_F16) {
      const int tile = R / I.nt, rr = R - tile * I.nt;
      const size_t off = ((size_t)(tile * (I.k_pad >> 6) + (C >> 6)) * I.nt + rr) * 128 + ((((C & 63) >> 3) ^ (rr & 7)) << 4) + ((C & 7) << 1);
      unsigned short h;
      asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
      *reinterpret_cast<unsigned short*>(static_cast<unsigned char*>(I.dst) + off) = h;
    } else {
      const int tile = R / I.nt, rr = R - tile * I.nt;
      const size_t off = ((size_t)(tile * (I.k_pad >> 5) + (C >> 5)) * I.nt + rr) * 128 + ((((C & 31) >> 2) ^ (rr & 7)) << 4) + ((C & 3) << 2);
      const unsigned int bits = (__float_as_uint(v) + 0x1000u) & ~0x1FFFu;       // round to nearest tf32, ties away (cvt.rna)
      *reinterpret_cast<unsigned int*>(static_cast<unsigned char*>(I.dst) + off) = bits;
    }
  }
}

}  // namespace

cudaError_t launch_pack_items(const jodo_pack_item* items, const int* blk_item, const int* blk_first, int n_blocks, cudaStream_t st) {
  if (n_blocks <= 0) return cudaSuccess;
  k_pack_items<<<n_blocks, PK_BLOCK, 0, st>>>(items, blk_item, blk_first);
  return cudaGetLastError();
}

}  // namespace jodo
