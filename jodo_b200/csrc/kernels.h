// Host-side launch interfaces of the jodo_b200 kernels (internal; the public C ABI is include/jodo_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jodo {

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2 };
enum { EPI_STORE = 0, EPI_ACT = 1, EPI_ADD = 2, EPI_GATED_RES = 3 };

struct RowLinearArgs {
  const float* A; int lda; int M; int K;        // activations, row-major, K % 32 == 0 (zero padded)
  const float* Wimg;                            // weight image [N/NT][K/32][NT][128 B]
  const float* bias;                            // [N] or null
  float* C; int ldc; int N; int NT;             // output, row-major
  int act_in;                                   // applied to A on load
  int epi; int act_out;
  const float* aux; int ld_aux;                 // EPI_ADD addend / EPI_GATED_RES residual
  const float* gate; int ld_gate;               // EPI_GATED_RES: gate[row_mol[row], col]
  const int* row_mol;
};
const char* check_rowlinear(const RowLinearArgs& a);
cudaError_t launch_rowlinear(const RowLinearArgs& a, cudaStream_t stream);

// ---- per-molecule AdaLN table layout (floats from the start of a molecule's table row) -------------
// [0,2)  model-level GBF (scale, shift); then per layer l at TAB_HEAD + l*tab_layer_stride(D):
//   node  shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp   (6 x D)
//   edge  shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp   (6 x ed)
//   equi  shift, scale                                                      (2 x D)
//   gbf   scale, shift                                                      (2, padded to 16)
constexpr int TAB_HEAD = 16;
__host__ __device__ constexpr int tab_layer_stride(int D) { return 6 * D + 6 * (D / 4) + 2 * D + 16; }
__host__ __device__ constexpr int tab_node(int D) { return 0; }
__host__ __device__ constexpr int tab_edge(int D) { return 6 * D; }
__host__ __device__ constexpr int tab_equi(int D) { return 6 * D + 6 * (D / 4); }
__host__ __device__ constexpr int tab_gbf(int D) { return 6 * D + 6 * (D / 4) + 2 * D; }

// ---- varlen plan (built by the host once per node mask) -------------------------------------------
// Atoms are packed (padding removed): node index in [0, Nn).  Directed edges are laid out in tiles of
// 128 rows; every group = all (n-1) partners of one atom, never split across tiles.
struct Plan {
  int B, Nn, n_tiles, N;                 // molecules, packed atoms, edge tiles, dense padded size
  const int* node_mol;                   // [Nn] molecule of a packed atom
  const int* node_dense;                 // [Nn] b*N + i
  const int* mol_start;                  // [B+1] first packed atom of a molecule
  const int* row_g;                      // [n_tiles*128] packed atom that owns the row's group, -1 = padding
  const int* row_j;                      // [n_tiles*128] the partner atom
  const uint32_t* row_meta;              // group start row (8b) | group length (8b) << 8 | group index in tile (8b) << 16
  const int* tile_ngroups;               // [n_tiles]
};

struct ModelDims {
  int D, ed, T, L, r, S, sc, qk, C, inn, ch, cn, ce, cond_ch;
  int ld_tab;                            // floats per molecule in the table buffer
  int ld_ah;                             // row stride of the concatenated atom hidden buffer
  int keh;                               // columns (multiple of 32) of the concatenated edge hidden image
};

}  // namespace jodo
