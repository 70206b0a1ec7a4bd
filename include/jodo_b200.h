/* jodo_b200 — C ABI of the B200-native DGT denoiser kernels (libjodo_b200.so).
 *
 * This is the boundary a host language binds (ctypes in jodo_b200/_lib.py).
 * The reference (GRAPH-0/JODO) is pure Python/PyTorch and has no FFI of its own; each entry point
 * below names the reference code whose work it replaces (paths relative to the reference root).
 * Conventions: device pointers, sizes in elements, `stream` is a cudaStream_t passed as void*,
 * every call is asynchronous on that stream, allocates nothing, keeps no global mutable state, and
 * returns JODO_OK or an error code; jodo_last_error_string() describes the last failure of the
 * calling thread.
 */
#ifndef JODO_B200_H
#define JODO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JODO_ABI_VERSION 1

#define JODO_OK 0
#define JODO_ERR_ARG 1   /* invalid argument (shape, alignment, unsupported size) */
#define JODO_ERR_CUDA 2  /* CUDA runtime error at launch */

/* activations / epilogues of jodo_rowlinear */
#define JODO_ACT_NONE 0
#define JODO_ACT_SILU 1
#define JODO_ACT_GELU 2
#define JODO_EPI_STORE 0      /* C = acc + bias */
#define JODO_EPI_ACT 1        /* C = act_out(acc + bias) */
#define JODO_EPI_ADD 2        /* C = acc + bias + aux */
#define JODO_EPI_GATED_RES 3  /* C = aux + gate[row_mol[row]] * (acc + bias) */

const char* jodo_last_error_string(void);
int jodo_abi_version(void);

/* C[M,N] = epi(act_in(A[M,K]) * W^T + bias) on the tcgen05 tensor cores (tf32 operands, fp32 accumulate).
 * Wimg is the pre-swizzled weight image built by jodo_b200.pack.weight_image ([N/NT][K/32][NT][32]).
 * Replaces every per-atom / per-molecule nn.Linear of the reference forward: time_mlp, cond_mlp, cond_lin
 * (models/mol_gnn.py:481-489, 679-684), node_emb (:556), lin_query/key/value (models/layers.py:147-149),
 * ff_linear1/2 (models/mol_gnn.py:262-264), node_i (:567), node_pred_mlp (:573) and the hoisted per-atom
 * parts of node2edge_lin (:304-305) and input_lin (:73,79), plus the per-molecule AdaLN tables the
 * reference evaluates per edge (node_time_mlp/edge_time_mlp :291-294, equi time_mlp :78, GBF time_mlp
 * models/layers.py:330). */
int jodo_rowlinear(const float* A, int lda, int M, int K, const float* Wimg, const float* bias, float* C, int ldc,
                   int N, int NT, int act_in, int epi, int act_out, const float* aux, int ld_aux, const float* gate,
                   int ld_gate, const int* row_mol, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JODO_B200_H */
