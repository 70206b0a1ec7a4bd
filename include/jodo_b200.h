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

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JODO_ABI_VERSION 17

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
#define JODO_EPI_LN_MOD 4     /* jodo_imglinear only: image = LayerNorm(acc + bias) * (1 + scale) + shift per row (see ln_* fields) */

const char* jodo_last_error_string(void);
int jodo_abi_version(void);

/* C[M,N] = epi(act_in(A[M,K]) * W^T + bias) on the tcgen05 tensor cores (fp16 operands -- the mantissa of tf32, saturating at +-65504 -- fp32 accumulate).
 * Wimg is the pre-swizzled fp16 weight image built by jodo_b200.pack.weight_image_h ([N/NT][K/64][NT][64]); K % 64 == 0.
 * Replaces every per-atom / per-molecule nn.Linear of the reference forward: time_mlp, cond_mlp, cond_lin
 * (models/mol_gnn.py:481-489, 679-684), node_emb (:556), lin_query/key/value (models/layers.py:147-149),
 * ff_linear1/2 (models/mol_gnn.py:262-264), node_i (:567), node_pred_mlp (:573) and the hoisted per-atom
 * parts of node2edge_lin (:304-305) and input_lin (:73,79), plus the per-molecule AdaLN tables the
 * reference evaluates per edge (node_time_mlp/edge_time_mlp :291-294, equi time_mlp :78, GBF time_mlp
 * models/layers.py:330).
 * out_f16 != 0 (JODO_EPI_STORE only): C is an fp16 row-major matrix (ldc in elements) -- used for the per-atom
 * operands the edge kernels gather (q | k | v, the hoisted input_lin parts, the hoisted node2edge_lin part). */
int jodo_rowlinear(const float* A, int lda, int M, int K, const void* Wimg, const float* bias, void* C, int ldc,
                   int N, int NT, int act_in, int epi, int act_out, const float* aux, int ld_aux, const float* gate,
                   int ld_gate, const int* row_mol, int out_f16, const int* skip_if_zero, void* stream);
/* skip_if_zero (may be null): device flag of jodo_uniform_flag; when it reads 0 the launch does nothing -- under uniform
 * conditioning the per-molecule rows (noise-level embedding, AdaLN table) are all equal to row 0, which jodo_row0_linear
 * computes and every consumer reads. */


/* Persistent, TMA-fed variant of jodo_rowlinear for the per-atom GEMMs of a DGT block: the activation operand is
 * an fp16 operand image in HBM, [ceil(M/128)][K/64][128 rows][128 bytes] (K-major SWIZZLE_128B, the layout
 * jodo_ln_mod_img and this kernel's own image output write), so both operands are bulk-copied into shared memory
 * and the epilogue of one 128 x NT tile overlaps the tcgen05 main loop of the next (two TMEM accumulators).
 * Replaces the same reference nn.Linear call sites as jodo_rowlinear (models/layers.py:147-149,
 * models/mol_gnn.py:262-264, 304-311, 567, 73-79).  Any subset of the three outputs may be requested. */
typedef struct jodo_imglinear_args {
  const void* Aimg; int M, K;             /* fp16 activation image; M real rows */
  const void* Wimg; const float* bias;    /* fp16 weight image [N/NT][K/64][NT][128 B]; bias [N] or null */
  int N, NT;
  int epi, act_out;                       /* JODO_EPI_STORE | JODO_EPI_ACT | JODO_EPI_GATED_RES */
  const float* aux; int ld_aux;           /* GATED_RES: residual rows */
  const float* gate; int ld_gate;         /* GATED_RES: gate[row_mol[row], col] */
  const int* row_mol;
  const int* nonuni;                      /* device flag of jodo_uniform_flag or null: 0 = read gate row 0 for every row */
  const int* skip_if_zero;                /* device flag or null: when it reads 0 the launch does nothing (the per-molecule
                                             AdaLN table under uniform conditioning: row 0 comes from jodo_rowlinear) */
  float* C32; int ldc32;                  /* fp32 row-major output or null */
  void* C16; int ldc16;                   /* fp16 row-major output or null (ld in elements) */
  int c16_piece_major;                    /* != 0: C16 is [N/8][ldc16 rows][8] -- 16-byte column pieces with the rows of one
                                             piece contiguous, so that lanes gathering consecutive atoms read whole lines */
  void* Cimg;                             /* fp16 operand image output [ceil(M/128)][N/64][128][128 B] or null */
  /* optional placement of the image output inside a wider image (0 = the defaults N, 0, N): the destination has cimg_k
   * columns, output column c goes to column cimg_col0 + c, and only columns c < cimg_ncols are written (cimg_col0 % 8 == 0,
   * cimg_ncols % 4 == 0).  Cimg2 (may be null): a second image destination with its own placement.  Used by the wide path:
   * the edge FFN writes the new edge state straight into the [e | dist] operand of the next GEMMs and into its slot of the
   * edge heads' operand (no separate conversion pass). */
  int cimg_k, cimg_col0, cimg_ncols;
  void* Cimg2; int cimg2_k, cimg2_col0, cimg2_ncols;
  /* optional fused row dot products (JODO_EPI_ACT only, may be null): dot_out[row, 4 s + k] = sum over the columns of
   * slot s of act(acc + bias)[row, c] * dot_w[k * N + c], k < 3, slot s = (column tile) * 2 + (column half): the caller adds
   * the 2 N / NT slots.  coord_mlp.2 (three outputs, reference models/mol_gnn.py:66-69, 82) rides on coord_mlp.0's epilogue, so
   * the SiLU output never goes to HBM.  With dot_out no other output is required.  With dot_out as the ONLY output, SiLU,
   * NT >= 128 and N <= 512 the launch runs k_imglinear_dot2 (two 128-row tiles share every weight chunk: half the L2 bytes per
   * FLOP; 16 epilogue warps reading tensor memory directly); NT = 192 is accepted in that mode only. */
  const float* dot_w; float* dot_out; int ld_dot;
  /* JODO_EPI_LN_MOD (N = NT = 128, the only output is Cimg with the default placement): LayerNorm (eps 1e-6, no affine) over
   * the first ln_cols columns of acc + bias, then x * tab[mol, ln_off_scale + c] + tab[mol, ln_off_shift + c] with tab = gate /
   * ld_gate / row_mol / nonuni as for the gated epilogue (the scale columns hold 1 + scale); columns >= ln_cols and rows with
   * ln_valid[row] < 0 are written as zeros.  The wide path's block edge_emb + norm1_edge (reference models/mol_gnn.py:287,
   * 296-297) in one launch: the fp32 edge_emb output never goes to HBM. */
  const int* ln_valid; int ln_cols, ln_off_shift, ln_off_scale;
} jodo_imglinear_args;
int jodo_imglinear(const jodo_imglinear_args* a, void* stream);


/* ---- weight packing: reference parameters -> operand images / tables (replaces the per-tensor host code a framework
 * would run at load time; reference side: the state dict of DGT_concat, models/mol_gnn.py:414-489).  An item copies the
 * source sub-matrix src[rows, cols] (fp32, row-major, leading dimension src_ld) to (row0, col0) of a destination piece as
 * dst = src * scale + add.  Destination kinds: JODO_PACK_F32 -- fp32 row-major with leading dimension dst_ld;
 * JODO_PACK_IMG_F16 -- fp16 operand image [N/nt][k_pad/64][nt][128 B], K-major SWIZZLE_128B, saturating at +-65504;
 * JODO_PACK_IMG_TF32 -- fp32 image [N/nt][k_pad/32][nt][128 B] rounded to tf32.  Padding = zero-filled destination.
 * jodo_pack_weights runs ONE kernel over a device-resident item table; blk_item / blk_first (device, built by the host)
 * map every thread block of 2048 source elements to its item and the item's first block. */
#define JODO_PACK_F32 0
#define JODO_PACK_IMG_F16 1
#define JODO_PACK_IMG_TF32 2
#define JODO_PACK_ELEMS_PER_BLOCK 2048
typedef struct jodo_pack_item {
  const float* src; int src_ld, rows, cols;
  void* dst; int kind, dst_ld, nt, k_pad, row0, col0;
  float scale, add;
} jodo_pack_item;
int jodo_pack_weights(const jodo_pack_item* items_dev, const int* blk_item_dev, const int* blk_first_dev, int n_blocks, void* stream);

/* ---- varlen plan and argument blocks of the edge-tile kernels ------------------------------------
 * Atoms are packed (padding removed); directed edges live in tiles of 128 rows; a group = all partners
 * of one atom, never split across tiles (built by jodo_b200/plan.py once per node mask; replaces the
 * per-forward adj_mask.nonzero() + dense_to_sparse of reference models/mol_gnn.py:512-514). */
typedef struct jodo_plan {
  int B, Nn, n_tiles, N;                 /* molecules, packed atoms, edge tiles, dense padded size */
  const int* node_mol;                   /* [Nn] molecule of a packed atom */
  const int* node_dense;                 /* [Nn] b*N + i */
  const int* mol_start;                  /* [B+1] first packed atom of a molecule */
  const int* row_g;                      /* [n_tiles*128] packed atom that owns the row's group, -1 = padding */
  const int* row_j;                      /* [n_tiles*128] the partner atom */
  const uint32_t* row_meta;              /* group start row (8b) | group length (8b) << 8 | group index in tile (8b) << 16 */
  const int* tile_ngroups;               /* [n_tiles] */
  const int* row_mol;                    /* [n_tiles*128] molecule of the row's group atom (0 on padding rows) */
  const int* row_pair;                   /* [n_tiles*128] row of the unordered pair {g, j} in the PAIR plan (-1 = padding), or
                                            null.  The edge state is symmetric in (g, j) (the reference's edge ops are pointwise
                                            on symmetric inputs, models/mol_gnn.py:284-317), so it is stored once per pair:
                                            jodo_edge_embed / jodo_edge_update / jodo_edge_head run on a second jodo_plan whose
                                            rows are the pairs i < j of every molecule (row_g = i, row_j = j, no groups,
                                            row_pair = null), and jodo_attn / jodo_equi, which work per directed edge, gather
                                            their fp16 operand rows and adjacency bits through row_pair. */
} jodo_plan;

/* Edge state between kernels (per tile of 128 PAIR rows):
 *   e32  fp32 master copy, piece-major [16 pieces of 4 columns][128 rows][16 B] (32 KB per tile) -- the residual
 *        stream, written by jodo_edge_embed, read and rewritten in place by jodo_edge_update (row-per-thread, coalesced);
 *   e16  fp16 operand copy, image [128 rows][64 cols] (16 KB per tile) -- what the tensor-core kernels load;
 *   eh   fp16 image of the concatenated edge hiddens [keh/64 chunks][128][64]: chunk 0 = model-level embedding,
 *        then ce columns per block (edge_i projections), consumed by jodo_edge_head.
 * fp16 operands carry the same 10-bit mantissa as tf32 ones (values beyond +-65504 saturate). */
typedef struct jodo_edge_embed_args {                    /* model-level edge embedding (reference models/mol_gnn.py:517-557) */
  jodo_plan p;
  const float* edge_x; const float* cond_edge_x; const float* cond_x;   /* dense inputs; cond_* null on the first call */
  int ch, inn; float edge_th, spatial_cut;
  int* dist_flag;                         /* device int, set by the kernel's first phase: any cond distance != 0 */
  const float* tab; int ld_tab;           /* per-molecule tables (model-level GBF scale/shift at [0],[1]) */
  const float* gbf;                       /* GBF constants {mu, c1, c2} x 64 */
  const float* w_img; const float* bias;  /* edge_emb tf32 image (N=64, K=96: [dist 64 | edge_x | cond_edge_x | 0]), bias[64] */
  float* e32; void* e16;                  /* out: edge state */
  void* eh; size_t eh_tile_bytes;         /* out: chunk 0 of the edge-hidden image */
  uint8_t* extra;                         /* out: [R] bit0 = 2-D adjacency head, bit1 = spatial adjacency head */
  const int* nonuni;                      /* device flag of jodo_uniform_flag or null: 0 = read table row 0 */
  int* mol_bad;                           /* [B] or null: a non-finite edge_x / cond_edge_x / cond position entry is replaced by
                                             0 and marks its molecule (see jodo_gather_nodes) */
} jodo_edge_embed_args;

typedef struct jodo_attn_args {                         /* TransMixLayer on edge tiles (reference models/layers.py:131-186) */
  jodo_plan p;
  const void* e16;                        /* block input edge features */
  const float* pos;                       /* [Nn] float4 */
  const void* qkv; int ldq;               /* fp16 piece-major [96][ldq rows][8]: q | k | v of LN-modulated atoms (q, k in split-head layout) */
  const float* tab; int ld_tab; int tab_off;   /* table base of this layer */
  const uint8_t* extra;
  const void* w_emb_img;                  /* block edge_emb fp16 image (N=64, K=128: [dist | e]) */
  const void* w0_img; const void* w1_img; /* lin_edge0 (N=256 split-head, K=64), lin_edge1 (N=256, K=64), fp16 */
  float* hnode;                           /* out [Nn, 256] */
  const int* nonuni;                      /* see jodo_equi_args */
  /* per-column constants by value (constant-bank operands): */
  float gbf4[256];                        /* this layer's GBF constants {mu, c1, c2, 0} per feature column */
  float b_emb[64];                        /* block edge_emb bias */
} jodo_attn_args;

typedef struct jodo_edge_update_args {                   /* edge residual + FFN + edge_l (reference models/mol_gnn.py:304-305,313-317,568) */
  jodo_plan p;
  float* e32; void* e16;                  /* in/out fp32 state (in place, piece-major tiles), out fp16 operand copy */
  const void* P; int ldp;                 /* fp16 piece-major [8][ldp rows][8]: node2edge_lin(hnode) without bias */
  const float* tab; int ld_tab; int tab_off;
  int r;                                  /* mlp_ratio (2 or 4): hidden = 64 r */
  const void* w3_img;                     /* fp16 image (N=64 r in tiles of 128, K=64), pre-scaled by 1/2 */
  const void* w4_img;                     /* fp16 image (N=64, K=64 r) */
  const void* wl_img;                     /* edge_l fp16 image (N=16, K=64) */
  void* eh; size_t eh_tile_bytes; int eh_col; int ce;   /* out: columns [eh_col, eh_col+ce) of the edge-hidden image */
  const int* nonuni;                      /* see jodo_equi_args */
  /* per-column constants by value (constant-bank operands): */
  float b_n2e[64];                        /* node2edge_lin bias */
  float b3[256];                          /* ff_linear3 bias / 2 (first 64 r entries) */
  float b4[64];                           /* ff_linear4 bias */
  float bl[16];                           /* edge_l bias (first ce entries) */
} jodo_edge_update_args;

typedef struct jodo_equi_args {                         /* MultiCondEquiUpdate (reference models/mol_gnn.py:71-94) */
  jodo_plan p;
  const void* e16;                        /* updated edge features */
  const float* pos_in; float* pos_out;    /* [Nn] float4 */
  const void* AB; int ldab;               /* fp16 piece-major [64][ldab rows][8]: input_lin[:, :D] h + bias | input_lin[:, D:2D] h */
  const float* tab; int ld_tab; int tab_off;
  const uint8_t* extra;
  const void* win_img;                    /* input_lin edge part fp16 image (N=256, K=128: [e | dist]) */
  const void* wc0_img;                    /* coord_mlp.0 fp16 image (N=256, K=256), pre-scaled by 1/2 */
  const void* w2_img;                     /* coord_mlp.2 fp16 image (N=16: rows 0..2 real, K=256) */
  const void* w2_img32;                   /* the same padded to N=32 (the CTA-pair kernel feeds this MMA's A operand from tensor
                                             memory, which needs N >= 32 with cta_group::2); null = single-CTA kernel only */
  float coord_scale;                      /* CoorsNorm.scale */
  const int* nonuni;                      /* device flag written by jodo_uniform_flag: 0 = every molecule has the same
                                             conditioning row (fast path reads row 0 through constant memory); may be null */
  /* per-column constants passed BY VALUE so that they reach the FMA pipe as constant-bank operands: */
  float gbf4[256];                        /* GBF constants, {mu, c1, c2, 0} per feature column (entry 0 unused) */
  float b0h[256];                         /* coord_mlp.0 bias / 2 */
  int skip_if_uniform;                    /* != 0: do nothing when *nonuni == 0 (jodo_equi_lin covers that case) */
} jodo_equi_args;

/* The same update under UNIFORM conditioning (every molecule carries the same noise level: always in unconditional
 * sampling, reference sampling.py:549) with coord_mlp.0 composed into input_lin: LayerNorm without affine is C x / sigma
 * (C = centering), so  coord_mlp.0(LN(x)(1 + scale) + shift) = (M x) / sigma + d  with  M = W0 diag(1 + scale) C,
 * d = W0 shift + b0, and M x is linear in [h_row | h_col | e | dist].  jodo_equi_compose forms M W_in for every block
 * from the step's table row (one launch per call); jodo_equi_lin then needs one K = 128 tensor-core product (N = 512:
 * x for the row statistics, y = (M W_e)[e | dist]) per edge tile instead of the K = 128 and K = 256 products, the
 * LayerNorm operand image and the coord_mlp.2 product of jodo_equi.  Both launches return at once when *nonuni != 0. */
typedef struct jodo_equi_compose_item {   /* one per block; device array */
  const float* w0; const float* b0;       /* coord_mlp.0 weight [256, 256] / bias, fp32 */
  const float* wi; const float* bi;       /* input_lin weight [256, 640] ([h_row | h_col | e | dist]) / bias, fp32 */
  const float* w2;                        /* coord_mlp.2 weight [3, 256], fp32 */
  int tab_off;                            /* floats from the start of a table row to this block's equi (shift | 1 + scale) rows */
  void* wce_img;                          /* out: fp16 image of (M W_e) / 2 (N = 256, K = 128), the layout of win_img */
  void* ab_img;                           /* in/out: the per-atom GEMM's weight image (N = 1024 in tiles of 256, K = 256); tiles 2, 3
                                             receive (M W_A) / 2 and (M W_B) / 2 */
  float* ab_bias;                         /* in/out: its bias [1024]; entries [512, 768) receive (M b_in) / 2 */
  float* consts;                          /* out [1026]: d / 2 | coord_mlp.2 rows | GBF (1 + scale, shift) of the step */
} jodo_equi_compose_item;
int jodo_equi_compose(const jodo_equi_compose_item* items_dev, int n_blocks, const float* tab_row0, const int* nonuni, void* stream);

typedef struct jodo_equi_lin_args {
  jodo_plan p;
  const void* e16;
  const float* pos_in; float* pos_out;
  const void* AB; int ldab;               /* fp16 piece-major [128][ldab rows][8]: A | B | YA | YB (the per-atom GEMM with N = 1024) */
  const uint8_t* extra;
  const void* win_img; const void* wce_img;
  const float* consts;                    /* the block's jodo_equi_compose_item.consts */
  float coord_scale;
  const int* nonuni;                      /* required */
  float gbf4[256];
} jodo_equi_lin_args;
int jodo_equi_lin(const jodo_equi_lin_args* a, void* stream);

typedef struct jodo_edge_head_args {                     /* edge_exist_mlp | edge_type_mlp (reference models/mol_gnn.py:466-479,574-578) */
  jodo_plan p;
  const void* eh; size_t eh_tile_bytes; int keh;         /* concatenated edge hiddens (keh = 192) */
  const void* w0_img; const float* b0;    /* [exist.0 ; type.0]  fp16 image (N=128, K=keh) */
  const void* w2_img; const float* b2;    /* block-diag [exist.2 ; type.2] fp16 image (N=64, K=128) */
  const float* w4; const float* b4;       /* [ch, 32] rows: exist.4, type.4...; bias [ch] */
  int ch;
  float* out_dense;                       /* [B,N,N,ch], zero-filled by the caller; a pair row writes [b,i,j,:] and [b,j,i,:]
                                             (the reference's 0.5 (e + e^T), models/mol_gnn.py:579, of two identical values) */
  const int* mol_bad;                     /* [B] or null: molecules marked by jodo_gather_nodes / jodo_edge_embed get NaN */
} jodo_edge_head_args;


/* per-atom / per-molecule elementwise kernels */
int jodo_time_features(const float* noise_level, const float* w8, float* feat64, int B, void* stream);  /* [B, 64] */
int jodo_cond_in(const float* ctx, const float* w0, const float* b0, float* out, int rows, int D, void* stream);
/* mol_bad ([B] ints, may be null): NaN isolation.  The reference lets a non-finite input poison its own molecule and
 * then zeroes ALL positions of the batch (models/mol_gnn.py:587-589).  The edge-tile kernels reduce over the rows of
 * a tile with a tensor-core product, where 0 x NaN would leak into other molecules that share the tile, so non-finite
 * inputs are replaced by 0 at the two input kernels, their molecule is marked here, and jodo_node_out / jodo_edge_head
 * write NaN for marked molecules and treat any mark as the batch-global NaN condition. */
int jodo_gather_nodes(const float* xh, const float* cond_x, const jodo_plan* p, int inn, int kin, float* xin, float* pos4,
                      int* mol_bad, void* stream);
int jodo_ln_mod(int D, const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab, int off_gate,
                int off_shift, int off_scale, const jodo_plan* p, float* out, int ldo, void* stream);
/* jodo_ln_mod with fp16 operand-image outputs (D = 256): out_img = image of LN(x + gate*y)*(1+scale)+shift,
 * out32 (optional) the same rows in fp32, y_img (optional) the image of y itself (the attention output that
 * node2edge_lin consumes, reference models/mol_gnn.py:304-305).  Rows beyond Nn of the last tile are zeroed. */
int jodo_ln_mod_img(const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab, int off_gate,
                    int off_shift, int off_scale, const jodo_plan* p, float* out32, int ldo, void* out_img, void* y_img,
                    const int* nonuni, void* stream);
/* fp16 operand image of act(rows[M, K]) (row-major fp32, ld in elements, K % 64 == 0) for jodo_imglinear */
int jodo_act_image(const float* rows, int ld, int M, int K, int act, void* img, void* stream);
/* nonuni[0] = 1 if any row of rows[B, T] differs (bitwise) from row 0, else 0.  rows = the conditioning embedding
 * temb (noise level [+ context], reference models/mol_gnn.py:534, 728-734): the samplers broadcast one noise level
 * over the batch (sampling.py:549), in which case every per-molecule AdaLN row is the same row. */
int jodo_uniform_flag(const float* rows, int B, int T, int* nonuni, void* stream);
/* out[n] = act_out(act_in(A[0, :]) . W[n, :] + bias[n]) + aux[n] for n < N: ROW 0 of the rowlinear product as a matrix-vector kernel on the same
 * fp16 weight image (activation rounded to fp16 like the GEMM's operand), executed only while *run_if_zero == 0 (null: always).  The per-molecule AdaLN table under uniform
 * conditioning (every molecule carries the same noise level and context, as in sampling: reference sampling.py:549), where
 * all consumers read row 0; the all-rows GEMM (jodo_imglinear with skip_if_zero) covers the other case. */
int jodo_row0_linear(const float* A, int K, const void* Wimg, int NT, int N, const float* bias, int act_in, int act_out,
                     const float* aux, float* out, const int* run_if_zero, void* stream);   /* act_out: none / GELU; aux (may be null): row added after the activation */
int jodo_com(float* pos4, const jodo_plan* p, void* stream);
int jodo_node_out(const float* pos4, const float* atom_pred, int ldp, const jodo_plan* p, int* nan_flag, const int* mol_bad,
                  int inn, float* out_dense, void* stream);
int jodo_sym_edges(const float* tmp, float* out, int B, int N, int ch, void* stream);

/* Fused posterior-mean update + noise of one ancestral reverse step (reference sampling.py:569-589 with the noise
 * construction of models/utils.py:67-99): dense [B,N,F] / [B,N,N,ch] tensors, raw standard-normal draws supplied by
 * the caller ([B,N,3], [B,N,F-3], [B,ch,N,N]); writes x_new, x_mean, e_new, e_mean.  coef_dev, when not null, is a
 * device array {c_x, c_pred, sigma} that overrides the by-value coefficients: a captured CUDA graph of one reverse step
 * is replayed with per-step coefficients written into it (jodo_b200/sampler.py, graph=True). */
int jodo_ancestral_update(const float* x, const float* pred, const float* raw_pos, const float* raw_feat,
                          const float* node_mask, const float* edge_x, const float* edge_pred, const float* raw_edge,
                          const float* edge_mask, int B, int N, int F, int ch, float c_x, float c_pred, float sigma,
                          const float* coef_dev, float* x_new, float* x_mean, float* e_new, float* e_mean, void* stream);

/* The same update with the noise drawn IN the kernels: Philox4x32-10 (counter = element index, step,
 * stream; key = seed) + Box-Muller, so a step needs no raw-draw buffers and no generator launches.  This is a different
 * random stream than the reference's torch.randn by construction; its parity chain is oracle/philox_ref.py (pinned on the
 * published Philox known-answer vectors) -> jodo_philox_normal -> this call against jodo_ancestral_update fed with the
 * oracle's draws.  coef_dev, when not null: {c_x, c_pred, sigma, (unused), step as a float}. */
int jodo_ancestral_update_philox(const float* x, const float* pred, const float* node_mask, const float* edge_x,
                                 const float* edge_pred, const float* edge_mask, int B, int N, int F, int ch, float c_x,
                                 float c_pred, float sigma, const float* coef_dev, unsigned long long seed, unsigned int step,
                                 float* x_new, float* x_mean, float* e_new, float* e_mean, void* stream);
/* out[4 i .. 4 i + 3] = the four standard normals of counter (i, step, stream_id) under `seed`, i < n4 (tests, tools) */
int jodo_philox_normal(unsigned long long n4, unsigned long long seed, unsigned int step, unsigned int stream_id, float* out,
                       void* stream);

/* Tensor-core operands are fp16 (the mantissa of tf32) and saturate at +-65504 instead of overflowing.  The kernels that write
 * the operands with an unbounded range -- the per-atom GEMM outputs the edge kernels gather (q | k | v, hoisted input_lin and
 * node2edge_lin parts, activation images) and the fp16 copy of the edge state -- count every clamped store.  Synchronises the
 * current device; *out = clamped stores since the last reset. */
int jodo_saturation_count(unsigned long long* out, int reset);

/* Fused update of the DPM-Solver++ singlestep for joint 2D & 3D generation (reference mix_dpm_solver.py:93-150 with the
 * stochastic position update of :44-59): atoms and bonds  out = a start - b P0 - c (P1 - P0)  (P1 = null: the intermediate
 * update, c unused), positions  out = cx pos_in + cp pos_pred + sigma z  with z the CoM-free masked normal built from the
 * caller's raw draws [B,N,3] (pos_pred = the positions of P1 when given, else of P0).  coef_dev = device array
 * {a, b, c, cx, cp, sigma}.  pos_in has row stride ld_pos (3 for a bare [B,N,3] tensor, F for the first columns of x). */
int jodo_dpm_update(const float* x_start, const float* pos_in, int ld_pos, const float* pred0, const float* pred1,
                    const float* raw_pos, const float* node_mask, const float* edge_start, const float* edge_pred0,
                    const float* edge_pred1, int B, int N, int F, int ch, const float* coef_dev, float* x_out, float* edge_out,
                    void* stream);

/* edge-tile kernels (tcgen05 + TMEM + bulk-copied operand images); see the structs above */
int jodo_edge_embed(const jodo_edge_embed_args* a, void* stream);
int jodo_attn(const jodo_attn_args* a, void* stream);
int jodo_edge_update(const jodo_edge_update_args* a, void* stream);
int jodo_equi(const jodo_equi_args* a, void* stream);
int jodo_edge_head(const jodo_edge_head_args* a, void* stream);

/* ---- wide path: row kernels between the GEMMs for hidden sizes the fused edge-tile kernels above are not built
 * for (model.nf = 384, reference README.md:156,168).  Every linear layer runs through jodo_imglinear; these kernels do
 * the row-local work on the plan's edge rows (R = n_tiles * 128) or on packed atoms.  Operand images are the fp16
 * K-major SWIZZLE_128B images jodo_imglinear reads: [rows / 128][K / 64][128 rows][128 B]. */
typedef struct jodo_wide_embed_args {                   /* model-level edge inputs (reference models/mol_gnn.py:517-557) */
  jodo_plan p;
  const float* edge_x; const float* cond_edge_x; const float* cond_x;   /* dense inputs; cond_* null on the first call; cond_x
                                             also null for 2-D models (no coordinates: ed = 0, no distance part) */
  int ch, inn, ed; float edge_th, spatial_cut;
  int* dist_flag;                         /* device int (out): any cond distance != 0 (models/mol_gnn.py:544) */
  const float* tab; int ld_tab;           /* per-molecule tables (model-level GBF 1 + scale, shift at [0], [1]) */
  const float* gbf; int ld_gbf;           /* GBF constants {mu, c1, c2} x ld_gbf */
  void* img; int K;                       /* out: image [dist0 (ed) | edge_x (ch) | cond_edge_x (ch) | 0], K columns */
  uint8_t* extra;                         /* out: [R] bit0 = 2-D adjacency head, bit1 = spatial adjacency head */
} jodo_wide_embed_args;

typedef struct jodo_wide_ln_args {                      /* LayerNorm(eps 1e-6) + modulation of a = x + gate (y[yi] + y2[y2i] + ybias) */
  int M, W, Kimg;                         /* rows; real columns (W % 8 == 0, <= 512); image columns (>= W, % 64 == 0) */
  const float* x; int ldx; const int* xi; /* x rows (gathered through xi when given: directed rows reading their pair's row) */
  const float* y; int ldy; const int* yi; /* optional addend rows (gathered through yi when given) */
  const float* y2; int ldy2; const int* y2i;
  const float* ybias;                     /* optional [W] */
  const float* tab; int ld_tab; const int* row_mol;       /* per-molecule table row of every row */
  int off_gate, off_shift, off_scale;     /* table columns; off_gate < 0: the addend enters with weight 1; scale holds 1 + scale */
  const int* valid;                       /* optional: rows with valid[row] < 0 are padding (zero outputs) */
  float* out32; int ldo;                  /* optional modulated fp32 rows (columns [W, min(Kimg, ldo)) are zeroed) */
  void* out_img; void* y_img;             /* fp16 images of the result / of y (either may be null) */
  int x_f16, y_f16;                       /* != 0: x / (y and y2) are fp16 rows (strides in elements, % 8 == 0) */
  const int* nonuni;                      /* device flag of jodo_uniform_flag or null: 0 = every row reads table row 0 (uniform
                                             conditioning: the rows are L1-resident and no longer depend on the row_mol load) */
} jodo_wide_ln_args;

typedef struct jodo_wide_attn_args {                    /* TransMixLayer message + aggregation (reference models/layers.py:157-186) */
  int Nn, D, H, X, sc;                    /* atoms, hidden, heads, extra (adjacency) heads, channels per learned q/k head */
  const int* grp_row0; const int* grp_len; const int* row_j;   /* first edge row / partner count of every atom; partner per row */
  const uint16_t* qkv; int ldq, k_off, v_off;                  /* fp16 rows per atom: q at 0, k at k_off, v at v_off */
  const uint16_t* G; int ldg, g1_off;                          /* fp16 rows per edge: tanh(lin_edge0) at 0, tanh(lin_edge1) at g1_off */
  const uint8_t* extra;
  const int* row_pair;                    /* optional: G and extra are stored per unordered pair; row_pair[row] is the pair row of edge row `row` */
  float* hnode;                           /* out [Nn, D] */
  int max_gl;                             /* largest partner count (sizes the per-CTA logit buffer; <= 255) */
  const int* mol_start; int B, n_max;     /* optional ([B + 1] first packed atom per molecule, molecules, largest molecule): enables
                                             the molecule-staged kernel (k | v of a molecule in shared memory, online softmax) for
                                             the head layouts it is built for; otherwise one CTA per target atom */
} jodo_wide_attn_args;

typedef struct jodo_wide_equi_args {     /* fused coordinate branch: LN(input_lin) + modulation -> coord_mlp.0 -> SiLU -> coord_mlp.2 (mol_gnn.py:71-82) */
  int M, D;                               /* directed edge rows; hidden size (256 or 384) */
  const void* U; int ldu; const int* xi;  /* fp16 rows of input_lin's edge part (per pair when xi = row_pair is given) */
  const void* AB; int ldab;               /* fp16 rows per atom: input_lin[:, :D] h + bias | input_lin[:, D:2D] h */
  const int* row_g; const int* row_j; const int* row_mol;    /* row atom (< 0: padding row), col atom, molecule of every row */
  const float* tab; int ld_tab, off_shift, off_scale;        /* per-molecule table (scale column holds 1 + scale) */
  const void* Wimg; const float* bias;    /* coord_mlp.0: fp16 image [D / 128][D / 64][128][128 B], bias [D] */
  const float* dot_w;                     /* coord_mlp.2 rows [3][D] fp32 (rows beyond 1 + X zero) */
  float* out; int ld_out;                 /* out[row, 4 s + k], s < 2: partial coord_mlp.2 outputs of the two column halves */
} jodo_wide_equi_args;
int jodo_wide_equi(const jodo_wide_equi_args* a, void* stream);
typedef struct jodo_wide_ffn_args {      /* edge FFN of one block on pair rows, one kernel (csrc/wide_ffn.cu): norm2_edge + modulation of
                                             e + gate * (P[i] + P[j] + b), ff_linear3, SiLU, ff_linear4, gated residual (mol_gnn.py:304-317) */
  int M, ed, H;                           /* pair rows (whole 128-row tiles), edge width (32 / 64 / 96), hidden width r * ed, r = 2 or 4 */
  float* e32; int lde;                    /* in / out: fp32 edge state rows */
  const float* P; int ldp;                /* hoisted node2edge_lin per atom: fp32 rows without the bias */
  const int* pair_i; const int* pair_j; const int* pair_mol;   /* atoms of the pair (pair_i < 0: padding row -> zeros), molecule */
  const float* n2e_bias;                  /* [ed] */
  const float* tab; int ld_tab, off_gate, off_shift, off_scale, off_gate2;   /* per-molecule table (scale column holds 1 + scale) */
  const void* w3_img; const float* b3;    /* ff_linear3: fp16 image [ceil(ed / 64)][H][128 B] and bias [H], both pre-scaled by 1/2 (SiLU = h + h tanh h) */
  const void* w4_img; const float* b4;    /* ff_linear4: fp16 image [H / 64][ed][128 B], bias [ed] */
  void* img1; int k1, col1;               /* fp16 copies of the new state: columns [col, col + ed) of operand images with k columns (may be null) */
  void* img2; int k2, col2;
  const int* nonuni;                      /* device flag of jodo_uniform_flag or null: 0 = every row reads table row 0 */
} jodo_wide_ffn_args;
int jodo_wide_edge_ffn(const jodo_wide_ffn_args* a, void* stream);
int jodo_wide_embed_in(const jodo_wide_embed_args* a, void* stream);
int jodo_wide_put(const float* src, int ld, int M, int W, const int* valid, void* img1, int K1, int col1, void* img2, int K2,
                  int col2, void* img3, int K3, int col3, void* stream);   /* fp32 rows -> fp16 columns [col, col + W) of up to three images */
int jodo_wide_dist(const jodo_plan* p, const float* pos4, const float* tab, int ld_tab, int off_gbf, const float* gbf,
                   int ld_gbf, int ed, void* img1, int K1, int col1, void* img2, int K2, int col2, void* stream);   /* mol_gnn.py:284-286 */
int jodo_wide_ln(const jodo_wide_ln_args* a, void* stream);
int jodo_wide_attn(const jodo_wide_attn_args* a, void* stream);
int jodo_wide_equi_out(const int* grp_row0, const int* grp_len, const int* row_j, const float* c3, int ldc, int nslots,
                       const uint8_t* extra, const int* row_pair, int X, float coord_scale, const float* pos_in4, float* pos_out4,
                       int Nn, void* stream);                   /* mol_gnn.py:82-92; extra[row_pair[row]] when row_pair is given;
                                                                   c3[row, 4 s + k], s < nslots: partial coord_mlp.2 outputs to add */
int jodo_wide_head_out(const jodo_plan* p, const float* x, int ldx, int hw, const float* w4, const float* b4, int ch,
                       int both, float* out_dense, void* stream);   /* mol_gnn.py:574-578; both != 0: p is the pair plan, every
                                                                     row writes e_hat[b, i, j] and e_hat[b, j, i] */

/* ---- property classifier (reference cond_gen/model.py:26-220, EGNN of E_GCL_mask layers; called once per batch of
 * conditional samples, sampling.py:363-367).  The linear layers run through jodo_rowlinear (packed atoms) and
 * jodo_imglinear (the plan's directed edge rows: only real ordered pairs i != j, where the reference enumerates the full
 * padded n x n graph with a Python triple loop, cond_gen/utils.py:18-40, and multiplies by the edge mask); these are
 * the row kernels between them. */
/* SiLU(P[g] + Q[j] + w_r |x_g - x_j|^2) -> fp16 operand image [n_tiles * 128][H]: edge_mlp.0 hoisted to per-atom
 * parts PQ = [P | Q] (fp32 rows, leading dimension ldpq >= 2H), model.py:93-95, 126-131, 164-167.  H % 64 == 0, H <= 256. */
int jodo_egnn_edge_in(const jodo_plan* p, const float* pos4, const float* PQ, int ldpq, int H, const float* wr, void* img,
                      void* stream);
/* agg[g] = sum over the rows of atom g of m * sigmoid(w_a . m + b_a) (w_a null: no attention gate), m = fp16 rows
 * [rows][ldm] written by the edge_mlp.2 GEMM; model.py:132-139, 207.  agg: fp32 rows, 16-byte aligned. */
int jodo_egnn_agg(const int* grp_row0, const int* grp_len, const void* M16, int ldm, int H, const float* wa, float ba,
                  float* agg, int ldagg, int Nn, void* stream);
/* out[b, :W] = sum of x[v, :W] over the packed atoms v of molecule b (model.py:66-68) */
int jodo_mol_sum(const float* x, int ldx, int W, const int* mol_start, int B, float* out, int ldo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JODO_B200_H */
