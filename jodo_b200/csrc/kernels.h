// Host-side launch interfaces of the jodo_b200 kernels (internal; the public C ABI is include/jodo_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/jodo_b200.h"

namespace jodo {

// ---- per-device launch state -------------------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: every launcher remembers, per
// device, the opt-in it has already made (a process may drive several devices, e.g. under torch.nn.DataParallel or
// when a test moves a model from cuda:0 to cuda:1).
constexpr int MAX_DEVICES = 64;
struct DevAttr { int bytes[MAX_DEVICES]; };
template <typename F>
inline cudaError_t ensure_dyn_smem(F kernel, int bytes, DevAttr& cache) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < MAX_DEVICES && cache.bytes[dev] >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < MAX_DEVICES) cache.bytes[dev] = bytes;
  return e;
}
// The uniform-conditioning rows of the edge kernels live in __constant__ tables (one copy per device, shared by all
// streams) that are refreshed by a device-to-device copy in front of each launch.  const_tables_acquire makes `st` wait
// for the last launch that read the tables when that was enqueued on ANOTHER stream; const_tables_release records the
// launch just made.  Two streams driving jodo_b200 models concurrently are therefore serialised at these launches
// instead of reading each other's rows.  Inside a stream capture neither call adds anything (cross-stream event
// dependencies would break capture isolation): a captured graph must not be replayed concurrently with other
// jodo_b200 work on the same device.
cudaError_t const_tables_acquire(cudaStream_t st);
cudaError_t const_tables_release(cudaStream_t st);

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_TANH = 3 };
enum { EPI_STORE = 0, EPI_ACT = 1, EPI_ADD = 2, EPI_GATED_RES = 3, EPI_LN_MOD = 4 };

struct RowLinearArgs {
  const float* A; int lda; int M; int K;        // activations, row-major, K % 32 == 0 (zero padded)
  const float* Wimg;                            // weight image [N/NT][K/32][NT][128 B]
  const float* bias;                            // [N] or null
  void* C; int ldc; int N; int NT;              // output, row-major fp32 (or fp16 when out_f16; ldc in elements)
  int act_in;                                   // applied to A on load
  int epi; int act_out;
  const float* aux; int ld_aux;                 // EPI_ADD addend / EPI_GATED_RES residual
  const float* gate; int ld_gate;               // EPI_GATED_RES: gate[row_mol[row], col]
  const int* row_mol;
  int out_f16;                                  // EPI_STORE only: write fp16 (saturating) instead of fp32
  const int* skip_if_zero;                      // device flag: 0 = the launch does nothing
};
const char* check_rowlinear(const RowLinearArgs& a);
cudaError_t launch_rowlinear(const RowLinearArgs& a, cudaStream_t stream);

using ImgLinearArgs = ::jodo_imglinear_args;
const char* check_imglinear(const ImgLinearArgs& a);
cudaError_t launch_imglinear(const ImgLinearArgs& a, int num_sms, cudaStream_t stream);

cudaError_t sat_count_imglinear(unsigned int* out, bool reset);
cudaError_t sat_count_edge_update(unsigned int* out, bool reset);
cudaError_t launch_pack_items(const jodo_pack_item* items, const int* blk_item, const int* blk_first, int n_blocks, cudaStream_t st);

// ---- per-molecule AdaLN table layout (floats from the start of a molecule's table row) -------------
// Every `scale` entry holds 1 + scale (the packer adds 1 to the bias of those columns), so modulation is one FMA.
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
using Plan = ::jodo_plan;

// ---- node / molecule elementwise kernels (node_kernels.cu) -----------------------------------------
cudaError_t launch_time_features(const float* nl, const float* w, float* feat, int B, cudaStream_t st);
cudaError_t launch_cond_in(const float* ctx, const float* w0, const float* b0, float* out, int rows, int D, cudaStream_t st);
cudaError_t launch_gather_nodes(const float* xh, const float* cond_x, const Plan& p, int inn, int kin, float* xin,
                                float* pos, int* mol_bad, cudaStream_t st);
cudaError_t launch_ln_mod(int D, const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab,
                          int off_gate, int off_shift, int off_scale, const Plan& p, float* out, int ldo,
                          cudaStream_t st);
cudaError_t launch_ln_mod_img(const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab, int off_gate,
                              int off_shift, int off_scale, const Plan& p, float* out32, int ldo, void* out_img,
                              void* y_img, const int* nonuni, cudaStream_t st);
cudaError_t launch_act_image(const float* rows, int ld, int M, int K, int act, void* img, cudaStream_t st);
cudaError_t launch_uniform_flag(const float* rows, int B, int T, int* nonuni, cudaStream_t st);
cudaError_t launch_row0_linear(const float* A, int K, const void* Wimg, int nt, int N, const float* bias, int act_in, int act_out,
                               const float* aux, float* out, const int* run_if_zero, cudaStream_t st);
cudaError_t launch_com(float* pos_new, const Plan& p, cudaStream_t st);
cudaError_t launch_nan_flag(const float* pos, int Nn, const int* mol_bad, int B, int* flag, cudaStream_t st);
cudaError_t launch_node_out(const float* pos, const float* atom_pred, int ldp, const Plan& p, const int* nan_flag,
                            const int* mol_bad, int inn, float* out, cudaStream_t st);
cudaError_t launch_sym_edges(const float* tmp, float* out, int B, int N, int ch, cudaStream_t st);

cudaError_t launch_ancestral_update(const float* x, const float* pred, const float* raw_pos, const float* raw_feat,
                                    const float* node_mask, const float* ex, const float* epred, const float* raw_edge,
                                    const float* edge_mask, int B, int N, int F, int ch, float c_x, float c_p, float sigma,
                                    const float* coef, int philox, unsigned long long seed, unsigned int step, float* x_new,
                                    float* x_mean, float* e_new, float* e_mean, cudaStream_t st);
cudaError_t launch_philox_normal(unsigned long long n4, unsigned long long seed, unsigned int step, unsigned int stream_id,
                                 float* out, cudaStream_t st);

cudaError_t launch_dpm_update(const float* x_start, const float* pos_in, int ld_pos, const float* p0, const float* p1,
                              const float* raw_pos, const float* node_mask, const float* e_start, const float* e0,
                              const float* e1, int B, int N, int F, int ch, const float* coef, float* x_out, float* e_out,
                              cudaStream_t st);

// ---- edge-tile kernels (edge_kernels.cu); all built for nf = 256 (ed = 64, 14+2 heads) --------------
using EdgeEmbedArgs = ::jodo_edge_embed_args;
cudaError_t launch_dist_flag(const EdgeEmbedArgs& a, cudaStream_t st);
cudaError_t launch_edge_embed(const EdgeEmbedArgs& a, int num_sms, cudaStream_t st);

using AttnArgs = ::jodo_attn_args;
cudaError_t launch_attn(const AttnArgs& a, int num_sms, cudaStream_t st);

using EdgeUpdateArgs = ::jodo_edge_update_args;
cudaError_t launch_edge_update(const EdgeUpdateArgs& a, int num_sms, cudaStream_t st);

using EquiArgs = ::jodo_equi_args;
cudaError_t launch_equi(const EquiArgs& a, int num_sms, cudaStream_t st);      // one tile in flight per SM (equi.cu)
cudaError_t launch_equi2(const EquiArgs& a, int num_sms, cudaStream_t st);     // CTA pairs, two-stage pipeline (equi2.cu)

using EquiLinArgs = ::jodo_equi_lin_args;
cudaError_t launch_equi_lin(const EquiLinArgs& a, int num_sms, cudaStream_t st);            // uniform conditioning (equi_lin.cu)
cudaError_t launch_equi_compose(const jodo_equi_compose_item* items_dev, int L, const float* tab_row0, const int* nonuni,
                                cudaStream_t st);

using EdgeHeadArgs = ::jodo_edge_head_args;
cudaError_t launch_edge_head(const EdgeHeadArgs& a, int num_sms, cudaStream_t st);

// ---- wide path (wide.cu): row kernels between the GEMMs for nf = 384
using WideEmbedArgs = ::jodo_wide_embed_args;
using WideLnArgs = ::jodo_wide_ln_args;
using WideAttnArgs = ::jodo_wide_attn_args;
cudaError_t launch_wide_embed_in(const WideEmbedArgs& a, cudaStream_t st);
cudaError_t launch_wide_put(const float* src, int ld, int M, int W, const int* valid, void* img1, int K1, int col1,
                            void* img2, int K2, int col2, void* img3, int K3, int col3, cudaStream_t st);
cudaError_t launch_wide_dist(const Plan& p, const float* pos, const float* tab, int ld_tab, int off_gbf, const float* gbf,
                             int ld_gbf, int ed, void* img1, int K1, int col1, void* img2, int K2, int col2, cudaStream_t st);
cudaError_t launch_wide_ln(const WideLnArgs& a, cudaStream_t st);
cudaError_t launch_wide_attn(const WideAttnArgs& a, cudaStream_t st);
using WideFfnArgs = ::jodo_wide_ffn_args;
const char* check_wide_ffn(const WideFfnArgs& a);
cudaError_t launch_wide_ffn(const WideFfnArgs& a, int num_sms, cudaStream_t st);             // wide_ffn.cu
cudaError_t sat_count_wide_ffn(unsigned int* out, bool reset);
using WideEquiArgs = ::jodo_wide_equi_args;
const char* check_wide_equi(const WideEquiArgs& a);
cudaError_t launch_wide_equi(const WideEquiArgs& a, int num_sms, cudaStream_t st);           // wide_equi.cu
bool wide_attn_mol_ok(const WideAttnArgs& a);                                   // wide_attn.cu: molecule-staged variant
cudaError_t launch_wide_attn_mol(const WideAttnArgs& a, cudaStream_t st);
cudaError_t launch_wide_equi_out(const int* grp_row0, const int* grp_len, const int* row_j, const float* c3, int ldc, int nslots,
                                 const uint8_t* extra, const int* row_pair, int X, float coord_scale, const float* pos_in,
                                 float* pos_out, int Nn, cudaStream_t st);
cudaError_t launch_wide_head_out(const Plan& p, const float* x, int ldx, int hw, const float* w4, const float* b4, int ch,
                                 int both, float* out_dense, cudaStream_t st);

// ---- property classifier (egnn.cu)
cudaError_t launch_egnn_edge_in(const Plan& p, const float* pos4, const float* PQ, int ldpq, int H, const float* wr, void* img,
                                cudaStream_t st);
cudaError_t launch_egnn_agg(const int* grp_row0, const int* grp_len, const void* M16, int ldm, int H, const float* wa, float ba,
                            float* agg, int ldagg, int Nn, cudaStream_t st);
cudaError_t launch_mol_sum(const float* x, int ldx, int W, const int* mol_start, int B, float* out, int ldo, cudaStream_t st);

}  // namespace jodo
