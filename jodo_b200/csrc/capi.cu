// C ABI of libjodo_b200.so (declared in include/jodo_b200.h): plain pointers and sizes, no torch types.
#include "../../include/jodo_b200.h"
#include "kernels.h"

#include <cstdio>
#include <cstring>

namespace {
thread_local char g_err[512] = "";
int fail(const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return JODO_ERR_ARG;
}
int cuda_fail(cudaError_t e, const char* where) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return JODO_ERR_CUDA;
}
}  // namespace

extern "C" {

const char* jodo_last_error_string(void) { return g_err; }
int jodo_abi_version(void) { return JODO_ABI_VERSION; }

int jodo_rowlinear(const float* A, int lda, int M, int K, const void* Wimg, const float* bias, void* C, int ldc,
                   int N, int NT, int act_in, int epi, int act_out, const float* aux, int ld_aux, const float* gate,
                   int ld_gate, const int* row_mol, int out_f16, const int* only_row0_if_zero, void* stream) {
  jodo::RowLinearArgs a{A, lda, M, K, static_cast<const float*>(Wimg), bias, C, ldc, N, NT, act_in, epi, act_out, aux, ld_aux,
                        gate, ld_gate, row_mol, out_f16, only_row0_if_zero};
  if (const char* m = jodo::check_rowlinear(a)) return fail(m);
  cudaError_t e = jodo::launch_rowlinear(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? JODO_OK : cuda_fail(e, "jodo_rowlinear");
}

#define JODO_LAUNCH(expr, name)                                  \
  do {                                                           \
    cudaError_t e__ = (expr);                                    \
    return e__ == cudaSuccess ? JODO_OK : cuda_fail(e__, name);  \
  } while (0)

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

int jodo_time_features(const float* nl, const float* w8, float* feat64, int B, void* stream) {
  if (B <= 0) return fail("jodo_time_features: B <= 0");
  JODO_LAUNCH(jodo::launch_time_features(nl, w8, feat64, B, S(stream)), "jodo_time_features");
}
int jodo_cond_in(const float* ctx, const float* w0, const float* b0, float* out, int rows, int D, void* stream) {
  if (rows <= 0 || D <= 0) return fail("jodo_cond_in: bad sizes");
  JODO_LAUNCH(jodo::launch_cond_in(ctx, w0, b0, out, rows, D, S(stream)), "jodo_cond_in");
}
int jodo_gather_nodes(const float* xh, const float* cond_x, const jodo_plan* p, int inn, int kin, float* xin, float* pos4,
                      void* stream) {
  if (!p || p->Nn <= 0 || kin < 2 * inn) return fail("jodo_gather_nodes: bad sizes");
  JODO_LAUNCH(jodo::launch_gather_nodes(xh, cond_x, *p, inn, kin, xin, pos4, S(stream)), "jodo_gather_nodes");
}
int jodo_ln_mod(int D, const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab, int off_gate,
                int off_shift, int off_scale, const jodo_plan* p, float* out, int ldo, void* stream) {
  if (D != 256 && D != 384) return fail("jodo_ln_mod: D must be 256 or 384");
  JODO_LAUNCH(jodo::launch_ln_mod(D, x, ldx, y, ldy, tab, ld_tab, off_gate, off_shift, off_scale, *p, out, ldo, S(stream)),
              "jodo_ln_mod");
}
int jodo_ln_mod_img(const float* x, int ldx, const float* y, int ldy, const float* tab, int ld_tab, int off_gate,
                    int off_shift, int off_scale, const jodo_plan* p, float* out32, int ldo, void* out_img, void* y_img,
                    const int* nonuni, void* stream) {
  if (!p || !x || !tab || !out_img) return fail("jodo_ln_mod_img: null pointer");
  if ((ldx % 4) || (y && (ldy % 4)) || (out32 && (ldo % 4)) || (ld_tab % 4) || (off_gate % 4) || (off_shift % 4) || (off_scale % 4))
    return fail("jodo_ln_mod_img: strides and offsets must be multiples of 4");
  JODO_LAUNCH(jodo::launch_ln_mod_img(x, ldx, y, ldy, tab, ld_tab, off_gate, off_shift, off_scale, *p, out32, ldo, out_img,
                                      y_img, nonuni, S(stream)),
              "jodo_ln_mod_img");
}
int jodo_imglinear(const jodo_imglinear_args* a, void* stream) {
  if (!a) return fail("jodo_imglinear: null args");
  if (const char* m = jodo::check_imglinear(*a)) return fail(m);
  JODO_LAUNCH(jodo::launch_imglinear(*a, num_sms(), S(stream)), "jodo_imglinear");
}
int jodo_act_image(const float* rows, int ld, int M, int K, int act, void* img, void* stream) {
  if (!rows || !img || M <= 0 || K <= 0 || (K % 64) || (ld % 4)) return fail("jodo_act_image: bad arguments");
  JODO_LAUNCH(jodo::launch_act_image(rows, ld, M, K, act, img, S(stream)), "jodo_act_image");
}
int jodo_uniform_flag(const float* rows, int B, int T, int* nonuni, void* stream) {
  if (!rows || !nonuni || B <= 0 || T <= 0) return fail("jodo_uniform_flag: bad arguments");
  JODO_LAUNCH(jodo::launch_uniform_flag(rows, B, T, nonuni, S(stream)), "jodo_uniform_flag");
}
int jodo_com(float* pos4, const jodo_plan* p, void* stream) { JODO_LAUNCH(jodo::launch_com(pos4, *p, S(stream)), "jodo_com"); }
int jodo_node_out(const float* pos4, const float* atom_pred, int ldp, const jodo_plan* p, int* nan_flag, int inn,
                  float* out_dense, void* stream) {
  cudaError_t e = cudaMemsetAsync(nan_flag, 0, sizeof(int), S(stream));
  if (e != cudaSuccess) return cuda_fail(e, "jodo_node_out");
  e = jodo::launch_nan_flag(pos4, p->Nn, nan_flag, S(stream));
  if (e != cudaSuccess) return cuda_fail(e, "jodo_node_out");
  JODO_LAUNCH(jodo::launch_node_out(pos4, atom_pred, ldp, *p, nan_flag, inn, out_dense, S(stream)), "jodo_node_out");
}
int jodo_sym_edges(const float* tmp, float* out, int B, int N, int ch, void* stream) {
  JODO_LAUNCH(jodo::launch_sym_edges(tmp, out, B, N, ch, S(stream)), "jodo_sym_edges");
}

int jodo_ancestral_update(const float* x, const float* pred, const float* raw_pos, const float* raw_feat,
                          const float* node_mask, const float* edge_x, const float* edge_pred, const float* raw_edge,
                          const float* edge_mask, int B, int N, int F, int ch, float c_x, float c_pred, float sigma,
                          float* x_new, float* x_mean, float* e_new, float* e_mean, void* stream) {
  if (B <= 0 || N <= 0 || F <= 3 || ch <= 0) return fail("jodo_ancestral_update: bad sizes");
  if (!x || !pred || !raw_pos || !raw_feat || !node_mask || !edge_x || !edge_pred || !raw_edge || !edge_mask || !x_new ||
      !x_mean || !e_new || !e_mean)
    return fail("jodo_ancestral_update: null pointer");
  JODO_LAUNCH(jodo::launch_ancestral_update(x, pred, raw_pos, raw_feat, node_mask, edge_x, edge_pred, raw_edge, edge_mask, B, N,
                                            F, ch, c_x, c_pred, sigma, x_new, x_mean, e_new, e_mean, S(stream)),
              "jodo_ancestral_update");
}

static const char* check_plan(const jodo_plan& p) {
  if (p.B <= 0 || p.Nn <= 0 || p.n_tiles <= 0 || p.N <= 0) return "plan: empty";
  if (!p.node_mol || !p.node_dense || !p.mol_start || !p.row_g || !p.row_j || !p.row_meta || !p.tile_ngroups || !p.row_mol)
    return "plan: null pointer";
  return nullptr;
}

int jodo_edge_embed(const jodo_edge_embed_args* a, void* stream) {
  if (!a) return fail("jodo_edge_embed: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  if (a->ch < 1 || a->ch > 8 || 2 * a->ch > 32) return fail("jodo_edge_embed: edge_ch must be in [1, 8]");
  if ((a->cond_x == nullptr) != (a->cond_edge_x == nullptr)) return fail("jodo_edge_embed: cond_x / cond_edge_x must come together");
  cudaError_t e = jodo::launch_dist_flag(*a, S(stream));
  if (e != cudaSuccess) return cuda_fail(e, "jodo_edge_embed(dist_flag)");
  JODO_LAUNCH(jodo::launch_edge_embed(*a, num_sms(), S(stream)), "jodo_edge_embed");
}
int jodo_attn(const jodo_attn_args* a, void* stream) {
  if (!a) return fail("jodo_attn: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  if (a->ldq < a->p.Nn || a->ld_tab % 4 || a->tab_off % 4) return fail("jodo_attn: bad strides");
  if (!a->e16 || !a->hnode) return fail("jodo_attn: null buffer");
  JODO_LAUNCH(jodo::launch_attn(*a, num_sms(), S(stream)), "jodo_attn");
}
int jodo_edge_update(const jodo_edge_update_args* a, void* stream) {
  if (!a) return fail("jodo_edge_update: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  if (a->ldp < a->p.Nn || a->ld_tab % 4 || a->tab_off % 4) return fail("jodo_edge_update: bad strides");
  if (a->r != 2 && a->r != 4) return fail("jodo_edge_update: mlp_ratio must be 2 or 4");
  if (!a->e32 || !a->e16 || !a->eh) return fail("jodo_edge_update: null buffer");
  if (a->eh_col < 0 || a->eh_col + a->ce > 192) return fail("jodo_edge_update: edge-hidden slice out of range");
  JODO_LAUNCH(jodo::launch_edge_update(*a, num_sms(), S(stream)), "jodo_edge_update");
}
int jodo_equi(const jodo_equi_args* a, void* stream) {
  if (!a) return fail("jodo_equi: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  if (a->ldab < a->p.Nn || a->ld_tab % 4 || a->tab_off % 4) return fail("jodo_equi: bad strides");
  JODO_LAUNCH(jodo::launch_equi(*a, num_sms(), S(stream)), "jodo_equi");
}
int jodo_edge_head(const jodo_edge_head_args* a, void* stream) {
  if (!a) return fail("jodo_edge_head: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  JODO_LAUNCH(jodo::launch_edge_head(*a, num_sms(), S(stream)), "jodo_edge_head");
}

}  // extern "C"
