// C ABI of libjodo_b200.so (declared in include/jodo_b200.h): plain pointers and sizes, no torch types.
#include "../../include/jodo_b200.h"
#include "kernels.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

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

// ---- constant-table stream guard (kernels.h) ---------------------------------------------------------
namespace jodo {
namespace {
struct ConstGuard { cudaStream_t last; cudaEvent_t ev; bool have; };
ConstGuard g_guard[MAX_DEVICES] = {};
std::mutex g_guard_mu;
bool capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
}
}  // namespace
cudaError_t const_tables_acquire(cudaStream_t st) {
  if (capturing(st)) return cudaSuccess;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return cudaSuccess;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  ConstGuard& g = g_guard[dev];
  if (g.have && g.last != st) return cudaStreamWaitEvent(st, g.ev, 0);
  return cudaSuccess;
}
cudaError_t const_tables_release(cudaStream_t st) {
  if (capturing(st)) return cudaSuccess;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return cudaSuccess;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  ConstGuard& g = g_guard[dev];
  if (!g.ev) {
    cudaError_t e = cudaEventCreateWithFlags(&g.ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(g.ev, st);
  if (e != cudaSuccess) return e;
  g.last = st;
  g.have = true;
  return cudaSuccess;
}
}  // namespace jodo

extern "C" {

const char* jodo_last_error_string(void) { return g_err; }
int jodo_abi_version(void) { return JODO_ABI_VERSION; }

int jodo_rowlinear(const float* A, int lda, int M, int K, const void* Wimg, const float* bias, void* C, int ldc,
                   int N, int NT, int act_in, int epi, int act_out, const float* aux, int ld_aux, const float* gate,
                   int ld_gate, const int* row_mol, int out_f16, const int* skip_if_zero, void* stream) {
  jodo::RowLinearArgs a{A, lda, M, K, static_cast<const float*>(Wimg), bias, C, ldc, N, NT, act_in, epi, act_out, aux, ld_aux,
                        gate, ld_gate, row_mol, out_f16, skip_if_zero};
  if (const char* m = jodo::check_rowlinear(a)) return fail(m);
  cudaError_t e = jodo::launch_rowlinear(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? JODO_OK : cuda_fail(e, "jodo_rowlinear");
}

#define JODO_LAUNCH(expr, name)                                  \
  do {                                                           \
    cudaError_t e__ = (expr);                                    \
    return e__ == cudaSuccess ? JODO_OK : cuda_fail(e__, name);  \
  } while (0)

static int num_sms() {                      // of the CURRENT device (cached per device)
  static int n[jodo::MAX_DEVICES] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= jodo::MAX_DEVICES) dev = 0;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
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
                      int* mol_bad, void* stream) {
  if (!p || p->Nn <= 0 || kin < 2 * inn) return fail("jodo_gather_nodes: bad sizes");
  JODO_LAUNCH(jodo::launch_gather_nodes(xh, cond_x, *p, inn, kin, xin, pos4, mol_bad, S(stream)), "jodo_gather_nodes");
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
int jodo_row0_linear(const float* A, int K, const void* Wimg, int NT, int N, const float* bias, int act_in, int act_out,
                     const float* aux, float* out, const int* run_if_zero, void* stream) {
  if (!A || !Wimg || !out || K <= 0 || (K % 64) || K > 8192 || N <= 0 || NT <= 0 || (NT % 8) || (N % NT))
    return fail("jodo_row0_linear: bad arguments (K a multiple of 64 up to 8192, N a multiple of NT)");
  if (reinterpret_cast<uintptr_t>(Wimg) & 15) return fail("jodo_row0_linear: the weight image must be 16-byte aligned");
  if (act_in != JODO_ACT_NONE && act_in != JODO_ACT_SILU) return fail("jodo_row0_linear: act_in must be none or SiLU");
  if (act_out != JODO_ACT_NONE && act_out != JODO_ACT_GELU) return fail("jodo_row0_linear: act_out must be none or GELU");
  JODO_LAUNCH(jodo::launch_row0_linear(A, K, Wimg, NT, N, bias, act_in, act_out, aux, out, run_if_zero, S(stream)), "jodo_row0_linear");
}
int jodo_uniform_flag(const float* rows, int B, int T, int* nonuni, void* stream) {
  if (!rows || !nonuni || B <= 0 || T <= 0) return fail("jodo_uniform_flag: bad arguments");
  JODO_LAUNCH(jodo::launch_uniform_flag(rows, B, T, nonuni, S(stream)), "jodo_uniform_flag");
}
int jodo_com(float* pos4, const jodo_plan* p, void* stream) { JODO_LAUNCH(jodo::launch_com(pos4, *p, S(stream)), "jodo_com"); }
int jodo_node_out(const float* pos4, const float* atom_pred, int ldp, const jodo_plan* p, int* nan_flag, const int* mol_bad,
                  int inn, float* out_dense, void* stream) {
  cudaError_t e = cudaMemsetAsync(nan_flag, 0, sizeof(int), S(stream));
  if (e != cudaSuccess) return cuda_fail(e, "jodo_node_out");
  e = jodo::launch_nan_flag(pos4, p->Nn, mol_bad, p->B, nan_flag, S(stream));
  if (e != cudaSuccess) return cuda_fail(e, "jodo_node_out");
  JODO_LAUNCH(jodo::launch_node_out(pos4, atom_pred, ldp, *p, nan_flag, mol_bad, inn, out_dense, S(stream)), "jodo_node_out");
}
int jodo_sym_edges(const float* tmp, float* out, int B, int N, int ch, void* stream) {
  JODO_LAUNCH(jodo::launch_sym_edges(tmp, out, B, N, ch, S(stream)), "jodo_sym_edges");
}

int jodo_ancestral_update(const float* x, const float* pred, const float* raw_pos, const float* raw_feat,
                          const float* node_mask, const float* edge_x, const float* edge_pred, const float* raw_edge,
                          const float* edge_mask, int B, int N, int F, int ch, float c_x, float c_pred, float sigma,
                          const float* coef_dev, float* x_new, float* x_mean, float* e_new, float* e_mean, void* stream) {
  if (B <= 0 || N <= 0 || F <= 3 || ch <= 0) return fail("jodo_ancestral_update: bad sizes");
  if (!x || !pred || !raw_pos || !raw_feat || !node_mask || !edge_x || !edge_pred || !raw_edge || !edge_mask || !x_new ||
      !x_mean || !e_new || !e_mean)
    return fail("jodo_ancestral_update: null pointer");
  JODO_LAUNCH(jodo::launch_ancestral_update(x, pred, raw_pos, raw_feat, node_mask, edge_x, edge_pred, raw_edge, edge_mask, B, N,
                                            F, ch, c_x, c_pred, sigma, coef_dev, 0, 0ull, 0u, x_new, x_mean, e_new, e_mean, S(stream)),
              "jodo_ancestral_update");
}
int jodo_ancestral_update_philox(const float* x, const float* pred, const float* node_mask, const float* edge_x,
                                 const float* edge_pred, const float* edge_mask, int B, int N, int F, int ch, float c_x,
                                 float c_pred, float sigma, const float* coef_dev, unsigned long long seed, unsigned int step,
                                 float* x_new, float* x_mean, float* e_new, float* e_mean, void* stream) {
  if (B <= 0 || N <= 0 || F <= 3 || ch <= 0) return fail("jodo_ancestral_update_philox: bad sizes");
  if (!x || !pred || !node_mask || !edge_x || !edge_pred || !edge_mask || !x_new || !x_mean || !e_new || !e_mean)
    return fail("jodo_ancestral_update_philox: null pointer");
  JODO_LAUNCH(jodo::launch_ancestral_update(x, pred, nullptr, nullptr, node_mask, edge_x, edge_pred, nullptr, edge_mask, B, N, F, ch,
                                            c_x, c_pred, sigma, coef_dev, 1, seed, step, x_new, x_mean, e_new, e_mean, S(stream)),
              "jodo_ancestral_update_philox");
}
int jodo_philox_normal(unsigned long long n4, unsigned long long seed, unsigned int step, unsigned int stream_id, float* out,
                       void* stream) {
  if (!out || n4 == 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return fail("jodo_philox_normal: bad arguments");
  JODO_LAUNCH(jodo::launch_philox_normal(n4, seed, step, stream_id, out, S(stream)), "jodo_philox_normal");
}

int jodo_saturation_count(unsigned long long* out, int reset) {
  unsigned int a = 0, b = 0, c = 0;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = jodo::sat_count_imglinear(&a, reset != 0);
  if (e == cudaSuccess) e = jodo::sat_count_edge_update(&b, reset != 0);
  if (e == cudaSuccess) e = jodo::sat_count_wide_ffn(&c, reset != 0);
  if (e != cudaSuccess) return cuda_fail(e, "jodo_saturation_count");
  if (out) *out = (unsigned long long)a + b + c;
  return JODO_OK;
}

int jodo_dpm_update(const float* x_start, const float* pos_in, int ld_pos, const float* pred0, const float* pred1,
                    const float* raw_pos, const float* node_mask, const float* edge_start, const float* edge_pred0,
                    const float* edge_pred1, int B, int N, int F, int ch, const float* coef_dev, float* x_out, float* edge_out,
                    void* stream) {
  if (B <= 0 || N <= 0 || F <= 3 || ch <= 0 || ld_pos < 3) return fail("jodo_dpm_update: bad sizes");
  if (!x_start || !pos_in || !pred0 || !raw_pos || !node_mask || !edge_start || !edge_pred0 || !coef_dev || !x_out || !edge_out)
    return fail("jodo_dpm_update: null pointer");
  if ((pred1 == nullptr) != (edge_pred1 == nullptr)) return fail("jodo_dpm_update: pred1 / edge_pred1 must come together");
  JODO_LAUNCH(jodo::launch_dpm_update(x_start, pos_in, ld_pos, pred0, pred1, raw_pos, node_mask, edge_start, edge_pred0,
                                      edge_pred1, B, N, F, ch, coef_dev, x_out, edge_out, S(stream)),
              "jodo_dpm_update");
}

int jodo_pack_weights(const jodo_pack_item* items_dev, const int* blk_item_dev, const int* blk_first_dev, int n_blocks, void* stream) {
  if (n_blocks < 0 || (n_blocks > 0 && (!items_dev || !blk_item_dev || !blk_first_dev))) return fail("jodo_pack_weights: bad arguments");
  JODO_LAUNCH(jodo::launch_pack_items(items_dev, blk_item_dev, blk_first_dev, n_blocks, S(stream)), "jodo_pack_weights");
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
  if (!a->p.row_pair) return fail("jodo_attn: the plan needs row_pair (the edge state is stored per unordered pair)");
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
  if (!a->p.row_pair) return fail("jodo_equi: the plan needs row_pair (the edge state is stored per unordered pair)");
  // The CTA-pair kernel (equi2.cu: cta_group::2, all weights resident, two-stage pipeline) is correct but measured no faster
  // than the single-CTA kernel (0.41-0.45 ms vs 0.40 ms per launch at QM9 B = 2500: three pair rendezvous per tile at
  // ~1.2 us each eat the overlap).  It stays selectable for A/B runs: JODO_EQUI_PAIR=1.
  static const bool pair = std::getenv("JODO_EQUI_PAIR") != nullptr;
  if (a->w2_img32 && pair) JODO_LAUNCH(jodo::launch_equi2(*a, num_sms(), S(stream)), "jodo_equi");
  JODO_LAUNCH(jodo::launch_equi(*a, num_sms(), S(stream)), "jodo_equi");
}
int jodo_equi_compose(const jodo_equi_compose_item* items_dev, int n_blocks, const float* tab_row0, const int* nonuni, void* stream) {
  if (!items_dev || n_blocks <= 0 || n_blocks > 64 || !tab_row0 || !nonuni) return fail("jodo_equi_compose: bad arguments");
  JODO_LAUNCH(jodo::launch_equi_compose(items_dev, n_blocks, tab_row0, nonuni, S(stream)), "jodo_equi_compose");
}
int jodo_equi_lin(const jodo_equi_lin_args* a, void* stream) {
  if (!a) return fail("jodo_equi_lin: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  if (a->ldab < a->p.Nn) return fail("jodo_equi_lin: bad strides");
  if (!a->p.row_pair) return fail("jodo_equi_lin: the plan needs row_pair (the edge state is stored per unordered pair)");
  if (!a->e16 || !a->pos_in || !a->pos_out || !a->AB || !a->extra || !a->win_img || !a->wce_img || !a->consts || !a->nonuni)
    return fail("jodo_equi_lin: null buffer");
  JODO_LAUNCH(jodo::launch_equi_lin(*a, num_sms(), S(stream)), "jodo_equi_lin");
}
int jodo_edge_head(const jodo_edge_head_args* a, void* stream) {
  if (!a) return fail("jodo_edge_head: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  JODO_LAUNCH(jodo::launch_edge_head(*a, num_sms(), S(stream)), "jodo_edge_head");
}

// ---- wide path (nf = 384)
static bool img_ok(const void* img, int K, int col, int W) {
  return img && !(reinterpret_cast<uintptr_t>(img) & 127) && K > 0 && K % 64 == 0 && col >= 0 && col % 8 == 0 && W % 8 == 0 && col + W <= K;
}
int jodo_wide_embed_in(const jodo_wide_embed_args* a, void* stream) {
  if (!a) return fail("jodo_wide_embed_in: null args");
  if (const char* m = check_plan(a->p)) return fail(m);
  if (a->ch < 1 || a->ed < 0 || a->ed % 8 || a->ed + 2 * a->ch > a->K) return fail("jodo_wide_embed_in: bad sizes");
  if (!img_ok(a->img, a->K, 0, a->K) || !a->edge_x || !a->extra || !a->dist_flag || !a->tab || !a->gbf) return fail("jodo_wide_embed_in: bad buffers");
  if (a->cond_x && !a->cond_edge_x) return fail("jodo_wide_embed_in: cond_x needs cond_edge_x");     /* 2-D models: cond_edge_x alone */
  JODO_LAUNCH(jodo::launch_wide_embed_in(*a, S(stream)), "jodo_wide_embed_in");
}
int jodo_wide_put(const float* src, int ld, int M, int W, const int* valid, void* img1, int K1, int col1, void* img2, int K2,
                  int col2, void* img3, int K3, int col3, void* stream) {
  if (!src || M <= 0 || W <= 0 || (ld % 4) || ld < W) return fail("jodo_wide_put: bad source");
  if (!img_ok(img1, K1, col1, W) || (img2 && !img_ok(img2, K2, col2, W)) || (img3 && !img_ok(img3, K3, col3, W))) return fail("jodo_wide_put: bad image");
  JODO_LAUNCH(jodo::launch_wide_put(src, ld, M, W, valid, img1, K1, col1, img2, K2, col2, img3, K3, col3, S(stream)), "jodo_wide_put");
}
int jodo_wide_dist(const jodo_plan* p, const float* pos4, const float* tab, int ld_tab, int off_gbf, const float* gbf,
                   int ld_gbf, int ed, void* img1, int K1, int col1, void* img2, int K2, int col2, void* stream) {
  if (!p) return fail("jodo_wide_dist: null plan");
  if (const char* m = check_plan(*p)) return fail(m);
  if (!pos4 || !tab || !gbf || ed <= 0 || ed % 8 || ld_gbf < ed || ld_gbf % 4) return fail("jodo_wide_dist: bad arguments");
  if (!img_ok(img1, K1, col1, ed) || (img2 && !img_ok(img2, K2, col2, ed))) return fail("jodo_wide_dist: bad image");
  JODO_LAUNCH(jodo::launch_wide_dist(*p, pos4, tab, ld_tab, off_gbf, gbf, ld_gbf, ed, img1, K1, col1, img2, K2, col2, S(stream)),
              "jodo_wide_dist");
}
int jodo_wide_ln(const jodo_wide_ln_args* a, void* stream) {
  if (!a) return fail("jodo_wide_ln: null args");
  if (a->M <= 0 || a->W <= 0 || a->W % 8 || a->W > 512 || a->Kimg < a->W || a->Kimg % 64 || a->Kimg > 512) return fail("jodo_wide_ln: bad sizes");
  if ((a->x_f16 && (a->ldx % 8)) || (a->y_f16 && a->y && ((a->ldy % 8) || (a->y2 && (a->ldy2 % 8))))) return fail("jodo_wide_ln: fp16 rows need strides % 8 == 0");
  if (!a->x || !a->tab || !a->row_mol || (a->ldx % 4) || (a->ld_tab % 4) || (a->off_shift % 4) || (a->off_scale % 4)) return fail("jodo_wide_ln: bad inputs");
  if (a->y && (a->ldy % 4)) return fail("jodo_wide_ln: bad addend stride");
  if (a->y2 && (!a->y || (a->ldy2 % 4))) return fail("jodo_wide_ln: y2 needs y");
  if (!a->out_img && !a->out32) return fail("jodo_wide_ln: no output");
  if (a->out32 && (a->ldo % 4)) return fail("jodo_wide_ln: bad output stride");
  JODO_LAUNCH(jodo::launch_wide_ln(*a, S(stream)), "jodo_wide_ln");
}
int jodo_wide_edge_ffn(const jodo_wide_ffn_args* a, void* stream) {
  if (!a) return fail("jodo_wide_edge_ffn: null args");
  if (const char* m = jodo::check_wide_ffn(*a)) return fail(m);
  JODO_LAUNCH(jodo::launch_wide_ffn(*a, num_sms(), S(stream)), "jodo_wide_edge_ffn");
}
int jodo_wide_equi(const jodo_wide_equi_args* a, void* stream) {
  if (!a) return fail("jodo_wide_equi: null args");
  if (const char* m = jodo::check_wide_equi(*a)) return fail(m);
  JODO_LAUNCH(jodo::launch_wide_equi(*a, num_sms(), S(stream)), "jodo_wide_equi");
}
int jodo_wide_attn(const jodo_wide_attn_args* a, void* stream) {
  if (!a) return fail("jodo_wide_attn: null args");
  if (a->Nn <= 0 || a->D <= 0 || a->H <= 0 || a->H > 32 || a->X < 0 || a->X >= a->H || a->X > 8 || a->D % a->H || a->sc <= 0 ||
      (a->D / a->H) % 4 || (a->ldq % 4) || (a->ldg % 4) || (a->k_off % 2) || (a->v_off % 4) || (a->g1_off % 4))
    return fail("jodo_wide_attn: bad sizes (needs D / H % 4 == 0 and 8-byte aligned row parts, q/k parts padded to an even width)");
  if (a->max_gl < 1 || a->max_gl > 255) return fail("jodo_wide_attn: max_gl must be in [1, 255]");
  if (!a->grp_row0 || !a->grp_len || !a->row_j || !a->qkv || !a->G || !a->extra || !a->hnode) return fail("jodo_wide_attn: null buffer");
  JODO_LAUNCH(jodo::launch_wide_attn(*a, S(stream)), "jodo_wide_attn");
}
int jodo_wide_equi_out(const int* grp_row0, const int* grp_len, const int* row_j, const float* c3, int ldc, int nslots,
                       const uint8_t* extra, const int* row_pair, int X, float coord_scale, const float* pos_in4, float* pos_out4,
                       int Nn, void* stream) {
  if (!grp_row0 || !grp_len || !row_j || !c3 || !extra || !pos_in4 || !pos_out4 || Nn <= 0 || X < 0 || X > 2 || nslots < 1 ||
      ldc < 4 * nslots || ldc < 1 + X)
    return fail("jodo_wide_equi_out: bad arguments");
  JODO_LAUNCH(jodo::launch_wide_equi_out(grp_row0, grp_len, row_j, c3, ldc, nslots, extra, row_pair, X, coord_scale, pos_in4, pos_out4, Nn, S(stream)),
              "jodo_wide_equi_out");
}
int jodo_wide_head_out(const jodo_plan* p, const float* x, int ldx, int hw, const float* w4, const float* b4, int ch,
                       int both, float* out_dense, void* stream) {
  if (!p) return fail("jodo_wide_head_out: null plan");
  if (const char* m = check_plan(*p)) return fail(m);
  if (!x || !w4 || !b4 || !out_dense || hw <= 0 || ch < 1 || ldx < 2 * hw) return fail("jodo_wide_head_out: bad arguments");
  if (ch > 8 || (hw % 4) || (ldx % 4) || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w4)) & 15))
    return fail("jodo_wide_head_out: at most 8 outputs; hw, ldx multiples of 4; x and w4 16-byte aligned");
  JODO_LAUNCH(jodo::launch_wide_head_out(*p, x, ldx, hw, w4, b4, ch, both, out_dense, S(stream)), "jodo_wide_head_out");
}

// ---- property classifier
int jodo_egnn_edge_in(const jodo_plan* p, const float* pos4, const float* PQ, int ldpq, int H, const float* wr, void* img,
                      void* stream) {
  if (!p) return fail("jodo_egnn_edge_in: null plan");
  if (const char* m = check_plan(*p)) return fail(m);
  if (H <= 0 || H % 64 || H > 256 || ldpq < 2 * H || ldpq % 4) return fail("jodo_egnn_edge_in: H must be a multiple of 64 up to 256, ldpq >= 2 H");
  if (!pos4 || !PQ || !wr || !img || (reinterpret_cast<uintptr_t>(img) & 127) || (reinterpret_cast<uintptr_t>(PQ) & 15) ||
      (reinterpret_cast<uintptr_t>(wr) & 15))
    return fail("jodo_egnn_edge_in: bad buffers");
  JODO_LAUNCH(jodo::launch_egnn_edge_in(*p, pos4, PQ, ldpq, H, wr, img, S(stream)), "jodo_egnn_edge_in");
}
int jodo_egnn_agg(const int* grp_row0, const int* grp_len, const void* M16, int ldm, int H, const float* wa, float ba,
                  float* agg, int ldagg, int Nn, void* stream) {
  if (!grp_row0 || !grp_len || !M16 || !agg || Nn <= 0) return fail("jodo_egnn_agg: null buffer");
  if (H <= 0 || H % 8 || H > 256 || ldm < H || ldm % 8 || ldagg < H || ldagg % 4) return fail("jodo_egnn_agg: bad sizes (H % 8 == 0, H <= 256)");
  if ((reinterpret_cast<uintptr_t>(M16) | reinterpret_cast<uintptr_t>(agg) | reinterpret_cast<uintptr_t>(wa)) & 15) return fail("jodo_egnn_agg: pointers must be 16-byte aligned");
  JODO_LAUNCH(jodo::launch_egnn_agg(grp_row0, grp_len, M16, ldm, H, wa, ba, agg, ldagg, Nn, S(stream)), "jodo_egnn_agg");
}
int jodo_mol_sum(const float* x, int ldx, int W, const int* mol_start, int B, float* out, int ldo, void* stream) {
  if (!x || !mol_start || !out || B <= 0 || W <= 0 || ldx < W || ldo < W) return fail("jodo_mol_sum: bad arguments");
  JODO_LAUNCH(jodo::launch_mol_sum(x, ldx, W, mol_start, B, out, ldo, S(stream)), "jodo_mol_sum");
}

}  // extern "C"
