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

int jodo_rowlinear(const float* A, int lda, int M, int K, const float* Wimg, const float* bias, float* C, int ldc,
                   int N, int NT, int act_in, int epi, int act_out, const float* aux, int ld_aux, const float* gate,
                   int ld_gate, const int* row_mol, void* stream) {
  jodo::RowLinearArgs a{A, lda, M, K, Wimg, bias, C, ldc, N, NT, act_in, epi, act_out, aux, ld_aux, gate, ld_gate, row_mol};
  if (const char* m = jodo::check_rowlinear(a)) return fail(m);
  cudaError_t e = jodo::launch_rowlinear(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? JODO_OK : cuda_fail(e, "jodo_rowlinear");
}

}  // extern "C"
