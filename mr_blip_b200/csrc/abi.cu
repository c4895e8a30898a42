// Error plumbing for the C-ABI (include/mrblip_b200.h).
#include "common.cuh"
#include <string.h>

static thread_local char g_err[256] = "";

extern "C" int mrb_set_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", static_cast<int>(e), cudaGetErrorString(e));
  return MRB_ERR_CUDA;
}
extern "C" const char* mrb_last_error(void) { return g_err; }
extern "C" int mrb_abi_version(void) { return 1; }
