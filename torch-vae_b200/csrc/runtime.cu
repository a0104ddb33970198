// Error plumbing and library identification for the vae_b200 C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace vae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return VAE_ERR_CUDA;
  }
  return VAE_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

}  // namespace vae

extern "C" int vae_abi_version(void) { return VAE_ABI_VERSION; }
extern "C" const char* vae_last_error(void) { return vae::g_err; }
