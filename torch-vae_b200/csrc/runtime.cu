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

static int g_reserved_sms = 0;

// SMs the persistent kernels size their grids for.  In data-parallel runs with LARGE gradient buckets a few SMs can be
// left to the NCCL kernels (vae_set_reserved_sms): the persistent conv / output-block kernels hold one CTA with
// ~200 KB of shared memory and most registers on EVERY SM, so a concurrent all-reduce otherwise only advances in the
// gaps between them.
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  const int left = n - g_reserved_sms;
  return left >= 8 ? left : 8;
}

}  // namespace vae

/* Leave `n` SMs out of the grids of the persistent kernels (for concurrent NCCL kernels in data-parallel runs with large
 * gradient buckets); returns the previous setting.  0 (default): every SM. */
extern "C" int vae_set_reserved_sms(int n) {
  int old = vae::g_reserved_sms;
  vae::g_reserved_sms = n < 0 ? 0 : n;
  return old;
}

extern "C" int vae_abi_version(void) { return VAE_ABI_VERSION; }
extern "C" const char* vae_last_error(void) { return vae::g_err; }
