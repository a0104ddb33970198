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

constexpr int kMaxDevices = 64;
static uintptr_t g_arena_lo[kMaxDevices] = {0}, g_arena_hi[kMaxDevices] = {0};

cudaError_t zero_scratch(void* p, size_t bytes, cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if (a >= g_arena_lo[dev] && a + bytes <= g_arena_hi[dev]) return cudaSuccess;   // the caller cleared it with its arena
  }
  return cudaMemsetAsync(p, 0, bytes, st);
}

}  // namespace vae

/* Register [base, base + bytes) on the CURRENT device as a scratch arena the caller keeps zeroed: buffers inside it that
 * a kernel accumulates into (BatchNorm statistics, split-K workspaces, loss scratch) are NOT cleared by the library --
 * the caller clears what it handed out once per step (ops.ZeroArena).  base = NULL removes the registration. */
extern "C" int vae_set_zero_arena(void* base, long long bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= vae::kMaxDevices || bytes < 0) {
    vae::set_error("vae_set_zero_arena: bad device or size");
    return VAE_ERR_ARG;
  }
  vae::g_arena_lo[dev] = reinterpret_cast<uintptr_t>(base);
  vae::g_arena_hi[dev] = base ? reinterpret_cast<uintptr_t>(base) + (uintptr_t)bytes : 0;
  return VAE_OK;
}

/* Leave `n` SMs out of the grids of the persistent kernels (for concurrent NCCL kernels in data-parallel runs with large
 * gradient buckets); returns the previous setting.  0 (default): every SM. */
extern "C" int vae_set_reserved_sms(int n) {
  int old = vae::g_reserved_sms;
  vae::g_reserved_sms = n < 0 ? 0 : n;
  return old;
}

extern "C" int vae_abi_version(void) { return VAE_ABI_VERSION; }
extern "C" const char* vae_last_error(void) { return vae::g_err; }
