// Fused AdamW for the reference's two parameter groups (train.py:210-228, stepped at :656).
// torch.optim.AdamW's foreach path spends ~14 launches of ~20 us per step on 28 small tensors
// (0.78 M parameters); here ONE launch updates all of them: blockIdx.y selects the tensor from a
// by-value table, blockIdx.x strides its elements.  Arithmetic follows torch's single-tensor
// AdamW (decoupled weight decay, bias-corrected moments, eps added after the sqrt):
//   p   *= 1 - lr * wd
//   m    = b1 * m + (1 - b1) * g            v = b2 * v + (1 - b2) * g * g
//   p   -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)          t = step + 1
// The per-parameter step counters stay on the device (CUDA-graph capturable) and are bumped by a
// second, one-block kernel so that every block of the update reads the same pre-step value.
#include "common.cuh"

namespace vae {

constexpr int kAdamMaxTensors = 32;

struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  float* step[kAdamMaxTensors];
  int n[kAdamMaxTensors];
  float lr[kAdamMaxTensors];
  float wd[kAdamMaxTensors];
};

__global__ void __launch_bounds__(256)
adamw_kernel(const __grid_constant__ AdamTable T, float beta1, float beta2, float omb1, float omb2, float eps) {
  const int t = blockIdx.y;
  const int n = T.n[t];
  const float lr = T.lr[t], wd = T.wd[t];
  const float step = T.step[t][0] + 1.0f;
  const float bc1 = 1.0f - powf(beta1, step);
  const float bc2 = 1.0f - powf(beta2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  const float decay = 1.0f - lr * wd;
  float* __restrict__ p = T.p[t];
  const float* __restrict__ g = T.g[t];
  float* __restrict__ m = T.m[t];
  float* __restrict__ v = T.v[t];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float gi = g[i];
    const float mi = fmaf(beta1, m[i], omb1 * gi);   // == lerp(m, g, 1 - b1) up to rounding
    const float vi = fmaf(beta2, v[i], omb2 * gi * gi);
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = fmaf(-step_size, mi / denom, p[i] * decay);
  }
}

__global__ void adamw_bump_steps_kernel(const __grid_constant__ AdamTable T, int count) {
  const int t = threadIdx.x;
  if (t < count) T.step[t][0] += 1.0f;
}

}  // namespace vae

using namespace vae;

/* One AdamW step over `count` fp32 tensors described by a HOST array (plain pointers and sizes):
 * p, g, m (exp_avg), v (exp_avg_sq) device pointers of n elements, step a device float scalar
 * (torch's capturable state layout), lr / wd per tensor.
 * replaces: optimizer.step() of torch.optim.AdamW over the encoder / decoder groups (train.py:656) */
extern "C" int vae_adamw_step(const vae_adamw_tensor* tensors, int count, double beta1, double beta2, double eps,
                              vae_stream_t stream) {
  VAE_REQUIRE(tensors && count >= 0, "adamw_step: bad argument");
  cudaStream_t st = as_stream(stream);
  for (int base = 0; base < count; base += kAdamMaxTensors) {
    const int c = count - base < kAdamMaxTensors ? count - base : kAdamMaxTensors;
    AdamTable T{};
    long long max_n = 0;
    for (int i = 0; i < c; ++i) {
      const vae_adamw_tensor& e = tensors[base + i];
      VAE_REQUIRE(e.p && e.g && e.m && e.v && e.step && e.n > 0 && e.n < (1LL << 31), "adamw_step: bad tensor %d", base + i);
      T.p[i] = (float*)e.p; T.g[i] = (const float*)e.g; T.m[i] = (float*)e.m; T.v[i] = (float*)e.v;
      T.step[i] = e.step; T.n[i] = (int)e.n; T.lr[i] = e.lr; T.wd[i] = e.wd;
      if (e.n > max_n) max_n = e.n;
    }
    // four elements per thread of the largest tensor: the loop is a chain of dependent loads, so few
    // iterations per thread matter more than few blocks (blocks past a small tensor's end exit at once)
    int bx = (int)((max_n + 256 * 4 - 1) / (256 * 4));
    if (bx < 1) bx = 1;
    if (bx > 512) bx = 512;
    adamw_kernel<<<dim3(bx, c), 256, 0, st>>>(T, (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),
                                               (float)eps);
    int rc = check_launch("adamw");
    if (rc != 0) return rc;
    adamw_bump_steps_kernel<<<1, kAdamMaxTensors, 0, st>>>(T, c);
    rc = check_launch("adamw_bump_steps");
    if (rc != 0) return rc;
  }
  return VAE_OK;
}
