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

// Same update with the hyper-parameters read from DEVICE memory: hyper[group] = {lr, beta1, beta2,
// weight_decay, 1 - beta1, 1 - beta2, -, -} (the complements are formed in double precision by whoever
// writes the table, as torch forms `1 - beta` from Python floats).  A captured CUDA graph then follows a per-batch schedule (OneCycleLR cycles lr and
// beta1, train.py:232-238,659) instead of replaying the values of capture time.
constexpr int kAdamHyperStride = 8;

struct AdamGroupTable {
  int grp[kAdamMaxTensors];
};

__global__ void __launch_bounds__(256)
adamw_dev_kernel(const __grid_constant__ AdamTable T, const __grid_constant__ AdamGroupTable G,
                 const float* __restrict__ hyper, float eps) {
  const int t = blockIdx.y;
  const int n = T.n[t];
  const float* h = hyper + kAdamHyperStride * G.grp[t];
  const float lr = h[0], beta1 = h[1], beta2 = h[2], wd = h[3];
  const float omb1 = h[4], omb2 = h[5];
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
    const float mi = fmaf(beta1, m[i], omb1 * gi);
    const float vi = fmaf(beta2, v[i], omb2 * gi * gi);
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = fmaf(-step_size, mi / denom, p[i] * decay);
  }
}

// torch.optim.lr_scheduler.OneCycleLR (two phases) advanced on the device.  One thread per group.
//   phase 0: steps [0, pct_start*total - 1]      lr initial -> max      momentum max_m -> base_m
//   phase 1: steps (.., total - 1]               lr max -> min          momentum base_m -> max_m
// anneal 0 = cos: end + (start - end)/2 * (cos(pi*pct) + 1); 1 = linear: (end - start)*pct + start.
struct OneCycleTable {
  double initial_lr[8], max_lr[8], min_lr[8], base_m[8], max_m[8];
};

__global__ void onecycle_kernel(float* hyper, long long* counter, const __grid_constant__ OneCycleTable T, int groups,
                                long long total_steps, double pct_start, int anneal, int cycle_momentum) {
  const int g = threadIdx.x;
  const long long step = counter[0] + 1;   // last_epoch after this scheduler.step()
  if (g < groups) {
    const double end0 = pct_start * (double)total_steps - 1.0, end1 = (double)total_steps - 1.0;
    const double s = (double)(step <= total_steps ? step : total_steps);   // torch raises past total_steps; the device clamps
    double lr, mom;
    auto ann = [&](double a, double b, double pct) {
      return anneal == 0 ? b + (a - b) / 2.0 * (cos(3.14159265358979323846 * pct) + 1.0) : (b - a) * pct + a;
    };
    if (s <= end0) {
      const double pct = s / end0;
      lr = ann(T.initial_lr[g], T.max_lr[g], pct);
      mom = ann(T.max_m[g], T.base_m[g], pct);
    } else {
      const double pct = (s - end0) / (end1 - end0);
      lr = ann(T.max_lr[g], T.min_lr[g], pct);
      mom = ann(T.base_m[g], T.max_m[g], pct);
    }
    hyper[kAdamHyperStride * g] = (float)lr;
    if (cycle_momentum) {
      hyper[kAdamHyperStride * g + 1] = (float)mom;
      hyper[kAdamHyperStride * g + 4] = (float)(1.0 - mom);
    }
  }
  __syncthreads();
  if (g == 0) counter[0] = step;
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

/* vae_adamw_step with the hyper-parameters in DEVICE memory: hyper[8*g .. 8*g+5] = {lr, beta1, beta2,
 * weight_decay, 1 - beta1, 1 - beta2} of parameter group g (fp32, rows of 8), group_of[i] (host) = group of tensor i.  The lr / wd fields of
 * the tensor records are ignored.  CUDA-graph replays follow whatever vae_onecycle_step (or a host copy)
 * last wrote into `hyper`.
 * replaces: optimizer.step() (train.py:656) under a per-batch scheduler (train.py:232-238,659) */
extern "C" int vae_adamw_step_dev(const vae_adamw_tensor* tensors, const int* group_of, int count, const float* hyper,
                                  int n_groups, double eps, vae_stream_t stream) {
  VAE_REQUIRE(tensors && group_of && hyper && count >= 0 && n_groups > 0, "adamw_step_dev: bad argument");
  cudaStream_t st = as_stream(stream);
  for (int base = 0; base < count; base += kAdamMaxTensors) {
    const int c = count - base < kAdamMaxTensors ? count - base : kAdamMaxTensors;
    AdamTable T{};
    AdamGroupTable G{};
    long long max_n = 0;
    for (int i = 0; i < c; ++i) {
      const vae_adamw_tensor& e = tensors[base + i];
      VAE_REQUIRE(e.p && e.g && e.m && e.v && e.step && e.n > 0 && e.n < (1LL << 31), "adamw_step_dev: bad tensor %d", base + i);
      VAE_REQUIRE(group_of[base + i] >= 0 && group_of[base + i] < n_groups, "adamw_step_dev: bad group of tensor %d", base + i);
      T.p[i] = (float*)e.p; T.g[i] = (const float*)e.g; T.m[i] = (float*)e.m; T.v[i] = (float*)e.v;
      T.step[i] = e.step; T.n[i] = (int)e.n;
      G.grp[i] = group_of[base + i];
      if (e.n > max_n) max_n = e.n;
    }
    int bx = (int)((max_n + 256 * 4 - 1) / (256 * 4));
    if (bx < 1) bx = 1;
    if (bx > 512) bx = 512;
    adamw_dev_kernel<<<dim3(bx, c), 256, 0, st>>>(T, G, hyper, (float)eps);
    int rc = check_launch("adamw_dev");
    if (rc != 0) return rc;
    adamw_bump_steps_kernel<<<1, kAdamMaxTensors, 0, st>>>(T, c);
    rc = check_launch("adamw_bump_steps");
    if (rc != 0) return rc;
  }
  return VAE_OK;
}

/* One scheduler.step() of torch.optim.lr_scheduler.OneCycleLR (two-phase) on the device: counter[0] (int64,
 * = last_epoch) += 1, then hyper[8*g] = lr and, with cycle_momentum, hyper[8*g+1] = beta1, hyper[8*g+4] = 1 - beta1 for g < n_groups (<= 8).
 * per_group (host) = n_groups records {initial_lr, max_lr, min_lr, base_momentum, max_momentum} as doubles.
 * replaces: scheduler.step() (train.py:659) of the OneCycleLR built at train.py:232-238 */
extern "C" int vae_onecycle_step(float* hyper, long long* counter, const double* per_group, int n_groups,
                                 long long total_steps, double pct_start, int anneal, int cycle_momentum,
                                 vae_stream_t stream) {
  VAE_REQUIRE(hyper && counter && per_group && n_groups > 0 && n_groups <= 8 && total_steps > 1 && pct_start > 0.0 &&
                  pct_start * (double)total_steps > 1.0 && (anneal == 0 || anneal == 1),
              "onecycle_step: bad argument");
  OneCycleTable T{};
  for (int g = 0; g < n_groups; ++g) {
    T.initial_lr[g] = per_group[5 * g]; T.max_lr[g] = per_group[5 * g + 1]; T.min_lr[g] = per_group[5 * g + 2];
    T.base_m[g] = per_group[5 * g + 3]; T.max_m[g] = per_group[5 * g + 4];
  }
  onecycle_kernel<<<1, 32, 0, as_stream(stream)>>>(hyper, counter, T, n_groups, total_steps, pct_start, anneal, cycle_momentum);
  return check_launch("onecycle");
}
