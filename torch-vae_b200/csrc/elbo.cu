// Fused ELBO: BCE(sigmoid(logits), x) with the reference's log clamp (-100),
// KL(N(mu, e^lv) || N(0,1)), the three LossOutput scalars, and the gradient.
// HBM-bound: one coalesced 16-byte-vectorised pass, warp-shuffle + one double
// atomic per block for the reduction, last-block finalisation (no second launch).
//
// Reference semantics reproduced exactly (models.py:208-225 on top of ATen):
//   p      = 1 / (1 + exp(-logit))                       aten::sigmoid (models.py:81)
//   bce    = (t-1)*max(log1p(-p), -100) - t*max(log(p), -100)     binary_cross_entropy
//   dbce/dp= (p - t) / max(p*(1-p), 1e-12)               binary_cross_entropy_backward
//   dp/dl  = p*(1-p)                                     sigmoid_backward
// so the saturation of fp32 sigmoid for |logit| >~ 17 behaves as in the reference.
#include "common.cuh"

namespace vae {

__device__ __forceinline__ float bce_term(float p, float t) {
  float l1 = fmaxf(log1pf(-p), -100.0f);
  float l0 = fmaxf(logf(p), -100.0f);
  return (t - 1.0f) * l1 - t * l0;
}

__device__ __forceinline__ float prob_of(float v, int input_is_prob) {
  return input_is_prob ? v : 1.0f / (1.0f + expf(-v));
}

template <typename T>
__global__ void __launch_bounds__(256)
elbo_fwd_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                const float* __restrict__ mu, const float* __restrict__ logvar,
                float* __restrict__ losses, double* __restrict__ scratch, long long n_elem, int B,
                int L, float kld_weight, int input_is_prob) {
  __shared__ double red[2][8];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  float bce = 0.0f;
  const long long n_vec = n_elem / 8;
  for (long long i = tid; i < n_vec; i += nthreads) {
    Vec8<T> lg;
    Vec8<float> tg;
    lg.load(logits + i * 8);
    tg.load(target + i * 8);
    float part = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) part += bce_term(prob_of(lg.v[k], input_is_prob), tg.v[k]);
    bce += part;
  }
  for (long long i = n_vec * 8 + tid; i < n_elem; i += nthreads)
    bce += bce_term(prob_of(to_f(logits[i]), input_is_prob), target[i]);

  float kl = 0.0f;
  const long long n_lat = (long long)B * L;
  for (long long i = tid; i < n_lat; i += nthreads) {
    float m = mu[i], lv = logvar[i];
    kl += 1.0f + lv - m * m - expf(lv);
  }

  double dbce = warp_sum((double)bce), dkl = warp_sum((double)kl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = dbce;
    red[1][warp] = dkl;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      s0 += red[0][w];
      s1 += red[1][w];
    }
    atomicAdd(&scratch[0], s0);
    atomicAdd(&scratch[1], s1);
    __threadfence();
    unsigned int* counter = reinterpret_cast<unsigned int*>(&scratch[2]);
    unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      double t0 = atomicAdd(&scratch[0], 0.0), t1 = atomicAdd(&scratch[1], 0.0);
      float recon = (float)(t0 / (double)n_elem);
      float kld = (float)(-0.5 * t1 / (double)B);
      losses[0] = recon + kld_weight * kld;
      losses[1] = recon;
      losses[2] = -kld;  // reference returns the sign-flipped value (models.py:224)
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
elbo_bwd_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                const float* __restrict__ mu, const float* __restrict__ logvar,
                const float* __restrict__ gscale, T* __restrict__ dlogits,
                float* __restrict__ dmu, float* __restrict__ dlogvar, long long n_elem, int B,
                int L, float kld_weight, int input_is_prob) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const float g = gscale != nullptr ? *gscale : 1.0f;
  const float gn = g / (float)n_elem;
  auto grad_of = [&](float v, float t) {
    float p = prob_of(v, input_is_prob);
    float pq = p * (1.0f - p);
    float d = gn * (p - t) / fmaxf(pq, 1e-12f);
    return input_is_prob ? d : d * pq;
  };
  const long long n_vec = n_elem / 8;
  for (long long i = tid; i < n_vec; i += nthreads) {
    Vec8<T> lg;
    Vec8<float> tg;
    lg.load(logits + i * 8);
    tg.load(target + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) lg.v[k] = grad_of(lg.v[k], tg.v[k]);
    lg.store(dlogits + i * 8);
  }
  for (long long i = n_vec * 8 + tid; i < n_elem; i += nthreads)
    dlogits[i] = from_f<T>(grad_of(to_f(logits[i]), target[i]));

  if (dmu != nullptr) {
    const long long n_lat = (long long)B * L;
    const float kb = g * kld_weight / (float)B;
    for (long long i = tid; i < n_lat; i += nthreads) {
      dmu[i] = kb * mu[i];
      dlogvar[i] = kb * 0.5f * (expf(logvar[i]) - 1.0f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward and gradient in ONE pass (north-star item 2): the loss is the root of the graph, so
// dlogits = (p - x) / n (clamp / saturation emulated), dmu and dlogvar are known as soon as the
// logits are read.  Bytes per element: logits + target in, dlogits out (3*H*W*sizeof
// with an fp32 target: 2 + 4 + 2 B in bf16).  The gradients are written for an upstream grad of 1;
// elbo_scale_grads_kernel rescales them in place only when the actual grad differs.
// FAST (bf16 path, tolerance 2e-2): bare MUFU exp2 / log2 / reciprocal -- four special-function
// ops per element instead of ~60 instructions of exact libm code, which made the exact kernel
// issue-bound at 1.7 TB/s.  The clamps and the fp32 saturation points are the same (p rounds to
// 1 for logit > 17.3 and log(0) = -inf is clamped to -100 in both); vectors holding a |logit| > 80
// (subnormal intermediates) take the exact code.  The fp32 instantiation keeps the exact functions
// of the two-kernel path (1e-5 parity).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// FAST: bare MUFU ops, no range fix-ups.  Valid for |logit| <= 80, where no intermediate is subnormal;
// elbo_vec8 routes anything larger to the exact code, so the saturation (p == 1 for logit > 17.3,
// log(0) -> -100) and the clamp behave as in the reference everywhere.
template <bool FAST, bool PROB>
__device__ __forceinline__ void elbo_point(float v, float t, float gn, float& bce, float& d) {
  constexpr int input_is_prob = PROB ? 1 : 0;
  if (FAST && !PROB) {
    const float e = mufu_ex2(v * -1.4426950408889634f);
    const float p = mufu_rcp(1.0f + e);
    const float om = 1.0f - p;
    const float l1 = fmaxf(mufu_lg2(om) * 0.6931471805599453f, -100.0f);
    const float l0 = fmaxf(mufu_lg2(p) * 0.6931471805599453f, -100.0f);
    bce = (t - 1.0f) * l1 - t * l0;
    const float pq = p * om, dd = gn * (p - t);
    d = pq >= 1e-12f ? dd : dd * pq * 1e12f;
  } else {
    const float p = prob_of(v, input_is_prob);
    bce = bce_term(p, t);
    const float pq = p * (1.0f - p);
    const float dd = gn * (p - t) / fmaxf(pq, 1e-12f);
    d = input_is_prob ? dd : dd * pq;
  }
}

// eight elements in place (v: logits in, dlogits out); returns their BCE sum.  MUFU path unless one of
// them is out of its range.
template <bool FAST, bool PROB>
__device__ __forceinline__ float elbo_vec8(float* v, const float* t, float gn) {
  float part = 0.0f;
  bool fast = false;
  if (FAST && !PROB) {
    float m = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, fabsf(v[k]));
    fast = m <= 80.0f;
  }
  if (fast) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float b_, d_;
      elbo_point<true, false>(v[k], t[k], gn, b_, d_);
      part += b_;
      v[k] = d_;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float b_, d_;
      elbo_point<false, PROB>(v[k], t[k], gn, b_, d_);
      part += b_;
      v[k] = d_;
    }
  }
  return part;
}

template <typename T, bool FAST, bool PROB>   // PROB: `logits` holds probabilities (compile-time: no per-element branch)
__global__ void __launch_bounds__(256)
elbo_fused_kernel(const T* __restrict__ logits, const float* __restrict__ target, const float* __restrict__ mu,
                  const float* __restrict__ logvar, float* __restrict__ losses, double* __restrict__ scratch,
                  T* __restrict__ dlogits, float* __restrict__ dmu, float* __restrict__ dlogvar, long long n_elem,
                  int B, int L, float kld_weight) {
  __shared__ double red[2][8];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const float gn = 1.0f / (float)n_elem;
  float bce = 0.0f;
  const long long n_vec = n_elem / 8;
  // two independent 8-element vectors in flight per thread
  for (long long i = tid; i < n_vec; i += 2 * nthreads) {
    const long long i1 = i + nthreads;
    const bool two = i1 < n_vec;
    Vec8<T> lg0, lg1;
    Vec8<float> tg0, tg1;
    lg0.load(logits + i * 8);
    tg0.load(target + i * 8);
    if (two) {
      lg1.load(logits + i1 * 8);
      tg1.load(target + i1 * 8);
    }
    float part = elbo_vec8<FAST, PROB>(lg0.v, tg0.v, gn);
    lg0.store(dlogits + i * 8);
    if (two) {
      part += elbo_vec8<FAST, PROB>(lg1.v, tg1.v, gn);
      lg1.store(dlogits + i1 * 8);
    }
    bce += part;
  }
  for (long long i = n_vec * 8 + tid; i < n_elem; i += nthreads) {
    float b_, d_;
    elbo_point<false, PROB>(to_f(logits[i]), target[i], gn, b_, d_);
    bce += b_;
    dlogits[i] = from_f<T>(d_);
  }

  float kl = 0.0f;
  const long long n_lat = (long long)B * L;
  const float kb = kld_weight / (float)B;
  for (long long i = tid; i < n_lat; i += nthreads) {
    const float m = mu[i], lv = logvar[i], ev = expf(lv);
    kl += 1.0f + lv - m * m - ev;
    dmu[i] = kb * m;
    dlogvar[i] = kb * 0.5f * (ev - 1.0f);
  }

  double dbce = warp_sum((double)bce), dkl = warp_sum((double)kl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = dbce;
    red[1][warp] = dkl;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      s0 += red[0][w];
      s1 += red[1][w];
    }
    atomicAdd(&scratch[0], s0);
    atomicAdd(&scratch[1], s1);
    __threadfence();
    unsigned int* counter = reinterpret_cast<unsigned int*>(&scratch[2]);
    unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      double t0 = atomicAdd(&scratch[0], 0.0), t1 = atomicAdd(&scratch[1], 0.0);
      float recon = (float)(t0 / (double)n_elem);
      float kld = (float)(-0.5 * t1 / (double)B);
      losses[0] = recon + kld_weight * kld;
      losses[1] = recon;
      losses[2] = -kld;
    }
  }
}

// in-place rescale of the cached gradients by the upstream grad; exits at once when it is 1
template <typename T>
__global__ void __launch_bounds__(256)
elbo_scale_grads_kernel(const float* __restrict__ gscale, T* __restrict__ dlogits, float* __restrict__ dmu,
                        float* __restrict__ dlogvar, long long n_elem, long long n_lat) {
  const float g = *gscale;
  if (g == 1.0f) return;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long n_vec = n_elem / 8;
  for (long long i = tid; i < n_vec; i += nthreads) {
    Vec8<T> v;
    v.load(dlogits + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] *= g;
    v.store(dlogits + i * 8);
  }
  for (long long i = n_vec * 8 + tid; i < n_elem; i += nthreads) dlogits[i] = from_f<T>(to_f(dlogits[i]) * g);
  for (long long i = tid; i < n_lat; i += nthreads) {
    dmu[i] *= g;
    dlogvar[i] *= g;
  }
}

static int elbo_grid(long long n_elem) {
  long long blocks = (n_elem / 8 + 255) / 256;
  long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace vae

using namespace vae;

extern "C" int vae_elbo_fwd(const void* logits, const float* target, const float* mu,
                            const float* logvar, float* losses, double* scratch, long long n_elem,
                            int B, int L, float kld_weight, int dtype, int input_is_prob,
                            vae_stream_t stream) {
  VAE_REQUIRE(logits && target && mu && logvar && losses && scratch && n_elem > 0 && B > 0 && L > 0,
              "elbo_fwd: bad argument");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(scratch, 4 * sizeof(double), st));
  int grid = elbo_grid(n_elem);
  if (dtype == VAE_BF16)
    elbo_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)logits, target, mu, logvar, losses,
                                                scratch, n_elem, B, L, kld_weight, input_is_prob);
  else
    elbo_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)logits, target, mu, logvar, losses,
                                                 scratch, n_elem, B, L, kld_weight, input_is_prob);
  return check_launch("elbo_fwd");
}

extern "C" int vae_elbo_bwd(const void* logits, const float* target, const float* mu,
                            const float* logvar, const float* gscale, void* dlogits, float* dmu,
                            float* dlogvar, long long n_elem, int B, int L, float kld_weight,
                            int dtype, int input_is_prob, vae_stream_t stream) {
  VAE_REQUIRE(logits && target && dlogits && n_elem > 0 && B > 0 && L > 0 &&
                  ((dmu == nullptr) == (dlogvar == nullptr)) && (dmu == nullptr || (mu && logvar)),
              "elbo_bwd: bad argument");
  cudaStream_t st = as_stream(stream);
  int grid = elbo_grid(n_elem);
  if (dtype == VAE_BF16)
    elbo_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)logits, target, mu, logvar, gscale,
                                                (bf16*)dlogits, dmu, dlogvar, n_elem, B, L,
                                                kld_weight, input_is_prob);
  else
    elbo_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)logits, target, mu, logvar, gscale,
                                                 (float*)dlogits, dmu, dlogvar, n_elem, B, L,
                                                 kld_weight, input_is_prob);
  return check_launch("elbo_bwd");
}

/* vae_elbo_fwd and vae_elbo_bwd in one pass for an upstream gradient of 1: losses as vae_elbo_fwd,
 * dlogits / dmu / dlogvar as vae_elbo_bwd with *gscale == 1.  vae_elbo_scale_grads applies the
 * actual upstream gradient afterwards (a no-op launch when it is 1).
 * replaces: models.py:208-216 and their autograd nodes (train.py:650) */
extern "C" int vae_elbo_fused(const void* logits, const float* target, const float* mu, const float* logvar,
                              float* losses, double* scratch, void* dlogits, float* dmu, float* dlogvar,
                              long long n_elem, int B, int L, float kld_weight, int dtype, int input_is_prob,
                              vae_stream_t stream) {
  VAE_REQUIRE(logits && target && mu && logvar && losses && scratch && dlogits && dmu && dlogvar && n_elem > 0 &&
                  B > 0 && L > 0,
              "elbo_fused: bad argument");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(scratch, 4 * sizeof(double), st));
  // every thread walks the same number (+-1) of vector pairs; 4 blocks per SM stay resident
  long long blocks = (n_elem / 16 + 255) / 256;
  const long long cap = (long long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int grid = (int)blocks;
#define VAE_ELBO_GO(T, FAST, PROB)                                                                               \
  elbo_fused_kernel<T, FAST, PROB><<<grid, 256, 0, st>>>((const T*)logits, target, mu, logvar, losses, scratch,   \
                                                         (T*)dlogits, dmu, dlogvar, n_elem, B, L, kld_weight)
  if (dtype == VAE_BF16) {
    if (input_is_prob) VAE_ELBO_GO(bf16, true, true);
    else VAE_ELBO_GO(bf16, true, false);
  } else {
    if (input_is_prob) VAE_ELBO_GO(float, false, true);
    else VAE_ELBO_GO(float, false, false);
  }
#undef VAE_ELBO_GO
  return check_launch("elbo_fused");
}

/* dlogits, dmu, dlogvar *= *gscale, in place; returns without touching memory when *gscale == 1 */
extern "C" int vae_elbo_scale_grads(const float* gscale, void* dlogits, float* dmu, float* dlogvar, long long n_elem,
                                    long long n_lat, int dtype, vae_stream_t stream) {
  VAE_REQUIRE(gscale && dlogits && dmu && dlogvar && n_elem > 0 && n_lat > 0, "elbo_scale_grads: bad argument");
  cudaStream_t st = as_stream(stream);
  const int grid = elbo_grid(n_elem);
  if (dtype == VAE_BF16)
    elbo_scale_grads_kernel<bf16><<<grid, 256, 0, st>>>(gscale, (bf16*)dlogits, dmu, dlogvar, n_elem, n_lat);
  else
    elbo_scale_grads_kernel<float><<<grid, 256, 0, st>>>(gscale, (float*)dlogits, dmu, dlogvar, n_elem, n_lat);
  return check_launch("elbo_scale_grads");
}
