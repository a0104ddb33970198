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
  VAE_CUDA(cudaMemsetAsync(scratch, 0, 4 * sizeof(double), st));
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
