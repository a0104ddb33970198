// HBM-bound glue kernels: BatchNorm finalize / apply / backward, layout
// transposes, reparameterisation, Philox, sigmoid, weight packing, piano-roll
// transforms.  All are coalesced, 16-byte vectorised along the NHWC channel axis
// and sized in multiples of the SM count.
#include "common.cuh"

namespace vae {

static inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)sm_count() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------- BN finalize
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean,
                                   float* running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, int C, double count,
                                   float momentum, float eps, int training) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, rstd;
  if (training) {
    double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
      double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean = running_mean[c];
    rstd = 1.0f / sqrtf(running_var[c] + eps);
  }
  float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
}

// u = (y - mean) * (gamma * rstd) + beta, in ATen's order of operations: the
// subtraction comes first and is exact where u changes sign (y ~ mean when beta ~ 0),
// so LeakyReLU's branch agrees with the reference's even for |u| ~ 1e-7.
__device__ __forceinline__ float bn_affine(float y, float mean, float scale, float beta) {
  return fmaf(y - mean, scale, beta);
}

// ------------------------------------------------------------ BN apply + lrelu
// Raw (unconverted) channel vector: loads stay 4 / 8 registers wide until they are consumed, so
// several of them can be in flight per thread without the fp32 copies occupying registers.
template <typename T, int V>
struct RawVec;
template <>
struct RawVec<bf16, 8> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float* v) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};
template <>
struct RawVec<float, 8> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void get(float* v) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <typename T>
struct RawVec<T, 1> {
  T r;
  __device__ __forceinline__ void load(const T* p) { r = p[0]; }
  __device__ __forceinline__ void get(float* v) const { v[0] = to_f(r); }
};
template <typename T>
struct BnUnroll {
  static constexpr int U = sizeof(T) == 2 ? 4 : 2;
};

// Thread layout of the vectorised BatchNorm passes: a block owns `lanes` pixel lanes times G = C/8
// channel groups; a thread keeps its 8 channels for the whole kernel (per-channel constants in
// registers) and has kBnUnroll independent 16-byte loads per tensor in flight per iteration.
constexpr int kBnUnroll = 4;

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_lrelu_vec8(const T* __restrict__ y, const float* __restrict__ scale,
                    const float* __restrict__ mean, const float* __restrict__ beta,
                    T* __restrict__ act, long long n_pix, int C, float slope) {
  const int G = C / 8;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  if (lane >= lanes) return;
  float sc[8], mu[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = scale[g * 8 + k];
    mu[k] = mean[g * 8 + k];
    be[k] = beta[g * 8 + k];
  }
  const long long step = (long long)gridDim.x * lanes;
  for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < n_pix; p0 += step * kBnUnroll) {
    RawVec<T, 8> v[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long long p = p0 + u * step;
      if (p < n_pix) v[u].load(y + p * C + g * 8);
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long long p = p0 + u * step;
      if (p < n_pix) {
        Vec8<T> o;
        v[u].get(o.v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float uu = bn_affine(o.v[k], mu[k], sc[k], be[k]);
          o.v[k] = uu > 0.0f ? uu : uu * slope;
        }
        o.store(act + p * C + g * 8);
      }
    }
  }
}

// finalize + apply in one launch: every thread derives the constants of ITS 8 channels from the
// fp64 statistics (a dozen flops), block 0 also publishes the (scale, shift, mean, rstd) vectors the
// backward pass needs and performs the running-statistics update exactly as bn_finalize_kernel.
template <typename T>
__global__ void __launch_bounds__(256)
bn_finalize_apply_vec8(const T* __restrict__ y, const double* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float* __restrict__ vec_out, T* __restrict__ act,
                       long long n_pix, int C, double count, float momentum, float eps, int training, float slope) {
  const int G = C / 8;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  float sc[8], mu[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = g * 8 + k;
    float mean, rstd;
    if (training) {
      const double m = stats[c] / count;
      double var = stats[C + c] / count - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)eps));
    } else {
      mean = running_mean[c];
      rstd = 1.0f / sqrtf(running_var[c] + eps);
    }
    mu[k] = mean;
    sc[k] = gamma[c] * rstd;
    be[k] = beta[c];
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float mean, rstd;
      if (training) {
        const double m = stats[c] / count;
        double var = stats[C + c] / count - m * m;
        if (var < 0.0) var = 0.0;
        mean = (float)m;
        rstd = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean != nullptr) {
          const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
          running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
      } else {
        mean = running_mean[c];
        rstd = 1.0f / sqrtf(running_var[c] + eps);
      }
      const float s_ = gamma[c] * rstd;
      vec_out[c] = s_;
      vec_out[C + c] = beta[c] - mean * s_;
      vec_out[2 * C + c] = mean;
      vec_out[3 * C + c] = rstd;
    }
    if (threadIdx.x == 0 && training && num_batches_tracked != nullptr) num_batches_tracked[0] += 1;
  }
  if (lane >= lanes) return;
  const long long step = (long long)gridDim.x * lanes;
  for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < n_pix; p0 += step * kBnUnroll) {
    RawVec<T, 8> v[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long long p = p0 + u * step;
      if (p < n_pix) v[u].load(y + p * C + g * 8);
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long long p = p0 + u * step;
      if (p < n_pix) {
        Vec8<T> o;
        v[u].get(o.v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float uu = bn_affine(o.v[k], mu[k], sc[k], be[k]);
          o.v[k] = uu > 0.0f ? uu : uu * slope;
        }
        o.store(act + p * C + g * 8);
      }
    }
  }
}

template <typename T>
__global__ void bn_apply_lrelu_scalar(const T* __restrict__ y, const float* __restrict__ scale,
                                      const float* __restrict__ mean, const float* __restrict__ beta,
                                      T* __restrict__ act, long long n, int C, float slope) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c = (int)(i % C);
    float u = bn_affine(to_f(y[i]), mean[c], scale[c], beta[c]);
    act[i] = from_f<T>(u > 0.0f ? u : u * slope);
  }
}

// ----------------------------------------------------- BN + lrelu backward
// Thread layout shared by both passes: a block owns `lanes` pixel lanes times
// G = C/V channel groups; a thread keeps its V channels for the whole kernel so
// the per-channel constants sit in registers.
template <typename T, int V>
struct ChanVec {
  float v[V];
  __device__ __forceinline__ void load(const T* p) {
    if constexpr (V == 8) {
      Vec8<T> t;
      t.load(p);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = t.v[k];
    } else {
      v[0] = to_f(p[0]);
    }
  }
  __device__ __forceinline__ void store(T* p) const {
    if constexpr (V == 8) {
      Vec8<T> t;
#pragma unroll
      for (int k = 0; k < 8; ++k) t.v[k] = v[k];
      t.store(p);
    } else {
      p[0] = from_f<T>(v[0]);
    }
  }
};

template <typename T, int V>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_kernel(const T* __restrict__ dact, const T* __restrict__ y,
                     const float* __restrict__ scale, const float* __restrict__ beta,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     double* __restrict__ dstats, long long n_pix, int C, float slope) {
  extern __shared__ float red[];  // [lanes][G][2V]
  const int G = C / V;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  float s1[V], s2[V];
#pragma unroll
  for (int k = 0; k < V; ++k) s1[k] = s2[k] = 0.0f;
  if (lane < lanes) {
    float sc[V], sh[V], mu[V], rs[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      sc[k] = scale[g * V + k];
      sh[k] = beta[g * V + k];
      mu[k] = mean[g * V + k];
      rs[k] = rstd[g * V + k];
    }
    const long long step = (long long)gridDim.x * lanes;
    constexpr int U = BnUnroll<T>::U;
    for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < n_pix; p0 += step * U) {
      RawVec<T, V> d[U], yy[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * step;
        if (p < n_pix) {
          d[u].load(dact + p * C + g * V);
          yy[u].load(y + p * C + g * V);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (p0 + u * step < n_pix) {
          float dv[V], yv[V];
          d[u].get(dv);
          yy[u].get(yv);
#pragma unroll
          for (int k = 0; k < V; ++k) {
            float uu = bn_affine(yv[k], mu[k], sc[k], sh[k]);
            float du = uu > 0.0f ? dv[k] : dv[k] * slope;
            float xh = (yv[k] - mu[k]) * rs[k];
            s1[k] += du;
            s2[k] = fmaf(du, xh, s2[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      red[(lane * G + g) * 2 * V + k] = s1[k];
      red[(lane * G + g) * 2 * V + V + k] = s2[k];
    }
  }
  __syncthreads();
  // 2C outputs, each summed over lanes
  for (int o = threadIdx.x; o < 2 * C; o += blockDim.x) {
    int which = o / C, c = o % C;
    int gg = c / V, k = c % V;
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += (double)red[(l * G + gg) * 2 * V + which * V + k];
    atomicAdd(&dstats[which * C + c], acc);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const T* __restrict__ dact, const T* __restrict__ y,
                    const float* __restrict__ scale, const float* __restrict__ beta,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const double* __restrict__ dstats,
                    T* __restrict__ dy, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    float* __restrict__ dbias, long long n_pix, int C, float slope, int flags) {
  extern __shared__ float red[];  // [lanes][G][V]
  const int training = flags & 1;
  const bool is_du = (flags & 2) != 0;   // `dact` already holds du (a conv epilogue applied LeakyReLU')
  const int G = C / V;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  const double inv_m = 1.0 / (double)n_pix;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = (float)dstats[c];
      dgamma[c] = (float)dstats[C + c];
    }
  }
  float sdy[V];
#pragma unroll
  for (int k = 0; k < V; ++k) sdy[k] = 0.0f;
  if (lane < lanes) {
    float sc[V], sh[V], mu[V], rs[V], c1[V], c2[V], a[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      int c = g * V + k;
      sc[k] = scale[c];
      sh[k] = beta[c];
      mu[k] = mean[c];
      rs[k] = rstd[c];
      a[k] = gamma[c] * rs[k];
      c1[k] = training ? (float)(dstats[c] * inv_m) : 0.0f;
      c2[k] = training ? (float)(dstats[C + c] * inv_m) : 0.0f;
    }
    const long long step = (long long)gridDim.x * lanes;
    constexpr int U = BnUnroll<T>::U;
    for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < n_pix; p0 += step * U) {
      RawVec<T, V> d[U], yy[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * step;
        if (p < n_pix) {
          d[u].load(dact + p * C + g * V);
          yy[u].load(y + p * C + g * V);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * step;
        if (p < n_pix) {
          float dv[V], yv[V];
          d[u].get(dv);
          yy[u].get(yv);
          ChanVec<T, V> o;
#pragma unroll
          for (int k = 0; k < V; ++k) {
            float uu = bn_affine(yv[k], mu[k], sc[k], sh[k]);
            float du = (is_du || uu > 0.0f) ? dv[k] : dv[k] * slope;
            float xh = (yv[k] - mu[k]) * rs[k];
            float r = a[k] * (du - c1[k] - xh * c2[k]);
            o.v[k] = r;
            sdy[k] += r;
          }
          o.store(dy + p * C + g * V);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) red[(lane * G + g) * V + k] = sdy[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int gg = c / V, k = c % V;
    float acc = 0.0f;
    for (int l = 0; l < lanes; ++l) acc += red[(l * G + gg) * V + k];
    atomicAdd(&dbias[c], acc);
  }
}

// Both passes in ONE cooperative launch for tensors that fit the L2 twice over (d_act + y <= ~96 MB): every block
// reduces its pixels, the grid meets at a counter, then every block applies the gradient to THE SAME pixels, which it
// now finds in the L2 / L1 instead of HBM.  dstats: [2C] sums + one 64-bit arrival counter (zeroed by the caller).
template <typename T, int V>
__global__ void __launch_bounds__(256, 2)
bn_bwd_fused_kernel(const T* __restrict__ dact, const T* __restrict__ y,
                    const float* __restrict__ scale, const float* __restrict__ beta,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, double* __restrict__ dstats,
                    T* __restrict__ dy, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    float* __restrict__ dbias, long long n_pix, int C, float slope, int training) {
  extern __shared__ float red[];  // [lanes][G][2V]
  const int G = C / V;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  const bool active = lane < lanes;
  float sc[V], sh[V], mu[V], rs[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    sc[k] = active ? scale[g * V + k] : 0.0f;
    sh[k] = active ? beta[g * V + k] : 0.0f;
    mu[k] = active ? mean[g * V + k] : 0.0f;
    rs[k] = active ? rstd[g * V + k] : 0.0f;
  }
  const long long step = (long long)gridDim.x * lanes;
  constexpr int U = BnUnroll<T>::U;
  // ---- pass 1: sum du, sum du * xhat
  {
    float s1[V], s2[V];
#pragma unroll
    for (int k = 0; k < V; ++k) s1[k] = s2[k] = 0.0f;
    if (active) {
      for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < n_pix; p0 += step * U) {
        RawVec<T, V> d[U], yy[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long p = p0 + u * step;
          if (p < n_pix) {
            d[u].load(dact + p * C + g * V);
            yy[u].load(y + p * C + g * V);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (p0 + u * step < n_pix) {
            float dv[V], yv[V];
            d[u].get(dv);
            yy[u].get(yv);
#pragma unroll
            for (int k = 0; k < V; ++k) {
              const float uu = bn_affine(yv[k], mu[k], sc[k], sh[k]);
              const float du = uu > 0.0f ? dv[k] : dv[k] * slope;
              const float xh = (yv[k] - mu[k]) * rs[k];
              s1[k] += du;
              s2[k] = fmaf(du, xh, s2[k]);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        red[(lane * G + g) * 2 * V + k] = s1[k];
        red[(lane * G + g) * 2 * V + V + k] = s2[k];
      }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * C; o += blockDim.x) {
      const int which = o / C, c = o % C;
      const int gg = c / V, k = c % V;
      double acc = 0.0;
      for (int l = 0; l < lanes; ++l) acc += (double)red[(l * G + gg) * 2 * V + which * V + k];
      atomicAdd(&dstats[which * C + c], acc);
    }
  }
  // ---- the whole grid has contributed (cooperative launch: all blocks are resident)
  __threadfence();
  __syncthreads();
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(dstats + 2 * C);
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1ULL);
    while (*reinterpret_cast<volatile unsigned long long*>(counter) < (unsigned long long)gridDim.x) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
  // ---- pass 2: dy (+ the small gradients)
  const double inv_m = 1.0 / (double)n_pix;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = (float)__ldcg(&dstats[c]);
      dgamma[c] = (float)__ldcg(&dstats[C + c]);
    }
  }
  float sdy[V];
#pragma unroll
  for (int k = 0; k < V; ++k) sdy[k] = 0.0f;
  __syncthreads();   // red is reused below
  if (active) {
    float c1[V], c2[V], a[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int c = g * V + k;
      a[k] = gamma[c] * rs[k];
      c1[k] = training ? (float)(__ldcg(&dstats[c]) * inv_m) : 0.0f;
      c2[k] = training ? (float)(__ldcg(&dstats[C + c]) * inv_m) : 0.0f;
    }
    for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < n_pix; p0 += step * U) {
      RawVec<T, V> d[U], yy[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * step;
        if (p < n_pix) {
          d[u].load(dact + p * C + g * V);
          yy[u].load(y + p * C + g * V);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * step;
        if (p < n_pix) {
          float dv[V], yv[V];
          d[u].get(dv);
          yy[u].get(yv);
          ChanVec<T, V> o;
#pragma unroll
          for (int k = 0; k < V; ++k) {
            const float uu = bn_affine(yv[k], mu[k], sc[k], sh[k]);
            const float du = uu > 0.0f ? dv[k] : dv[k] * slope;
            const float xh = (yv[k] - mu[k]) * rs[k];
            const float r = a[k] * (du - c1[k] - xh * c2[k]);
            o.v[k] = r;
            sdy[k] += r;
          }
          o.store(dy + p * C + g * V);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) red[(lane * G + g) * V + k] = sdy[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int gg = c / V, k = c % V;
    float acc = 0.0f;
    for (int l = 0; l < lanes; ++l) acc += red[(l * G + gg) * V + k];
    atomicAdd(&dbias[c], acc);
  }
}

// ------------------------------------------------------------------ transposes
// src[b][r][c] (R x Cc) -> dst[b][c][r]
template <typename TS, typename TD>
__global__ void transpose_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int R, int Cc) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * R * Cc;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = to_f(src[base + (long long)r * Cc + c]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) dst[base + (long long)c * R + r] = from_f<TD>(tile[threadIdx.x][i]);
  }
}

// bf16 -> bf16 with both extents multiples of 64: 16-byte global accesses on both sides
__global__ void __launch_bounds__(256)
transpose_bf16_tile64(const bf16* __restrict__ src, bf16* __restrict__ dst, int R, int Cc) {
  __shared__ __align__(16) bf16 tile[64][72];
  const long long base = (long long)blockIdx.z * R * Cc;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int r = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(&tile[r][ch * 8]) =
        *reinterpret_cast<const uint4*>(src + base + (long long)(r0 + r) * Cc + c0 + ch * 8);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int c = i & 63, ch = i >> 6;   // consecutive threads: consecutive columns (conflict-free 2-byte reads)
    uint4 pk;
    bf16* o = reinterpret_cast<bf16*>(&pk);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = tile[ch * 8 + k][c];
    *reinterpret_cast<uint4*>(dst + base + (long long)(c0 + c) * R + r0 + ch * 8) = pk;
  }
}

// ------------------------------------------------------- reparam / rng / sigmoid
__global__ void philox_randn_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                    unsigned long long offset) {
  Philox rng(seed);
  long long n4 = (n + 3) / 4;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    uint4 r = rng(offset + (unsigned long long)i);
    float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i * 4 + k < n) out[i * 4 + k] = v[k];
  }
}

__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                   const float* __restrict__ eps_in, float* __restrict__ eps_out,
                                   float* __restrict__ z, long long n, unsigned long long seed,
                                   unsigned long long offset,
                                   const long long* __restrict__ dev_offset) {
  Philox rng(seed);
  if (dev_offset != nullptr) offset += (unsigned long long)(*dev_offset);
  long long n4 = (n + 3) / 4;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float e[4];
    if (eps_in == nullptr) {
      uint4 r = rng(offset + (unsigned long long)i);
      float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
      e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      long long j = i * 4 + k;
      if (j < n) {
        float ee = eps_in != nullptr ? eps_in[j] : e[k];
        if (eps_out != nullptr) eps_out[j] = ee;
        z[j] = fmaf(ee, expf(0.5f * logvar[j]), mu[j]);
      }
    }
  }
}

__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ eps,
                                   const float* __restrict__ logvar, float* __restrict__ dmu,
                                   float* __restrict__ dlogvar, long long n, int accumulate) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = dz[i];
    float dl = g * eps[i] * 0.5f * expf(0.5f * logvar[i]);
    if (accumulate) {
      dmu[i] += g;
      dlogvar[i] += dl;
    } else {
      dmu[i] = g;
      dlogvar[i] = dl;
    }
  }
}

template <typename T>
__global__ void sigmoid_kernel(const T* __restrict__ logits, float* __restrict__ probs, long long n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    probs[i] = 1.0f / (1.0f + expf(-to_f(logits[i])));
}

template <typename T>
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dprobs, const float* __restrict__ probs,
                                   T* __restrict__ dlogits, long long n, int accumulate) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float p = probs[i];
    float g = dprobs[i] * p * (1.0f - p);
    if (accumulate) g += to_f(dlogits[i]);
    dlogits[i] = from_f<T>(g);
  }
}

// ------------------------------------------------------------- weight packing
template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ wp,
                                        T* __restrict__ wt, int Co, int Ci) {
  long long n = (long long)Co * Ci * 9;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int t = (int)(i % 9);
    int ci = (int)((i / 9) % Ci);
    int co = (int)(i / (9LL * Ci));
    T v = from_f<T>(w[i]);
    if (wp != nullptr) wp[((long long)co * 9 + t) * Ci + ci] = v;
    if (wt != nullptr) wt[((long long)ci * 9 + t) * Co + co] = v;
  }
}

// ---------------------------------------------------------------- colsum
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ A, float* __restrict__ out, int M, int N,
                              int rows_per_block) {
  // block handles 32 columns x rows_per_block rows; threads (32, 8)
  __shared__ float red[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  int m0 = blockIdx.y * rows_per_block;
  int m1 = min(M, m0 + rows_per_block);
  float acc = 0.0f;
  if (n < N)
    for (int m = m0 + threadIdx.y; m < m1; m += 8) acc += to_f(A[(long long)m * N + n]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    atomicAdd(&out[n], s);
  }
}

// ------------------------------------------------------- piano-roll transform
__global__ void roll_transform_kernel(const uint8_t* __restrict__ rolls, float* __restrict__ out,
                                      int Hs, int Ws, int H, int W, int crop_mode,
                                      int max_pitch_shift, float thr, float mean, float inv_std,
                                      unsigned long long seed, unsigned long long offset) {
  const int b = blockIdx.y;
  // per-sample augmentation parameters: one Philox draw per sample
  Philox rng(seed);
  uint4 r = rng(offset + (unsigned long long)b, 0x5eedu);
  int oy, ox, shift = 0;
  if (crop_mode == 1) {
    oy = (int)(r.x % (uint32_t)(Hs - H + 1));
    ox = (int)(r.y % (uint32_t)(Ws - W + 1));
  } else {
    // torchvision CenterCrop: top = round((Hs - H) / 2)
    oy = (int)((Hs - H) / 2.0f + 0.5f);
    ox = (int)((Ws - W) / 2.0f + 0.5f);
    if (((Hs - H) & 1) && (((Hs - H) / 2) & 1) == 0) oy = (Hs - H) / 2;  // round-half-even
    if (((Ws - W) & 1) && (((Ws - W) / 2) & 1) == 0) ox = (Ws - W) / 2;
  }
  if (max_pitch_shift > 0) shift = (int)(r.z % (uint32_t)(2 * max_pitch_shift + 1)) - max_pitch_shift;
  const uint8_t* src = rolls + (long long)b * Hs * Ws;
  float* dst = out + (long long)b * H * W;
  const int n = H * W;
  if ((W & 7) == 0) {
    // a thread owns 8 consecutive time steps of one pitch row: 8 byte loads, two 16-byte stores
    const int wv = W >> 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * wv; i += gridDim.x * blockDim.x) {
      const int h = i / wv, w0 = (i - h * wv) << 3;
      const int hs = h + shift;  // pitch transpose inside the cropped window, zero fill
      float v[8];
      if (hs >= 0 && hs < H) {
        const uint8_t* row = src + (long long)(oy + hs) * Ws + (ox + w0);
        uint8_t raw[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) raw[k] = __ldg(row + k);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (float)raw[k] * (1.0f / 255.0f);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.0f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (thr >= 0.0f) v[k] = v[k] > thr ? 1.0f : 0.0f;
        v[k] = (v[k] - mean) * inv_std;
      }
      float4* o = reinterpret_cast<float4*>(dst + (long long)h * W + w0);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int h = i / W, w = i % W;
    int hs = h + shift;  // pitch transpose inside the cropped window, zero fill
    float v = 0.0f;
    if (hs >= 0 && hs < H) v = (float)src[(long long)(oy + hs) * Ws + (ox + w)] * (1.0f / 255.0f);
    if (thr >= 0.0f) v = v > thr ? 1.0f : 0.0f;
    dst[i] = (v - mean) * inv_std;
  }
}

}  // namespace vae

using namespace vae;

// ============================================================ C ABI wrappers
extern "C" int vae_bn_finalize(const double* stats, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float* scale, float* shift,
                               float* mean, float* rstd, int C, long long count, float momentum,
                               float eps, int training, vae_stream_t stream) {
  VAE_REQUIRE(C > 0 && gamma && beta && scale && shift && mean && rstd, "bn_finalize: bad argument");
  VAE_REQUIRE(training ? (stats != nullptr && count > 0) : (running_mean && running_var),
              "bn_finalize: training needs stats, eval needs running buffers");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(
      stats, gamma, beta, running_mean, running_var, scale, shift, mean, rstd, C, (double)count,
      momentum, eps, training);
  return check_launch("bn_finalize");
}

extern "C" int vae_bn_finalize_apply_lrelu(const void* y, const double* stats, const float* gamma, const float* beta,
                                           float* running_mean, float* running_var, long long* num_batches_tracked,
                                           float* vec_out, void* act, long long n_pix, int C, long long count,
                                           float momentum, float eps, int training, float slope, int dtype,
                                           vae_stream_t stream) {
  VAE_REQUIRE(y && gamma && beta && vec_out && act && n_pix > 0 && C > 0 && C % 8 == 0 && C / 8 <= 256,
              "bn_finalize_apply_lrelu: bad argument (C must be a multiple of 8, <= 2048)");
  VAE_REQUIRE(training ? (stats != nullptr && count > 0) : (running_mean && running_var),
              "bn_finalize_apply_lrelu: training needs stats, eval needs running buffers");
  cudaStream_t st = as_stream(stream);
  const int G = C / 8, lanes = 256 / G;
  long long blocks = (n_pix + (long long)lanes * kBnUnroll - 1) / ((long long)lanes * kBnUnroll);
  const long long cap = (long long)sm_count() * 8;
  const int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  if (dtype == VAE_BF16)
    bn_finalize_apply_vec8<bf16><<<grid, 256, 0, st>>>((const bf16*)y, stats, gamma, beta, running_mean, running_var,
                                                       num_batches_tracked, vec_out, (bf16*)act, n_pix, C, (double)count,
                                                       momentum, eps, training, slope);
  else
    bn_finalize_apply_vec8<float><<<grid, 256, 0, st>>>((const float*)y, stats, gamma, beta, running_mean, running_var,
                                                        num_batches_tracked, vec_out, (float*)act, n_pix, C, (double)count,
                                                        momentum, eps, training, slope);
  return check_launch("bn_finalize_apply_lrelu");
}

extern "C" int vae_bn_apply_lrelu(const void* y, const float* scale, const float* mean,
                                  const float* beta, void* act, long long n_pix, int C, float slope,
                                  int dtype, vae_stream_t stream) {
  VAE_REQUIRE(y && act && scale && mean && beta && n_pix > 0 && C > 0, "bn_apply_lrelu: bad argument");
  long long n = n_pix * C;
  cudaStream_t st = as_stream(stream);
  if (C % 8 == 0) {
    const int G = C / 8;
    if (G <= 256) {
      const int lanes = 256 / G;
      long long blocks = (n_pix + (long long)lanes * kBnUnroll - 1) / ((long long)lanes * kBnUnroll);
      const long long cap = (long long)sm_count() * 8;
      const int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
      if (dtype == VAE_BF16)
        bn_apply_lrelu_vec8<bf16><<<grid, 256, 0, st>>>((const bf16*)y, scale, mean, beta, (bf16*)act, n_pix, C, slope);
      else
        bn_apply_lrelu_vec8<float><<<grid, 256, 0, st>>>((const float*)y, scale, mean, beta, (float*)act, n_pix, C, slope);
      return check_launch("bn_apply_lrelu");
    }
  }
  {
    int grid = grid_for(n, 256);
    if (dtype == VAE_BF16)
      bn_apply_lrelu_scalar<bf16><<<grid, 256, 0, st>>>((const bf16*)y, scale, mean, beta, (bf16*)act, n, C, slope);
    else
      bn_apply_lrelu_scalar<float><<<grid, 256, 0, st>>>((const float*)y, scale, mean, beta, (float*)act, n, C, slope);
  }
  return check_launch("bn_apply_lrelu");
}

namespace {
struct BwdCfg {
  int V, G, lanes, grid;
};
static bool bwd_cfg(long long n_pix, int C, BwdCfg* cfg) {
  cfg->V = (C % 8 == 0) ? 8 : 1;
  cfg->G = C / cfg->V;
  if (cfg->G > 256) return false;
  cfg->lanes = 256 / cfg->G;
  long long blocks = (n_pix + cfg->lanes * 4 - 1) / (cfg->lanes * 4);
  long long cap = (long long)sm_count() * 2;
  cfg->grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  return true;
}
}  // namespace

extern "C" int vae_bn_lrelu_bwd_reduce(const void* dact, const void* y, const float* scale,
                                       const float* beta, const float* mean, const float* rstd,
                                       double* dstats, long long n_pix, int C, float slope,
                                       int dtype, vae_stream_t stream) {
  VAE_REQUIRE(dact && y && scale && beta && mean && rstd && dstats && n_pix > 0 && C > 0,
              "bn_lrelu_bwd_reduce: bad argument");
  BwdCfg cfg;
  VAE_REQUIRE(bwd_cfg(n_pix, C, &cfg), "bn_lrelu_bwd_reduce: unsupported channel count %d", C);
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(dstats, sizeof(double) * 2 * C, st));
  size_t smem = sizeof(float) * cfg.lanes * cfg.G * 2 * cfg.V;
#define LAUNCH(T, V)                                                                       \
  bn_bwd_reduce_kernel<T, V><<<cfg.grid, 256, smem, st>>>((const T*)dact, (const T*)y, scale, \
                                                          beta, mean, rstd, dstats, n_pix, C, slope)
  if (dtype == VAE_BF16) {
    if (cfg.V == 8) LAUNCH(bf16, 8); else LAUNCH(bf16, 1);
  } else {
    if (cfg.V == 8) LAUNCH(float, 8); else LAUNCH(float, 1);
  }
#undef LAUNCH
  return check_launch("bn_lrelu_bwd_reduce");
}

extern "C" int vae_bn_lrelu_bwd_apply(const void* dact, const void* y, const float* scale,
                                      const float* beta, const float* mean, const float* rstd,
                                      const float* gamma, const double* dstats, void* dy,
                                      float* dgamma, float* dbeta, float* dbias,
                                      long long n_pix, int C, float slope, int training, int dtype,
                                      vae_stream_t stream) {
  VAE_REQUIRE(dact && y && scale && beta && mean && rstd && gamma && dstats && dy && dgamma &&
                  dbeta && dbias && n_pix > 0 && C > 0,
              "bn_lrelu_bwd_apply: bad argument");
  BwdCfg cfg;
  VAE_REQUIRE(bwd_cfg(n_pix, C, &cfg), "bn_lrelu_bwd_apply: unsupported channel count %d", C);
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(dbias, sizeof(float) * C, st));
  size_t smem = sizeof(float) * cfg.lanes * cfg.G * cfg.V;
#define LAUNCH(T, V)                                                                          \
  bn_bwd_apply_kernel<T, V><<<cfg.grid, 256, smem, st>>>((const T*)dact, (const T*)y, scale,    \
                                                         beta, mean, rstd, gamma, dstats,      \
                                                         (T*)dy, dgamma, dbeta, dbias, n_pix, C, \
                                                         slope, training)
  if (dtype == VAE_BF16) {
    if (cfg.V == 8) LAUNCH(bf16, 8); else LAUNCH(bf16, 1);
  } else {
    if (cfg.V == 8) LAUNCH(float, 8); else LAUNCH(float, 1);
  }
#undef LAUNCH
  return check_launch("bn_lrelu_bwd_apply");
}

/* Both passes of the BatchNorm + LeakyReLU backward in ONE cooperative launch (reduce, grid-wide arrival counter,
 * apply on the same pixels, now cache-resident).  dstats: [2C + 1] doubles (the last one is the arrival counter),
 * zeroed by this call.  Returns VAE_ERR_UNSUPPORTED when the shape should take the two-pass path (tensor too large to
 * stay in the L2, or the grid cannot be made co-resident).
 * replaces: aten::leaky_relu_backward + aten::native_batch_norm_backward (train.py:650) */
extern "C" int vae_bn_lrelu_bwd_fused(const void* dact, const void* y, const float* scale, const float* beta,
                                      const float* mean, const float* rstd, const float* gamma, double* dstats, void* dy,
                                      float* dgamma, float* dbeta, float* dbias, long long n_pix, int C, float slope,
                                      int training, int dtype, vae_stream_t stream) {
  VAE_REQUIRE(dact && y && scale && beta && mean && rstd && gamma && dstats && dy && dgamma && dbeta && dbias && n_pix > 0 &&
                  C > 0,
              "bn_lrelu_bwd_fused: bad argument");
  BwdCfg cfg;
  const size_t elem = dtype == VAE_BF16 ? 2 : 4;
  if (!bwd_cfg(n_pix, C, &cfg) || (size_t)n_pix * C * elem * 2 > (size_t)96 << 20) return VAE_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const size_t smem = sizeof(float) * cfg.lanes * cfg.G * 2 * cfg.V;
  void* kern = nullptr;
  if (dtype == VAE_BF16) kern = cfg.V == 8 ? (void*)bn_bwd_fused_kernel<bf16, 8> : (void*)bn_bwd_fused_kernel<bf16, 1>;
  else kern = cfg.V == 8 ? (void*)bn_bwd_fused_kernel<float, 8> : (void*)bn_bwd_fused_kernel<float, 1>;
  int per_sm = 0;
  VAE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
  if (per_sm < 1) return VAE_ERR_UNSUPPORTED;
  int grid = cfg.grid;
  if (grid > per_sm * sm_count()) grid = per_sm * sm_count();
  VAE_CUDA(zero_scratch(dstats, sizeof(double) * (2 * C + 1), st));
  VAE_CUDA(zero_scratch(dbias, sizeof(float) * C, st));
  void* args[] = {(void*)&dact, (void*)&y, (void*)&scale, (void*)&beta, (void*)&mean, (void*)&rstd, (void*)&gamma,
                  (void*)&dstats, (void*)&dy, (void*)&dgamma, (void*)&dbeta, (void*)&dbias, (void*)&n_pix, (void*)&C,
                  (void*)&slope, (void*)&training};
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(256);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all blocks resident at once: the arrival counter cannot deadlock
  attr[0].val.cooperative = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  VAE_CUDA(cudaLaunchKernelExC(&lc, kern, args));
  return check_launch("bn_lrelu_bwd_fused");
}

template <typename TS, typename TD>
static int launch_transpose(const void* src, void* dst, int B, int R, int Cc, cudaStream_t st) {
  dim3 grid(ceil_div(Cc, 32), ceil_div(R, 32), B), block(32, 8);
  transpose_kernel<TS, TD><<<grid, block, 0, st>>>((const TS*)src, (TD*)dst, R, Cc);
  return check_launch("transpose");
}

static int transpose_dispatch(const void* src, void* dst, int B, int R, int Cc, int dt_src,
                              int dt_dst, cudaStream_t st) {
  if (dt_src == VAE_BF16 && dt_dst == VAE_BF16 && R % 64 == 0 && Cc % 64 == 0) {
    dim3 grid(Cc / 64, R / 64, B);
    transpose_bf16_tile64<<<grid, 256, 0, st>>>((const bf16*)src, (bf16*)dst, R, Cc);
    return check_launch("transpose_bf16_tile64");
  }
  if (dt_src == VAE_F32 && dt_dst == VAE_F32) return launch_transpose<float, float>(src, dst, B, R, Cc, st);
  if (dt_src == VAE_BF16 && dt_dst == VAE_BF16) return launch_transpose<bf16, bf16>(src, dst, B, R, Cc, st);
  if (dt_src == VAE_F32 && dt_dst == VAE_BF16) return launch_transpose<float, bf16>(src, dst, B, R, Cc, st);
  if (dt_src == VAE_BF16 && dt_dst == VAE_F32) return launch_transpose<bf16, float>(src, dst, B, R, Cc, st);
  set_error("transpose: bad dtype");
  return VAE_ERR_ARG;
}

extern "C" int vae_nhwc_to_nchw(const void* src, void* dst, int B, int HW, int C, int dt_src,
                                int dt_dst, vae_stream_t stream) {
  VAE_REQUIRE(src && dst && B > 0 && HW > 0 && C > 0 && B <= 65535, "nhwc_to_nchw: bad argument");
  return transpose_dispatch(src, dst, B, HW, C, dt_src, dt_dst, as_stream(stream));
}

extern "C" int vae_nchw_to_nhwc(const void* src, void* dst, int B, int HW, int C, int dt_src,
                                int dt_dst, vae_stream_t stream) {
  VAE_REQUIRE(src && dst && B > 0 && HW > 0 && C > 0 && B <= 65535, "nchw_to_nhwc: bad argument");
  return transpose_dispatch(src, dst, B, C, HW, dt_src, dt_dst, as_stream(stream));
}

extern "C" int vae_philox_randn(float* out, long long n, unsigned long long seed,
                                unsigned long long offset, vae_stream_t stream) {
  VAE_REQUIRE(out && n > 0, "philox_randn: bad argument");
  philox_randn_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(out, n, seed, offset);
  return check_launch("philox_randn");
}

extern "C" int vae_reparam_fwd(const float* mu, const float* logvar, const float* eps_in,
                               float* eps_out, float* z, long long n, unsigned long long seed,
                               unsigned long long offset, const long long* dev_offset,
                               vae_stream_t stream) {
  VAE_REQUIRE(mu && logvar && z && n > 0, "reparam_fwd: bad argument");
  reparam_fwd_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(
      mu, logvar, eps_in, eps_out, z, n, seed, offset, dev_offset);
  return check_launch("reparam_fwd");
}

extern "C" int vae_reparam_bwd(const float* dz, const float* eps, const float* logvar, float* dmu,
                               float* dlogvar, long long n, int accumulate, vae_stream_t stream) {
  VAE_REQUIRE(dz && eps && logvar && dmu && dlogvar && n > 0, "reparam_bwd: bad argument");
  reparam_bwd_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(dz, eps, logvar, dmu, dlogvar,
                                                                      n, accumulate);
  return check_launch("reparam_bwd");
}

extern "C" int vae_sigmoid(const void* logits, float* probs, long long n, int dtype,
                           vae_stream_t stream) {
  VAE_REQUIRE(logits && probs && n > 0, "sigmoid: bad argument");
  int grid = grid_for(n, 256);
  if (dtype == VAE_BF16)
    sigmoid_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)logits, probs, n);
  else
    sigmoid_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)logits, probs, n);
  return check_launch("sigmoid");
}

extern "C" int vae_sigmoid_bwd(const float* dprobs, const float* probs, void* dlogits, long long n,
                               int dtype, int accumulate, vae_stream_t stream) {
  VAE_REQUIRE(dprobs && probs && dlogits && n > 0, "sigmoid_bwd: bad argument");
  int grid = grid_for(n, 256);
  if (dtype == VAE_BF16)
    sigmoid_bwd_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(dprobs, probs, (bf16*)dlogits, n, accumulate);
  else
    sigmoid_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(dprobs, probs, (float*)dlogits, n, accumulate);
  return check_launch("sigmoid_bwd");
}

extern "C" int vae_pack_conv_weight(const float* w_oihw, void* wp, void* wt, int Co, int Ci,
                                    int dtype, vae_stream_t stream) {
  VAE_REQUIRE(w_oihw && (wp || wt) && Co > 0 && Ci > 0, "pack_conv_weight: bad argument");
  long long n = (long long)Co * Ci * 9;
  int grid = grid_for(n, 256);
  if (dtype == VAE_BF16)
    pack_conv_weight_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, (bf16*)wp, (bf16*)wt, Co, Ci);
  else
    pack_conv_weight_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, (float*)wp, (float*)wt, Co, Ci);
  return check_launch("pack_conv_weight");
}

extern "C" int vae_colsum(const void* A, float* out, int M, int N, int dtype, vae_stream_t stream) {
  VAE_REQUIRE(A && out && M > 0 && N > 0, "colsum: bad argument");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(out, sizeof(float) * N, st));
  int col_blocks = ceil_div(N, 32);
  int want = ceil_div(4 * sm_count(), col_blocks);
  int row_blocks = want < 1 ? 1 : want;
  int rows_per_block = ceil_div(M, row_blocks);
  if (rows_per_block < 8) rows_per_block = 8;
  row_blocks = ceil_div(M, rows_per_block);
  dim3 grid(col_blocks, row_blocks), block(32, 8);
  if (dtype == VAE_BF16)
    colsum_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)A, out, M, N, rows_per_block);
  else
    colsum_kernel<float><<<grid, block, 0, st>>>((const float*)A, out, M, N, rows_per_block);
  return check_launch("colsum");
}

extern "C" int vae_roll_transform(const uint8_t* rolls, float* out, int B, int Hs, int Ws, int H,
                                  int W, int crop_mode, int max_pitch_shift,
                                  float binarize_threshold, float mean, float std,
                                  unsigned long long seed, unsigned long long offset,
                                  vae_stream_t stream) {
  VAE_REQUIRE(rolls && out && B > 0 && B <= 65535 && H > 0 && W > 0 && Hs >= H && Ws >= W &&
                  std != 0.0f && (crop_mode == 0 || crop_mode == 1) && max_pitch_shift >= 0,
              "roll_transform: bad argument");
  int bx = ceil_div((long long)H * W, 256);
  int cap = ceil_div(8LL * sm_count(), B);
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(bx, B);
  roll_transform_kernel<<<grid, 256, 0, as_stream(stream)>>>(rolls, out, Hs, Ws, H, W, crop_mode,
                                                              max_pitch_shift, binarize_threshold,
                                                              mean, 1.0f / std, seed, offset);
  return check_launch("roll_transform");
}
