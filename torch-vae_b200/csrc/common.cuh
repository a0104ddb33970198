// Shared device/host helpers for the vae_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vae_b200.h"

namespace vae {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VAE_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::vae::set_error(__VA_ARGS__);  \
      return VAE_ERR_ARG;             \
    }                                 \
  } while (0)

#define VAE_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::vae::set_error("%s failed: %s", #call, cudaGetErrorString(e__));       \
      return VAE_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

static inline cudaStream_t as_stream(vae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// Zero a scratch buffer (statistics, split-K workspaces) before a kernel accumulates into it -- unless it lies inside the
// caller's pre-zeroed arena (vae_set_zero_arena), whose one clear per step replaces ~25 memset nodes of a train step.
cudaError_t zero_scratch(void* p, size_t bytes, cudaStream_t st);

// ---- scalar type traits ---------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(uint8_t v) { return (float)v; }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 8-wide channel vectors (16 B of bf16 / 2x16 B of fp32) -----------------
template <typename T>
struct Vec8;

template <>
struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

template <>
struct Vec8<bf16> {
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
};

// ---- reductions -----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- Philox4x32-10 (counter-based RNG; same rounds/constants as the
// published algorithm, Salmon et al. SC'11) ----------------------------------
struct Philox {
  uint32_t key0, key1;
  __device__ __forceinline__ Philox(unsigned long long seed)
      : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(unsigned long long ctr, uint32_t stream_id = 0) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream_id, c3 = 0;
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// two uniforms in (0,1] -> two N(0,1) (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float kInv = 2.3283064365386963e-10f;  // 2^-32
  float u1 = ((float)a + 1.0f) * kInv;         // (0,1]
  float u2 = (float)b * kInv;                  // [0,1)
  u1 = fminf(u1, 1.0f);
  float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace vae
