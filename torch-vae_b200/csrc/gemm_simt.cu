// Strided fp32-accumulate GEMM for the three Linear layers (fc_mu, fc_var,
// decoder_input; models.py:55-59) and their gradients.  At the default widths
// these are <0.3 % of the step's FLOPs and purely HBM/latency bound, so a
// CUDA-core tile kernel with split-K is the right tool; operands may be fp32
// (master weights, latents) or bf16 (activations) independently.
#include "common.cuh"

namespace vae {

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p, long long i) { return to_f(p[i]); }

// A_K: threads walk K when loading A (sak == 1), else they walk M (sam == 1 or anything).
// B_N: threads walk N when loading B (sbn == 1), else they walk K.
template <typename TA, typename TB, typename TC, bool A_K, bool B_N>
__global__ void __launch_bounds__(256)
gemm_kernel(const TA* __restrict__ A, const TB* __restrict__ Bm, TC* __restrict__ C,
            const float* __restrict__ bias, int M, int N, int K, long long sam, long long sak,
            long long sbk, long long sbn, long long scm, long long scn, int accumulate,
            int k_per_split, int use_atomics) {
  constexpr int BM = 64, BN = 64, BK = 32, LD = 68;
  __shared__ __align__(16) float As[BK][LD];
  __shared__ __align__(16) float Bs[BK][LD];
  const int tid = threadIdx.x;
  const int m_base = blockIdx.x * BM, n_base = blockIdx.y * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
  const int tx = tid & 15, ty = tid >> 4;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk, mm;
      if (A_K) { kk = tid & 31; mm = (tid >> 5) + 8 * i; }
      else     { mm = tid & 63; kk = (tid >> 6) + 4 * i; }
      int k = k0 + kk, m = m_base + mm;
      float v = 0.0f;
      if (k < k_end && m < M) v = ld_as_float(A, (long long)m * sam + (long long)k * sak);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int kk, nn;
      if (B_N) { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
      else     { kk = tid & 31; nn = (tid >> 5) + 8 * i; }
      int k = k0 + kk, n = n_base + nn;
      float v = 0.0f;
      if (k < k_end && n < N) v = ld_as_float(Bm, (long long)k * sbk + (long long)n * sbn);
      Bs[kk][nn] = v;
    }
    __syncthreads();
    // two-level summation: a fresh 32-term partial per k-block, then one add into the
    // running total -- shortens the fp32 rounding chains from K to 32 + K/32 terms
    float part[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) part[r][cc] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) part[r][cc] = fmaf(av[r], bv[cc], part[r][cc]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) acc[r][cc] += part[r][cc];
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int m = m_base + ty * 4 + r;
    if (m >= M) continue;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      int n = n_base + tx * 4 + cc;
      if (n >= N) continue;
      float v = acc[r][cc];
      if (bias != nullptr && blockIdx.z == 0) v += bias[n];
      long long idx = (long long)m * scm + (long long)n * scn;
      if (use_atomics) {
        if constexpr (sizeof(TC) == 4) atomicAdd(reinterpret_cast<float*>(C) + idx, v);
      } else {
        if (accumulate) v += to_f(C[idx]);
        C[idx] = from_f<TC>(v);
      }
    }
  }
}

}  // namespace vae

using namespace vae;

namespace {

template <typename TA, typename TB, typename TC>
int gemm_launch(const void* A, const void* Bm, void* C, const float* bias, int M, int N, int K,
                long long sam, long long sak, long long sbk, long long sbn, long long scm,
                long long scn, int accumulate, cudaStream_t st) {
  int tiles = ceil_div(M, 64) * ceil_div(N, 64);
  int splits = 1;
  if (sizeof(TC) == 4 && K >= 256 && tiles < 2 * sm_count()) {
    splits = ceil_div(4LL * sm_count(), tiles);
    int max_splits = K / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int kps = ceil_div(K, splits);
  kps = (kps + 31) / 32 * 32;
  splits = ceil_div(K, kps);
  int use_atomics = splits > 1;
  if (use_atomics && !accumulate) {
    if (scn == 1 && scm == N) {
      cudaError_t e = zero_scratch(C, sizeof(float) * (size_t)M * N, st);
      if (e != cudaSuccess) {
        set_error("gemm: memset failed: %s", cudaGetErrorString(e));
        return VAE_ERR_CUDA;
      }
    } else {  // strided output: no split-K
      splits = 1; kps = (K + 31) / 32 * 32; use_atomics = 0;
    }
  }
  dim3 grid(ceil_div(M, 64), ceil_div(N, 64), splits);
  const bool ak = (sak == 1), bn = (sbn == 1);
#define GO(AK, BN_)                                                                              \
  gemm_kernel<TA, TB, TC, AK, BN_><<<grid, 256, 0, st>>>((const TA*)A, (const TB*)Bm, (TC*)C, bias, \
                                                         M, N, K, sam, sak, sbk, sbn, scm, scn,  \
                                                         accumulate, kps, use_atomics)
  if (ak && bn) GO(true, true);
  else if (ak && !bn) GO(true, false);
  else if (!ak && bn) GO(false, true);
  else GO(false, false);
#undef GO
  return check_launch("gemm");
}

}  // namespace

extern "C" int vae_gemm(const void* A, const void* Bm, void* C, const float* bias, int M, int N,
                        int K, long long sam, long long sak, long long sbk, long long sbn,
                        long long scm, long long scn, int dtA, int dtB, int dtC, int accumulate,
                        vae_stream_t stream) {
  VAE_REQUIRE(A && Bm && C && M > 0 && N > 0 && K > 0, "gemm: bad argument");
  cudaStream_t st = as_stream(stream);
  int code = dtA * 4 + dtB * 2 + dtC;
#define CASE(code_, TA, TB, TC) \
  case code_:                   \
    return gemm_launch<TA, TB, TC>(A, Bm, C, bias, M, N, K, sam, sak, sbk, sbn, scm, scn, accumulate, st)
  switch (code) {
    CASE(0, float, float, float);
    CASE(1, float, float, bf16);
    CASE(2, float, bf16, float);
    CASE(3, float, bf16, bf16);
    CASE(4, bf16, float, float);
    CASE(5, bf16, float, bf16);
    CASE(6, bf16, bf16, float);
    CASE(7, bf16, bf16, bf16);
  }
#undef CASE
  set_error("gemm: bad dtype codes %d %d %d", dtA, dtB, dtC);
  return VAE_ERR_ARG;
}
