// Skinny GEMMs of the latent heads (models.py:55-59): fc_mu / fc_var (F -> L), decoder_input
// (L -> F) and their gradients, for SMALL L (<= 16; the reference default is 10).  One side of
// every product is only L wide, so these are streaming kernels over the [B, F] activation
// (16.8 MB at B = 512, F = 16 384) with the L-wide operand held in registers -- HBM / L2 bound,
// no tensor cores by design (a 64 x 64 tile GEMM wastes 84 % of its tile on N = 10).
//
//   skinny_reduce :  out[m][n]  = sum_k A[m][k] * W(n, k)  (+ bias[n])     fwd heads, dz
//   skinny_wgrad  :  G(n, k)    = sum_m D[m][n] * A[m][k]  (+ column sums) all three weight grads
//   skinny_expand :  out[m][k]  = sum_n D[m][n] * W(n, k)  (+ bias[k])     decoder_input fwd, d_pre
//
// W(n, k) is addressed as W[n * wsn + k * wsk] so both [L, F] (fc_mu) and [F, L] (decoder_input)
// master weights are used in place, in fp32, without packing.
#include "common.cuh"

namespace vae {

constexpr int kMaxL = 16;

// W(n, k .. k+7) for all n of a chunk of 8 consecutive k: vector loads for the two dense layouts
// ([N][K] rows: wsk == 1;  [K][N] rows: wsn == 1 && wsk == N), scalar loads otherwise.
template <int N>
struct WChunk {
  float w[N][8];
  __device__ __forceinline__ void load(const float* __restrict__ W, int k, long long wsn, long long wsk) {
    if (wsk == 1) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        Vec8<float> v;
        v.load(W + n * wsn + k);
#pragma unroll
        for (int j = 0; j < 8; ++j) w[n][j] = v.v[j];
      }
    } else if (wsn == 1 && wsk == N) {
      const float4* p = reinterpret_cast<const float4*>(W + (long long)k * N);  // 8*N floats, 32*N bytes from a 32*N-byte multiple
      float tmp[8 * N];
#pragma unroll
      for (int i = 0; i < 2 * N; ++i) {
        const float4 f = p[i];
        tmp[4 * i] = f.x; tmp[4 * i + 1] = f.y; tmp[4 * i + 2] = f.z; tmp[4 * i + 3] = f.w;
      }
#pragma unroll
      for (int n = 0; n < N; ++n)
#pragma unroll
        for (int j = 0; j < 8; ++j) w[n][j] = tmp[j * N + n];
    } else {
#pragma unroll
      for (int n = 0; n < N; ++n)
#pragma unroll
        for (int j = 0; j < 8; ++j) w[n][j] = W[n * wsn + (long long)(k + j) * wsk];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// out[m][n] += sum_{k in slice} A[m][k] * W(n, k): a warp owns 4 rows x a K slice; lanes stride K in
// 8-element chunks; warp-shuffle reduction, one atomic per (row, n).  `out` must be pre-filled
// (bias or zero) by the caller's init kernel.
// ---------------------------------------------------------------------------------------------
template <typename TA, int N>
__global__ void __launch_bounds__(256)
skinny_reduce_kernel(const TA* __restrict__ A, const float* __restrict__ W, float* __restrict__ out, int M, int K,
                     long long wsn, long long wsk, int k_per_warp) {
  const int warp_global = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int kslices = (K + k_per_warp - 1) / k_per_warp;
  const int mt = warp_global / kslices, ks = warp_global - mt * kslices;
  const int m0 = mt * 4;
  if (m0 >= M) return;
  const int k_begin = ks * k_per_warp, k_end = min(K, k_begin + k_per_warp);
  float acc[4][N];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int n = 0; n < N; ++n) acc[r][n] = 0.0f;
  for (int k = k_begin + lane * 8; k < k_end; k += 256) {
    Vec8<TA> a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (m0 + r < M) a[r].load(A + (long long)(m0 + r) * K + k);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[r].v[j] = 0.0f;
      }
    }
    WChunk<N> wc;
    wc.load(W, k, wsn, wsk);
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][n] = fmaf(a[r].v[j], wc.w[n][j], acc[r][n]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float v = warp_sum(acc[r][n]);
      if (lane == ((r * N + n) & 31) && m0 + r < M) atomicAdd(&out[(long long)(m0 + r) * N + n], v);
    }
}

__global__ void fill_rows_kernel(float* __restrict__ out, const float* __restrict__ bias, int M, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M * N) out[i] = bias != nullptr ? bias[i % N] : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// G(n, k) = sum_m D[m][n] * A[m][k]; optional colsum[k] = sum_m A[m][k].
// Block = 32 k-chunks (256 columns) x 8 row groups; rows of the block's M range are dealt to the
// row groups, partials are combined in shared memory, one atomic per (n, k) and block.
// ---------------------------------------------------------------------------------------------
template <typename TA, int N>
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(const float* __restrict__ D, const TA* __restrict__ A, float* __restrict__ G,
                    float* __restrict__ colsum, int M, int K, long long gsn, long long gsk, int m_per_block) {
  __shared__ __align__(16) float red[8][256];
  const int kc = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int k = (blockIdx.x * 32 + kc) * 8;
  const int m_begin = blockIdx.y * m_per_block, m_end = min(M, m_begin + m_per_block);
  float acc[N + 1][8];
#pragma unroll
  for (int n = 0; n <= N; ++n)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[n][j] = 0.0f;
  if (k < K) {
    for (int m = m_begin + rg; m < m_end; m += 8) {
      Vec8<TA> a;
      a.load(A + (long long)m * K + k);
      const float* d = D + (long long)m * N;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const float dv = __ldg(d + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[n][j] = fmaf(dv, a.v[j], acc[n][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[N][j] += a.v[j];
    }
  }
  // combine the 8 row groups: one output row n (256 columns) per round
  const int col = threadIdx.x;
  const int kk = blockIdx.x * 256 + col;
#pragma unroll
  for (int n = 0; n <= N; ++n) {
    if (n == N && colsum == nullptr) break;
    *reinterpret_cast<float4*>(&red[rg][kc * 8]) = make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]);
    *reinterpret_cast<float4*>(&red[rg][kc * 8 + 4]) = make_float4(acc[n][4], acc[n][5], acc[n][6], acc[n][7]);
    __syncthreads();
    if (kk < K) {
      float v = 0.0f;
#pragma unroll
      for (int g = 0; g < 8; ++g) v += red[g][col];
      if (n < N) atomicAdd(&G[n * gsn + (long long)kk * gsk], v);
      else atomicAdd(&colsum[kk], v);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// out[m][k] = bias[k] + sum_n D0[m][n] * W0(n, k) (+ sum_n D1[m][n] * W1(n, k)); a thread owns
// 4 rows x 8 columns so each weight chunk is loaded once per 4 rows.
// ---------------------------------------------------------------------------------------------
template <typename TO, int N>
__global__ void __launch_bounds__(256)
skinny_expand_kernel(const float* __restrict__ D0, const float* __restrict__ W0, const float* __restrict__ D1,
                     const float* __restrict__ W1, const float* __restrict__ bias, TO* __restrict__ out, int M, int K,
                     long long wsn, long long wsk) {
  const int kchunks = K / 8;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const int kc = (int)(t % kchunks);
  const int m0 = (int)(t / kchunks) * 4;
  if (m0 >= M) return;
  const int k = kc * 8;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float b = bias != nullptr ? bias[k + j] : 0.0f;
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r][j] = b;
  }
#pragma unroll
  for (int pair = 0; pair < 2; ++pair) {
    const float* D = pair == 0 ? D0 : D1;
    const float* W = pair == 0 ? W0 : W1;
    if (D == nullptr) continue;
    WChunk<N> wc;
    wc.load(W, k, wsn, wsk);
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float dv = m0 + r < M ? __ldg(D + (long long)(m0 + r) * N + n) : 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(dv, wc.w[n][j], acc[r][j]);
      }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (m0 + r >= M) break;
    Vec8<TO> o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = acc[r][j];
    o.store(out + (long long)(m0 + r) * K + k);
  }
}

}  // namespace vae

using namespace vae;

#define VAE_N_SWITCH(N, ...)                 \
  switch (N) {                               \
    case 1: { constexpr int NN = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int NN = 2; __VA_ARGS__; } break;   \
    case 3: { constexpr int NN = 3; __VA_ARGS__; } break;   \
    case 4: { constexpr int NN = 4; __VA_ARGS__; } break;   \
    case 5: { constexpr int NN = 5; __VA_ARGS__; } break;   \
    case 6: { constexpr int NN = 6; __VA_ARGS__; } break;   \
    case 7: { constexpr int NN = 7; __VA_ARGS__; } break;   \
    case 8: { constexpr int NN = 8; __VA_ARGS__; } break;   \
    case 9: { constexpr int NN = 9; __VA_ARGS__; } break;   \
    case 10: { constexpr int NN = 10; __VA_ARGS__; } break; \
    case 11: { constexpr int NN = 11; __VA_ARGS__; } break; \
    case 12: { constexpr int NN = 12; __VA_ARGS__; } break; \
    case 13: { constexpr int NN = 13; __VA_ARGS__; } break; \
    case 14: { constexpr int NN = 14; __VA_ARGS__; } break; \
    case 15: { constexpr int NN = 15; __VA_ARGS__; } break; \
    case 16: { constexpr int NN = 16; __VA_ARGS__; } break; \
  }

/* 1 when the skinny head kernels serve this (F, L): L <= 16 and F a multiple of 8. */
extern "C" int vae_heads_supported(int F, int L) { return (L >= 1 && L <= kMaxL && F % 8 == 0 && F >= 8) ? 1 : 0; }

/* out[M][N] (fp32) = A[M][K] (dtype dtA) . W + bias, W(n, k) = W[n*wsn + k*wsk] (fp32).
 * replaces: aten::addmm of fc_mu / fc_var (models.py:137,141; wsn = K, wsk = 1) and the latent
 * gradient of decoder_input (train.py:650; W = weight[F, L]: wsn = 1, wsk = L) */
extern "C" int vae_skinny_reduce(const void* A, const float* W, const float* bias, float* out, int M, int N, int K,
                                 long long wsn, long long wsk, int dtA, vae_stream_t stream) {
  VAE_REQUIRE(A && W && out && M > 0 && K > 0 && K % 8 == 0 && N >= 1 && N <= kMaxL, "skinny_reduce: bad argument");
  cudaStream_t st = as_stream(stream);
  fill_rows_kernel<<<(M * N + 255) / 256, 256, 0, st>>>(out, bias, M, N);
  int kpw = 1024;
  while (kpw > 256 && (long long)((M + 3) / 4) * ((K + kpw - 1) / kpw) < 32LL * sm_count()) kpw /= 2;
  const long long warps = (long long)((M + 3) / 4) * ((K + kpw - 1) / kpw);
  const int grid = (int)((warps + 7) / 8);
  VAE_N_SWITCH(N, {
    if (dtA == VAE_BF16)
      skinny_reduce_kernel<bf16, NN><<<grid, 256, 0, st>>>((const bf16*)A, W, out, M, K, wsn, wsk, kpw);
    else
      skinny_reduce_kernel<float, NN><<<grid, 256, 0, st>>>((const float*)A, W, out, M, K, wsn, wsk, kpw);
  });
  return check_launch("skinny_reduce");
}

/* G(n, k) = sum_m D[m][n] * A[m][k] written to G[n*gsn + k*gsk] (fp32, overwritten; the caller
 * passes a ZEROED G and colsum) and, when colsum != NULL, colsum[k] = sum_m A[m][k].
 * replaces: the weight / bias gradients of the three Linear layers (train.py:650) */
extern "C" int vae_skinny_wgrad(const float* D, const void* A, float* G, float* colsum, int M, int N, int K,
                                long long gsn, long long gsk, int dtA, vae_stream_t stream) {
  VAE_REQUIRE(D && A && G && M > 0 && K > 0 && K % 8 == 0 && N >= 1 && N <= kMaxL, "skinny_wgrad: bad argument");
  cudaStream_t st = as_stream(stream);
  const int kblocks = (K + 255) / 256;
  int msplit = (2 * sm_count() + kblocks - 1) / kblocks;
  if (msplit > (M + 15) / 16) msplit = (M + 15) / 16;
  if (msplit < 1) msplit = 1;
  int mpb = (M + msplit - 1) / msplit;
  mpb = (mpb + 7) / 8 * 8;
  msplit = (M + mpb - 1) / mpb;
  dim3 grid(kblocks, msplit);
  VAE_N_SWITCH(N, {
    if (dtA == VAE_BF16)
      skinny_wgrad_kernel<bf16, NN><<<grid, 256, 0, st>>>(D, (const bf16*)A, G, colsum, M, K, gsn, gsk, mpb);
    else
      skinny_wgrad_kernel<float, NN><<<grid, 256, 0, st>>>(D, (const float*)A, G, colsum, M, K, gsn, gsk, mpb);
  });
  return check_launch("skinny_wgrad");
}

/* out[M][K] (dtype dtO) = bias[k] + D0 . W0 (+ D1 . W1); D [M][N] fp32, W(n, k) = W[n*wsn + k*wsk].
 * replaces: aten::addmm of decoder_input (models.py:162; W = weight[F, L]: wsn = 1, wsk = L) and
 * the input gradient of fc_mu + fc_var (train.py:650; W = weight[L, F]: wsn = F, wsk = 1) */
extern "C" int vae_skinny_expand(const float* D0, const float* W0, const float* D1, const float* W1,
                                 const float* bias, void* out, int M, int N, int K, long long wsn, long long wsk,
                                 int dtO, vae_stream_t stream) {
  VAE_REQUIRE(D0 && W0 && out && M > 0 && K > 0 && K % 8 == 0 && N >= 1 && N <= kMaxL, "skinny_expand: bad argument");
  cudaStream_t st = as_stream(stream);
  const long long threads = (long long)((M + 3) / 4) * (K / 8);
  const int grid = (int)((threads + 255) / 256);
  VAE_N_SWITCH(N, {
    if (dtO == VAE_BF16)
      skinny_expand_kernel<bf16, NN><<<grid, 256, 0, st>>>(D0, W0, D1, W1, bias, (bf16*)out, M, K, wsn, wsk);
    else
      skinny_expand_kernel<float, NN><<<grid, 256, 0, st>>>(D0, W0, D1, W1, bias, (float*)out, M, K, wsn, wsk);
  });
  return check_launch("skinny_expand");
}
