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
// out[m][n] += sum_{k in slice} A[m][k] * W(n, k).  A block owns 8 warps x ROWS rows and one K
// slice whose weights it stages ONCE in shared memory; a warp streams its rows' 16-byte chunks
// (lanes stride K), accumulates over the whole slice in registers, and finishes with warp-shuffle
// sums + one atomic per (row, n).  The first chunk of every row is requested before the weights are
// staged, so the two latencies overlap.  Shared layout [N][2][ks_len / 2]: the low and the high four
// floats of a lane's 8-wide chunk live in separate planes, which makes both 16-byte reads of a warp
// contiguous (no bank conflicts).  `out` is zeroed by the caller (memset node); the slice-0 blocks add
// the bias.
// ---------------------------------------------------------------------------------------------
// NW = 2: two weight matrices / biases / outputs share ONE pass over A (fc_mu and fc_var read the same features)
template <typename TA, int N, int ROWS, int NW>
__global__ void __launch_bounds__(256, 2)
skinny_reduce_kernel(const TA* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                     float* __restrict__ out, const float* __restrict__ W1, const float* __restrict__ bias1,
                     float* __restrict__ out1, int M, int K, long long wsn, long long wsk, int ks_len, int nslice) {
  extern __shared__ __align__(16) float w_s[];  // [NW][N][2][ks_len / 2]
  // A block walks `nslice` consecutive K slices with the same 8 x ROWS rows and keeps the accumulators across them:
  // the warp reductions and the atomics at the end are paid once per block, not once per slice (with one slice per
  // block the M = 512, K = 16384 pair launch spent as long reducing as multiplying: 44 us).
  const int kslices = (K + ks_len - 1) / ks_len;
  const int kgroups = (kslices + nslice - 1) / nslice;
  const int kg = blockIdx.x % kgroups, mb = blockIdx.x / kgroups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (mb * 8 + warp) * ROWS;
  const int half_len = ks_len >> 1;
  float acc[NW][ROWS][N];
#pragma unroll
  for (int q = 0; q < NW; ++q)
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int n = 0; n < N; ++n) acc[q][r][n] = 0.0f;
  for (int s = 0; s < nslice; ++s) {
    const int ks = kg * nslice + s;
    if (ks >= kslices) break;   // uniform over the block
    const int k_begin = ks * ks_len, k_len = min(ks_len, K - k_begin);  // multiples of 8
    const TA* a_base = A + (long long)m0 * K + k_begin;
    Vec8<TA> a[ROWS];
    auto load_rows = [&](int kk) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (m0 + r < M && kk < k_len) a[r].load(a_base + (long long)r * K + kk);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[r].v[j] = 0.0f;
        }
      }
    };
    load_rows(lane * 8);
    if (s > 0) __syncthreads();   // every warp is done reading the previous slice's weights
#pragma unroll
    for (int q = 0; q < NW; ++q) {
      const float* __restrict__ Wq = q == 0 ? W : W1;
      float* ws = w_s + q * N * ks_len;
      if (wsk == 1 && (wsn & 3) == 0) {
        for (int kk = threadIdx.x * 4; kk < k_len; kk += 1024) {
          const int dst = ((kk >> 2) & 1) * half_len + (kk >> 3) * 4;
          float4 v[N];
#pragma unroll
          for (int n = 0; n < N; ++n) v[n] = *reinterpret_cast<const float4*>(Wq + n * wsn + k_begin + kk);
#pragma unroll
          for (int n = 0; n < N; ++n) *reinterpret_cast<float4*>(&ws[n * ks_len + dst]) = v[n];
        }
      } else {
        for (int i = threadIdx.x; i < N * k_len; i += 256) {
          const int kk = i / N, n = i - kk * N;
          ws[n * ks_len + ((kk >> 2) & 1) * half_len + (kk >> 3) * 4 + (kk & 3)] = Wq[n * wsn + (long long)(k_begin + kk) * wsk];
        }
      }
    }
    __syncthreads();
    if (m0 < M) {
      for (int kk = lane * 8; kk < k_len; kk += 256) {
        if (kk != lane * 8) load_rows(kk);
#pragma unroll
        for (int q = 0; q < NW; ++q)
#pragma unroll
          for (int n = 0; n < N; ++n) {
            const float* ws = w_s + (q * N + n) * ks_len;
            const float4 w0 = *reinterpret_cast<const float4*>(&ws[kk >> 1]);
            const float4 w1 = *reinterpret_cast<const float4*>(&ws[half_len + (kk >> 1)]);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[q][r][n] = fmaf(a[r].v[j], w[j], acc[q][r][n]);
          }
      }
    }
  }
  if (m0 >= M) return;
#pragma unroll
  for (int q = 0; q < NW; ++q) {
    const float* __restrict__ bq = q == 0 ? bias : bias1;
    float* __restrict__ oq = q == 0 ? out : out1;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        float v = warp_sum(acc[q][r][n]);
        if (lane == ((r * N + n) & 31) && m0 + r < M) {
          if (kg == 0 && bq != nullptr) v += bq[n];
          atomicAdd(&oq[(long long)(m0 + r) * N + n], v);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// G(n, k) = sum_m D[m][n] * A[m][k]; optional colsum[k] = sum_m A[m][k].
// Block = 32 k-chunks (256 columns) x 8 row groups; rows of the block's M range are dealt to the
// row groups, partials are combined in shared memory, one atomic per (n, k) and block.
// ---------------------------------------------------------------------------------------------
template <typename TA, int N>
__global__ void __launch_bounds__(256, (N <= 11 ? 2 : 1))   // (N + 1) x 8 accumulators per thread
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
    // four rows (one per 8-row stride) in flight per thread; same accumulation order as a plain loop
    int m = m_begin + rg;
    for (; m + 24 < m_end; m += 32) {
      Vec8<TA> a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u].load(A + (long long)(m + 8 * u) * K + k);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* d = D + (long long)(m + 8 * u) * N;
#pragma unroll
        for (int n = 0; n < N; ++n) {
          const float dv = __ldg(d + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[n][j] = fmaf(dv, a[u].v[j], acc[n][j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[N][j] += a[u].v[j];
      }
    }
    for (; m < m_end; m += 8) {
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
// out[m][k] = bias[k] + sum_n D0[m][n] * W0(n, k) (+ sum_n D1[m][n] * W1(n, k)); a thread owns 8
// columns, keeps their weight chunk(s) in registers and walks the block's rows, whose D rows sit
// in shared memory (broadcast reads); a warp stores 512 contiguous bytes per row.
// ---------------------------------------------------------------------------------------------
constexpr int kExpandRows = 32;

template <typename TO, int N, bool PAIR>
__global__ void __launch_bounds__(256)
skinny_expand_kernel(const float* __restrict__ D0, const float* __restrict__ W0, const float* __restrict__ D1,
                     const float* __restrict__ W1, const float* __restrict__ bias, TO* __restrict__ out, int M, int K,
                     long long wsn, long long wsk, int accumulate) {
  constexpr int NP = (N + 3) & ~3;
  __shared__ __align__(16) float d_s[PAIR ? 2 : 1][kExpandRows][NP];
  const int m_begin = blockIdx.y * kExpandRows;
  const int rows = min(kExpandRows, M - m_begin);
  for (int i = threadIdx.x; i < rows * N; i += 256) {
    const int r = i / N, n = i - r * N;
    d_s[0][r][n] = D0[(long long)(m_begin + r) * N + n];
    if (PAIR) d_s[PAIR ? 1 : 0][r][n] = D1[(long long)(m_begin + r) * N + n];
  }
  __syncthreads();
  const int k = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (k >= K) return;
  WChunk<N> w0, w1;
  w0.load(W0, k, wsn, wsk);
  if (PAIR) w1.load(W1, k, wsn, wsk);
  float b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = bias != nullptr ? bias[k + j] : 0.0f;
  TO* o = out + (long long)m_begin * K + k;
#pragma unroll 2
  for (int r = 0; r < rows; ++r) {
    float dv0[NP], dv1[NP];
#pragma unroll
    for (int q = 0; q < NP / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(&d_s[0][r][4 * q]);
      dv0[4 * q] = t.x; dv0[4 * q + 1] = t.y; dv0[4 * q + 2] = t.z; dv0[4 * q + 3] = t.w;
      if (PAIR) {
        const float4 u = *reinterpret_cast<const float4*>(&d_s[PAIR ? 1 : 0][r][4 * q]);
        dv1[4 * q] = u.x; dv1[4 * q + 1] = u.y; dv1[4 * q + 2] = u.z; dv1[4 * q + 3] = u.w;
      }
    }
    Vec8<TO> acc;
    if (accumulate) {
      acc.load(o + (long long)r * K);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[j] += b[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[j] = b[j];
    }
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(dv0[n], w0.w[n][j], acc.v[j]);
    if (PAIR) {
#pragma unroll
      for (int n = 0; n < N; ++n)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(dv1[n], w1.w[n][j], acc.v[j]);
    }
    acc.store(o + (long long)r * K);
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
static int skinny_reduce_launch(const void* A, const float* W, const float* bias, float* out, const float* W1,
                                const float* bias1, float* out1, int M, int N, int K, long long wsn, long long wsk, int dtA,
                                cudaStream_t st) {
  const bool pair = W1 != nullptr;
  VAE_CUDA(zero_scratch(out, sizeof(float) * (size_t)M * N, st));
  if (pair) VAE_CUDA(zero_scratch(out1, sizeof(float) * (size_t)M * N, st));
  const int rows = pair ? 2 : 4;   // the pair keeps two accumulator sets: half the rows per warp
  const int row_blocks = (M + 8 * rows - 1) / (8 * rows);
  int ks_len = pair ? 512 : 1024;  // K slice per block: one wave of two blocks per SM where the shape allows it
  while (ks_len > 256 && (long long)row_blocks * ((K + ks_len - 1) / ks_len) * 2 <= 2LL * sm_count()) ks_len /= 2;
  const int kslices = (K + ks_len - 1) / ks_len;
  // slices per block: as many as still leave one full wave of two resident blocks per SM
  int nslice = (int)(((long long)row_blocks * kslices + 2LL * sm_count() - 1) / (2LL * sm_count()));
  if (nslice < 1) nslice = 1;
  if (nslice > kslices) nslice = kslices;
  const int grid = row_blocks * ((kslices + nslice - 1) / nslice);
  const size_t smem = sizeof(float) * (size_t)N * ks_len * (pair ? 2 : 1);
#define VAE_SR_LAUNCH(TA, ROWS, NW)                                                                                   \
  do {                                                                                                                \
    auto kern = skinny_reduce_kernel<TA, NN, ROWS, NW>;                                                               \
    if (smem > 48 * 1024) VAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
    kern<<<grid, 256, smem, st>>>((const TA*)A, W, bias, out, W1, bias1, out1, M, K, wsn, wsk, ks_len, nslice);               \
  } while (0)
  VAE_N_SWITCH(N, {
    if (dtA == VAE_BF16) {
      if (pair) VAE_SR_LAUNCH(bf16, 2, 2); else VAE_SR_LAUNCH(bf16, 4, 1);
    } else {
      if (pair) VAE_SR_LAUNCH(float, 2, 2); else VAE_SR_LAUNCH(float, 4, 1);
    }
  });
#undef VAE_SR_LAUNCH
  return check_launch("skinny_reduce");
}

extern "C" int vae_skinny_reduce(const void* A, const float* W, const float* bias, float* out, int M, int N, int K,
                                 long long wsn, long long wsk, int dtA, vae_stream_t stream) {
  VAE_REQUIRE(A && W && out && M > 0 && K > 0 && K % 8 == 0 && N >= 1 && N <= kMaxL, "skinny_reduce: bad argument");
  return skinny_reduce_launch(A, W, bias, out, nullptr, nullptr, nullptr, M, N, K, wsn, wsk, dtA, as_stream(stream));
}

/* Two heads that share their input in ONE pass over it: out0 = A . W0 + bias0, out1 = A . W1 + bias1 (same shapes
 * and strides as vae_skinny_reduce).
 * replaces: the two aten::addmm of fc_mu and fc_var (models.py:137,141) */
extern "C" int vae_skinny_reduce_pair(const void* A, const float* W0, const float* bias0, float* out0, const float* W1,
                                      const float* bias1, float* out1, int M, int N, int K, long long wsn, long long wsk,
                                      int dtA, vae_stream_t stream) {
  VAE_REQUIRE(A && W0 && out0 && W1 && out1 && M > 0 && K > 0 && K % 8 == 0 && N >= 1 && N <= kMaxL,
              "skinny_reduce_pair: bad argument");
  return skinny_reduce_launch(A, W0, bias0, out0, W1, bias1, out1, M, N, K, wsn, wsk, dtA, as_stream(stream));
}

/* G(n, k) = sum_m D[m][n] * A[m][k] written to G[n*gsn + k*gsk] (fp32, overwritten; the caller
 * passes a ZEROED G and colsum) and, when colsum != NULL, colsum[k] = sum_m A[m][k].
 * replaces: the weight / bias gradients of the three Linear layers (train.py:650) */
extern "C" int vae_skinny_wgrad(const float* D, const void* A, float* G, float* colsum, int M, int N, int K,
                                long long gsn, long long gsk, int dtA, vae_stream_t stream) {
  VAE_REQUIRE(D && A && G && M > 0 && K > 0 && K % 8 == 0 && N >= 1 && N <= kMaxL, "skinny_wgrad: bad argument");
  cudaStream_t st = as_stream(stream);
  const int kblocks = (K + 255) / 256;
  int msplit = (N <= 11 ? 2 : 1) * sm_count() / kblocks;  // one wave of resident blocks
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
  dim3 grid((K / 8 + 255) / 256, (M + kExpandRows - 1) / kExpandRows);
  const bool pair = D1 != nullptr;
  VAE_REQUIRE(!pair || W1 != nullptr, "skinny_expand: D1 without W1");
  // two weight chunks in registers fit up to N = 10 (2 x 80); wider pairs run as two accumulating passes
  VAE_N_SWITCH(N, {
    constexpr bool kPairFits = NN <= 10;
    if (dtO == VAE_BF16) {
      bf16* o = (bf16*)out;
      if (pair && kPairFits) {
        skinny_expand_kernel<bf16, NN, kPairFits><<<grid, 256, 0, st>>>(D0, W0, D1, W1, bias, o, M, K, wsn, wsk, 0);
      } else {
        skinny_expand_kernel<bf16, NN, false><<<grid, 256, 0, st>>>(D0, W0, nullptr, nullptr, bias, o, M, K, wsn, wsk, 0);
        if (pair) skinny_expand_kernel<bf16, NN, false><<<grid, 256, 0, st>>>(D1, W1, nullptr, nullptr, nullptr, o, M, K, wsn, wsk, 1);
      }
    } else {
      float* o = (float*)out;
      if (pair && kPairFits) {
        skinny_expand_kernel<float, NN, kPairFits><<<grid, 256, 0, st>>>(D0, W0, D1, W1, bias, o, M, K, wsn, wsk, 0);
      } else {
        skinny_expand_kernel<float, NN, false><<<grid, 256, 0, st>>>(D0, W0, nullptr, nullptr, bias, o, M, K, wsn, wsk, 0);
        if (pair) skinny_expand_kernel<float, NN, false><<<grid, 256, 0, st>>>(D1, W1, nullptr, nullptr, nullptr, o, M, K, wsn, wsk, 1);
      }
    }
  });
  return check_launch("skinny_expand");
}
