// CUDA-core (FFMA, fp32-accumulate) convolution kernels.
//
// These are the fp32 path (tcgen05 has no fp32 MMA, and the parity bar in fp32
// is 1e-5 against a true-fp32 oracle), the bf16 path of the two edge layers that
// are not MMA-shaped (Cin = 1 and Cout = 1: HBM-bound direct kernels), and the
// on-device cross-check for the tcgen05 kernels in conv_umma.cu.
#include "common.cuh"
#include "conv_geom.cuh"

namespace vae {

// ============================================================================
// generic gather-GEMM: 64x64 output tile, K blocked by 32, 256 threads, 4x4/thread
// ============================================================================
template <typename TI, typename TW, typename TO>
__global__ void __launch_bounds__(256)
gather_gemm_kernel(const __grid_constant__ GatherGeom g, const TI* __restrict__ In,
                   const TW* __restrict__ Wq, const float* __restrict__ bias,
                   TO* __restrict__ Out, double* __restrict__ stats) {
  constexpr int BM = 64, BN = 64, BK = 32, LD = 68;
  __shared__ __align__(16) float As[BK][LD];
  __shared__ __align__(16) float Bs[BK][LD];
  __shared__ TapTable tt;
  const int tid = threadIdx.x;
  if (tid == 0) tt = g.phase[blockIdx.z];
  __syncthreads();

  const long long M_total = (long long)g.B * g.Hj * g.Wj;
  const int Kp = tt.ntaps * g.Cin;
  const long long m_base = (long long)blockIdx.x * BM;
  const int n_base = blockIdx.y * BN;

  // loader mapping: 32 consecutive threads walk K (= channels: coalesced), 8 rows per pass
  const int k_l = tid & 31, r_l = tid >> 5;
  int rb[8], rh[8], rw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m_base + r_l + 8 * i;
    if (m < M_total) {
      int jw = (int)(m % g.Wj);
      long long t = m / g.Wj;
      int jh = (int)(t % g.Hj);
      rb[i] = (int)(t / g.Hj);
      rh[i] = jh * g.in_mul;
      rw[i] = jw * g.in_mul;
    } else {
      rb[i] = -1; rh[i] = 0; rw[i] = 0;
    }
  }

  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
  const int tx = tid & 15, ty = tid >> 4;

  for (int k0 = 0; k0 < Kp; k0 += BK) {
    const int k = k0 + k_l;
    const bool kv = k < Kp;
    int tap_i = 0, c = 0, dh = 0, dw = 0, wt = 0;
    if (kv) {
      tap_i = k / g.Cin;
      c = k - tap_i * g.Cin;
      dh = tt.dh[tap_i]; dw = tt.dw[tap_i]; wt = tt.wt[tap_i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = 0.0f;
      if (kv && rb[i] >= 0) {
        int ih = rh[i] + dh, iw = rw[i] + dw;
        if (ih >= 0 && ih < g.Hin && iw >= 0 && iw < g.Win)
          v = to_f(In[(((long long)rb[i] * g.Hin + ih) * g.Win + iw) * g.Cin + c]);
      }
      As[k_l][r_l + 8 * i] = v;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int n = n_base + r_l + 8 * i;
      float v = 0.0f;
      if (kv && n < g.Nout) v = to_f(Wq[((long long)n * 9 + wt) * g.Cin + c]);
      Bs[k_l][r_l + 8 * i] = v;
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

  // epilogue: bias, store, BatchNorm statistics
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  float bv[4];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    int n = n_base + tx * 4 + cc;
    bv[cc] = (bias != nullptr && n < g.Nout) ? bias[n] : 0.0f;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    long long m = m_base + ty * 4 + r;
    if (m >= M_total) continue;
    int jw = (int)(m % g.Wj);
    long long t = m / g.Wj;
    int jh = (int)(t % g.Hj);
    int b = (int)(t / g.Hj);
    int oh = jh * g.out_mul + tt.ph, ow = jw * g.out_mul + tt.pw;
    long long base = (((long long)b * g.Hout + oh) * g.Wout + ow) * g.Nout;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      int n = n_base + tx * 4 + cc;
      if (n < g.Nout) {
        float v = acc[r][cc] + bv[cc];
        Out[base + n] = from_f<TO>(v);
        s1[cc] += v;
        s2[cc] = fmaf(v, v, s2[cc]);
      }
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      As[ty][tx * 4 + cc] = s1[cc];
      Bs[ty][tx * 4 + cc] = s2[cc];
    }
    __syncthreads();
    if (tid < 64) {
      int n = n_base + tid;
      if (n < g.Nout) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int y = 0; y < 16; ++y) {
          a += (double)As[y][tid];
          b += (double)Bs[y][tid];
        }
        atomicAdd(&stats[n], a);
        atomicAdd(&stats[g.Nout + n], b);
      }
    }
  }
}

// ============================================================================
// generic weight gradient: dW[n][c][tap] (OIHW fp32) += sum_j Lo[j][n] * Hi[s*j+d(tap)][c]
// GEMM with M = Co, N = 9*Ci, K = pixels; split-K over blockIdx.z, fp32 atomics.
// ============================================================================
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
wgrad_kernel(const TX* __restrict__ X, const TY* __restrict__ dY, float* __restrict__ dW, int B,
             int Hi, int Wi, int Ci, int Co, int s, long long pix_per_split) {
  constexpr int BM = 64, BN = 64, BK = 32, LD = 68;
  __shared__ __align__(16) float As[BK][LD];  // [pixel][co]
  __shared__ __align__(16) float Bs[BK][LD];  // [pixel][(tap,ci)]
  __shared__ int pb[BK], ph[BK], pw[BK];
  const int Ho = Hi / s, Wo = Wi / s;
  const long long J = (long long)B * Ho * Wo;
  const int tid = threadIdx.x;
  const int NQ = 9 * Ci;
  const int co_base = blockIdx.x * BM, q_base = blockIdx.y * BN;
  const long long j_begin = (long long)blockIdx.z * pix_per_split;
  long long j_end = j_begin + pix_per_split;
  if (j_end > J) j_end = J;

  // loader mapping: 64 consecutive threads walk the contiguous channel axis, 4 pixels per pass
  const int col = tid & 63, prow = tid >> 6;
  const int co = co_base + col;
  const int q = q_base + col;
  int tap = 0, ci = 0, kh = 0, kw = 0;
  const bool qv = q < NQ;
  if (qv) {
    tap = q / Ci; ci = q - tap * Ci;
    kh = tap / 3; kw = tap - kh * 3;
  }
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
  const int tx = tid & 15, ty = tid >> 4;

  for (long long j0 = j_begin; j0 < j_end; j0 += BK) {
    if (tid < BK) {
      long long j = j0 + tid;
      if (j < j_end) {
        int jw = (int)(j % Wo);
        long long t = j / Wo;
        pb[tid] = (int)(t / Ho);
        ph[tid] = (int)(t % Ho);
        pw[tid] = jw;
      } else {
        pb[tid] = -1; ph[tid] = 0; pw[tid] = 0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int p = prow + 4 * i;
      int b = pb[p];
      float va = 0.0f, vb = 0.0f;
      if (b >= 0) {
        long long j = j0 + p;
        if (co < Co) va = to_f(dY[j * Co + co]);
        if (qv) {
          int ih = ph[p] * s + kh - 1, iw = pw[p] * s + kw - 1;
          if (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi)
            vb = to_f(X[(((long long)b * Hi + ih) * Wi + iw) * Ci + ci]);
        }
      }
      As[p][col] = va;
      Bs[p][col] = vb;
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
    int n = co_base + ty * 4 + r;
    if (n >= Co) continue;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      int qq = q_base + tx * 4 + cc;
      if (qq < NQ) {
        int t = qq / Ci, c = qq - t * Ci;
        atomicAdd(&dW[((long long)n * Ci + c) * 9 + t], acc[r][cc]);
      }
    }
  }
}

// ============================================================================
// direct kernels for the two edge layers (HBM-bound; no tensor cores by design)
// ============================================================================

// (a) single-channel input -> N channels:  Out[out_pix(j)][n] = bias[n] +
//     sum_t In1[in_mul*j + d_t] * Wq[n][wt_t]      (Wq is [N][9], Cin == 1)
// Serves enc.conv0 forward (models.py:45, Cin=1) and the data gradient of the
// final 32->1 conv (models.py:80).  One thread = one pixel x 8 channels, stores
// are fully coalesced 16/32-byte vectors; BatchNorm statistics are kept in
// registers across a grid-stride loop and reduced once per block.
template <typename TI, typename TW, typename TO>
__global__ void __launch_bounds__(256)
cin1_gather_kernel(const __grid_constant__ GatherGeom g, const TI* __restrict__ In,
                   const TW* __restrict__ Wq, const float* __restrict__ bias,
                   TO* __restrict__ Out, double* __restrict__ stats) {
  extern __shared__ float red[];  // [lanes][G][16]
  const TapTable& tt = g.phase[0];
  const int G = g.Nout / 8;
  const int lanes = blockDim.x / G;
  const int gidx = threadIdx.x % G, lane = threadIdx.x / G;
  const long long M_total = (long long)g.B * g.Hj * g.Wj;
  float w[9][8], bv[8], s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1[k] = s2[k] = 0.0f;
    bv[k] = bias != nullptr ? bias[gidx * 8 + k] : 0.0f;
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) w[t][k] = to_f(Wq[(gidx * 8 + k) * 9 + tt.wt[t]]);

  if (lane < lanes) {
    for (long long m = (long long)blockIdx.x * lanes + lane; m < M_total;
         m += (long long)gridDim.x * lanes) {
      int jw = (int)(m % g.Wj);
      long long t2 = m / g.Wj;
      int jh = (int)(t2 % g.Hj);
      int b = (int)(t2 / g.Hj);
      float xin[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        int ih = jh * g.in_mul + tt.dh[t], iw = jw * g.in_mul + tt.dw[t];
        xin[t] = (ih >= 0 && ih < g.Hin && iw >= 0 && iw < g.Win)
                     ? to_f(In[((long long)b * g.Hin + ih) * g.Win + iw])
                     : 0.0f;
      }
      Vec8<TO> o;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = bv[k];
#pragma unroll
        for (int t = 0; t < 9; ++t) v = fmaf(xin[t], w[t][k], v);
        o.v[k] = v;
        s1[k] += v;
        s2[k] = fmaf(v, v, s2[k]);
      }
      int oh = jh * g.out_mul + tt.ph, ow = jw * g.out_mul + tt.pw;
      o.store(Out + (((long long)b * g.Hout + oh) * g.Wout + ow) * g.Nout + gidx * 8);
    }
  }
  if (stats != nullptr) {
    if (lane < lanes) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        red[(lane * G + gidx) * 16 + k] = s1[k];
        red[(lane * G + gidx) * 16 + 8 + k] = s2[k];
      }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * g.Nout; o += blockDim.x) {
      int which = o / g.Nout, c = o % g.Nout;
      double a = 0.0;
      for (int l = 0; l < lanes; ++l) a += (double)red[(l * G + c / 8) * 16 + which * 8 + (c & 7)];
      atomicAdd(&stats[which * g.Nout + c], a);
    }
  }
}

// (b) weight gradient when one side has a single channel:
//     dW[n*9 + t] = sum_p Mc[p][n] * S1[mul*p + d_t]
// enc.conv0 wgrad: Mc = dY (Co ch), S1 = x, mul = 2, d = (kh-1, kw-1);
// final conv wgrad: Mc = act (Ci ch), S1 = dlogits, mul = 1, d = (1-kh, 1-kw).
template <typename TM, typename TS>
__global__ void __launch_bounds__(256)
c1_wgrad_kernel(const TM* __restrict__ Mc, const TS* __restrict__ S1, float* __restrict__ dW,
                int B, int Hm, int Wm, int N, int Hs, int Ws, int mul, int flip) {
  extern __shared__ float red[];  // [lanes][G][72]
  const int G = N / 8;
  const int lanes = blockDim.x / G;
  const int gidx = threadIdx.x % G, lane = threadIdx.x / G;
  const long long P = (long long)B * Hm * Wm;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.0f;
  if (lane < lanes) {
    for (long long p = (long long)blockIdx.x * lanes + lane; p < P;
         p += (long long)gridDim.x * lanes) {
      int pw_ = (int)(p % Wm);
      long long t2 = p / Wm;
      int ph_ = (int)(t2 % Hm);
      int b = (int)(t2 / Hm);
      Vec8<TM> mv;
      mv.load(Mc + p * N + gidx * 8);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        int kh = t / 3, kw = t % 3;
        int dh = flip ? 1 - kh : kh - 1, dw = flip ? 1 - kw : kw - 1;
        int ih = ph_ * mul + dh, iw = pw_ * mul + dw;
        float sv = (ih >= 0 && ih < Hs && iw >= 0 && iw < Ws)
                       ? to_f(S1[((long long)b * Hs + ih) * Ws + iw])
                       : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(mv.v[k], sv, acc[t][k]);
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[((lane * G + gidx) * 8 + k) * 9 + t] = acc[t][k];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < N * 9; o += blockDim.x) {
    int n = o / 9, t = o % 9;
    float a = 0.0f;
    for (int l = 0; l < lanes; ++l) a += red[((l * G + n / 8) * 8 + (n & 7)) * 9 + t];
    atomicAdd(&dW[o], a);
  }
}

// (c) N channels -> single output channel, stride 1 (final conv, models.py:80):
//     Out[p] = bias + sum_t sum_c In[p + (kh-1,kw-1)][c] * w[t][c]
// Block = 32x8 pixel patch so the 3x3 halo is re-read from L1, not HBM.
template <typename T>
__global__ void __launch_bounds__(256)
cout1_fprop_kernel(const T* __restrict__ In, const T* __restrict__ Wp, const float* __restrict__ bias,
                   T* __restrict__ Out, int H, int W, int C) {
  extern __shared__ float wsm[];  // [9][C]
  for (int i = threadIdx.y * 32 + threadIdx.x; i < 9 * C; i += 256) wsm[i] = to_f(Wp[i]);
  __syncthreads();
  const int b = blockIdx.z;
  const int w0 = blockIdx.x * 32 + threadIdx.x, h0 = blockIdx.y * 8 + threadIdx.y;
  if (w0 >= W || h0 >= H) return;
  float acc = bias != nullptr ? bias[0] : 0.0f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    int ih = h0 + t / 3 - 1, iw = w0 + t % 3 - 1;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
    const T* src = In + (((long long)b * H + ih) * W + iw) * C;
    const float* wr = wsm + t * C;
    for (int c = 0; c < C; c += 8) {
      Vec8<T> v;
      v.load(src + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(v.v[k], wr[c + k], acc);
    }
  }
  Out[((long long)b * H + h0) * W + w0] = from_f<T>(acc);
}

}  // namespace vae

using namespace vae;

// ============================================================================
// host dispatch
// ============================================================================
namespace {

template <typename TI, typename TW, typename TO>
int launch_gather(const GatherGeom& g, const void* in, const void* wq, const float* bias, void* out,
                  double* stats, cudaStream_t st) {
  long long M = (long long)g.B * g.Hj * g.Wj;
  dim3 grid(ceil_div(M, 64), ceil_div(g.Nout, 64), g.nphase);
  gather_gemm_kernel<TI, TW, TO><<<grid, 256, 0, st>>>(g, (const TI*)in, (const TW*)wq, bias,
                                                      (TO*)out, stats);
  return check_launch("gather_gemm");
}

template <typename TI, typename TW, typename TO>
int launch_cin1(const GatherGeom& g, const void* in, const void* wq, const float* bias, void* out,
                double* stats, cudaStream_t st) {
  int G = g.Nout / 8, lanes = 256 / G;
  long long M = (long long)g.B * g.Hj * g.Wj;
  long long blocks = (M + lanes * 4 - 1) / (lanes * 4);
  long long cap = (long long)sm_count() * 4;
  int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  size_t smem = sizeof(float) * lanes * G * 16;
  cin1_gather_kernel<TI, TW, TO><<<grid, 256, smem, st>>>(g, (const TI*)in, (const TW*)wq, bias,
                                                         (TO*)out, stats);
  return check_launch("cin1_gather");
}

template <typename TM, typename TS>
int launch_c1_wgrad(const void* mc, const void* s1, float* dw, int B, int Hm, int Wm, int N, int Hs,
                    int Ws, int mul, int flip, cudaStream_t st) {
  int G = N / 8, lanes = 256 / G;
  long long P = (long long)B * Hm * Wm;
  long long blocks = (P + lanes * 8 - 1) / (lanes * 8);
  long long cap = (long long)sm_count() * 2;
  int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  size_t smem = sizeof(float) * lanes * G * 72;
  auto kern = c1_wgrad_kernel<TM, TS>;
  static bool attr_set = false;  // once per instantiation (not a stream op; keeps graph capture clean)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) {
      set_error("c1_wgrad: smem attribute: %s", cudaGetErrorString(e));
      return VAE_ERR_CUDA;
    }
    attr_set = true;
  }
  kern<<<grid, 256, smem, st>>>((const TM*)mc, (const TS*)s1, dw, B, Hm, Wm, N, Hs, Ws, mul, flip);
  return check_launch("c1_wgrad");
}

bool edge_ok(int n) { return n % 8 == 0 && n / 8 <= 256 && n >= 8; }

}  // namespace

// set by conv_umma.cu when the tcgen05 path is compiled in and enabled
namespace vae {
int umma_conv_fprop(const void*, const void*, const float*, void*, double*, int, int, int, int, int,
                    int, cudaStream_t, bool*);
int umma_conv_dgrad(const void*, const void*, const float*, void*, double*, int, int, int, int, int,
                    int, cudaStream_t, bool*);
int umma_conv_wgrad(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t,
                    bool*);
// conv_edge.cu: strip-mined edge-layer kernels; return 1 when the shape is not served
int edge_cin1(const GatherGeom& g, bool is_dgrad, const void* in, const void* wq, const float* bias, void* out,
              double* stats, int dt_in, int dt_out, cudaStream_t st);
int edge_c1_wgrad(const void* mc, const void* s1, float* dw, int B, int Hm, int Wm, int N, int Hs, int Ws, int mul,
                  int flip, int dt_m, int dt_s, cudaStream_t st);
int edge_cin1_tc(const GatherGeom& g, const void* in, const void* wq, const float* bias, void* out, double* stats,
                 int dt_in, int dt_out, cudaStream_t st);
}  // namespace vae

extern "C" int vae_conv_fprop(const void* x, const void* wp, const float* bias, void* y,
                              double* stats, int B, int Hi, int Wi, int Ci, int Co, int stride,
                              int dt_in, int dt_out, vae_stream_t stream) {
  VAE_REQUIRE(x && wp && y && B > 0 && Ci > 0 && Co > 0 && (stride == 1 || stride == 2) &&
                  Hi > 0 && Wi > 0 && Hi % stride == 0 && Wi % stride == 0,
              "conv_fprop: bad argument (B=%d Hi=%d Wi=%d Ci=%d Co=%d s=%d)", B, Hi, Wi, Ci, Co, stride);
  cudaStream_t st = as_stream(stream);
  if (stats) VAE_CUDA(zero_scratch(stats, sizeof(double) * 2 * Co, st));
  GatherGeom g = geom_fprop(B, Hi, Wi, Ci, Co, stride);
  const bool bf = dt_out == VAE_BF16;
  if (Ci == 1 && edge_ok(Co)) {
    const int rt = edge_cin1_tc(g, x, wp, bias, y, stats, dt_in, dt_out, st);   // tensor cores: N = 32 / 64, stride 2, bf16 out
    if (rt != 1) return rt;
    const int rc = edge_cin1(g, false, x, wp, bias, y, stats, dt_in, dt_out, st);
    if (rc != 1) return rc;
    if (dt_in == VAE_F32 && !bf) return launch_cin1<float, float, float>(g, x, wp, bias, y, stats, st);
    if (dt_in == VAE_F32 && bf) return launch_cin1<float, bf16, bf16>(g, x, wp, bias, y, stats, st);
    if (dt_in == VAE_BF16 && bf) return launch_cin1<bf16, bf16, bf16>(g, x, wp, bias, y, stats, st);
  }
  if (Co == 1 && stride == 1 && Ci % 8 == 0 && dt_in == dt_out && stats == nullptr && B <= 65535) {
    dim3 grid(ceil_div(Wi, 32), ceil_div(Hi, 8), B), block(32, 8);
    size_t smem = sizeof(float) * 9 * Ci;
    if (bf)
      cout1_fprop_kernel<bf16><<<grid, block, smem, st>>>((const bf16*)x, (const bf16*)wp, bias, (bf16*)y, Hi, Wi, Ci);
    else
      cout1_fprop_kernel<float><<<grid, block, smem, st>>>((const float*)x, (const float*)wp, bias, (float*)y, Hi, Wi, Ci);
    return check_launch("cout1_fprop");
  }
  if (bf && dt_in == VAE_BF16) {
    bool handled = false;
    int rc = umma_conv_fprop(x, wp, bias, y, stats, B, Hi, Wi, Ci, Co, stride, st, &handled);
    if (handled) return rc;
    return launch_gather<bf16, bf16, bf16>(g, x, wp, bias, y, stats, st);
  }
  if (dt_in == VAE_F32 && !bf) return launch_gather<float, float, float>(g, x, wp, bias, y, stats, st);
  if (dt_in == VAE_F32 && bf) return launch_gather<float, bf16, bf16>(g, x, wp, bias, y, stats, st);
  set_error("conv_fprop: unsupported dtype combination %d -> %d", dt_in, dt_out);
  return VAE_ERR_UNSUPPORTED;
}

extern "C" int vae_conv_dgrad(const void* dy, const void* wt, const float* bias, void* dx,
                              double* stats, int B, int Hi, int Wi, int Ci, int Co, int stride,
                              int dtype, vae_stream_t stream) {
  VAE_REQUIRE(dy && wt && dx && B > 0 && Ci > 0 && Co > 0 && (stride == 1 || stride == 2) &&
                  Hi > 0 && Wi > 0 && Hi % stride == 0 && Wi % stride == 0,
              "conv_dgrad: bad argument (B=%d Hi=%d Wi=%d Ci=%d Co=%d s=%d)", B, Hi, Wi, Ci, Co, stride);
  cudaStream_t st = as_stream(stream);
  if (stats) VAE_CUDA(zero_scratch(stats, sizeof(double) * 2 * Ci, st));
  GatherGeom g = geom_dgrad(B, Hi, Wi, Ci, Co, stride);
  const bool bf = dtype == VAE_BF16;
  if (Co == 1 && stride == 1 && edge_ok(Ci)) {
    if (bf) return launch_cin1<bf16, bf16, bf16>(g, dy, wt, bias, dx, stats, st);
    return launch_cin1<float, float, float>(g, dy, wt, bias, dx, stats, st);
  }
  if (bf) {
    bool handled = false;
    int rc = umma_conv_dgrad(dy, wt, bias, dx, stats, B, Hi, Wi, Ci, Co, stride, st, &handled);
    if (handled) return rc;
    return launch_gather<bf16, bf16, bf16>(g, dy, wt, bias, dx, stats, st);
  }
  return launch_gather<float, float, float>(g, dy, wt, bias, dx, stats, st);
}

extern "C" int vae_conv_wgrad(const void* x, const void* dy, float* dw, int B, int Hi, int Wi,
                              int Ci, int Co, int stride, int dt_x, int dt_dy,
                              vae_stream_t stream) {
  VAE_REQUIRE(x && dy && dw && B > 0 && Ci > 0 && Co > 0 && (stride == 1 || stride == 2) &&
                  Hi > 0 && Wi > 0 && Hi % stride == 0 && Wi % stride == 0,
              "conv_wgrad: bad argument (B=%d Hi=%d Wi=%d Ci=%d Co=%d s=%d)", B, Hi, Wi, Ci, Co, stride);
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(dw, sizeof(float) * 9 * (size_t)Ci * Co, st));
  const int Ho = Hi / stride, Wo = Wi / stride;
  if (Ci == 1 && edge_ok(Co) && Co <= 64) {  // enc.conv0: Mc = dY, S1 = x
    const int rc = edge_c1_wgrad(dy, x, dw, B, Ho, Wo, Co, Hi, Wi, stride, 0, dt_dy, dt_x, st);
    if (rc != 1) return rc;
    if (dt_x == VAE_F32 && dt_dy == VAE_F32)
      return launch_c1_wgrad<float, float>(dy, x, dw, B, Ho, Wo, Co, Hi, Wi, stride, 0, st);
    if (dt_x == VAE_F32 && dt_dy == VAE_BF16)
      return launch_c1_wgrad<bf16, float>(dy, x, dw, B, Ho, Wo, Co, Hi, Wi, stride, 0, st);
  }
  if (Co == 1 && stride == 1 && edge_ok(Ci) && Ci <= 64 && dt_x == dt_dy) {  // final conv: Mc = act, S1 = dlogits
    if (dt_x == VAE_BF16) return launch_c1_wgrad<bf16, bf16>(x, dy, dw, B, Hi, Wi, Ci, Hi, Wi, 1, 1, st);
    return launch_c1_wgrad<float, float>(x, dy, dw, B, Hi, Wi, Ci, Hi, Wi, 1, 1, st);
  }
  if (dt_x == VAE_BF16 && dt_dy == VAE_BF16) {
    bool handled = false;
    int rc = umma_conv_wgrad(x, dy, dw, B, Hi, Wi, Ci, Co, stride, st, &handled);
    if (handled) return rc;
  }
  const long long J = (long long)B * Ho * Wo;
  int tiles = ceil_div(Co, 64) * ceil_div(9 * Ci, 64);
  int splits = ceil_div(4LL * sm_count(), tiles);
  long long max_splits = (J + 31) / 32;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  long long pps = (J + splits - 1) / splits;
  pps = (pps + 31) / 32 * 32;
  splits = (int)((J + pps - 1) / pps);
  dim3 grid(ceil_div(Co, 64), ceil_div(9 * Ci, 64), splits);
#define WG(TX, TY) \
  wgrad_kernel<TX, TY><<<grid, 256, 0, st>>>((const TX*)x, (const TY*)dy, dw, B, Hi, Wi, Ci, Co, stride, pps)
  if (dt_x == VAE_F32 && dt_dy == VAE_F32) WG(float, float);
  else if (dt_x == VAE_BF16 && dt_dy == VAE_BF16) WG(bf16, bf16);
  else if (dt_x == VAE_F32 && dt_dy == VAE_BF16) WG(float, bf16);
  else {
    set_error("conv_wgrad: unsupported dtype combination");
    return VAE_ERR_UNSUPPORTED;
  }
#undef WG
  return check_launch("conv_wgrad");
}
