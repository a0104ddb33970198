// HBM-bound direct kernels for the two layers that are not MMA-shaped:
//   enc.conv0   Conv2d(1 -> C, k3, s2, p1)   (models.py:45 with in_channels = 1)
//   final conv  Conv2d(C -> 1, k3, s1, p1)   (models.py:80)
// forward, data gradient and weight gradient.  With one channel on one side the
// arithmetic intensity is ~8 flop/B: no tensor cores by design.
// Each thread owns 8 channels x a strip of 4 consecutive pixels along W, so the 3x3
// window of the single-channel tensor is loaded once per strip (18 / 27 scalars for 4
// outputs) and every global access of the multi-channel tensor is a 16-byte vector.
#include "common.cuh"
#include "conv_geom.cuh"

namespace vae {

constexpr int kStrip = 4;

// 16-byte shared-memory load the compiler may not hoist out of the strip loop: keeping the 9 x 8
// tap weights in registers costs 72 registers per thread and halves the resident warps, while
// re-reading them is one broadcast wavefront per load.
__device__ __forceinline__ float4 lds128_volatile(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

// tap index (kh*3+kw) that pairs with window offset (a, b) in {-1,0,1}^2:
// forward conv: (kh, kw) = (a+1, b+1); data gradient of a stride-1 conv: (1-a, 1-b)
template <bool FLIP>
__device__ __forceinline__ constexpr int tap_of(int a, int b) {
  return FLIP ? (1 - a) * 3 + (1 - b) : (a + 1) * 3 + (b + 1);
}

// ---------------------------------------------------------------------------
// (1) single-channel input -> N channels.
//     Out[b][out_mul... == 1][jh][jw][n] = bias[n] + sum_{a,b} In1[MUL*jh + a][MUL*jw + b] * W[n][tap(a,b)]
// ---------------------------------------------------------------------------
template <typename TI, typename TW, typename TO, int MUL, bool FLIP>
__global__ void __launch_bounds__(256, 2)
cin1_strip_kernel(const TI* __restrict__ In, const TW* __restrict__ Wq, const float* __restrict__ bias,
                  TO* __restrict__ Out, double* __restrict__ stats, int B, int Hin, int Win, int Hj, int Wj, int N) {
  constexpr int NC = MUL * (kStrip - 1) + 3;  // window columns a strip touches
  extern __shared__ float smem_f[];
  float* wsm = smem_f;           // [9][N]
  float* red = smem_f + 9 * N;   // [spb][G][16]
  const int G = N / 8;
  const int spb = blockDim.x / G;
  const int gidx = threadIdx.x % G, sl = threadIdx.x / G;
  for (int i = threadIdx.x; i < 9 * N; i += blockDim.x) {
    const int t = i / N, n = i - t * N;
    wsm[i] = to_f(Wq[n * 9 + t]);
  }
  __syncthreads();
  float bv[8], s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    bv[k] = bias != nullptr ? bias[gidx * 8 + k] : 0.0f;
    s1[k] = s2[k] = 0.0f;
  }
  const int wstrips = Wj / kStrip;
  const long long nstrips = (long long)B * Hj * wstrips;
  if (sl < spb) {
    for (long long strip = (long long)blockIdx.x * spb + sl; strip < nstrips; strip += (long long)gridDim.x * spb) {
      const int sw = (int)(strip % wstrips);
      const long long t2 = strip / wstrips;
      const int jh = (int)(t2 % Hj), b = (int)(t2 / Hj);
      const int jw0 = sw * kStrip;
      float acc[kStrip][8];
#pragma unroll
      for (int i = 0; i < kStrip; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][k] = bv[k];
#pragma unroll
      for (int a = -1; a <= 1; ++a) {
        const int ih = jh * MUL + a;
        const bool rowok = ih >= 0 && ih < Hin;
        const TI* row = In + ((long long)b * Hin + (rowok ? ih : 0)) * Win;
        float xin[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int iw = jw0 * MUL + c - 1;
          xin[c] = (rowok && iw >= 0 && iw < Win) ? to_f(row[iw]) : 0.0f;
        }
#pragma unroll
        for (int bb = -1; bb <= 1; ++bb) {
          const float* wp = wsm + tap_of<FLIP>(a, bb) * N + gidx * 8;
          const float4 w0 = lds128_volatile(wp), w1 = lds128_volatile(wp + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < kStrip; ++i) {
            const float v = xin[MUL * i + bb + 1];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(v, wv[k], acc[i][k]);
          }
        }
      }
      TO* dst = Out + (((long long)b * Hj + jh) * Wj + jw0) * N + gidx * 8;
#pragma unroll
      for (int i = 0; i < kStrip; ++i) {
        Vec8<TO> o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          o.v[k] = acc[i][k];
          s1[k] += acc[i][k];
          s2[k] = fmaf(acc[i][k], acc[i][k], s2[k]);
        }
        o.store(dst + (long long)i * N);
      }
    }
  }
  if (stats != nullptr) {
    if (sl < spb) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        red[(sl * G + gidx) * 16 + k] = s1[k];
        red[(sl * G + gidx) * 16 + 8 + k] = s2[k];
      }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * N; o += blockDim.x) {
      const int which = o / N, c = o - which * N;
      double a = 0.0;
      for (int l = 0; l < spb; ++l) a += (double)red[(l * G + c / 8) * 16 + which * 8 + (c & 7)];
      atomicAdd(&stats[which * N + c], a);
    }
  }
}

// ---------------------------------------------------------------------------
// (2) weight gradient with one single-channel side:
//     dW[n][tap(a,b)] = sum_p Mc[p][n] * S1[MUL*p + (a, b)]
//     enc.conv0: Mc = dY, S1 = x, MUL = 2, FLIP = false; final conv: Mc = act, S1 = dlogits, MUL = 1, FLIP = true
// ---------------------------------------------------------------------------
template <typename TM, typename TS, int MUL, bool FLIP>
__global__ void __launch_bounds__(256, 2)
c1_wgrad_strip_kernel(const TM* __restrict__ Mc, const TS* __restrict__ S1, float* __restrict__ dW, int B, int Hm,
                      int Wm, int N, int Hs, int Ws) {
  constexpr int NC = MUL * (kStrip - 1) + 3;
  extern __shared__ float red[];  // [spb][G][72]
  const int G = N / 8;
  const int spb = blockDim.x / G;
  const int gidx = threadIdx.x % G, sl = threadIdx.x / G;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.0f;
  const int wstrips = Wm / kStrip;
  const long long nstrips = (long long)B * Hm * wstrips;
  if (sl < spb) {
    for (long long strip = (long long)blockIdx.x * spb + sl; strip < nstrips; strip += (long long)gridDim.x * spb) {
      const int sw = (int)(strip % wstrips);
      const long long t2 = strip / wstrips;
      const int ph = (int)(t2 % Hm), b = (int)(t2 / Hm);
      const int pw0 = sw * kStrip;
      const TM* src = Mc + (((long long)b * Hm + ph) * Wm + pw0) * N + gidx * 8;
      Vec8<TM> mv[kStrip];
#pragma unroll
      for (int i = 0; i < kStrip; ++i) mv[i].load(src + (long long)i * N);
#pragma unroll
      for (int a = -1; a <= 1; ++a) {
        const int ih = ph * MUL + a;
        const bool rowok = ih >= 0 && ih < Hs;
        const TS* row = S1 + ((long long)b * Hs + (rowok ? ih : 0)) * Ws;
        float win[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int iw = pw0 * MUL + c - 1;
          win[c] = (rowok && iw >= 0 && iw < Ws) ? to_f(row[iw]) : 0.0f;
        }
#pragma unroll
        for (int bb = -1; bb <= 1; ++bb) {
          constexpr int dummy = 0;
          (void)dummy;
          const int t = tap_of<FLIP>(a, bb);
#pragma unroll
          for (int i = 0; i < kStrip; ++i) {
            const float sv = win[MUL * i + bb + 1];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(mv[i].v[k], sv, acc[t][k]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[((sl * G + gidx) * 8 + k) * 9 + t] = acc[t][k];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < N * 9; o += blockDim.x) {
    const int n = o / 9, t = o - n * 9;
    float a = 0.0f;
    for (int l = 0; l < spb; ++l) a += red[((l * G + n / 8) * 8 + (n & 7)) * 9 + t];
    atomicAdd(&dW[o], a);
  }
}

// ---------------------------------------------------------------------------
// (3) N channels -> single output channel, stride 1 (final conv forward):
//     the block stages a (8+2) x (32+2) pixel patch in shared memory once; a thread owns
//     8 channels x 4 pixels and the channel groups are combined with warp shuffles.
// ---------------------------------------------------------------------------
template <typename T, int CG>  // CG = C / 8 channel groups per pixel (1, 2, 4 or 8)
__global__ void __launch_bounds__(256)
cout1_strip_kernel(const T* __restrict__ In, const T* __restrict__ Wp, const float* __restrict__ bias,
                   T* __restrict__ Out, int H, int W) {
  constexpr int C = CG * 8;
  constexpr int TW_ = 32;                      // tile width (pixels)
  constexpr int TH = (256 / CG) / (TW_ / kStrip);  // tile height: 256 threads = CG x 8 strips x TH rows
  constexpr int PW = TW_ + 2, PH = TH + 2;
  constexpr int PITCH = C + 4;                 // floats per patch pixel (+4: keeps LDS.128 conflict-free)
  extern __shared__ float sm[];
  float* patch = sm;                            // [PH][PW][PITCH]
  float* wsm = sm + PH * PW * PITCH;            // [9][C]
  const int b = blockIdx.z;
  const int w_base = blockIdx.x * TW_, h_base = blockIdx.y * TH;
  for (int i = threadIdx.x; i < 9 * C; i += 256) wsm[i] = to_f(Wp[i]);
  // stage the patch: 16-byte chunks, zero outside the image
  for (int i = threadIdx.x; i < PH * PW * CG; i += 256) {
    const int cg = i % CG, pix = i / CG;
    const int pw = pix % PW, ph = pix / PW;
    const int ih = h_base + ph - 1, iw = w_base + pw - 1;
    Vec8<T> v;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      v.load(In + (((long long)b * H + ih) * W + iw) * C + cg * 8);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] = 0.0f;
    }
    float* d = patch + (ph * PW + pw) * PITCH + cg * 8;
    *reinterpret_cast<float4*>(d) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
  }
  __syncthreads();
  const int cg = threadIdx.x % CG;
  const int strip = (threadIdx.x / CG) % (TW_ / kStrip);
  const int row = threadIdx.x / (CG * (TW_ / kStrip));
  float acc[kStrip] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float wv[3][8];
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) {
      const float4* w4 = reinterpret_cast<const float4*>(wsm + (a * 3 + bb) * C + cg * 8);
      const float4 w0 = w4[0], w1 = w4[1];
      wv[bb][0] = w0.x; wv[bb][1] = w0.y; wv[bb][2] = w0.z; wv[bb][3] = w0.w;
      wv[bb][4] = w1.x; wv[bb][5] = w1.y; wv[bb][6] = w1.z; wv[bb][7] = w1.w;
    }
    const float* prow = patch + ((row + a) * PW + strip * kStrip) * PITCH + cg * 8;
#pragma unroll
    for (int c = 0; c < kStrip + 2; ++c) {
      const float4 x0 = *reinterpret_cast<const float4*>(prow + c * PITCH);
      const float4 x1 = *reinterpret_cast<const float4*>(prow + c * PITCH + 4);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const int i = c - bb;  // output pixel whose tap bb reads patch column c
        if (i >= 0 && i < kStrip) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[i] = fmaf(xv[k], wv[bb][k], acc[i]);
        }
      }
    }
  }
  // combine the CG channel groups of a pixel (adjacent lanes)
#pragma unroll
  for (int off = CG / 2; off >= 1; off >>= 1)
#pragma unroll
    for (int i = 0; i < kStrip; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
  if (cg == 0) {
    const int oh = h_base + row, ow = w_base + strip * kStrip;
    if (oh < H) {
      const float bvl = bias != nullptr ? bias[0] : 0.0f;
      T* dst = Out + ((long long)b * H + oh) * W + ow;
#pragma unroll
      for (int i = 0; i < kStrip; ++i)
        if (ow + i < W) dst[i] = from_f<T>(acc[i] + bvl);
    }
  }
}

// ============================================================================
// host launchers (called from conv_simt.cu's entry points)
// ============================================================================
static int strip_grid(long long nstrips, int spb) {
  long long blocks = (nstrips + spb * 2 - 1) / (spb * 2);
  long long cap = (long long)sm_count() * 6;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

template <typename TI, typename TW, typename TO>
static int cin1_launch(const GatherGeom& g, const void* in, const void* wq, const float* bias, void* out,
                       double* stats, bool flip, cudaStream_t st) {
  const int N = g.Nout, G = N / 8, spb = 256 / G;
  const long long nstrips = (long long)g.B * g.Hj * (g.Wj / kStrip);
  const int grid = strip_grid(nstrips, spb);
  const size_t smem = sizeof(float) * (9 * N + spb * G * 16);
#define GO(MUL, FLIP)                                                                                     \
  cin1_strip_kernel<TI, TW, TO, MUL, FLIP><<<grid, 256, smem, st>>>((const TI*)in, (const TW*)wq, bias,      \
                                                                  (TO*)out, stats, g.B, g.Hin, g.Win, g.Hj, \
                                                                  g.Wj, N)
  if (g.in_mul == 2 && !flip) GO(2, false);
  else if (g.in_mul == 1 && flip) GO(1, true);
  else if (g.in_mul == 1 && !flip) GO(1, false);
  else return 1;
#undef GO
  return check_launch("cin1_strip");
}

// returns 1 if the shape is not served (caller falls back), 0 / negative otherwise
int edge_cin1(const GatherGeom& g, bool is_dgrad, const void* in, const void* wq, const float* bias, void* out,
              double* stats, int dt_in, int dt_out, cudaStream_t st) {
  const int N = g.Nout;
  if (g.Cin != 1 || N % 8 || N / 8 > 64 || g.Wj % kStrip || g.out_mul != 1 || g.nphase != 1) return 1;
  const bool flip = is_dgrad;
  if (dt_in == VAE_F32 && dt_out == VAE_F32) return cin1_launch<float, float, float>(g, in, wq, bias, out, stats, flip, st);
  if (dt_in == VAE_F32 && dt_out == VAE_BF16) return cin1_launch<float, bf16, bf16>(g, in, wq, bias, out, stats, flip, st);
  if (dt_in == VAE_BF16 && dt_out == VAE_BF16) return cin1_launch<bf16, bf16, bf16>(g, in, wq, bias, out, stats, flip, st);
  return 1;
}

template <typename TM, typename TS>
static int c1_wgrad_launch(const void* mc, const void* s1, float* dw, int B, int Hm, int Wm, int N, int Hs, int Ws,
                           int mul, int flip, cudaStream_t st) {
  const int G = N / 8, spb = 256 / G;
  const long long nstrips = (long long)B * Hm * (Wm / kStrip);
  long long blocks = (nstrips + spb * 4 - 1) / (spb * 4);
  const long long cap = (long long)sm_count() * 2;
  const int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  const size_t smem = sizeof(float) * spb * G * 72;
#define GO(MUL, FLIP)                                                                                         \
  do {                                                                                                        \
    auto kern = c1_wgrad_strip_kernel<TM, TS, MUL, FLIP>;                                                      \
    static bool attr = false;                                                                                 \
    if (!attr) {                                                                                              \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) { \
        set_error("c1_wgrad_strip: smem attribute");                                                          \
        return VAE_ERR_CUDA;                                                                                  \
      }                                                                                                       \
      attr = true;                                                                                            \
    }                                                                                                         \
    kern<<<grid, 256, smem, st>>>((const TM*)mc, (const TS*)s1, dw, B, Hm, Wm, N, Hs, Ws);                      \
  } while (0)
  if (mul == 2 && !flip) GO(2, false);
  else if (mul == 1 && flip) GO(1, true);
  else return 1;
#undef GO
  return check_launch("c1_wgrad_strip");
}

int edge_c1_wgrad(const void* mc, const void* s1, float* dw, int B, int Hm, int Wm, int N, int Hs, int Ws, int mul,
                  int flip, int dt_m, int dt_s, cudaStream_t st) {
  if (N % 8 || N / 8 > 64 || Wm % kStrip) return 1;
  if (dt_m == VAE_F32 && dt_s == VAE_F32) return c1_wgrad_launch<float, float>(mc, s1, dw, B, Hm, Wm, N, Hs, Ws, mul, flip, st);
  if (dt_m == VAE_BF16 && dt_s == VAE_F32) return c1_wgrad_launch<bf16, float>(mc, s1, dw, B, Hm, Wm, N, Hs, Ws, mul, flip, st);
  if (dt_m == VAE_BF16 && dt_s == VAE_BF16) return c1_wgrad_launch<bf16, bf16>(mc, s1, dw, B, Hm, Wm, N, Hs, Ws, mul, flip, st);
  return 1;
}

template <typename T, int CG>
static int cout1_launch(const void* in, const void* wp, const float* bias, void* out, int B, int H, int W,
                        cudaStream_t st) {
  constexpr int C = CG * 8, TH = (256 / CG) / 8, PITCH = C + 4;
  const size_t smem = sizeof(float) * ((TH + 2) * 34 * PITCH + 9 * C);
  auto kern = cout1_strip_kernel<T, CG>;
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("cout1_strip: smem attribute");
      return VAE_ERR_CUDA;
    }
    attr = true;
  }
  dim3 grid(ceil_div(W, 32), ceil_div(H, TH), B);
  kern<<<grid, 256, smem, st>>>((const T*)in, (const T*)wp, bias, (T*)out, H, W);
  return check_launch("cout1_strip");
}

int edge_cout1(const void* in, const void* wp, const float* bias, void* out, int B, int H, int W, int C, int dtype,
               cudaStream_t st) {
  if (B > 65535 || W % kStrip) return 1;
  const bool bf = dtype == VAE_BF16;
#define GO(CG) return bf ? cout1_launch<bf16, CG>(in, wp, bias, out, B, H, W, st) : cout1_launch<float, CG>(in, wp, bias, out, B, H, W, st)
  switch (C) {
    case 8: GO(1);
    case 16: GO(2);
    case 32: GO(4);
    case 64: GO(8);
  }
#undef GO
  return 1;
}

}  // namespace vae
