// Fused kernels for the output end of the decoder (models.py:76-82, run at :171):
//
//     y = ConvT(...)  ->  BatchNorm2d  ->  LeakyReLU  ->  Conv2d(C -> 1, k3, s1, p1)  ->  Sigmoid
//
// y [B,H,W,C] is the LARGEST tensor of the network (51 % of all activation elements,
// SURVEY.md section 8a a10) and everything here is HBM-bound (the C -> 1 conv has an
// arithmetic intensity of ~9 flop/B), so the three kernels are built around touching y
// as few times as possible -- the normalised activation and its gradient are never
// stored:
//
//   forward            logits = conv(lrelu(bn(y)))           reads y once, writes logits + probs
//   backward, pass 1   du = (dlogits (*) w) * lrelu'(bn(y)); reduces  sum du, sum du*xhat  (BatchNorm
//                      backward), the final conv's weight gradient  sum act (x) dlogits  and its
//                      bias gradient                          reads y once, writes nothing big
//   backward, pass 2   dy = gamma*rstd*(du - mean(du) - xhat*mean(du*xhat))
//                                                            reads y once, writes dy once
//
// against seven passes over tensors of that size in the unfused schedule (bn_apply, conv
// forward, conv dgrad, conv wgrad, bn_bwd_reduce, bn_bwd_apply read/write).
// CUDA cores by design: the single-channel side makes these GEMMs N = 1.
#include "common.cuh"

namespace vae {

constexpr int kFS = 4;  // pixels per thread strip (along W)

__device__ __forceinline__ float bn_affine_f(float y, float mean, float scale, float beta) {
  return fmaf(y - mean, scale, beta);  // ATen's order of operations (see elementwise.cu)
}

// ---------------------------------------------------------------------------------------------
// forward: the block stages a (TH+2) x (TW+2) pixel patch of act = lrelu(bn(y)) in shared memory
// (the transform is applied ONCE per element while staging; zero outside the image = the conv's
// padding), then every thread owns 8 channels x 4 consecutive pixels, the 3x3xC weights live in
// registers, and the C/8 channel groups of a pixel are combined with warp shuffles.
// ---------------------------------------------------------------------------------------------
template <typename T, int CG>
__global__ void __launch_bounds__(256)
final_conv_fwd_kernel(const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ scale,
                      const float* __restrict__ beta, const float* __restrict__ w, const float* __restrict__ bias,
                      T* __restrict__ logits, float* __restrict__ probs, int H, int W, float slope) {
  constexpr int C = CG * 8;
  constexpr int TW = 32, NS = TW / kFS, TH = 256 / (CG * NS);
  constexpr int PW = TW + 2, PH = TH + 2;
  constexpr int PITCH = C + 16 / (int)sizeof(T);  // +16 B per pixel: conflict-free 16-byte shared loads
  extern __shared__ __align__(16) uint8_t fsm_raw[];
  T* patch = reinterpret_cast<T*>(fsm_raw);                            // [PH][PW][PITCH]
  float* wsm = reinterpret_cast<float*>(fsm_raw + sizeof(T) * PH * PW * PITCH);  // [9][C]
  const int b = blockIdx.z;
  const int w_base = blockIdx.x * TW, h_base = blockIdx.y * TH;
  for (int i = threadIdx.x; i < 9 * C; i += 256) {
    const int t = i / C, c = i - t * C;
    wsm[i] = w[c * 9 + t];  // OIHW with O == 1
  }
  {
    const int cgs = threadIdx.x % CG;  // 256 % CG == 0: a thread stages the same channel group every pass
    float mu[8], sc[8], be[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mu[k] = mean[cgs * 8 + k];
      sc[k] = scale[cgs * 8 + k];
      be[k] = beta[cgs * 8 + k];
    }
    for (int i = threadIdx.x; i < PH * PW * CG; i += 256) {
      const int pix = i / CG;
      const int pw = pix % PW, ph = pix / PW;
      const int ih = h_base + ph - 1, iw = w_base + pw - 1;
      Vec8<T> v;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        v.load(y + (((long long)b * H + ih) * W + iw) * C + cgs * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float u = bn_affine_f(v.v[k], mu[k], sc[k], be[k]);
          v.v[k] = u > 0.0f ? u : u * slope;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] = 0.0f;
      }
      v.store(patch + pix * PITCH + cgs * 8);
    }
  }
  __syncthreads();
  const int cg = threadIdx.x % CG;
  const int strip = (threadIdx.x / CG) % NS;
  const int row = threadIdx.x / (CG * NS);
  float acc[kFS];
#pragma unroll
  for (int i = 0; i < kFS; ++i) acc[i] = 0.0f;
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
    const T* prow = patch + ((row + a) * PW + strip * kFS) * PITCH + cg * 8;
#pragma unroll
    for (int c = 0; c < kFS + 2; ++c) {
      Vec8<T> x;
      x.load(prow + c * PITCH);
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const int i = c - bb;  // output pixel whose tap bb reads patch column c
        if (i >= 0 && i < kFS) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[i] = fmaf(x.v[k], wv[bb][k], acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int off = CG / 2; off >= 1; off >>= 1)
#pragma unroll
    for (int i = 0; i < kFS; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
  const int oh = h_base + row, ow = w_base + strip * kFS;
  if (cg == 0 && oh < H && ow < W) {  // W % 4 == 0 (host-checked): a strip is all in or all out
    const float bv = bias != nullptr ? bias[0] : 0.0f;
    const long long o = ((long long)b * H + oh) * W + ow;
    float pr[kFS];
    if constexpr (sizeof(T) == 2) {
      __nv_bfloat162 l01 = __floats2bfloat162_rn(acc[0] + bv, acc[1] + bv);
      __nv_bfloat162 l23 = __floats2bfloat162_rn(acc[2] + bv, acc[3] + bv);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&l01);
      pk.y = *reinterpret_cast<uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(logits + o) = pk;
      const float2 f01 = __bfloat1622float2(l01), f23 = __bfloat1622float2(l23);
      pr[0] = f01.x; pr[1] = f01.y; pr[2] = f23.x; pr[3] = f23.y;
    } else {
#pragma unroll
      for (int i = 0; i < kFS; ++i) pr[i] = acc[i] + bv;
      *reinterpret_cast<float4*>(logits + o) = make_float4(pr[0], pr[1], pr[2], pr[3]);
    }
    if (probs != nullptr) {
#pragma unroll
      for (int i = 0; i < kFS; ++i) pr[i] = 1.0f / (1.0f + expf(-pr[i]));  // aten::sigmoid (models.py:81)
      *reinterpret_cast<float4*>(probs + o) = make_float4(pr[0], pr[1], pr[2], pr[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, shared by both passes: a thread owns 8 channels x a strip of 4 pixels and reads the
// 3 x 6 window of dlogits around the strip straight from global memory (single channel: L1 hits).
// window[a][j] = dlogits[h + a - 1][w0 + j - 1]; the tap that pairs with window offset (a, b)
// in BOTH the data gradient and the weight gradient is (kh, kw) = (2 - a, 2 - b), i.e. 8 - (3a+b).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_dl_window(const T* __restrict__ dl, int b, int h, int w0, int H, int W,
                                               float win[3][kFS + 2]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int ih = h + a - 1;
    if (ih < 0 || ih >= H) {
#pragma unroll
      for (int j = 0; j < kFS + 2; ++j) win[a][j] = 0.0f;
      continue;
    }
    const T* row = dl + ((long long)b * H + ih) * W;
    win[a][0] = w0 > 0 ? to_f(row[w0 - 1]) : 0.0f;
    if constexpr (sizeof(T) == 2) {
      const uint2 raw = *reinterpret_cast<const uint2*>(row + w0);
      const float2 f01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      const float2 f23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
      win[a][1] = f01.x; win[a][2] = f01.y; win[a][3] = f23.x; win[a][4] = f23.y;
    } else {
      const float4 f = *reinterpret_cast<const float4*>(row + w0);
      win[a][1] = f.x; win[a][2] = f.y; win[a][3] = f.z; win[a][4] = f.w;
    }
    win[a][kFS + 1] = w0 + kFS < W ? to_f(row[w0 + kFS]) : 0.0f;
  }
}

struct StripIter {
  // strips are numbered (b, h, w-strip, channel group) with the channel group fastest, so the
  // 32 lanes of a warp read 32 consecutive 16-byte chunks of y
  int CG, wstrips, H;
  long long n;
  __device__ __forceinline__ void decode(long long s, int& cg, int& b, int& h, int& w0) const {
    cg = (int)(s % CG);
    long long t = s / CG;
    w0 = (int)(t % wstrips) * kFS;
    t /= wstrips;
    h = (int)(t % H);
    b = (int)(t / H);
  }
};

// pass 1
template <typename T, int CG>
__global__ void __launch_bounds__(256)
final_bwd_reduce_kernel(const T* __restrict__ dl, const T* __restrict__ y, const float* __restrict__ mean,
                        const float* __restrict__ scale, const float* __restrict__ beta,
                        const float* __restrict__ rstd, const float* __restrict__ w, double* __restrict__ dstats,
                        float* __restrict__ dw_last, int B, int H, int W, float slope) {
  constexpr int C = CG * 8;
  constexpr int LANES = 256 / CG;
  extern __shared__ __align__(16) float rsm[];  // wsm [9][C]; later the reduction scratch
  float* wsm = rsm;
  for (int i = threadIdx.x; i < 9 * C; i += 256) {
    const int t = i / C, c = i - t * C;
    wsm[i] = w[c * 9 + t];
  }
  __syncthreads();
  const int cg = threadIdx.x % CG;  // fixed per thread: gridDim.x * 256 % CG == 0
  float mu[8], sc[8], be[8], rs[8], s1[8], s2[8], acc[9][8];
  float sdl = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mu[k] = mean[cg * 8 + k];
    sc[k] = scale[cg * 8 + k];
    be[k] = beta[cg * 8 + k];
    rs[k] = rstd[cg * 8 + k];
    s1[k] = s2[k] = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t][k] = 0.0f;
  }
  StripIter it{CG, W / kFS, H, (long long)B * H * (W / kFS) * CG};
  for (long long s = (long long)blockIdx.x * 256 + threadIdx.x; s < it.n; s += (long long)gridDim.x * 256) {
    int cg_, b, h, w0;
    it.decode(s, cg_, b, h, w0);
    const T* src = y + (((long long)b * H + h) * W + w0) * C + cg * 8;
    Vec8<T> yv[kFS];
#pragma unroll
    for (int i = 0; i < kFS; ++i) yv[i].load(src + (long long)i * C);
    float win[3][kFS + 2];
    load_dl_window(dl, b, h, w0, H, W, win);
    if (cg == 0) sdl += win[1][1] + win[1][2] + win[1][3] + win[1][4];
    float da[kFS][8];
#pragma unroll
    for (int i = 0; i < kFS; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) da[i][k] = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const float4* w4 = reinterpret_cast<const float4*>(wsm + (8 - (a * 3 + bb)) * C + cg * 8);
        const float4 w0v = w4[0], w1v = w4[1];
        const float wv[8] = {w0v.x, w0v.y, w0v.z, w0v.w, w1v.x, w1v.y, w1v.z, w1v.w};
#pragma unroll
        for (int i = 0; i < kFS; ++i) {
          const float v = win[a][i + bb];
#pragma unroll
          for (int k = 0; k < 8; ++k) da[i][k] = fmaf(v, wv[k], da[i][k]);
        }
      }
#pragma unroll
    for (int i = 0; i < kFS; ++i) {
      float act[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yy = yv[i].v[k];
        const float u = bn_affine_f(yy, mu[k], sc[k], be[k]);
        const float du = u > 0.0f ? da[i][k] : da[i][k] * slope;
        const float xh = (yy - mu[k]) * rs[k];
        act[k] = u > 0.0f ? u : u * slope;
        s1[k] += du;
        s2[k] = fmaf(du, xh, s2[k]);
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
          const float v = win[a][i + bb];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[8 - (a * 3 + bb)][k] = fmaf(act[k], v, acc[8 - (a * 3 + bb)][k]);
        }
    }
  }
  // block reduction over the LANES threads of each channel group, 8 values per round
  __syncthreads();
  float* red = rsm;  // [LANES][CG][8]
  const int lane_id = threadIdx.x / CG;
  auto reduce8 = [&](const float* vals, int round_kind, int t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[(lane_id * CG + cg) * 8 + k] = vals[k];
    __syncthreads();
    if (threadIdx.x < C) {
      const int c = threadIdx.x, g = c / 8, k = c % 8;
      double a = 0.0;
      for (int l = 0; l < LANES; ++l) a += (double)red[(l * CG + g) * 8 + k];
      if (round_kind == 0) atomicAdd(&dstats[c], a);
      else if (round_kind == 1) atomicAdd(&dstats[C + c], a);
      else atomicAdd(&dw_last[c * 9 + t], (float)a);
    }
    __syncthreads();
  };
  reduce8(s1, 0, 0);
  reduce8(s2, 1, 0);
#pragma unroll
  for (int t = 0; t < 9; ++t) reduce8(acc[t], 2, t);
  // bias gradient of the final conv: sum of dlogits
  double d = warp_sum((double)sdl);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (float)d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += (double)red[i];
    atomicAdd(&dstats[2 * C], a);
  }
}

// pass 2
template <typename T, int CG>
__global__ void __launch_bounds__(256)
final_bwd_apply_kernel(const T* __restrict__ dl, const T* __restrict__ y, const float* __restrict__ mean,
                       const float* __restrict__ scale, const float* __restrict__ beta,
                       const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ w,
                       const double* __restrict__ dstats, T* __restrict__ dy, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dbias, float* __restrict__ db_last, int B, int H,
                       int W, float slope, int training) {
  constexpr int C = CG * 8;
  constexpr int LANES = 256 / CG;
  extern __shared__ __align__(16) float rsm[];
  float* wsm = rsm;
  for (int i = threadIdx.x; i < 9 * C; i += 256) {
    const int t = i / C, c = i - t * C;
    wsm[i] = w[c * 9 + t];
  }
  __syncthreads();
  const double inv_m = 1.0 / ((double)B * H * W);
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += 256) {
      dbeta[c] = (float)dstats[c];
      dgamma[c] = (float)dstats[C + c];
    }
    if (threadIdx.x == 0) db_last[0] = (float)dstats[2 * C];
  }
  const int cg = threadIdx.x % CG;
  float mu[8], sc[8], be[8], rs[8], ga[8], c1[8], c2[8], sdy[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cg * 8 + k;
    mu[k] = mean[c];
    sc[k] = scale[c];
    be[k] = beta[c];
    rs[k] = rstd[c];
    ga[k] = gamma[c] * rs[k];
    c1[k] = training ? (float)(dstats[c] * inv_m) : 0.0f;
    c2[k] = training ? (float)(dstats[C + c] * inv_m) : 0.0f;
    sdy[k] = 0.0f;
  }
  StripIter it{CG, W / kFS, H, (long long)B * H * (W / kFS) * CG};
  for (long long s = (long long)blockIdx.x * 256 + threadIdx.x; s < it.n; s += (long long)gridDim.x * 256) {
    int cg_, b, h, w0;
    it.decode(s, cg_, b, h, w0);
    const long long off = (((long long)b * H + h) * W + w0) * C + cg * 8;
    Vec8<T> yv[kFS];
#pragma unroll
    for (int i = 0; i < kFS; ++i) yv[i].load(y + off + (long long)i * C);
    float win[3][kFS + 2];
    load_dl_window(dl, b, h, w0, H, W, win);
    float da[kFS][8];
#pragma unroll
    for (int i = 0; i < kFS; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) da[i][k] = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const float4* w4 = reinterpret_cast<const float4*>(wsm + (8 - (a * 3 + bb)) * C + cg * 8);
        const float4 w0v = w4[0], w1v = w4[1];
        const float wv[8] = {w0v.x, w0v.y, w0v.z, w0v.w, w1v.x, w1v.y, w1v.z, w1v.w};
#pragma unroll
        for (int i = 0; i < kFS; ++i) {
          const float v = win[a][i + bb];
#pragma unroll
          for (int k = 0; k < 8; ++k) da[i][k] = fmaf(v, wv[k], da[i][k]);
        }
      }
#pragma unroll
    for (int i = 0; i < kFS; ++i) {
      Vec8<T> o;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yy = yv[i].v[k];
        const float u = bn_affine_f(yy, mu[k], sc[k], be[k]);
        const float du = u > 0.0f ? da[i][k] : da[i][k] * slope;
        const float xh = (yy - mu[k]) * rs[k];
        const float r = ga[k] * (du - c1[k] - xh * c2[k]);
        o.v[k] = r;
        sdy[k] += r;
      }
      o.store(dy + off + (long long)i * C);
    }
  }
  __syncthreads();
  float* red = rsm;
  const int lane_id = threadIdx.x / CG;
#pragma unroll
  for (int k = 0; k < 8; ++k) red[(lane_id * CG + cg) * 8 + k] = sdy[k];
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x, g = c / 8, k = c % 8;
    float a = 0.0f;
    for (int l = 0; l < LANES; ++l) a += red[(l * CG + g) * 8 + k];
    atomicAdd(&dbias[c], a);
  }
}

// ============================================================================================
template <typename K>
static int persistent_grid(K kern, size_t smem, long long strips) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long blocks = (strips + 255) / 256;
  long long cap = (long long)sm_count() * per_sm;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

static bool final_shape_ok(int B, int H, int W, int C) {
  return B > 0 && B <= 65535 && H > 0 && W > 0 && W % kFS == 0 && (C == 8 || C == 16 || C == 32 || C == 64);
}

template <typename T, int CG>
static int fwd_launch(const void* y, const float* mean, const float* scale, const float* beta, const float* w,
                      const float* bias, void* logits, float* probs, int B, int H, int W, float slope,
                      cudaStream_t st) {
  constexpr int C = CG * 8, TH = 256 / (CG * 8), PITCH = C + 16 / (int)sizeof(T);
  const size_t smem = sizeof(T) * (TH + 2) * 34 * PITCH + sizeof(float) * 9 * C;
  auto kern = final_conv_fwd_kernel<T, CG>;
  static bool attr = false;
  if (!attr) {
    VAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid(ceil_div(W, 32), ceil_div(H, TH), B);
  kern<<<grid, 256, smem, st>>>((const T*)y, mean, scale, beta, w, bias, (T*)logits, probs, H, W, slope);
  return check_launch("final_conv_fwd");
}

template <typename T, int CG>
static int reduce_launch(const void* dl, const void* y, const float* mean, const float* scale, const float* beta,
                         const float* rstd, const float* w, double* dstats, float* dw_last, int B, int H, int W,
                         float slope, cudaStream_t st) {
  constexpr int C = CG * 8;
  const size_t smem = sizeof(float) * (256 * 8 > 9 * C ? 256 * 8 : 9 * C);
  auto kern = final_bwd_reduce_kernel<T, CG>;
  const int grid = persistent_grid(kern, smem, (long long)B * H * (W / kFS) * CG);
  kern<<<grid, 256, smem, st>>>((const T*)dl, (const T*)y, mean, scale, beta, rstd, w, dstats, dw_last, B, H, W,
                                slope);
  return check_launch("final_bwd_reduce");
}

template <typename T, int CG>
static int apply_launch(const void* dl, const void* y, const float* mean, const float* scale, const float* beta,
                        const float* rstd, const float* gamma, const float* w, const double* dstats, void* dy,
                        float* dgamma, float* dbeta, float* dbias, float* db_last, int B, int H, int W, float slope,
                        int training, cudaStream_t st) {
  constexpr int C = CG * 8;
  const size_t smem = sizeof(float) * (256 * 8 > 9 * C ? 256 * 8 : 9 * C);
  auto kern = final_bwd_apply_kernel<T, CG>;
  const int grid = persistent_grid(kern, smem, (long long)B * H * (W / kFS) * CG);
  kern<<<grid, 256, smem, st>>>((const T*)dl, (const T*)y, mean, scale, beta, rstd, gamma, w, dstats, (T*)dy, dgamma,
                                dbeta, dbias, db_last, B, H, W, slope, training);
  return check_launch("final_bwd_apply");
}

}  // namespace vae

using namespace vae;

#define VAE_FINAL_DISPATCH(FN, ...)                                   \
  do {                                                                \
    const bool bf = dtype == VAE_BF16;                                \
    switch (C) {                                                      \
      case 8: return bf ? FN<bf16, 1>(__VA_ARGS__) : FN<float, 1>(__VA_ARGS__);   \
      case 16: return bf ? FN<bf16, 2>(__VA_ARGS__) : FN<float, 2>(__VA_ARGS__);  \
      case 32: return bf ? FN<bf16, 4>(__VA_ARGS__) : FN<float, 4>(__VA_ARGS__);  \
      case 64: return bf ? FN<bf16, 8>(__VA_ARGS__) : FN<float, 8>(__VA_ARGS__);  \
    }                                                                 \
  } while (0)

/* 1 when the fused final-block kernels serve this shape (else the caller composes
 * vae_bn_apply_lrelu + vae_conv_fprop / dgrad / wgrad + vae_bn_lrelu_bwd_*). */
extern "C" int vae_final_block_supported(int B, int H, int W, int C) { return final_shape_ok(B, H, W, C) ? 1 : 0; }

extern "C" int vae_final_conv_fwd(const void* y, const float* mean, const float* scale, const float* beta,
                                  const float* w_last, const float* b_last, void* logits, float* probs, int B, int H,
                                  int W, int C, float slope, int dtype, vae_stream_t stream) {
  VAE_REQUIRE(y && mean && scale && beta && w_last && logits, "final_conv_fwd: null argument");
  VAE_REQUIRE(final_shape_ok(B, H, W, C), "final_conv_fwd: shape B=%d H=%d W=%d C=%d not served", B, H, W, C);
  VAE_REQUIRE(dtype == VAE_F32 || dtype == VAE_BF16, "final_conv_fwd: bad dtype");
  cudaStream_t st = as_stream(stream);
  VAE_FINAL_DISPATCH(fwd_launch, y, mean, scale, beta, w_last, b_last, logits, probs, B, H, W, slope, st);
  return VAE_ERR_UNSUPPORTED;
}

extern "C" int vae_final_bwd_reduce(const void* dlogits, const void* y, const float* mean, const float* scale,
                                    const float* beta, const float* rstd, const float* w_last, double* dstats,
                                    float* dw_last, int B, int H, int W, int C, float slope, int dtype,
                                    vae_stream_t stream) {
  VAE_REQUIRE(dlogits && y && mean && scale && beta && rstd && w_last && dstats && dw_last,
              "final_bwd_reduce: null argument");
  VAE_REQUIRE(final_shape_ok(B, H, W, C), "final_bwd_reduce: shape B=%d H=%d W=%d C=%d not served", B, H, W, C);
  VAE_REQUIRE(dtype == VAE_F32 || dtype == VAE_BF16, "final_bwd_reduce: bad dtype");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(dstats, sizeof(double) * (2 * C + 1), st));
  VAE_CUDA(zero_scratch(dw_last, sizeof(float) * 9 * C, st));
  VAE_FINAL_DISPATCH(reduce_launch, dlogits, y, mean, scale, beta, rstd, w_last, dstats, dw_last, B, H, W, slope, st);
  return VAE_ERR_UNSUPPORTED;
}

extern "C" int vae_final_bwd_apply(const void* dlogits, const void* y, const float* mean, const float* scale,
                                   const float* beta, const float* rstd, const float* gamma, const float* w_last,
                                   const double* dstats, void* dy, float* dgamma, float* dbeta, float* dbias,
                                   float* db_last, int B, int H, int W, int C, float slope, int training, int dtype,
                                   vae_stream_t stream) {
  VAE_REQUIRE(dlogits && y && mean && scale && beta && rstd && gamma && w_last && dstats && dy && dgamma && dbeta &&
                  dbias && db_last,
              "final_bwd_apply: null argument");
  VAE_REQUIRE(final_shape_ok(B, H, W, C), "final_bwd_apply: shape B=%d H=%d W=%d C=%d not served", B, H, W, C);
  VAE_REQUIRE(dtype == VAE_F32 || dtype == VAE_BF16, "final_bwd_apply: bad dtype");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(dbias, sizeof(float) * C, st));
  VAE_FINAL_DISPATCH(apply_launch, dlogits, y, mean, scale, beta, rstd, gamma, w_last, dstats, dy, dgamma, dbeta,
                     dbias, db_last, B, H, W, slope, training, st);
  return VAE_ERR_UNSUPPORTED;
}
