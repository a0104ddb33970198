// TMA-fed tcgen05 / TMEM implicit-GEMM convolution (bf16 x bf16 -> fp32), sm_100a.
//
// Same GEMM as conv_umma.cu (D[128 pixels x N] += A[128 x K] * B[N x K]^T, geometry
// in conv_geom.cuh) but the A operand is staged by the TMA engine instead of by
// cp.async threads.  The trick that makes a plain TILED tensor map do the im2col
// gather: a 128-pixel GEMM tile is made of whole output rows (128 = n x h x w with
// w = min(Wj,128)), so for one filter tap its input pixels form a strided BOX of
// the NHWC tensor  {C: 64 channels, W: w pixels step in_mul, H: h rows step in_mul,
// N: n images}  starting at (in_mul*jw0 + dw, in_mul*jh0 + dh).  One
// cp.async.bulk.tensor.4d per tap moves it into the 128B/64B-swizzled K-major
// sub-tile UMMA reads; coordinates outside the image (the 3x3 padding, or the far
// edge of a transposed conv) are zero-filled by the TMA unit.  A single thread
// issues everything: no address arithmetic, no LSU instructions, no registers.
//
//   warp 0     producer: per stage wait(empty) -> expect_tx -> 1-2 tensor TMA loads
//              (A sub-tiles) + 1 bulk copy (pre-swizzled weight image)
//   warp 1     MMA issuer (one elected thread), TMEM accumulator double-buffered
//   warps 2-9  epilogue: two sets of four warps, one per TMEM accumulator buffer (even /
//              odd tiles): tcgen05.ld, bias, BatchNorm statistics, bf16 NHWC store
//
// Grid = #SMs, persistent static tile schedule.  Layers whose extents do not
// decompose into whole rows fall back to the cp.async gather kernel.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "conv_geom.cuh"
#include "umma.cuh"
#include "tma_util.cuh"

namespace vae {

using namespace umma;

constexpr int kTStagesMax = 10;
constexpr int kTTileM = 128;
constexpr int kTThreads = 320;  // producer warp, MMA warp, 2 x 4 epilogue warps (one set per TMEM accumulator)

long long* umma_trace_ptr();

// fused-phase transposed conv (conv_umma_tmat.cu): serves mode 1 with 64 / 128 output channels behind this file's entry points
bool tmat_channels_ok(int Ci, int Co, int stride, int mode);
long long tmat_packed_elems(int Ci, int Co);
int tmat_pack(const float* w_oihw, void* out, int Ci, int Co, cudaStream_t st);
int tmat_launch(const void* in, const void* wpk, const float* bias, void* out, double* stats, int B, int Hi, int Wi, int Ci,
                int Co, cudaStream_t st, const void* bn_y, const float* const* bn_vec, float slope, const float* const* act);

struct TmaConvParams {
  CUtensorMap tmap;
  GatherGeom g;
  int nkb[4];          // stages per phase
  int b_off[4];        // element offset of the phase's first weight image
  int tiles_per_phase, total_tiles;
  int N;               // output channels
  int kc;              // channels per sub-tile: 32 (64-byte swizzle) or 64 (128-byte swizzle)
  int cin_chunks;      // sub-tiles per tap = Cin / kc
  int sps;             // sub-tiles per stage (2 when kc == 32, else 1)
  int sub_a_bytes;     // 128 * kc * 2
  int sub_b_elems;     // N * kc
  int a_stage_bytes, b_stage_bytes;
  int stages, tmem_cols;
  int rows_w, rows_h, rows_n;  // 128 = rows_n * rows_h * rows_w (whole output rows)
  const bf16* wpk;
  const float* bias;
  bf16* out;
  int ldo;             // channels per pixel of the produced tensor (> N when the launch writes a channel slice)
  double* stats;       // [N] sums of this launch's channels
  double* stats2;      // [N] sums of squares
  // fused BatchNorm-backward prologue of the next hop (see conv_umma_rw.cu): with bn_y set the epilogue writes
  // du = acc * LeakyReLU'(u(y)) and the statistics become sum du / sum du * xhat
  const bf16* bn_y;
  const float *bn_mean, *bn_scale, *bn_beta, *bn_rstd;   // [N] each (scale = gamma * rstd)
  float slope;
  const float *act_mean, *act_scale, *act_beta;   // inference epilogue: eval-mode BatchNorm + LeakyReLU (see conv_umma_rw.cu)
  long long* trace;  // optional [6][512] clock64 samples of CTA 0 (tuning aid)
};

// weight images for this kernel: per phase, per stage: [sub][N rows][kc] bf16, rows
// swizzled like the A sub-tiles (64B: chunk ^= (n>>1)&3, 128B: chunk ^= n&7).
// sub-k-block q = stage*sps + sub -> tap = q / cin_chunks, channel chunk = q % cin_chunks.
__global__ void pack_tma_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out,
                                       const __grid_constant__ TmaConvParams P, int Co, int Ci, int mode) {
  const int phase = blockIdx.y;
  const TapTable& tt = P.g.phase[phase];
  const int stage_elems = P.sps * P.sub_b_elems;
  const long long n_elem = (long long)P.nkb[phase] * stage_elems;
  bf16* dst = out + P.b_off[phase];
  const int cpr = P.kc / 8;  // 16-byte chunks per row
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % P.kc);
    const int n = (int)((i / P.kc) % P.N);
    const int sub = (int)((i / P.sub_b_elems) % P.sps);
    const int st = (int)(i / stage_elems);
    const int q = st * P.sps + sub;
    const int tap_i = q / P.cin_chunks;
    float v = 0.0f;
    if (tap_i < tt.ntaps) {
      const int c = (q - tap_i * P.cin_chunks) * P.kc + e;
      const int t = tt.wt[tap_i];
      const int co = mode == 0 ? n : c, ci = mode == 0 ? c : n;
      v = w[((long long)co * Ci + ci) * 9 + t];
    }
    const int chunk = e >> 3;
    const int sw = cpr == 4 ? (chunk ^ ((n >> 1) & 3)) : (chunk ^ (n & 7));
    dst[(long long)st * stage_elems + (long long)sub * P.sub_b_elems + (long long)n * P.kc + sw * 8 + (e & 7)] =
        __float2bfloat16_rn(v);
  }
}

#define VAE_TTRACE(role, idx)                                                             \
  do {                                                                                   \
    if (P.trace != nullptr && blockIdx.x == 0 && lane == 0 && (idx) < 512)                \
      P.trace[(role) * 512 + (idx)] = clock64();                                          \
  } while (0)

// EPI (plain epilogue, !BNB): 0 = nothing added (data gradients), 1 = bias + per-chunk BatchNorm statistics, 2 = bias +
// eval-mode BatchNorm + LeakyReLU (inference), 3 = bias + statistics kept in registers for the whole kernel (N <= 32).
template <bool BNB, int EPI>
__global__ void __launch_bounds__(kTThreads, 1)
umma_tma_conv_kernel(const __grid_constant__ TmaConvParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int stage_bytes = P.a_stage_bytes + P.b_stage_bytes;
  uint8_t* tail = smem + P.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kTStagesMax;
  uint64_t* tfull_bar = empty_bar + kTStagesMax;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_stats = reinterpret_cast<float*>(tmem_ptr + 4);  // [2][256]
  float* s_bias = s_stats + 512;                            // [256]
  uint32_t* s_taps = reinterpret_cast<uint32_t*>(s_bias + 256);  // [4 phases][dw pack, dh pack, sub-tiles, -]
  float4* s_bn = reinterpret_cast<float4*>(s_taps + 16);         // [256] {mean, scale, beta, rstd} (BNB only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatherGeom& g = P.g;
  const int N = P.N;
  const int kStages = P.stages;

  for (int i = threadIdx.x; i < 512; i += kTThreads) s_stats[i] = 0.0f;
  for (int i = threadIdx.x; i < 256; i += kTThreads) s_bias[i] = (P.bias != nullptr && i < N) ? P.bias[i] : 0.0f;
  if constexpr (BNB)
    for (int i = threadIdx.x; i < N; i += kTThreads) s_bn[i] = make_float4(P.bn_mean[i], P.bn_scale[i], P.bn_beta[i], P.bn_rstd[i]);
  else if (EPI == 2)   // inference: u = y * scale + shift, two channels per float4
    for (int i = threadIdx.x; i < N; i += kTThreads)
      reinterpret_cast<float2*>(s_bn)[i] = make_float2(P.act_scale[i], fmaf(-P.act_mean[i], P.act_scale[i], P.act_beta[i]));
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kTStagesMax; ++s) {
        mbar_init(&full_bar[s], 1);   // the producer's expect_tx arrive; TMA completes the bytes
        mbar_init(&empty_bar[s], 1);  // one tcgen05.commit
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, (uint32_t)P.tmem_cols);
  }
  if (warp == 0 && lane == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap) : "memory");
  if (threadIdx.x < 4) {  // tap offsets of a phase, 2 bits each (value + 1)
    const TapTable& tt = P.g.phase[threadIdx.x];
    uint32_t dwp = 0, dhp = 0;
    for (int t = 0; t < tt.ntaps; ++t) {
      dwp |= (uint32_t)(tt.dw[t] + 1) << (2 * t);
      dhp |= (uint32_t)(tt.dh[t] + 1) << (2 * t);
    }
    s_taps[threadIdx.x * 4] = dwp;
    s_taps[threadIdx.x * 4 + 1] = dhp;
    s_taps[threadIdx.x * 4 + 2] = (uint32_t)(tt.ntaps * P.cin_chunks);
    s_taps[threadIdx.x * 4 + 3] = 0u;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const long long M_total = (long long)g.B * g.Hj * g.Wj;

  if (warp == 0) {
    // ============================== producer ==============================
    // One thread issues every load; its instruction stream is latency-bound (one warp, dependent
    // scalar code), so everything that can be hoisted is: parameters live in registers, the tap
    // offsets of a phase are two packed words, (tap, channel chunk) advance incrementally.
    if (lane == 0) {
      int stage = 0, seq = 0;
      uint32_t ph = 0;
      const uint32_t b_bytes = (uint32_t)P.b_stage_bytes, sub_a = (uint32_t)P.sub_a_bytes;
      const uint32_t a_stage = (uint32_t)P.a_stage_bytes, stage_b = (uint32_t)stage_bytes;
      const int sps = P.sps, cin_chunks = P.cin_chunks, kc = P.kc, in_mul = g.in_mul;
      const uint32_t Wj = (uint32_t)g.Wj, Hj = (uint32_t)g.Hj;
      const int tpp = P.tiles_per_phase, total = P.total_tiles;
      const uint32_t smem_a = smem_u32(smem);
      (void)smem_a;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int phase = tile / tpp;
        const int mt = tile - phase * tpp;
        const uint32_t dwp = s_taps[phase * 4], dhp = s_taps[phase * 4 + 1];
        int rem = (int)s_taps[phase * 4 + 2];  // sub-tiles of this tile: taps x channel chunks
        const uint32_t m0 = (uint32_t)mt * kTTileM;
        const uint32_t jw0 = m0 % Wj, t2 = m0 / Wj;
        const int jh0 = (int)(t2 % Hj), b0 = (int)(t2 / Hj);
        const int cw = (int)jw0 * in_mul, ch = jh0 * in_mul;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.wpk + P.b_off[phase]);
        int tap_i = 0, cc = 0;
        while (rem > 0) {
          VAE_TTRACE(0, seq);
          mbar_wait(&empty_bar[stage], ph ^ 1u);
          uint8_t* a_st = smem + (uint32_t)stage * stage_b;
          const int nsub = rem < sps ? rem : sps;
          rem -= nsub;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)nsub * sub_a + b_bytes);
          for (int sub = 0; sub < nsub; ++sub) {
            const int dw = (int)((dwp >> (2 * tap_i)) & 3u) - 1, dh = (int)((dhp >> (2 * tap_i)) & 3u) - 1;
            tma_load_4d(a_st + (uint32_t)sub * sub_a, &P.tmap, &full_bar[stage], cc * kc, cw + dw, ch + dh, b0);
            if (++cc == cin_chunks) { cc = 0; ++tap_i; }
          }
          bulk_g2s(a_st + a_stage, wsrc, b_bytes, &full_bar[stage]);
          wsrc += b_bytes;
          VAE_TTRACE(1, seq);
          ++seq;
          if (++stage == kStages) { stage = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(kTTileM, (uint32_t)N, 0, 0);
    const uint32_t layout = P.kc == 32 ? 4u : 2u;        // SWIZZLE_64B : SWIZZLE_128B
    const uint32_t sbo = P.kc == 32 ? 512u : 1024u;      // 8 rows of 64 / 128 bytes
    // descriptor = {hi: SBO | version | swizzle, lo: address >> 4 | LBO(=1) << 16}; addresses < 256 KB, so
    // a descriptor is advanced with one 32-bit add on its low word
    const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    const int ksteps = P.kc / 16, sps = P.sps;
    const uint32_t sub_a16 = (uint32_t)P.sub_a_bytes >> 4, sub_b16 = (uint32_t)(P.sub_b_elems * 2) >> 4;
    const uint32_t a_stage = (uint32_t)P.a_stage_bytes, stage_b = (uint32_t)stage_bytes;
    const uint32_t smem_a = smem_u32(smem);
    const int tpp = P.tiles_per_phase, total = P.total_tiles;
    int stage = 0, seq = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int phase = tile / tpp;
      int rem = (int)s_taps[phase * 4 + 2];
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
      uint32_t accum = 0u;
      while (rem > 0) {
        mbar_wait(&full_bar[stage], ph);
        VAE_TTRACE(2, seq);
        tc_fence_after();
        const int nsub = rem < sps ? rem : sps;
        rem -= nsub;
        if (elect_one()) {
          const uint32_t a_addr = smem_a + (uint32_t)stage * stage_b;
          uint32_t lo_a = (a_addr >> 4) | (1u << 16);
          uint32_t lo_b = ((a_addr + a_stage) >> 4) | (1u << 16);
          for (int sub = 0; sub < nsub; ++sub) {
            for (int k = 0; k < ksteps; ++k) {
              mma_bf16_ss(d_tmem, ((uint64_t)hi << 32) | (lo_a + 2u * k), ((uint64_t)hi << 32) | (lo_b + 2u * k), idesc,
                          accum);
              accum = 1u;
            }
            lo_a += sub_a16;
            lo_b += sub_b16;
          }
          mma_commit(&empty_bar[stage]);
          if (rem == 0) mma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        VAE_TTRACE(3, seq);
        ++seq;
        if (++stage == kStages) { stage = 0; ph ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int ew = warp & 3;           // TMEM lane group = warp % 4
    const int eset = (warp - 2) >> 2;  // warps 2-5 serve accumulator 0 (even tiles), 6-9 accumulator 1
    const int r = ew * 32 + lane;
    const bool has_bias = P.bias != nullptr;
    const uint32_t s_bn_a = smem_u32(s_bn);
    constexpr int RUN = (!BNB && EPI == 3) ? 32 : 1;   // sums kept for the whole kernel: only when one chunk covers N
    float run1[RUN], run2[RUN];
#pragma unroll
    for (int i = 0; i < RUN; ++i) run1[i] = run2[i] = 0.0f;
    for (int it = eset, tile = blockIdx.x + eset * gridDim.x; tile < P.total_tiles; tile += 2 * gridDim.x, it += 2) {
      const int phase = tile / P.tiles_per_phase;
      const int mt = tile - phase * P.tiles_per_phase;
      const TapTable& tt = g.phase[phase];
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const long long m = (long long)mt * kTTileM + r;
      const bool valid = m < M_total;
      bf16* dst = nullptr;
      if (valid) {
        const uint32_t mu = (uint32_t)m;
        const uint32_t jw = mu % (uint32_t)g.Wj, t2 = mu / (uint32_t)g.Wj;
        const uint32_t jh = t2 % (uint32_t)g.Hj, b = t2 / (uint32_t)g.Hj;
        const int oh = (int)jh * g.out_mul + tt.ph, ow = (int)jw * g.out_mul + tt.pw;
        dst = P.out + (((long long)b * g.Hout + oh) * g.Wout + ow) * P.ldo;
      }
      // fused BN backward: y of this thread's pixel, 32 channels at a time, requested one chunk ahead
      uint4 ycur[4], ynext[4];
      const bf16* ysrc = nullptr;
      if constexpr (BNB) {
        ysrc = P.bn_y + (dst != nullptr ? (long long)(dst - P.out) : 0);
#pragma unroll
        for (int q = 0; q < 4; ++q) ycur[q] = ldg_nc_v4_pinned(reinterpret_cast<const uint4*>(ysrc) + q);
      }
      mbar_wait(&tfull_bar[acc], acc_ph);
      if (ew == 0) VAE_TTRACE(4, it);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * N);
      if constexpr (BNB) {
        for (int c0 = 0; c0 < N; c0 += 32) {
          if constexpr (BNB) {
            if (c0 + 32 < N) {
  #pragma unroll
              for (int q = 0; q < 4; ++q) ynext[q] = ldg_nc_v4_pinned(reinterpret_cast<const uint4*>(ysrc + c0 + 32) + q);
            }
          }
          float v[32];
          tmem_ld_32x32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          if (c0 + 32 >= N) {  // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (ew == 0) VAE_TTRACE(5, it);
          }
          const int ncol = min(32, N - c0);
          if (has_bias) {   // uniform: the data-gradient launches of the backward pass have no bias
  #pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[(c0 + i) & 255]);
              v[i] += b4.x;
              v[i + 1] += b4.y;
              v[i + 2] += b4.z;
              v[i + 3] += b4.w;
            }
          }
          if (!valid || ncol < 32) {   // rows past the end of the tensor / columns past N: keep them out of the statistics
  #pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!valid || i >= ncol) v[i] = 0.0f;
          }
          float second[32];   // second statistic: y^2 (forward) or du * xhat (fused BN backward)
          if constexpr (BNB) {
  #pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t w4[4] = {ycur[q].x, ycur[q].y, ycur[q].z, ycur[q].w};
  #pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int i = q * 8 + k;
                const float yv = __uint_as_float((k & 1) ? (w4[k >> 1] & 0xFFFF0000u) : (w4[k >> 1] << 16));
                const float4 c4 = lds_f4_addr(s_bn_a + (uint32_t)(((c0 + i) & 255) * 16));   // mean, scale, beta, rstd
                const float t = yv - c4.x;
                const float u = fmaf(t, c4.y, c4.z);           // same expression as the forward apply (bn_affine)
                const float du = u > 0.0f ? v[i] : v[i] * P.slope;
                v[i] = du;
                second[i] = du * (t * c4.w);
              }
            }
  #pragma unroll
            for (int q = 0; q < 4; ++q) ycur[q] = ynext[q];
          } else {
  #pragma unroll
            for (int i = 0; i < 32; ++i) second[i] = v[i] * v[i];
          }
          if (valid) {
  #pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (i < ncol) {
                uint4 pk;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
  #pragma unroll
                for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[i + 2 * k], v[i + 2 * k + 1]);
                *reinterpret_cast<uint4*>(dst + c0 + i) = pk;
              }
            }
          }
          if (P.stats != nullptr) {
            if (!BNB && N <= 32) {
              if constexpr (!BNB) {
  #pragma unroll
                for (int i = 0; i < 32; ++i) {
                  run1[i] += v[i];
                  run2[i] += second[i];
                }
              }
            } else {
              warp_transpose_reduce32(v, lane);
              warp_transpose_reduce32(second, lane);
              if (lane < ncol) {
                atomicAdd(&s_stats[c0 + lane], v[0]);
                atomicAdd(&s_stats[256 + c0 + lane], second[0]);
              }
            }
          }
        }
      } else {
        // Plain epilogue, one flat pipeline over the 32-column chunks: bias requested before the TMEM wait, the NEXT
        // chunk's TMEM load issued as soon as this one is packed, 32-byte stores (whole sectors) from registers.
        const uint32_t s_bias_a = smem_u32(s_bias);
        const int n_ch = (N + 31) >> 5;
        float v[32];
        tmem_ld_32x32(t_row, v);
        for (int ch = 0; ch < n_ch; ++ch) {
          const int c0 = ch << 5;
          float4 b4[8];
          if constexpr (EPI != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) b4[i] = lds_f4_addr(s_bias_a + (uint32_t)(((c0 + 4 * i) & 255) * 4));
          }
          tmem_ld_wait_regs(v);
          if (ch == n_ch - 1) {   // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (ew == 0) VAE_TTRACE(5, it);
          }
          const int ncol = min(32, N - c0);
          if constexpr (EPI != 0) {   // zeros when the launch has no bias
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w;
            }
          }
          if constexpr (EPI == 2) {   // inference: eval-mode BatchNorm + LeakyReLU on the accumulator row
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float4 c4 = lds_f4_addr(s_bn_a + (uint32_t)(((c0 + i) & 255) * 8));   // scale, shift of two channels
              const float u0 = fmaf(v[i], c4.x, c4.y), u1 = fmaf(v[i + 1], c4.z, c4.w);
              v[i] = u0 > 0.0f ? u0 : u0 * P.slope;
              v[i + 1] = u1 > 0.0f ? u1 : u1 * P.slope;
            }
          }
          if (!valid || ncol < 32) {   // rows past the end of the tensor / columns past N: keep them out of the statistics
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!valid || i >= ncol) v[i] = 0.0f;
          }
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          float t1[32], t2[32];
          if constexpr (EPI == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              t1[i] = v[i];
              t2[i] = v[i] * v[i];
            }
          }
          if constexpr (EPI == 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              run1[i] += v[i];
              run2[i] = fmaf(v[i], v[i], run2[i]);
            }
          }
          if (ch + 1 < n_ch) tmem_ld_32x32(t_row + (uint32_t)(c0 + 32), v);
          if (valid) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
              if (16 * q < ncol)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + c0 + 16 * q), "r"(w[8 * q]),
                             "r"(w[8 * q + 1]), "r"(w[8 * q + 2]), "r"(w[8 * q + 3]), "r"(w[8 * q + 4]), "r"(w[8 * q + 5]),
                             "r"(w[8 * q + 6]), "r"(w[8 * q + 7]) : "memory");
          }
          if constexpr (EPI == 1) {
            if (P.stats != nullptr) {
              warp_transpose_reduce32(t1, lane);
              warp_transpose_reduce32(t2, lane);
              if (lane < ncol) {
                atomicAdd(&s_stats[c0 + lane], t1[0]);
                atomicAdd(&s_stats[256 + c0 + lane], t2[0]);
              }
            }
          }
        }
      }
    }
    if constexpr (!BNB && EPI == 3) {
      if (P.stats != nullptr) {
        warp_transpose_reduce32(run1, lane);
        warp_transpose_reduce32(run2, lane);
        if (lane < N) {
          atomicAdd(&s_stats[lane], run1[0]);
          atomicAdd(&s_stats[256 + lane], run2[0]);
        }
      }
    }
    if (P.stats != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
      for (int i = threadIdx.x - 64; i < N; i += 256) {
        atomicAdd(&P.stats[i], (double)s_stats[i]);
        atomicAdd(&P.stats2[i], (double)s_stats[256 + i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// ============================================================================
// host side
// ============================================================================
static bool g_tma_enabled = true;

static bool tma_plan(TmaConvParams& P, const GatherGeom& g) {
  if (!g_tma_enabled) return false;
  if (!(g.Cin == 32 || g.Cin % 64 == 0) || g.Cin > 2048) return false;
  if (g.Nout % 16 || g.Nout < 16 || g.Nout > 256) return false;
  if (!tile_rows(g, &P.rows_w, &P.rows_h, &P.rows_n)) return false;
  if (P.rows_w * g.in_mul > 256 || P.rows_h * g.in_mul > 256 || P.rows_n > 256) return false;
  P.g = g;
  P.N = g.Nout;
  P.kc = g.Cin == 32 ? 32 : 64;
  P.cin_chunks = g.Cin / P.kc;
  P.sub_a_bytes = kTTileM * P.kc * 2;
  P.sub_b_elems = g.Nout * P.kc;
  // sub-tiles per stage: as many as keep >= 3 stages in flight (each producer / issuer iteration costs
  // several hundred cycles of single-thread issue latency, so stages should carry 32 KB or more)
  const int sub_bytes = P.sub_a_bytes + P.sub_b_elems * 2;
  // Stage size: the pipeline's period is set by the two single-thread control loops (producer: wait + expect-tx +
  // copies; issuer: wait + MMAs + commit -- ~800 cycles per stage whatever it carries: skipping the A loads or the
  // weight copies in a timing-only build changed the run time by < 8 %, `profiles/r03f_skip_copies.log`), so a stage should carry as much as leaves
  // three stages in flight (measured in the step, B = 512: 64->128 fprop 87 -> 78 us for two launches, 128->256
  // transposed 99 -> 90 us; one sub-tile per stage everywhere: +25 .. +35 %; two stages of four sub-tiles: worse on the
  // N = 64 and N = 256 layers).
  const int budget = 218 * 1024;   // 227 KB - alignment slack - barriers / statistics / bias / BN vectors
  P.sps = P.kc == 32 ? 4 : 2;
  {
    static int env_sps = -1;
    if (env_sps < 0) { const char* e = getenv("VAE_TMA_SPS"); env_sps = e ? atoi(e) : 0; }
    if (env_sps > 0) P.sps = env_sps;
  }
  while (P.sps > 1 && budget / (P.sps * sub_bytes) < 3) P.sps /= 2;
  P.a_stage_bytes = P.sps * P.sub_a_bytes;
  P.b_stage_bytes = (P.sps * P.sub_b_elems * 2 + 1023) / 1024 * 1024;
  P.stages = budget / (P.a_stage_bytes + P.b_stage_bytes);
  if (P.kc == 64 && P.stages > 3) P.stages = 3;   // a fourth stage of 48 KB measured 3 - 5 % slower in the step (N = 64, N = 256)
  if (P.stages > kTStagesMax) P.stages = kTStagesMax;
  if (P.stages < 2) return false;
  int cols = 32;
  while (cols < 2 * g.Nout) cols *= 2;
  P.tmem_cols = cols;
  long long M = (long long)g.B * g.Hj * g.Wj;
  if (M >= (1LL << 31)) return false;
  P.tiles_per_phase = (int)((M + kTTileM - 1) / kTTileM);
  P.total_tiles = P.tiles_per_phase * g.nphase;
  int off = 0;
  for (int p = 0; p < 4; ++p) {
    P.b_off[p] = off;
    if (p < g.nphase) {
      P.nkb[p] = (g.phase[p].ntaps * P.cin_chunks + P.sps - 1) / P.sps;
      off += P.nkb[p] * (P.b_stage_bytes / 2);
    } else {
      P.nkb[p] = 0;
    }
  }
  return true;
}

// A forward conv with more than 256 output channels runs as Co / 256 launches, each producing a 256-channel
// slice of the output (the UMMA N extent and the TMEM double buffer end at 256 columns); the packed weight
// images of the slices follow one another.  OIHW weights of a slice are contiguous (co is outermost).
static int tma_nsplit(int Co, int mode) { return (mode == 0 && Co > 256 && Co % 256 == 0) ? Co / 256 : 1; }

static long long tma_packed_elems(const TmaConvParams& P) {
  long long n = 0;
  for (int p = 0; p < P.g.nphase; ++p) n += (long long)P.nkb[p] * (P.b_stage_bytes / 2);
  return n;
}

}  // namespace vae

using namespace vae;

extern "C" int vae_set_umma_tma(int enable) {
  int old = g_tma_enabled ? 1 : 0;
  g_tma_enabled = enable != 0;
  return old;
}

/* 1 when vae_conv_umma_tma serves this layer: channel counts fit the swizzled sub-tiles and
 * the spatial extents decompose into whole output rows per 128-pixel tile. */
extern "C" int vae_umma_tma_eligible(int B, int Hi, int Wi, int Ci, int Co, int stride, int mode) {
  if (B <= 0 || Hi <= 0 || Wi <= 0 || (stride != 1 && stride != 2) || Hi % stride || Wi % stride) return 0;
  const int co = Co / tma_nsplit(Co, mode);
  GatherGeom g = mode == 0 ? geom_fprop(B, Hi, Wi, Ci, co, stride) : geom_dgrad(B, Hi, Wi, Ci, co, stride);
  TmaConvParams P{};
  return tma_plan(P, g) ? 1 : 0;
}

extern "C" long long vae_umma_tma_packed_elems(int Ci, int Co, int stride, int mode) {
  if (tmat_channels_ok(Ci, Co, stride, mode)) return tmat_packed_elems(Ci, Co);
  const int ns = tma_nsplit(Co, mode), co = Co / ns;
  GatherGeom g = mode == 0 ? geom_fprop(1, 128 * stride, 128 * stride, Ci, co, stride)
                           : geom_dgrad(1, 128 * stride, 128 * stride, Ci, co, stride);
  TmaConvParams P{};
  if (!tma_plan(P, g)) return 0;
  return ns * tma_packed_elems(P);
}

extern "C" int vae_umma_tma_pack_weight(const float* w_oihw, void* out, int Ci, int Co, int stride, int mode,
                                        vae_stream_t stream) {
  VAE_REQUIRE(w_oihw && out && (mode == 0 || mode == 1) && (stride == 1 || stride == 2),
              "umma_tma_pack_weight: bad argument");
  if (tmat_channels_ok(Ci, Co, stride, mode)) return tmat_pack(w_oihw, out, Ci, Co, as_stream(stream));
  const int ns = tma_nsplit(Co, mode), co = Co / ns;
  GatherGeom g = mode == 0 ? geom_fprop(1, 128 * stride, 128 * stride, Ci, co, stride)
                           : geom_dgrad(1, 128 * stride, 128 * stride, Ci, co, stride);
  TmaConvParams P{};
  VAE_REQUIRE(tma_plan(P, g), "umma_tma_pack_weight: shape not served by the TMA path");
  long long n_max = 0;  // one element per thread: the OIHW reads are scattered, parallelism hides them
  for (int ph = 0; ph < g.nphase; ++ph) {
    const long long n = (long long)P.nkb[ph] * P.sps * P.sub_b_elems;
    if (n > n_max) n_max = n;
  }
  dim3 grid((unsigned)((n_max + 255) / 256), g.nphase);
  const long long per = tma_packed_elems(P);
  for (int s = 0; s < ns; ++s) {
    pack_tma_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(w_oihw + (long long)s * co * Ci * 9, (bf16*)out + s * per, P,
                                                                 co, Ci, mode);
    int rc = check_launch("pack_tma_weight");
    if (rc != 0) return rc;
  }
  return 0;
}

/* Conv forward (mode 0) / data gradient == ConvTranspose forward (mode 1) with the A
 * operand staged by tiled TMA loads.  Same contract as vae_conv_umma (no prologue);
 * wpk from vae_umma_tma_pack_weight. */
static int tma_dispatch(const void* in, const void* wpk, const float* bias, void* out, double* stats, int B, int Hi, int Wi,
                        int Ci, int Co, int stride, int mode, vae_stream_t stream, const void* bn_y, const float* const* bn_vec,
                        float slope, const float* const* act = nullptr) {
  VAE_REQUIRE(in && wpk && out && B > 0 && (stride == 1 || stride == 2) && Hi % stride == 0 && Wi % stride == 0 &&
                  (mode == 0 || mode == 1),
              "conv_umma_tma: bad argument");
  if (tmat_channels_ok(Ci, Co, stride, mode)) {   // transposed conv, 64 / 128 output channels: the four phases fused
    const int rc = tmat_launch(in, wpk, bias, out, stats, B, Hi, Wi, Ci, Co, as_stream(stream), bn_y, bn_vec, slope, act);
    VAE_REQUIRE(rc != 1, "conv_umma_tma: shape not served by the fused transposed-conv kernel");
    return rc;
  }
  const int ns = tma_nsplit(Co, mode), co = Co / ns;
  GatherGeom g = mode == 0 ? geom_fprop(B, Hi, Wi, Ci, co, stride) : geom_dgrad(B, Hi, Wi, Ci, co, stride);
  TmaConvParams P{};
  VAE_REQUIRE(tma_plan(P, g), "conv_umma_tma: shape not served by the TMA path");
  const int n_total = g.Nout * ns;   // channels per pixel of the produced tensor
  EncodeTiledFn encode = get_encode_fn();
  VAE_REQUIRE(encode != nullptr, "conv_umma_tma: cuTensorMapEncodeTiled is not available in this driver");
  // NHWC tensor, innermost dimension first
  cuuint64_t gdim[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.Win, (cuuint64_t)g.Hin, (cuuint64_t)g.B};
  cuuint64_t gstr[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.Win * g.Cin * 2, (cuuint64_t)g.Hin * g.Win * g.Cin * 2};
  // with a traversal stride s the unit loads ceil(box / s) elements: box = count * s
  cuuint32_t box[4] = {(cuuint32_t)P.kc, (cuuint32_t)(P.rows_w * g.in_mul), (cuuint32_t)(P.rows_h * g.in_mul),
                       (cuuint32_t)P.rows_n};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.in_mul, (cuuint32_t)g.in_mul, 1};
  CUresult cr = encode(&P.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE,
                       P.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VAE_REQUIRE(cr == CUDA_SUCCESS, "conv_umma_tma: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  cudaStream_t st = as_stream(stream);
  if (stats) VAE_CUDA(zero_scratch(stats, sizeof(double) * 2 * n_total, st));
  P.ldo = n_total;
  P.trace = umma_trace_ptr();
  size_t smem = 1024 + (size_t)P.stages * (P.a_stage_bytes + P.b_stage_bytes) + 256 + 16 + 768 * 4 + 64 +
                ((bn_y != nullptr || act != nullptr) ? 256 * 16 : 0);
  VAE_REQUIRE(!(act != nullptr && (stats != nullptr || bn_y != nullptr)), "conv_umma_tma: the inference epilogue takes no statistics");
  const int which = bn_y != nullptr ? 4 : act != nullptr ? 2 : (stats == nullptr && bias == nullptr) ? 0 : (stats != nullptr && g.Nout <= 32) ? 3 : 1;
  auto kern = which == 4 ? umma_tma_conv_kernel<true, 1> : which == 3 ? umma_tma_conv_kernel<false, 3>
            : which == 2 ? umma_tma_conv_kernel<false, 2> : which == 1 ? umma_tma_conv_kernel<false, 1> : umma_tma_conv_kernel<false, 0>;
  static bool attr_set[5] = {false, false, false, false, false};
  if (!attr_set[which]) {
    VAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[which] = true;
  }
  P.bn_y = nullptr; P.slope = slope;
  int grid = sm_count();
  if (grid > P.total_tiles) grid = P.total_tiles;
  const long long per = tma_packed_elems(P);
  for (int s = 0; s < ns; ++s) {
    P.wpk = (const bf16*)wpk + s * per;
    P.bias = bias ? bias + s * g.Nout : nullptr;
    P.out = (bf16*)out + s * g.Nout;
    P.stats = stats ? stats + s * g.Nout : nullptr;
    P.stats2 = stats ? stats + n_total + s * g.Nout : nullptr;
    if (act != nullptr) {
      P.act_mean = act[0] + s * g.Nout; P.act_scale = act[1] + s * g.Nout; P.act_beta = act[2] + s * g.Nout;
    }
    if (bn_y != nullptr) {
      P.bn_y = (const bf16*)bn_y + s * g.Nout;
      P.bn_mean = bn_vec[0] + s * g.Nout; P.bn_scale = bn_vec[1] + s * g.Nout;
      P.bn_beta = bn_vec[2] + s * g.Nout; P.bn_rstd = bn_vec[3] + s * g.Nout;
    }
    kern<<<grid, kTThreads, smem, st>>>(P);
    int rc = check_launch("umma_tma_conv");
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int vae_conv_umma_tma(const void* in, const void* wpk, const float* bias, void* out, double* stats, int B,
                                 int Hi, int Wi, int Ci, int Co, int stride, int mode, vae_stream_t stream) {
  return tma_dispatch(in, wpk, bias, out, stats, B, Hi, Wi, Ci, Co, stride, mode, stream, nullptr, nullptr, 0.0f);
}

/* vae_conv_umma_tma with the first half of the BatchNorm backward of the layer below fused into the epilogue;
 * same contract as vae_conv_umma_rw_bnbwd (out = du, dstats = {sum du, sum du * xhat}).
 * replaces: the LeakyReLU / native_batch_norm_backward reductions of train.py:650 for the layer below */
extern "C" int vae_conv_umma_tma_bnbwd(const void* in, const void* wpk, void* out, const void* y, const float* mean,
                                       const float* scale, const float* beta, const float* rstd, double* dstats,
                                       float slope, int B, int Hi, int Wi, int Ci, int Co, int stride, int mode,
                                       vae_stream_t stream) {
  VAE_REQUIRE(y && mean && scale && beta && rstd && dstats, "conv_umma_tma_bnbwd: bad argument");
  const float* vec[4] = {mean, scale, beta, rstd};
  return tma_dispatch(in, wpk, nullptr, out, dstats, B, Hi, Wi, Ci, Co, stride, mode, stream, y, vec, slope);
}

/* vae_conv_umma_tma for INFERENCE: eval-mode BatchNorm + LeakyReLU in the epilogue (same contract as
 * vae_conv_umma_rw_act).
 * replaces: aten::convolution + aten::native_batch_norm (eval) + aten::leaky_relu of one block under model.eval() */
extern "C" int vae_conv_umma_tma_act(const void* in, const void* wpk, const float* bias, void* out, const float* mean,
                                     const float* scale, const float* beta, float slope, int B, int Hi, int Wi, int Ci,
                                     int Co, int stride, int mode, vae_stream_t stream) {
  VAE_REQUIRE(mean && scale && beta, "conv_umma_tma_act: bad argument");
  const float* act[3] = {mean, scale, beta};
  return tma_dispatch(in, wpk, bias, out, nullptr, B, Hi, Wi, Ci, Co, stride, mode, stream, nullptr, nullptr, slope, act);
}
