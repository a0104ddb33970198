// Resident-weight TMA / tcgen05 convolution for the NARROW, HBM-bound layers (32 / 64 channels
// at 64x64 .. 128x128: 80 % of the activation bytes of the network).
//
// Same GEMM and the same strided-box TMA gather as conv_umma_tma.cu, re-cut for layers whose
// K loop is so short (K = 32 .. 288) that per-stage issue latency, not bandwidth, was the limit:
//
//  * all nine 3x3 weight images (<= 36 KB) are loaded ONCE per CTA and stay in shared memory;
//    pipeline stages carry activations only, 24-32 KB each (3-4 filter taps per stage);
//  * producer and MMA issuer are straight-line template code: no divisions, no table look-ups,
//    descriptors are constant + immediate;
//  * MODE 1 (ConvTranspose2d k3 s2 p1 op1 == data gradient of the stride-2 conv) computes all
//    FOUR output-parity phases of a tile from the same four gathered taps (offsets {0,1}^2 of the
//    low-resolution input): 4 TMA boxes instead of 9, four TMEM accumulators, and every thread
//    stores 2 x 128 contiguous bytes (the 2x2 output pixels of its input pixel) instead of four
//    64-byte pieces at 128-byte stride written by four different tiles.
//
//   warp 0     producer (one thread): resident weights, then per stage wait(empty) -> expect_tx -> TMA boxes
//   warp 1     MMA issuer (one thread), TMEM accumulators double-buffered across tiles
//   warps 2-9  epilogue: two sets of four warps (even / odd tiles): tcgen05.ld, bias, BatchNorm
//              statistics, bf16 NHWC stores
#include "tma_util.cuh"

namespace vae {

constexpr int kRwThreads = 320;
constexpr int kRwMaxStages = 8;

struct RwParams {
  CUtensorMap tmap;
  int B, Hj, Wj;          // iteration grid (output pixels for MODE 0, input pixels for MODE 1)
  int in_mul;             // MODE 0: conv stride; MODE 1: 1
  int Hout, Wout;         // produced tensor
  int total_tiles, stages;
  const bf16* wpk;        // 9 swizzled [N][KC] images, tap-major
  const float* bias;
  bf16* out;
  double* stats;
  // Fused BatchNorm-backward prologue of the NEXT hop (optional): the produced tensor is the gradient
  // w.r.t. act = lrelu(bn(y)) of the layer below.  With bn_y set the epilogue reads y at the pixel it
  // writes, multiplies by LeakyReLU'(u) and accumulates  stats[c] = sum du,  stats[N + c] = sum du * xhat
  // (the two reductions BatchNorm backward needs) instead of the forward statistics: the tensor written
  // is du, and the separate reduce pass over (d_act, y) disappears.
  const bf16* bn_y;
  const float *bn_mean, *bn_scale, *bn_beta, *bn_rstd;   // [N] each (scale = gamma * rstd)
  float slope;
  // Inference epilogue (optional, no statistics, no backward): eval-mode BatchNorm + LeakyReLU applied to the
  // accumulator row, so the layer writes its ACTIVATION and the separate normalisation pass disappears:
  //   out = lrelu((acc + bias - act_mean) * act_scale + act_beta)        (the expression of the apply kernel)
  const float *act_mean, *act_scale, *act_beta;
};

template <int MODE, int KC, int N>
struct RwCfg {
  static constexpr int SUB = 128 * KC * 2;                        // one tap's 128-pixel sub-tile
  static constexpr int T = MODE == 0 ? 9 : 4;                     // sub-tiles per tile
  static constexpr int S = MODE == 0 ? 3 : (KC == 32 ? 4 : 2);    // sub-tiles per stage
  static constexpr int SPT = T / S;                               // stages per tile
  static constexpr int STAGE = S * SUB;
  static constexpr int IMG = N * KC * 2;
  static constexpr int WBYTES = 9 * IMG;
  static constexpr int ACC_COLS = MODE == 0 ? N : 4 * N;
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
  static constexpr uint32_t LAYOUT = KC == 32 ? 4u : 2u;          // SWIZZLE_64B : SWIZZLE_128B
  static constexpr uint32_t SBO = KC == 32 ? 512u : 1024u;
  static constexpr int KS = KC / 16;
  // MODE 1: per epilogue warp a [32 pixels][128 B + 16 B pad] staging tile that turns the per-thread
  // 128-byte rows of the accumulator into fully coalesced 512-byte global stores
  static constexpr int ROWB = MODE == 1 ? 128 : N * 2;   // contiguous global bytes a thread produces per output row
  static constexpr int EPI_PITCH = ROWB + 16;
  static constexpr int EPI = 8 * 32 * EPI_PITCH;
  static constexpr int TAIL = 8 * 24 + 16 + 4 * (2 * 64 + 64) + 64 * 16 + EPI;
  static size_t smem_bytes(int stages) { return 1024 + (size_t)stages * STAGE + WBYTES + TAIL; }
};

// weight images: tap t = kh*3+kw; image[n][c] (K-major rows of KC elements, swizzled like the A tiles)
//   MODE 0: B[n][c] = w[co = n][ci = c][t]       MODE 1: B[n][c] = w[co = c][ci = n][t]
__global__ void pack_rw_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Co, int Ci, int mode,
                                      int N, int KC) {
  const int total = 9 * N * KC;
  const int cpr = KC / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i % KC, n = (i / KC) % N, t = i / (KC * N);
    const int co = mode == 0 ? n : e, ci = mode == 0 ? e : n;
    const float v = w[((long long)co * Ci + ci) * 9 + t];
    const int chunk = e >> 3;
    const int sw = cpr == 4 ? (chunk ^ ((n >> 1) & 3)) : (chunk ^ (n & 7));
    out[(long long)t * N * KC + n * KC + sw * 8 + (e & 7)] = __float2bfloat16_rn(v);
  }
}

template <int MODE, int KC, int N, bool BNB>
__global__ void __launch_bounds__(kRwThreads, 1)
umma_rw_conv_kernel(const __grid_constant__ RwParams P) {
  using Cfg = RwCfg<MODE, KC, N>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int kStages = P.stages;
  uint8_t* wsm = smem + kStages * Cfg::STAGE;
  uint8_t* tail = wsm + Cfg::WBYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kRwMaxStages;
  uint64_t* tfull_bar = empty_bar + kRwMaxStages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full_bar + 24);  // barrier block padded to 192 B
  float* s_stats = reinterpret_cast<float*>(tmem_ptr + 4);  // [2][64]
  float* s_bias = s_stats + 128;                            // [64]
  float4* s_bn = reinterpret_cast<float4*>(s_bias + 64);     // [64] {mean, scale, beta, rstd} (fused BN backward)
  uint8_t* s_epi = reinterpret_cast<uint8_t*>(s_bn + 64);    // [8 epilogue warps][32 rows][EPI_PITCH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128; i += kRwThreads) s_stats[i] = 0.0f;
  for (int i = threadIdx.x; i < 64; i += kRwThreads) s_bias[i] = (P.bias != nullptr && i < N) ? P.bias[i] : 0.0f;
  if constexpr (BNB)
    for (int i = threadIdx.x; i < N; i += kRwThreads) s_bn[i] = make_float4(P.bn_mean[i], P.bn_scale[i], P.bn_beta[i], P.bn_rstd[i]);
  else if (P.act_scale != nullptr)
    for (int i = threadIdx.x; i < N; i += kRwThreads) s_bn[i] = make_float4(P.act_mean[i], P.act_scale[i], P.act_beta[i], 0.0f);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kRwMaxStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 128);
      }
      mbar_init(w_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, (uint32_t)Cfg::TMEM_COLS);
  }
  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t Wj = (uint32_t)P.Wj, Hj = (uint32_t)P.Hj;
  const long long M_total = (long long)P.B * P.Hj * P.Wj;

  if (warp == 0) {
    // ============================== producer ==============================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, (uint32_t)Cfg::WBYTES);
      bulk_g2s(wsm, P.wpk, (uint32_t)Cfg::WBYTES, w_bar);
      int stage = 0;
      uint32_t ph = 0;
      const int im = P.in_mul;
      // tile -> (image, row, column) of its first pixel: divisions only for the first tile, then the
      // per-CTA tile stride is added with carries (this lone thread pays ~150 cycles per division)
      int jw0, jh0, b0, dw0, dh0, db0;
      {
        const uint32_t m0 = (uint32_t)blockIdx.x * 128u;
        jw0 = (int)(m0 % Wj);
        const uint32_t t2 = m0 / Wj;
        jh0 = (int)(t2 % Hj);
        b0 = (int)(t2 / Hj);
        const uint32_t d0 = (uint32_t)gridDim.x * 128u;
        dw0 = (int)(d0 % Wj);
        const uint32_t d2 = d0 / Wj;
        dh0 = (int)(d2 % Hj);
        db0 = (int)(d2 / Hj);
      }
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const int cw = jw0 * im, ch = jh0 * im;
#pragma unroll
        for (int st = 0; st < Cfg::SPT; ++st) {
          mbar_wait(&empty_bar[stage], ph ^ 1u);
          uint8_t* a_st = smem + stage * Cfg::STAGE;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)Cfg::STAGE);
#pragma unroll
          for (int i = 0; i < Cfg::S; ++i) {
            const int o = st * Cfg::S + i;
            // MODE 0: tap (kh, kw) = (o / 3, o % 3) reads input pixel s*j + (kh-1, kw-1)
            // MODE 1: offset (dh, dw) = (o >> 1, o & 1) reads input pixel j + (dh, dw)
            const int dw = MODE == 0 ? (o % 3) - 1 : (o & 1);
            const int dh = MODE == 0 ? (o / 3) - 1 : (o >> 1);
            tma_load_4d(a_st + i * Cfg::SUB, &P.tmap, &full_bar[stage], 0, cw + dw, ch + dh, b0);
          }
          if (++stage == kStages) { stage = 0; ph ^= 1u; }
        }
        jw0 += dw0;
        if (jw0 >= (int)Wj) { jw0 -= (int)Wj; ++jh0; }
        jh0 += dh0;
        if (jh0 >= (int)Hj) { jh0 -= (int)Hj; ++b0; }
        b0 += db0;
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)N, 0, 0);
    const uint32_t a_base = smem_u32(smem), w_base = smem_u32(wsm);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
#pragma unroll
      for (int st = 0; st < Cfg::SPT; ++st) {
        mbar_wait(&full_bar[stage], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = a_base + (uint32_t)(stage * Cfg::STAGE);
#pragma unroll
          for (int i = 0; i < Cfg::S; ++i) {
            const int o = st * Cfg::S + i;
            if constexpr (MODE == 0) {
#pragma unroll
              for (int k = 0; k < Cfg::KS; ++k) {
                const uint64_t ad = make_desc_sw(a_addr + i * Cfg::SUB + k * 32, Cfg::SBO, Cfg::LAYOUT);
                const uint64_t bd = make_desc_sw(w_base + o * Cfg::IMG + k * 32, Cfg::SBO, Cfg::LAYOUT);
                mma_bf16_ss(d_tmem, ad, bd, idesc, (o > 0 || k > 0) ? 1u : 0u);
              }
            } else {
              const int dh = o >> 1, dw = o & 1;
#pragma unroll
              for (int php = dh; php < 2; ++php)
#pragma unroll
                for (int pwp = dw; pwp < 2; ++pwp) {
                  const int kh = php + 1 - 2 * dh, kw = pwp + 1 - 2 * dw;
#pragma unroll
                  for (int k = 0; k < Cfg::KS; ++k) {
                    const uint64_t ad = make_desc_sw(a_addr + i * Cfg::SUB + k * 32, Cfg::SBO, Cfg::LAYOUT);
                    const uint64_t bd = make_desc_sw(w_base + (kh * 3 + kw) * Cfg::IMG + k * 32, Cfg::SBO, Cfg::LAYOUT);
                    mma_bf16_ss(d_tmem + (uint32_t)((php * 2 + pwp) * N), ad, bd, idesc, (o > 0 || k > 0) ? 1u : 0u);
                  }
                }
            }
          }
          mma_commit(&empty_bar[stage]);
          if (st == Cfg::SPT - 1) mma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; ph ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int ew = warp & 3;           // TMEM lane group = warp % 4
    const int eset = (warp - 2) >> 2;  // warps 2-5: even tiles, 6-9: odd tiles
    const int r = ew * 32 + lane;
    const bool has_bias = P.bias != nullptr;
    const bool has_act = !BNB && P.act_scale != nullptr;
    const uint32_t s_bn_a = smem_u32(s_bn);
    float run1[BNB ? 1 : 32], run2[BNB ? 1 : 32];
#pragma unroll
    for (int i = 0; i < (BNB ? 1 : 32); ++i) run1[i] = run2[i] = 0.0f;
    for (int it = eset, tile = blockIdx.x + eset * gridDim.x; tile < P.total_tiles; tile += 2 * gridDim.x, it += 2) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const long long m = (long long)tile * 128 + r;
      const bool valid = m < M_total;
      bf16* dst = nullptr;
      if (valid) {
        const uint32_t mu = (uint32_t)m;
        const uint32_t jw = mu % Wj, t2 = mu / Wj;
        const uint32_t jh = t2 % Hj, b = t2 / Hj;
        if constexpr (MODE == 0)
          dst = P.out + (((long long)b * P.Hout + jh) * P.Wout + jw) * N;
        else
          dst = P.out + (((long long)b * P.Hout + 2 * jh) * P.Wout + 2 * jw) * N;
      }
      constexpr int NCHUNK = Cfg::ACC_COLS / 32;
      if constexpr (BNB) {
        // fused BN backward: this thread's y values (same element offsets as the row it writes), requested
        // BEFORE the wait for the accumulator so that the loads fly while the MMAs of the tile run
        uint4 yraw[BNB ? NCHUNK : 1][4];
        if constexpr (BNB) {
          const long long eoff = dst != nullptr ? (long long)(dst - P.out) : 0;
  #pragma unroll
          for (int ci = 0; ci < NCHUNK; ++ci) {
            // MODE 0: chunk ci = channels 32ci..; MODE 1: chunk ci = output pixel (2jh + (ci>>1), 2jw + (ci&1)), N = 32
            const long long o = MODE == 0 ? eoff + ci * 32 : eoff + (long long)(ci >> 1) * P.Wout * N + (ci & 1) * N;
            const uint4* src = reinterpret_cast<const uint4*>(P.bn_y + o);
  #pragma unroll
            for (int q = 0; q < 4; ++q) yraw[ci][q] = ldg_nc_v4_pinned(src + q);
          }
        }
        mbar_wait(&tfull_bar[acc], acc_ph);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);
  #pragma unroll
        for (int ci = 0; ci < NCHUNK; ++ci) {
          float v[32];
          tmem_ld_32x32(t_row + (uint32_t)(ci * 32), v);
          tmem_ld_wait();
          if (ci == NCHUNK - 1) {  // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
          }
          // channel base of this chunk
          const int cbase = MODE == 0 ? ci * 32 : 0;
          if (has_bias) {   // uniform: the data-gradient launches of the backward pass have no bias
  #pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[cbase + i]);
              v[i] += b4.x;
              v[i + 1] += b4.y;
              v[i + 2] += b4.z;
              v[i + 3] += b4.w;
            }
          }
          if (has_act) {    // inference: eval-mode BatchNorm + LeakyReLU on the accumulator row
  #pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float4 c4 = lds_f4_addr(s_bn_a + (uint32_t)((cbase + i) * 16));   // mean, scale, beta
              const float u = fmaf(v[i] - c4.x, c4.y, c4.z);
              v[i] = u > 0.0f ? u : u * P.slope;
            }
          }
          if (!valid) {     // rows past the end of the tensor (last tile only): keep them out of the statistics
  #pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.0f;
          }
          float second[32];   // what the second statistic sums: y^2 (forward) or du * xhat (fused BN backward)
          if constexpr (BNB) {
  #pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t w4[4] = {yraw[ci][q].x, yraw[ci][q].y, yraw[ci][q].z, yraw[ci][q].w};
  #pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int i = q * 8 + k;
                const float yv = __uint_as_float((k & 1) ? (w4[k >> 1] & 0xFFFF0000u) : (w4[k >> 1] << 16));
                const float4 c4 = lds_f4_addr(s_bn_a + (uint32_t)((cbase + i) * 16));   // mean, scale, beta, rstd
                const float t = yv - c4.x;
                const float u = fmaf(t, c4.y, c4.z);          // same expression as the forward apply (bn_affine)
                const float du = u > 0.0f ? v[i] : v[i] * P.slope;
                v[i] = du;
                second[i] = du * (t * c4.w);
              }
            }
          } else {
  #pragma unroll
            for (int i = 0; i < 32; ++i) second[i] = v[i] * v[i];
          }
          {
            // A thread's accumulator row is ROWB contiguous bytes of the output (MODE 0: its pixel's N channels;
            // MODE 1: the two column-parity pixels of one output row) and the 32 threads of a warp own 32 such
            // rows back to back.  Stage them, then let consecutive lanes store consecutive 16-byte pieces.
            uint8_t* st = s_epi + (warp - 2) * (32 * Cfg::EPI_PITCH);
            const int cpos = MODE == 1 ? (ci & 1) * 64 : ci * 64;   // byte position of this chunk in the row
  #pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 pk;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
  #pragma unroll
              for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[i + 2 * k], v[i + 2 * k + 1]);
              *reinterpret_cast<uint4*>(st + lane * Cfg::EPI_PITCH + cpos + i * 2) = pk;
            }
            const bool row_done = MODE == 1 ? (ci & 1) != 0 : ci == NCHUNK - 1;
            if (row_done) {
              __syncwarp();
              const unsigned long long my = reinterpret_cast<unsigned long long>(dst);  // 0 when the row is out of range
              const long long roff = MODE == 1 ? (long long)(ci >> 1) * P.Wout * N : 0;
              constexpr int PPR = Cfg::ROWB / 16;   // 16-byte pieces per row
  #pragma unroll
              for (int j = 0; j < PPR; ++j) {
                const int c = lane + 32 * j;
                const int src = c / PPR, sub = c % PPR;
                const unsigned long long d = __shfl_sync(0xffffffffu, my, src);
                const uint4 pk = *reinterpret_cast<const uint4*>(st + src * Cfg::EPI_PITCH + sub * 16);
                if (d != 0ull) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(d) + roff + sub * 8) = pk;
              }
              __syncwarp();
            }
          }
          if (P.stats != nullptr) {
            if constexpr (N <= 32 && !BNB) {   // BNB: the y prefetch needs the registers of the running sums
  #pragma unroll
              for (int i = 0; i < 32; ++i) {
                run1[i] += v[i];
                run2[i] += second[i];
              }
            } else {
              warp_transpose_reduce32(v, lane);
              warp_transpose_reduce32(second, lane);
              atomicAdd(&s_stats[cbase + lane], v[0]);
              atomicAdd(&s_stats[64 + cbase + lane], second[0]);
            }
          }
        }
      } else {
        // Plain epilogue.  One pipeline over the 32-column chunks: bias requested before the TMEM wait, the next chunk's
        // TMEM load issued as soon as this one is packed, the staging slab addressed as shared memory (st.shared /
        // ld.shared, not generic accesses), the output rows of the lanes this lane stores for fetched once per tile.
        constexpr int PPR = Cfg::ROWB / 16;   // 16-byte pieces per output row
        const uint32_t st_a = smem_u32(s_epi) + (uint32_t)((warp - 2) * (32 * Cfg::EPI_PITCH));
        const uint32_t s_bias_a = smem_u32(s_bias);
        // this thread's row as a 16-byte index into the output (all ones: out of range); rows (lane + 32 j) / PPR are
        // the ones this lane stores
        const uint32_t my16 = valid ? (uint32_t)((reinterpret_cast<const uint8_t*>(dst) - reinterpret_cast<const uint8_t*>(P.out)) >> 4) : 0xFFFFFFFFu;
        uint32_t d16[PPR];
#pragma unroll
        for (int j = 0; j < PPR; ++j) d16[j] = __shfl_sync(0xffffffffu, my16, (lane + 32 * j) / PPR);
        mbar_wait(&tfull_bar[acc], acc_ph);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);
        float v[32];
        tmem_ld_32x32(t_row, v);
#pragma unroll
        for (int ci = 0; ci < NCHUNK; ++ci) {
          const int cbase = MODE == 0 ? ci * 32 : 0;   // channel base of this chunk
          // bias (uniform: the data-gradient launches of the backward pass have none): requested before the wait when
          // the registers allow it (N > 32: no kernel-long running sums), else right after it
          constexpr bool kEarlyBias = N > 32;
          float4 b4[8];
          if (kEarlyBias && has_bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) b4[i] = lds_f4_addr(s_bias_a + (uint32_t)((cbase + 4 * i) * 4));
          }
          tmem_ld_wait_regs(v);
          if (ci == NCHUNK - 1) {  // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
          }
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (!kEarlyBias) b4[i] = lds_f4_addr(s_bias_a + (uint32_t)((cbase + 4 * i) * 4));
              v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w;
            }
          }
          if (has_act) {    // inference: eval-mode BatchNorm + LeakyReLU on the accumulator row
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float4 c4 = lds_f4_addr(s_bn_a + (uint32_t)((cbase + i) * 16));   // mean, scale, beta
              const float u = fmaf(v[i] - c4.x, c4.y, c4.z);
              v[i] = u > 0.0f ? u : u * P.slope;
            }
          }
          if (!valid) {     // rows past the end of the tensor (last tile only): keep them out of the statistics
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.0f;
          }
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          float t1[N <= 32 ? 1 : 32], t2[N <= 32 ? 1 : 32];
          if (P.stats != nullptr) {
            if constexpr (N <= 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                run1[i] += v[i];
                run2[i] = fmaf(v[i], v[i], run2[i]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                t1[i] = v[i];
                t2[i] = v[i] * v[i];
              }
            }
          }
          if (ci + 1 < NCHUNK) tmem_ld_32x32(t_row + (uint32_t)((ci + 1) * 32), v);
          const int cpos = MODE == 1 ? (ci & 1) * 64 : ci * 64;   // byte position of this chunk in the row
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(st_a + (uint32_t)(lane * Cfg::EPI_PITCH + cpos + 16 * q)),
                         "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3]) : "memory");
          const bool row_done = MODE == 1 ? (ci & 1) != 0 : ci == NCHUNK - 1;
          if (row_done) {
            __syncwarp();
            const uint32_t roff16 = MODE == 1 ? (uint32_t)(ci >> 1) * (uint32_t)(P.Wout * N * 2 / 16) : 0u;
            constexpr int G = PPR > 4 ? 4 : PPR;   // pieces in flight at a time (register budget)
#pragma unroll
            for (int j0 = 0; j0 < PPR; j0 += G) {
              uint4 pk[G];
#pragma unroll
              for (int j = 0; j < G; ++j) {
                const int c = lane + 32 * (j0 + j);
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pk[j].x), "=r"(pk[j].y), "=r"(pk[j].z), "=r"(pk[j].w)
                             : "r"(st_a + (uint32_t)((c / PPR) * Cfg::EPI_PITCH + (c % PPR) * 16)));
              }
#pragma unroll
              for (int j = 0; j < G; ++j)
                if (d16[j0 + j] != 0xFFFFFFFFu)
                  stg_v4(reinterpret_cast<uint4*>(P.out) + ((size_t)d16[j0 + j] + roff16 + (uint32_t)((lane + 32 * (j0 + j)) % PPR)), pk[j]);
            }
            __syncwarp();
          }
          if constexpr (N > 32) {
            if (P.stats != nullptr) {
              warp_transpose_reduce32(t1, lane);
              warp_transpose_reduce32(t2, lane);
              atomicAdd(&s_stats[cbase + lane], t1[0]);
              atomicAdd(&s_stats[64 + cbase + lane], t2[0]);
            }
          }
        }
      }
    }
    if (P.stats != nullptr) {
      if constexpr (N <= 32 && !BNB) {   // BNB: the y prefetch needs the registers of the running sums
        warp_transpose_reduce32(run1, lane);
        warp_transpose_reduce32(run2, lane);
        if (lane < N) {
          atomicAdd(&s_stats[lane], run1[0]);
          atomicAdd(&s_stats[64 + lane], run2[0]);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
      for (int i = threadIdx.x - 64; i < N; i += 256) {
        atomicAdd(&P.stats[i], (double)s_stats[i]);
        atomicAdd(&P.stats[N + i], (double)s_stats[64 + i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)Cfg::TMEM_COLS);
  }
}

// ============================================================================
// host side
// ============================================================================
static bool g_rw_enabled = true;

struct RwPlan {
  int mode, kc, n;
  int rows_w, rows_h, rows_n;
  int Hj, Wj, in_mul, Hout, Wout, Hin, Win;
};

// mode 0: X[B,Hi,Wi,Ci] -> Y[B,Hi/s,Wi/s,Co]; mode 1: dY[B,Hi/2,Wi/2,Co] -> dX[B,Hi,Wi,Ci]
static bool rw_plan(RwPlan& p, int B, int Hi, int Wi, int Ci, int Co, int stride, int mode) {
  if (!g_rw_enabled || B <= 0 || Hi <= 0 || Wi <= 0 || Hi % stride || Wi % stride) return false;
  GatherGeom g{};
  if (mode == 0) {
    if (Ci != 32 || !(Co == 32 || Co == 64) || !(stride == 1 || stride == 2)) return false;
    p.kc = 32; p.n = Co;
    p.Hj = Hi / stride; p.Wj = Wi / stride; p.in_mul = stride;
    p.Hout = p.Hj; p.Wout = p.Wj; p.Hin = Hi; p.Win = Wi;
  } else {
    if (stride != 2 || Ci != 32 || !(Co == 32 || Co == 64)) return false;
    p.kc = Co; p.n = Ci;
    p.Hj = Hi / 2; p.Wj = Wi / 2; p.in_mul = 1;
    p.Hout = Hi; p.Wout = Wi; p.Hin = p.Hj; p.Win = p.Wj;
  }
  p.mode = mode;
  g.Hj = p.Hj; g.Wj = p.Wj;
  if (!tile_rows(g, &p.rows_w, &p.rows_h, &p.rows_n)) return false;
  if (p.rows_w * p.in_mul > 256 || p.rows_h * p.in_mul > 256 || p.rows_n > 256) return false;
  if ((long long)B * p.Hj * p.Wj >= (1LL << 31)) return false;
  return get_encode_fn() != nullptr;
}

template <int MODE, int KC, int N>
static int rw_launch(const RwPlan& pl, const void* in, const void* wpk, const float* bias, void* out, double* stats,
                     int B, cudaStream_t st, const void* bn_y = nullptr, const float* const* bn_vec = nullptr, float slope = 0.0f,
                     const float* const* act = nullptr) {
  using Cfg = RwCfg<MODE, KC, N>;
  RwParams P{};
  cuuint64_t gdim[4] = {(cuuint64_t)KC, (cuuint64_t)pl.Win, (cuuint64_t)pl.Hin, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)KC * 2, (cuuint64_t)pl.Win * KC * 2, (cuuint64_t)pl.Hin * pl.Win * KC * 2};
  cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(pl.rows_w * pl.in_mul), (cuuint32_t)(pl.rows_h * pl.in_mul),
                       (cuuint32_t)pl.rows_n};
  cuuint32_t estr[4] = {1, (cuuint32_t)pl.in_mul, (cuuint32_t)pl.in_mul, 1};
  CUresult cr = get_encode_fn()(&P.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VAE_REQUIRE(cr == CUDA_SUCCESS, "conv_umma_rw: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  P.B = B; P.Hj = pl.Hj; P.Wj = pl.Wj; P.in_mul = pl.in_mul; P.Hout = pl.Hout; P.Wout = pl.Wout;
  const long long M = (long long)B * pl.Hj * pl.Wj;
  P.total_tiles = (int)((M + 127) / 128);
  int stages = (int)((226 * 1024 - 1024 - Cfg::WBYTES - Cfg::TAIL) / Cfg::STAGE);
  if (stages > kRwMaxStages) stages = kRwMaxStages;
  VAE_REQUIRE(stages >= 2, "conv_umma_rw: shared memory budget");
  P.stages = stages;
  P.wpk = (const bf16*)wpk; P.bias = bias; P.out = (bf16*)out; P.stats = stats;
  P.bn_y = (const bf16*)bn_y; P.slope = slope;
  if (bn_y != nullptr) { P.bn_mean = bn_vec[0]; P.bn_scale = bn_vec[1]; P.bn_beta = bn_vec[2]; P.bn_rstd = bn_vec[3]; }
  if (act != nullptr) { P.act_mean = act[0]; P.act_scale = act[1]; P.act_beta = act[2]; }
  auto kern = bn_y != nullptr ? umma_rw_conv_kernel<MODE, KC, N, true> : umma_rw_conv_kernel<MODE, KC, N, false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[bn_y != nullptr]) {
    VAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[bn_y != nullptr] = true;
  }
  int grid = sm_count();
  if (grid > P.total_tiles) grid = P.total_tiles;
  kern<<<grid, kRwThreads, Cfg::smem_bytes(stages), st>>>(P);
  return check_launch("umma_rw_conv");
}

}  // namespace vae

using namespace vae;

extern "C" int vae_set_umma_rw(int enable) {
  int old = g_rw_enabled ? 1 : 0;
  g_rw_enabled = enable != 0;
  return old;
}

/* 1 when vae_conv_umma_rw serves this layer (narrow layers: 32-channel conv inputs for
 * mode 0; 32-channel ConvTranspose outputs fed by 32 / 64 channels for mode 1). */
extern "C" int vae_umma_rw_eligible(int B, int Hi, int Wi, int Ci, int Co, int stride, int mode) {
  RwPlan p{};
  return rw_plan(p, B, Hi, Wi, Ci, Co, stride, mode) ? 1 : 0;
}

extern "C" long long vae_umma_rw_packed_elems(int Ci, int Co) { return 9LL * Ci * Co; }

extern "C" int vae_umma_rw_pack_weight(const float* w_oihw, void* out, int Ci, int Co, int mode, vae_stream_t stream) {
  VAE_REQUIRE(w_oihw && out && (mode == 0 || mode == 1), "umma_rw_pack_weight: bad argument");
  const int N = mode == 0 ? Co : Ci, KC = mode == 0 ? Ci : Co;
  VAE_REQUIRE((KC == 32 || KC == 64) && N % 16 == 0 && N <= 64, "umma_rw_pack_weight: shape not served");
  pack_rw_weight_kernel<<<(9 * N * KC + 255) / 256, 256, 0, as_stream(stream)>>>(w_oihw, (bf16*)out, Co, Ci, mode, N, KC);
  return check_launch("pack_rw_weight");
}

/* Conv forward (mode 0) / fused-phase ConvTranspose forward == data gradient (mode 1) with the
 * weights resident in shared memory; same contract as vae_conv_umma_tma, wpk from
 * vae_umma_rw_pack_weight.
 * replaces: aten::convolution (models.py:45, 66-68, 77) and the data-gradient half of
 * aten::convolution_backward (train.py:650) on the narrow layers */
static int rw_dispatch(const void* in, const void* wpk, const float* bias, void* out, double* stats, int B, int Hi, int Wi,
                       int Ci, int Co, int stride, int mode, vae_stream_t stream, const void* bn_y, const float* const* bn_vec,
                       float slope, const float* const* act = nullptr) {
  VAE_REQUIRE(in && wpk && out && (mode == 0 || mode == 1), "conv_umma_rw: bad argument");
  RwPlan pl{};
  VAE_REQUIRE(rw_plan(pl, B, Hi, Wi, Ci, Co, stride, mode), "conv_umma_rw: shape not served by this kernel");
  cudaStream_t st = as_stream(stream);
  if (stats) VAE_CUDA(zero_scratch(stats, sizeof(double) * 2 * pl.n, st));
  if (mode == 0 && pl.n == 32) return rw_launch<0, 32, 32>(pl, in, wpk, bias, out, stats, B, st, bn_y, bn_vec, slope, act);
  if (mode == 0 && pl.n == 64) return rw_launch<0, 32, 64>(pl, in, wpk, bias, out, stats, B, st, bn_y, bn_vec, slope, act);
  if (mode == 1 && pl.kc == 32) return rw_launch<1, 32, 32>(pl, in, wpk, bias, out, stats, B, st, bn_y, bn_vec, slope, act);
  if (mode == 1 && pl.kc == 64) return rw_launch<1, 64, 32>(pl, in, wpk, bias, out, stats, B, st, bn_y, bn_vec, slope, act);
  set_error("conv_umma_rw: no instantiation");
  return VAE_ERR_UNSUPPORTED;
}

extern "C" int vae_conv_umma_rw(const void* in, const void* wpk, const float* bias, void* out, double* stats, int B,
                                int Hi, int Wi, int Ci, int Co, int stride, int mode, vae_stream_t stream) {
  return rw_dispatch(in, wpk, bias, out, stats, B, Hi, Wi, Ci, Co, stride, mode, stream, nullptr, nullptr, 0.0f);
}

/* vae_conv_umma_rw producing the gradient w.r.t. an activation act = lrelu(bn(y)) (models.py:45-47,66-70) with the
 * first half of that BatchNorm's backward fused into the epilogue: `y` is the stored pre-normalisation tensor of
 * the layer below (same shape as `out`), mean / scale (= gamma*rstd) / beta / rstd = the layer's [channels] fp32 vectors.  `out`
 * receives du = d_act * LeakyReLU'(u) and dstats[0..C) = sum du, dstats[C..2C) = sum du * xhat (fp64), i.e. what
 * vae_bn_lrelu_bwd_reduce computes in a separate pass over (d_act, y).
 * replaces: the LeakyReLU / native_batch_norm_backward reductions of train.py:650 for the layer below */
extern "C" int vae_conv_umma_rw_bnbwd(const void* in, const void* wpk, void* out, const void* y, const float* mean,
                                      const float* scale, const float* beta, const float* rstd, double* dstats,
                                      float slope, int B, int Hi, int Wi, int Ci, int Co, int stride, int mode,
                                      vae_stream_t stream) {
  VAE_REQUIRE(y && mean && scale && beta && rstd && dstats, "conv_umma_rw_bnbwd: bad argument");
  const float* vec[4] = {mean, scale, beta, rstd};
  return rw_dispatch(in, wpk, nullptr, out, dstats, B, Hi, Wi, Ci, Co, stride, mode, stream, y, vec, slope);
}

/* vae_conv_umma_rw for INFERENCE: eval-mode BatchNorm (running statistics: mean, scale = gamma / sqrt(var + eps), beta)
 * + LeakyReLU applied in the epilogue, so `out` is the block's activation and no normalisation pass follows.
 * replaces: aten::convolution + aten::native_batch_norm (eval) + aten::leaky_relu of one block (models.py:45-47,66-70)
 * under model.eval() / torch.no_grad() (models.py:250-263, evaluation.py:42-65) */
extern "C" int vae_conv_umma_rw_act(const void* in, const void* wpk, const float* bias, void* out, const float* mean,
                                    const float* scale, const float* beta, float slope, int B, int Hi, int Wi, int Ci,
                                    int Co, int stride, int mode, vae_stream_t stream) {
  VAE_REQUIRE(mean && scale && beta, "conv_umma_rw_act: bad argument");
  const float* act[3] = {mean, scale, beta};
  return rw_dispatch(in, wpk, bias, out, nullptr, B, Hi, Wi, Ci, Co, stride, mode, stream, nullptr, nullptr, slope, act);
}
