// Streamed-weight tcgen05 kernel for the TRANSPOSED 3x3 stride-2 convolution (ConvTranspose2d(k3,s2,p1,op1) forward ==
// data gradient of Conv2d(k3,s2,p1); models.py:66-68, train.py:650) with the four output-parity phases FUSED,
// for layers too wide for resident weights (conv_umma_rw.cu): 64 / 128 output channels, Cin a multiple of 64.
//
// conv_umma_tma.cu runs the four phases as separate tiles with 1 / 2 / 2 / 4 taps each: NINE gathered input boxes per
// 128 input pixels, i.e. nine (x Cin / 64) pipeline stages -- and that kernel's period is ~800 cycles per STAGE
// whatever the stage carries (its two single-thread control loops are the limit).  Here a tile is 128
// INPUT pixels: the FOUR boxes at offsets {0,1}^2 feed all nine (phase, tap) pairs, accumulating into four TMEM
// accumulators (one per phase): 4 (x Cin / 64) stages per tile instead of 9, each carrying 4 / 2 / 2 / 1 weight images
// and 16 / 8 / 8 / 4 MMAs.
//
//   out[b][2jh+ph][2jw+pw][n] = bias[n] + sum_{dh <= ph, dw <= pw} sum_c in[b][jh+dh][jw+dw][c] * w[c][n][(ph+1-2dh)*3 + (pw+1-2dw)]
//
//   warp 0     producer (one thread): per stage wait(empty) -> expect_tx -> one tensor TMA box + one bulk weight copy
//   warp 1     MMA issuer: stage (chunk c, box (dh,dw)) -> for every phase (ph >= dh, pw >= dw): 4 x tcgen05.mma (K = 64)
//   warps 2-9  epilogue: two sets of four warps.  8N <= 512 TMEM columns (N = 64): accumulator sets double-buffered,
//              a set takes every other tile; else (N = 128) one accumulator set, the two epilogue sets split its output rows.
//              (Both sets on every tile, splitting the channel chunks, measured slower: 463 vs 415 us at the wide shape.)
//              Output rows leave through a swizzled shared-memory slab: one store instruction = 8 rows x 64 contiguous B.
#include <cuda.h>

#include "common.cuh"
#include "conv_geom.cuh"
#include "umma.cuh"
#include "tma_util.cuh"

namespace vae {

using namespace umma;

constexpr int kMtThreads = 320;
constexpr int kMtStagesMax = 8;

struct TmatParams {
  CUtensorMap tmap;
  int B, Hj, Wj, Cin, N;
  int Hout, Wout, ldo;
  int chunks;            // Cin / 64
  int total_tiles;
  int img_bytes;         // N * 64 * 2: one (phase, tap) weight image of one 64-channel chunk
  int slot_bytes, stages;
  int double_buf;
  int split;             // two accumulator sets: both epilogue sets on every tile, one 32-channel chunk each
  const bf16* wpk;       // [chunk][box][phase images][N][64] swizzled (pack_tmat_weight_kernel)
  const float* bias;
  bf16* out;
  double* stats;         // [N] sums, stats2 [N] sums of squares (or sum du / sum du * xhat with bn_y)
  double* stats2;
  const bf16* bn_y;      // fused BatchNorm-backward prologue of the next hop (see conv_umma_rw.cu), or null
  const float *bn_mean, *bn_scale, *bn_beta, *bn_rstd;
  float slope;
  const float *act_mean, *act_scale, *act_beta;   // inference epilogue: eval-mode BatchNorm + LeakyReLU (see conv_umma_rw.cu)
};

// images of box (dh, dw): phases (ph, pw) with ph >= dh, pw >= dw, ph-major
__host__ __device__ inline int tmat_nimg(int box) { return (2 - (box >> 1)) * (2 - (box & 1)); }

// packed weights: for chunk c, box b, image i: [n][e] = w[co = 64c + e][ci = n][tap], rows swizzled 128B (chunk ^= n & 7)
__global__ void pack_tmat_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Co, int Ci) {
  const int N = Ci, chunks = Co / 64;
  const long long total = 9LL * N * Co;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % 64);
    const int n = (int)((i / 64) % N);
    const int img = (int)((i / (64LL * N)) % 9);     // image index within a chunk: boxes in order, phases ph-major
    const int c = (int)(i / (64LL * N * 9));
    int box = 0, k = img;
    while (k >= tmat_nimg(box)) { k -= tmat_nimg(box); ++box; }
    const int dh = box >> 1, dw = box & 1;
    const int npw = 2 - dw;
    const int ph = dh + k / npw, pw = dw + k % npw;
    const int tap = (ph + 1 - 2 * dh) * 3 + (pw + 1 - 2 * dw);
    const float v = w[((long long)(c * 64 + e) * Ci + n) * 9 + tap];
    const int chunk16 = e >> 3;
    out[((long long)(c * 9 + img) * N + n) * 64 + ((chunk16 ^ (n & 7)) << 3) + (e & 7)] = __float2bfloat16_rn(v);
  }
  (void)chunks;
}

// EPI (plain epilogue only): 0 = bias, 1 = bias + BatchNorm statistics, 2 = bias + eval-mode BatchNorm + LeakyReLU (inference).
// Compile-time so that the piece loop carries no parameter loads or uniform branches.
template <bool BNB, int EPI>
__global__ void __launch_bounds__(kMtThreads, 1)
umma_tmat_conv_kernel(const __grid_constant__ TmatParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* tail = smem + P.stages * P.slot_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMtStagesMax;
  uint64_t* tfull_bar = empty_bar + kMtStagesMax;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_stats = reinterpret_cast<float*>(tmem_ptr + 4);   // [2][128]
  float* s_bias = s_stats + 256;                             // [128]
  float4* s_bn = reinterpret_cast<float4*>(s_bias + 128);    // [128] {mean, scale, beta, rstd} (BNB only)
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bn + 128); // [8 epilogue warps][32 rows][64 B], 16-byte pieces XOR-swizzled

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = P.N, kStages = P.stages, chunks = P.chunks;
  const bool dbl = P.double_buf != 0;

  for (int i = threadIdx.x; i < 256; i += kMtThreads) s_stats[i] = 0.0f;
  for (int i = threadIdx.x; i < 128; i += kMtThreads) s_bias[i] = (P.bias != nullptr && i < N) ? P.bias[i] : 0.0f;
  if constexpr (BNB)
    for (int i = threadIdx.x; i < N; i += kMtThreads) s_bn[i] = make_float4(P.bn_mean[i], P.bn_scale[i], P.bn_beta[i], P.bn_rstd[i]);
  else if (EPI == 2)   // inference: u = y * scale + shift, two channels per float4
    for (int i = threadIdx.x; i < N; i += kMtThreads)
      reinterpret_cast<float2*>(s_bn)[i] = make_float2(P.act_scale[i], fmaf(-P.act_mean[i], P.act_scale[i], P.act_beta[i]));
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kMtStagesMax; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], (dbl && (BNB || !P.split)) ? 128 : 256);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512u);
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
      int stage = 0;
      uint32_t ph = 0;
      const uint32_t img = (uint32_t)P.img_bytes, slot = (uint32_t)P.slot_bytes;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const uint32_t m0 = (uint32_t)tile * 128u;
        const int jw0 = (int)(m0 % Wj);
        const uint32_t t2 = m0 / Wj;
        const int jh0 = (int)(t2 % Hj), b0 = (int)(t2 / Hj);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.wpk);
        for (int c = 0; c < chunks; ++c) {
#pragma unroll
          for (int box = 0; box < 4; ++box) {
            const uint32_t wbytes = (uint32_t)tmat_nimg(box) * img;
            mbar_wait(&empty_bar[stage], ph ^ 1u);
            uint8_t* st = smem + (uint32_t)stage * slot;
            mbar_arrive_expect_tx(&full_bar[stage], 16384u + wbytes);
            tma_load_4d(st, &P.tmap, &full_bar[stage], c * 64, jw0 + (box & 1), jh0 + (box >> 1), b0);
            bulk_g2s(st + 16384, wsrc, wbytes, &full_bar[stage]);
            wsrc += wbytes;
            if (++stage == kStages) { stage = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)N, 0, 0);
    const uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);   // SBO = 1024, version, SWIZZLE_128B
    const uint64_t hi64 = (uint64_t)hi << 32;
    const uint32_t smem_a = smem_u32(smem), slot = (uint32_t)P.slot_bytes, img16 = (uint32_t)P.img_bytes >> 4;
    int stage = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int acc = dbl ? (it & 1) : 0;
      const uint32_t acc_ph = dbl ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(acc * 4 * N);
      for (int c = 0; c < chunks; ++c) {
#pragma unroll
        for (int box = 0; box < 4; ++box) {
          mbar_wait(&full_bar[stage], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_a + (uint32_t)stage * slot;
            const uint32_t lo_a = (a_addr >> 4) | (1u << 16);
            uint32_t lo_b = ((a_addr + 16384u) >> 4) | (1u << 16);
            const int dh = box >> 1, dw = box & 1;
#pragma unroll
            for (int php = dh; php < 2; ++php)
#pragma unroll
              for (int pwp = dw; pwp < 2; ++pwp) {
                const uint32_t d_tmem = d_base + (uint32_t)((php * 2 + pwp) * N);
                // every phase receives its first contribution from box (0, 0) of chunk 0
                mma_bf16_ss(d_tmem, hi64 | lo_a, hi64 | lo_b, idesc, (c > 0 || box > 0) ? 1u : 0u);
                mma_bf16_ss(d_tmem, hi64 | (lo_a + 2u), hi64 | (lo_b + 2u), idesc, 1u);
                mma_bf16_ss(d_tmem, hi64 | (lo_a + 4u), hi64 | (lo_b + 4u), idesc, 1u);
                mma_bf16_ss(d_tmem, hi64 | (lo_a + 6u), hi64 | (lo_b + 6u), idesc, 1u);
                lo_b += img16;
              }
            mma_commit(&empty_bar[stage]);
            if (c == chunks - 1 && box == 3) mma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int ew = warp & 3;           // TMEM lane group = warp % 4
    const int eset = (warp - 2) >> 2;  // double-buffered: set = accumulator set (tile parity); else: set = output-row parity
    const int r = ew * 32 + lane;
    const bool has_bias = P.bias != nullptr;
    const uint32_t s_bn_a = smem_u32(s_bn);
    // Work split between the two sets.  Two accumulator sets (N = 64): by default a set takes every other tile; with
    // P.split both sets work on every tile, each on its own 32-channel chunk -- the chunk's statistics then stay in
    // registers for the whole kernel (no transposes inside the tile loop).  One accumulator set (N = 128): the sets
    // split the output-row parity.
    const bool split = !BNB && dbl && P.split != 0;
    const int it0 = (dbl && !split) ? eset : 0, it_step = (dbl && !split) ? 2 : 1;
    const int php0 = dbl ? 0 : eset, php1 = dbl ? 2 : eset + 1;
    const int c_base = split ? eset * 32 : 0;
    float s1[32], s2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s1[i] = s2[i] = 0.0f;
    for (int it = it0, tile = blockIdx.x + it0 * gridDim.x; tile < P.total_tiles; tile += it_step * gridDim.x, it += it_step) {
      const int acc = dbl ? (it & 1) : 0;
      const uint32_t acc_ph = dbl ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      const long long m = (long long)tile * 128 + r;
      const bool valid = m < M_total;
      bf16* dst0 = nullptr;
      if (valid) {
        const uint32_t mu = (uint32_t)m;
        const uint32_t jw = mu % Wj, t2 = mu / Wj;
        const uint32_t jh = t2 % Hj, b = t2 / Hj;
        dst0 = P.out + (((long long)b * P.Hout + 2 * jh) * P.Wout + 2 * jw) * P.ldo;
      }
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * 4 * N);
      if constexpr (!BNB) {
        // Channel chunk outermost: the statistics of a chunk accumulate over its four phases in registers (one transpose
        // pair per chunk, not per phase), and every 32 x 64 B piece goes through a swizzled shared-memory slab so that a
        // store instruction writes 8 rows x 64 contiguous bytes (whole sectors) instead of 32 rows x 16 bytes.
        const uint32_t stg = smem_u32(s_stage) + (uint32_t)(warp - 2) * 2048u;
        const uint32_t st_w = stg + (uint32_t)lane * 64u, wx = ((uint32_t)lane >> 1) & 3u;
        const uint32_t st_r = stg + ((uint32_t)lane >> 2) * 64u + ((((uint32_t)lane & 3u) ^ (((uint32_t)lane >> 3) & 3u)) << 4);
        // output rows (lane / 4) + 8k are the ones this lane stores: their bases as 16-byte indices (all ones: out of range)
        const uint32_t my16 = valid ? (uint32_t)((reinterpret_cast<const uint8_t*>(dst0) - reinterpret_cast<const uint8_t*>(P.out)) >> 4) : 0xFFFFFFFFu;
        uint32_t d16[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) d16[k] = __shfl_sync(0xffffffffu, my16, (lane >> 2) + 8 * k);
        // One flat loop over (channel chunk, phase): the TMEM load of the NEXT piece is issued as soon as this one is packed
        // to bf16, so it flies under the shared-memory round trip and the stores.
        const int pcs = dbl ? 2 : 1, per_c = 1 << pcs;   // pieces per channel chunk: 4 phases, or this set's 2
        const int n_it = split ? per_c : (N >> 5) << pcs;
        const uint32_t s_bias_a = smem_u32(s_bias);
        const int Wout = P.Wout, ldo = P.ldo;
        const float slope = P.slope;
        float v[32];
        tmem_ld_32x32(t_row + (uint32_t)(php0 * 2 * N + c_base), v);
        for (int idx = 0; idx < n_it; ++idx) {
          const int c0 = c_base + ((idx >> pcs) << 5), rem = idx & (per_c - 1);
          const int php = php0 + (rem >> 1), pwp = rem & 1;
          float4 b4[8];   // requested before the wait: the latency hides under the TMEM load
#pragma unroll
          for (int i = 0; i < 8; ++i) b4[i] = lds_f4_addr(s_bias_a + (uint32_t)((c0 + 4 * i) * 4));
          tmem_ld_wait_regs(v);
          if (idx == n_it - 1) {   // last read of this set's accumulators
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {   // zeros when the launch has no bias
            v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w;
          }
          if constexpr (EPI == 2) {   // inference: eval-mode BatchNorm + LeakyReLU on the accumulator row
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float4 c4 = lds_f4_addr(s_bn_a + (uint32_t)((c0 + i) * 8));   // scale, shift of two channels
              const float u0 = fmaf(v[i], c4.x, c4.y), u1 = fmaf(v[i + 1], c4.z, c4.w);
              v[i] = u0 > 0.0f ? u0 : u0 * slope;
              v[i + 1] = u1 > 0.0f ? u1 : u1 * slope;
            }
          }
          if (EPI == 1 && valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s1[i] += v[i];
              s2[i] = fmaf(v[i], v[i], s2[i]);
            }
          }
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          if (idx + 1 < n_it) {
            const int c0n = c_base + (((idx + 1) >> pcs) << 5), remn = (idx + 1) & (per_c - 1);
            tmem_ld_32x32(t_row + (uint32_t)(((php0 + (remn >> 1)) * 2 + (remn & 1)) * N + c0n), v);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(st_w + (((uint32_t)q ^ wx) << 4)), "r"(w[4 * q]),
                         "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3]) : "memory");
          __syncwarp();
          const uint32_t off16 = (uint32_t)(((php * Wout + pwp) * ldo + c0) >> 3) + (uint32_t)(lane & 3);
          uint4 pk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pk[k].x), "=r"(pk[k].y), "=r"(pk[k].z), "=r"(pk[k].w) : "r"(st_r + (uint32_t)k * 512u));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (d16[k] != 0xFFFFFFFFu) stg_v4(reinterpret_cast<uint4*>(P.out) + ((size_t)d16[k] + off16), pk[k]);
          __syncwarp();
          if (EPI == 1 && !split && rem == per_c - 1) {
            warp_transpose_reduce32(s1, lane);
            warp_transpose_reduce32(s2, lane);
            atomicAdd(&s_stats[c0 + lane], s1[0]);
            atomicAdd(&s_stats[128 + c0 + lane], s2[0]);
#pragma unroll
            for (int i = 0; i < 32; ++i) s1[i] = s2[i] = 0.0f;
          }
        }
      } else {
      for (int php = php0; php < php1; ++php) {
        for (int pwp = 0; pwp < 2; ++pwp) {
          bf16* dst = dst0 != nullptr ? dst0 + ((long long)php * P.Wout + pwp) * P.ldo : nullptr;
          const bf16* ysrc = BNB ? P.bn_y + (dst != nullptr ? (long long)(dst - P.out) : 0) : nullptr;
          for (int c0 = 0; c0 < N; c0 += 32) {
            uint4 yraw[4];
            if constexpr (BNB) {
#pragma unroll
              for (int q = 0; q < 4; ++q) yraw[q] = ldg_nc_v4_pinned(reinterpret_cast<const uint4*>(ysrc + c0) + q);
            }
            float v[32];
            tmem_ld_32x32(t_row + (uint32_t)((php * 2 + pwp) * N + c0), v);
            tmem_ld_wait();
            if (php == php1 - 1 && pwp == 1 && c0 + 32 >= N) {   // last read of this set's accumulators
              tc_fence_before();
              mbar_arrive(&tempty_bar[acc]);
            }
            if (has_bias) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + i]);
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
            }
            if (!valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            float second[32];
            if constexpr (BNB) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t w4[4] = {yraw[q].x, yraw[q].y, yraw[q].z, yraw[q].w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const int i = q * 8 + k;
                  const float yv = __uint_as_float((k & 1) ? (w4[k >> 1] & 0xFFFF0000u) : (w4[k >> 1] << 16));
                  const float4 c4 = lds_f4_addr(s_bn_a + (uint32_t)((c0 + i) * 16));   // mean, scale, beta, rstd
                  const float t = yv - c4.x;
                  const float u = fmaf(t, c4.y, c4.z);
                  const float du = u > 0.0f ? v[i] : v[i] * P.slope;
                  v[i] = du;
                  second[i] = du * (t * c4.w);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) second[i] = v[i] * v[i];
            }
            if (valid) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 pk;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[i + 2 * k], v[i + 2 * k + 1]);
                *reinterpret_cast<uint4*>(dst + c0 + i) = pk;
              }
            }
            if (P.stats != nullptr) {
              warp_transpose_reduce32(v, lane);
              warp_transpose_reduce32(second, lane);
              atomicAdd(&s_stats[c0 + lane], v[0]);
              atomicAdd(&s_stats[128 + c0 + lane], second[0]);
            }
          }
        }
      }
      }
    }
    if (!BNB && EPI == 1 && split) {   // the chunk's sums of the whole kernel
      warp_transpose_reduce32(s1, lane);
      warp_transpose_reduce32(s2, lane);
      atomicAdd(&s_stats[c_base + lane], s1[0]);
      atomicAdd(&s_stats[128 + c_base + lane], s2[0]);
    }
    if (P.stats != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
      for (int i = threadIdx.x - 64; i < N; i += 256) {
        atomicAdd(&P.stats[i], (double)s_stats[i]);
        atomicAdd(&P.stats2[i], (double)s_stats[128 + i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

static bool g_tmat_enabled = true;

// channel-only rule (the weight pack must not depend on B, H, W): transposed stride-2 conv with 64 output channels
// (N = Ci of the equivalent conv), Cin (= its Co) a multiple of 64.  The kernel also runs N = 128 (one accumulator
// set of 512 TMEM columns, the epilogue sets split the output rows; unit-tested through g_tmat_wide) but without a
// second accumulator set the epilogue is not hidden under the next tile's MMAs: measured 102 us against 89 us for
// the per-phase kernel on the 128->256 layer at B = 512 (274 vs 236 us at the wide shape), so N = 128 stays there.
static bool g_tmat_wide = false;
bool tmat_channels_ok(int Ci, int Co, int stride, int mode) {
  return g_tmat_enabled && mode == 1 && stride == 2 && (Ci == 64 || (g_tmat_wide && Ci == 128)) && Co % 64 == 0 && Co >= 64 &&
         Co <= 2048;
}

long long tmat_packed_elems(int Ci, int Co) { return 9LL * Ci * Co; }

int tmat_pack(const float* w_oihw, void* out, int Ci, int Co, cudaStream_t st) {
  const long long total = 9LL * Ci * Co;
  pack_tmat_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w_oihw, (bf16*)out, Co, Ci);
  return check_launch("pack_tmat_weight");
}

// returns 1 if the geometry is not served (whole-row tiles), 0 / negative otherwise.  in: [B, Hi/2, Wi/2, Co] -> out [B, Hi, Wi, Ci]
int tmat_launch(const void* in, const void* wpk, const float* bias, void* out, double* stats, int B, int Hi, int Wi, int Ci,
                int Co, cudaStream_t st, const void* bn_y, const float* const* bn_vec, float slope, const float* const* act) {
  GatherGeom g = geom_dgrad(B, Hi, Wi, Ci, Co, 2);
  int rw, rh, rn;
  if (!tile_rows(g, &rw, &rh, &rn) || rw > 256 || rh > 256 || rn > 256) return 1;
  const long long M = (long long)B * g.Hj * g.Wj;
  if (M >= (1LL << 31) || get_encode_fn() == nullptr) return 1;
  TmatParams P{};
  cuuint64_t gdim[4] = {(cuuint64_t)Co, (cuuint64_t)g.Win, (cuuint64_t)g.Hin, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)Co * 2, (cuuint64_t)g.Win * Co * 2, (cuuint64_t)g.Hin * g.Win * Co * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)rw, (cuuint32_t)rh, (cuuint32_t)rn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = get_encode_fn()(&P.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("conv_umma_tmat: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return VAE_ERR_CUDA;
  }
  const int N = Ci;
  P.B = B; P.Hj = g.Hj; P.Wj = g.Wj; P.Cin = Co; P.N = N; P.Hout = Hi; P.Wout = Wi; P.ldo = N;
  P.chunks = Co / 64;
  P.total_tiles = (int)((M + 127) / 128);
  P.img_bytes = N * 64 * 2;
  P.slot_bytes = 16384 + 4 * P.img_bytes;
  P.stages = (202 * 1024) / P.slot_bytes;   // 16 KB of the budget is the epilogue's store slab
  if (P.stages > kMtStagesMax) P.stages = kMtStagesMax;
  if (P.stages < 2) return 1;
  P.double_buf = 8 * N <= 512 ? 1 : 0;
  // measured at the wide shape (64 -> 64 channels, 1x128x1024 rolls, B = 64): with statistics 369 -> 301 us, without 280 -> 287 us
  P.split = (stats != nullptr && bn_y == nullptr) ? 1 : 0;
  P.wpk = (const bf16*)wpk; P.bias = bias; P.out = (bf16*)out;
  P.stats = stats; P.stats2 = stats ? stats + N : nullptr;
  P.bn_y = (const bf16*)bn_y; P.slope = slope;
  if (bn_y != nullptr) { P.bn_mean = bn_vec[0]; P.bn_scale = bn_vec[1]; P.bn_beta = bn_vec[2]; P.bn_rstd = bn_vec[3]; }
  if (act != nullptr) { P.act_mean = act[0]; P.act_scale = act[1]; P.act_beta = act[2]; }
  if (stats && zero_scratch(stats, sizeof(double) * 2 * N, st) != cudaSuccess) {
    set_error("conv_umma_tmat: memset");
    return VAE_ERR_CUDA;
  }
  const size_t smem = 1024 + (size_t)P.stages * P.slot_bytes + 8 * (2 * kMtStagesMax + 4) + 16 + (256 + 128) * 4 + 128 * 16 + 8 * 2048 + 64;
  VAE_REQUIRE(!(act != nullptr && (stats != nullptr || bn_y != nullptr)), "conv_umma_tmat: the inference epilogue takes no statistics");
  const int which = bn_y != nullptr ? 3 : act != nullptr ? 2 : stats != nullptr ? 1 : 0;
  auto kern = which == 3 ? umma_tmat_conv_kernel<true, 1> : which == 2 ? umma_tmat_conv_kernel<false, 2>
            : which == 1 ? umma_tmat_conv_kernel<false, 1> : umma_tmat_conv_kernel<false, 0>;
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[which]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("conv_umma_tmat: shared memory attribute");
      return VAE_ERR_CUDA;
    }
    attr_set[which] = true;
  }
  int grid = sm_count();
  if (grid > P.total_tiles) grid = P.total_tiles;
  kern<<<grid, kMtThreads, smem, st>>>(P);
  return check_launch("umma_tmat_conv");
}

}  // namespace vae

/* Switch for the fused-phase transposed-conv kernel behind vae_conv_umma_tma (mode 1, 64 / 128 output channels); it also
 * selects the weight layout vae_umma_tma_pack_weight produces, so flip it BEFORE packing.  Returns the previous setting. */
extern "C" int vae_set_umma_tmat(int enable) {   // bit 0: the kernel (N = 64); bit 1: also N = 128 (slower, tests only)
  int old = (vae::g_tmat_enabled ? 1 : 0) | (vae::g_tmat_wide ? 2 : 0);
  vae::g_tmat_enabled = (enable & 1) != 0;
  vae::g_tmat_wide = (enable & 2) != 0;
  return old;
}
