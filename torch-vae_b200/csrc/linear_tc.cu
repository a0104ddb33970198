// Tensor-core kernels for the three Linear layers when the latent is WIDE (L = 64 .. 256; the long / wide
// configuration: F = 262 144 features, L = 256, batch 64 per GPU).  fc_mu / fc_var / decoder_input and their
// gradients (models.py:55-59,137,141,162; train.py:650) are GEMMs with one huge extent (F), one medium (L) and a
// small batch: each one streams a 268 MB fp32 master weight (or writes a 268 MB fp32 gradient) exactly once and
// is HBM-bound (~45 us at the measured copy peak) -- if the 8.6 GFLOP ride on tensor cores.  The CUDA-core GEMM
// they used before (gemm_simt.cu) ran them at 0.6 - 1.1 TB/s.
//
// No packing pass exists: the fp32 master weights are read in place (cp.async into shared memory), rounded to
// bf16 while the mma.sync fragments are formed (the same operand precision as the conv layers' weight images),
// accumulated in fp32.  Warp-level mma.sync.m16n8k16 (not tcgen05): the operands need a register-level
// conversion and the kernels are bandwidth-bound at < 15 % of the legacy tensor pipe.
//
//   reduce :  out[m][l]  = sum_f A[m][f] * W(l, f) (+ bias[l])      fc_mu / fc_var forward, dz      split-K over CTAs
//   expand :  out[m][f]  = sum_l D[m][l] * W(l, f) (+ bias[f])      decoder_input forward, d_pre    tiles of 64 f
//   wgrad  :  G(l, f)    = sum_m D[m][l] * A[m][f]                  all three weight gradients      tiles of 64 f
//
// W / G are addressed in one of two dense layouts: 0 = [L][F] (fc_mu, fc_var), 1 = [F][L] (decoder_input).
#include "common.cuh"

namespace vae {

constexpr int kLtThreads = 256;
constexpr int kLtMaxM = 64;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t a, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t a, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

// bf16 [rows][64] tile with 128-byte rows: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
__device__ __forceinline__ uint32_t swz128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// D [M][L] fp32 (global) -> bf16 [64][L] in shared memory (rows >= M zero), 16-byte chunks XOR-swizzled per row
// with (row & 7): element (m, l) at row m * (2L) + (((l >> 3) ^ (m & 7)) << 4) + (l & 7) * 2.
__device__ __forceinline__ void stage_d_bf16(uint8_t* dst, const float* __restrict__ D, int M, int L) {
  const int chunks = 64 * (L / 8);
  for (int c = threadIdx.x; c < chunks; c += kLtThreads) {
    const int m = c / (L / 8), q = c % (L / 8);
    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
    if (m < M) {
      const float4 a = *reinterpret_cast<const float4*>(D + (long long)m * L + q * 8);
      const float4 b = *reinterpret_cast<const float4*>(D + (long long)m * L + q * 8 + 4);
      pk = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    }
    *reinterpret_cast<uint4*>(dst + (size_t)m * (2 * L) + (((q ^ (m & 7))) << 4)) = pk;
  }
}

// =====================================================================================================
// reduce: out[m][l] += sum_{f in this CTA's range} A[m][f] * W(l, f)        (L == 8 warps x 32 columns max)
//   k-tiles of 32 features; per stage: A tile bf16 [64][32] (64-byte rows) + W tile fp32
//   layout 0: [L rows][32 f], pitch 40 floats    layout 1: [32 f rows][L], pitch L + 4 floats
// =====================================================================================================
constexpr int kRedKT = 32;
constexpr int kRedStages = 4;

template <int LAYOUT>
__global__ void __launch_bounds__(kLtThreads, 1)
linear_reduce_tc_kernel(const bf16* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                        float* __restrict__ out, int M, int L, int F, int tiles_total) {
  extern __shared__ __align__(128) uint8_t lt_smem[];
  const int w_pitch = LAYOUT == 0 ? 40 : L + 4;                       // floats
  const int w_bytes = (LAYOUT == 0 ? L : kRedKT) * w_pitch * 4;
  const int stage_bytes = w_bytes + 64 * 64;                          // + A tile [64][32] bf16
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_warps_used = L / 32;                                    // warps that own 32 output columns
  // contiguous k-tile range of this CTA
  const int per = (tiles_total + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(tiles_total, t_begin + per);
  const int n_tiles = t_end - t_begin;
  if (n_tiles <= 0) return;

  auto issue = [&](int tile, int stage) {
    uint8_t* st = lt_smem + (size_t)stage * stage_bytes;
    const long long f0 = (long long)tile * kRedKT;
    if (LAYOUT == 0) {
      const int chunks = L * 8;   // 8 x 16 B per row of 32 floats
      for (int c = threadIdx.x; c < chunks; c += kLtThreads) {
        const int n = c >> 3, q = c & 7;
        cp_async16(smem_addr(st + (n * w_pitch + q * 4) * 4), W + (long long)n * F + f0 + q * 4, true);
      }
    } else {
      const int cpr = L / 4;
      const int chunks = kRedKT * cpr;
      for (int c = threadIdx.x; c < chunks; c += kLtThreads) {
        const int k = c / cpr, q = c % cpr;
        cp_async16(smem_addr(st + (k * w_pitch + q * 4) * 4), W + (f0 + k) * L + q * 4, true);
      }
    }
    // A tile: 64 rows x 4 chunks of 16 B (8 bf16); 64-byte rows, chunk ^= (row >> 1) & 3
    {
      const int c = threadIdx.x;   // 256 chunks
      const int m = c >> 2, q = c & 3;
      cp_async16(smem_addr(st + w_bytes + m * 64 + ((q ^ ((m >> 1) & 3)) << 4)), A + (long long)min(m, M - 1) * F + f0 + q * 8, m < M);
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

  for (int s = 0; s < kRedStages - 1; ++s) {
    if (s < n_tiles) issue(t_begin + s, s);
    cp_async_commit();
  }
  for (int it = 0; it < n_tiles; ++it) {
    cp_async_wait<kRedStages - 2>();
    __syncthreads();
    {   // refill the stage that was consumed in the previous iteration
      const int nxt = it + kRedStages - 1;
      if (nxt < n_tiles) issue(t_begin + nxt, nxt % kRedStages);
      cp_async_commit();
    }
    if (warp < n_warps_used) {
      const uint8_t* st = lt_smem + (size_t)(it % kRedStages) * stage_bytes;
      const uint32_t w_a = smem_addr(st), a_a = smem_addr(st + w_bytes);
#pragma unroll
      for (int ks = 0; ks < kRedKT / 16; ++ks) {
        uint32_t af[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 16 + (lane & 15), chunk = ks * 2 + (lane >> 4);
          ldmatrix_x4(a_a + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4), af[i]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = warp * 32 + j * 8 + g, k = ks * 16 + 2 * t;
          uint32_t b0, b1;
          if (LAYOUT == 0) {
            const float2 lo = lds_f2(w_a + (n * w_pitch + k) * 4), hi = lds_f2(w_a + (n * w_pitch + k + 8) * 4);
            b0 = pack_bf16(lo.x, lo.y);
            b1 = pack_bf16(hi.x, hi.y);
          } else {
            b0 = pack_bf16(lds_f1(w_a + (k * w_pitch + n) * 4), lds_f1(w_a + ((k + 1) * w_pitch + n) * 4));
            b1 = pack_bf16(lds_f1(w_a + ((k + 8) * w_pitch + n) * 4), lds_f1(w_a + ((k + 9) * w_pitch + n) * 4));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) mma_16816(acc[i][j], af[i], b0, b1);
        }
      }
    }
  }
  cp_async_wait<0>();
  if (warp < n_warps_used) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = warp * 32 + j * 8 + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = i * 16 + g + 8 * h;
          if (m < M) {
            float v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
            if (blockIdx.x == 0 && bias != nullptr) { v0 += bias[n]; v1 += bias[n + 1]; }
            atomicAdd(&out[(long long)m * L + n], v0);
            atomicAdd(&out[(long long)m * L + n + 1], v1);
          }
        }
      }
  }
}

// =====================================================================================================
// expand: out[m][f] = bias[f] + sum_p sum_l D_p[m][l] * W_p(l, f)   for a tile of 64 features, p = 1 or 2 operands
//   D_p resident as bf16 [64][L]; per stage one W tile fp32:
//   layout 0: [L rows][64 f], pitch 68 floats (W[l][f])     layout 1: [64 f rows][L], pitch L + 8 floats (W[f][l])
// =====================================================================================================
constexpr int kExpNT = 64;

template <int LAYOUT, typename TO>
__global__ void __launch_bounds__(kLtThreads, 1)
linear_expand_tc_kernel(const float* __restrict__ D0, const float* __restrict__ W0, const float* __restrict__ D1,
                        const float* __restrict__ W1, const float* __restrict__ bias, TO* __restrict__ out, int M, int L,
                        int F) {
  extern __shared__ __align__(128) uint8_t lt_smem[];
  const int nops = D1 != nullptr ? 2 : 1;
  const int d_bytes = 64 * L * 2;
  const int w_pitch = LAYOUT == 0 ? kExpNT + 4 : L + 8;
  const int w_bytes = (LAYOUT == 0 ? L : kExpNT) * w_pitch * 4;
  uint8_t* d_s = lt_smem;                         // [nops][64][L] bf16
  uint8_t* w_s = lt_smem + nops * d_bytes;        // [2 stages][w_bytes]
  float* o_s = reinterpret_cast<float*>(w_s + 2 * w_bytes);   // [64][kExpNT + 4] fp32 output staging
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;   // 2 x 4 warps: 32 rows x 16 columns each
  stage_d_bf16(d_s, D0, M, L);
  if (nops == 2) stage_d_bf16(d_s + d_bytes, D1, M, L);

  const int tiles = F / kExpNT;
  const int steps_total = ((tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * nops;   // (tile, operand) pairs of this CTA
  auto issue = [&](int step, int stage) {
    const int tile = blockIdx.x + (step / nops) * gridDim.x, op = step % nops;
    const float* W = op == 0 ? W0 : W1;
    uint8_t* st = w_s + (size_t)stage * w_bytes;
    const long long f0 = (long long)tile * kExpNT;
    if (LAYOUT == 0) {
      const int chunks = L * (kExpNT / 4);
      for (int c = threadIdx.x; c < chunks; c += kLtThreads) {
        const int l = c / (kExpNT / 4), q = c % (kExpNT / 4);
        cp_async16(smem_addr(st + (l * w_pitch + q * 4) * 4), W + (long long)l * F + f0 + q * 4, true);
      }
    } else {
      const int cpr = L / 4;
      const int chunks = kExpNT * cpr;
      for (int c = threadIdx.x; c < chunks; c += kLtThreads) {
        const int n = c / cpr, q = c % cpr;
        cp_async16(smem_addr(st + (n * w_pitch + q * 4) * 4), W + (f0 + n) * L + q * 4, true);
      }
    }
  };
  if (steps_total > 0) issue(0, 0);
  cp_async_commit();
  float acc[2][2][4];
  for (int step = 0; step < steps_total; ++step) {
    const int op = step % nops;
    if (op == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;
    }
    cp_async_wait<0>();
    __syncthreads();   // W tile of this step landed; everybody finished the previous step (and the D staging)
    if (step + 1 < steps_total) issue(step + 1, (step + 1) & 1);
    cp_async_commit();
    const uint32_t w_a = smem_addr(w_s + (size_t)(step & 1) * w_bytes);
    const uint32_t d_a = smem_addr(d_s + op * d_bytes);
    for (int ks = 0; ks < L / 16; ++ks) {
      uint32_t af[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = wm * 32 + i * 16 + (lane & 15), chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(d_a + row * (2 * L) + ((chunk ^ (row & 7)) << 4), af[i]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = wn * 16 + j * 8 + g, k = ks * 16 + 2 * t;
        uint32_t b0, b1;
        if (LAYOUT == 0) {   // W[l][f]: k rows, n contiguous
          b0 = pack_bf16(lds_f1(w_a + (k * w_pitch + n) * 4), lds_f1(w_a + ((k + 1) * w_pitch + n) * 4));
          b1 = pack_bf16(lds_f1(w_a + ((k + 8) * w_pitch + n) * 4), lds_f1(w_a + ((k + 9) * w_pitch + n) * 4));
        } else {             // W[f][l]: n rows, k contiguous
          const float2 lo = lds_f2(w_a + (n * w_pitch + k) * 4), hi = lds_f2(w_a + (n * w_pitch + k + 8) * 4);
          b0 = pack_bf16(lo.x, lo.y);
          b1 = pack_bf16(hi.x, hi.y);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) mma_16816(acc[i][j], af[i], b0, b1);
      }
    }
    if (op == nops - 1) {   // tile finished: stage, add the bias, store 64 contiguous features per row
      const int tile = blockIdx.x + (step / nops) * gridDim.x;
      const long long f0 = (long long)tile * kExpNT;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = wm * 32 + i * 16 + g + 8 * h, n = wn * 16 + j * 8 + 2 * t;
            *reinterpret_cast<float2*>(&o_s[m * (kExpNT + 4) + n]) = make_float2(acc[i][j][2 * h], acc[i][j][2 * h + 1]);
          }
      __syncthreads();
      // 64 rows x 64 columns: 8 columns per thread, two passes of 256 threads
      for (int c = threadIdx.x; c < 64 * (kExpNT / 8); c += kLtThreads) {
        const int m = c / (kExpNT / 8), q = c % (kExpNT / 8);
        if (m < M) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = o_s[m * (kExpNT + 4) + q * 8 + e] + (bias != nullptr ? bias[f0 + q * 8 + e] : 0.0f);
          TO* dst = out + (long long)m * F + f0 + q * 8;
          if constexpr (sizeof(TO) == 2) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          } else {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
      // o_s is rewritten only after the next step's __syncthreads
    }
  }
  cp_async_wait<0>();
}

// =====================================================================================================
// wgrad: G(l, f) = sum_m D[m][l] * A[m][f] for a tile of 64 features (K = the batch, zero-padded to 64).
//   D resident as bf16 [64][L] (both operands are read "transposed": ldmatrix.trans); A tiles bf16 [64][64] streamed.
//   layout 0: G[l][f] -> M = l (8 warps x 32 rows), N = 64 f      layout 1: G[f][l] -> M = 64 f (2 warps), N = l (4 x 64)
// =====================================================================================================
constexpr int kWgStages = 4;

template <int LAYOUT>
__global__ void __launch_bounds__(kLtThreads, 1)
linear_wgrad_tc_kernel(const float* __restrict__ D, const bf16* __restrict__ A, float* __restrict__ G, int M, int L, int F) {
  extern __shared__ __align__(128) uint8_t lt_smem[];
  const int d_bytes = 64 * L * 2;
  uint8_t* d_s = lt_smem;                 // [64 m][L] bf16, chunk ^= (m & 7)
  uint8_t* a_s = lt_smem + d_bytes;       // [stages][64 m][64 f] bf16 (128-byte rows, chunk ^= (m & 7))
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  stage_d_bf16(d_s, D, M, L);
  const int tiles = F / 64;
  const int n_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i, int stage) {
    const long long f0 = (long long)(blockIdx.x + i * gridDim.x) * 64;
    uint8_t* st = a_s + (size_t)stage * 8192;
    for (int c = threadIdx.x; c < 512; c += kLtThreads) {
      const int m = c >> 3, q = c & 7;
      cp_async16(smem_addr(st + swz128(m, q)), A + (long long)min(m, M - 1) * F + f0 + q * 8, m < M);
    }
  };
  for (int s = 0; s < kWgStages - 1; ++s) {
    if (s < n_tiles) issue(s, s);
    cp_async_commit();
  }
  const uint32_t d_a = smem_addr(d_s);
  // fragment of a [k = m rows][cols] bf16 tile read transposed: ldmatrix.x4.trans, lane -> (k row, 16-byte column chunk)
  const int lk = (lane & 7) + ((lane >> 4) << 3);   // k row within the 16-row k step
  const int lc = (lane >> 3) & 1;                   // which of the two 8-column chunks
  for (int it = 0; it < n_tiles; ++it) {
    cp_async_wait<kWgStages - 2>();
    __syncthreads();
    {
      const int nxt = it + kWgStages - 1;
      if (nxt < n_tiles) issue(nxt, nxt % kWgStages);
      cp_async_commit();
    }
    const uint32_t a_a = smem_addr(a_s + (size_t)(it % kWgStages) * 8192);
    const long long f0 = (long long)(blockIdx.x + it * gridDim.x) * 64;
    if (LAYOUT == 0) {
      // M = l: this warp's rows l0 = warp * 32 .. + 31 (two m tiles); N = 64 f (eight n tiles); K = 64 m (four k steps)
      if (warp * 32 < L) {
        float acc[2][8][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t af[2][4];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            // A operand (m = l, k = batch) from D[batch][l]: matrices {k lo, l lo}, {k lo, l hi}, {k hi, l lo}, {k hi, l hi}
            const int krow = ks * 16 + lk, chunk = (warp * 32 + i * 16) / 8 + lc;
            ldmatrix_x4_trans(d_a + krow * (2 * L) + ((chunk ^ (krow & 7)) << 4), af[i]);
          }
#pragma unroll
          for (int jp = 0; jp < 4; ++jp) {
            // B operand (k = batch, n = f) from A[batch][f]: one x4.trans covers two n tiles: {k lo, n0}, {k hi, n0}, {k lo, n1}, {k hi, n1}
            uint32_t bf[4];
            const int krow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), chunk = jp * 2 + (lane >> 4);
            ldmatrix_x4_trans(a_a + swz128(krow, chunk), bf);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              mma_16816(acc[i][2 * jp], af[i], bf[0], bf[1]);
              mma_16816(acc[i][2 * jp + 1], af[i], bf[2], bf[3]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int l = warp * 32 + i * 16 + g + 8 * h;
              *reinterpret_cast<float2*>(G + (long long)l * F + f0 + j * 8 + 2 * t) = make_float2(acc[i][j][2 * h], acc[i][j][2 * h + 1]);
            }
      }
    } else {
      // M = 64 f (warp >> 2 selects 32 of them), N = l (warp & 3 selects a quarter of L); K = 64 m
      const int wm = warp >> 2, wn = warp & 3;
      const int nq = L / 4;   // columns per warp (multiple of 8)
      for (int n0 = 0; n0 < nq; n0 += 32) {   // 32 columns (four n tiles) at a time
        float acc[2][4][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;
        const int nb = wn * nq + n0;
        const int ncols = min(32, nq - n0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t af[2][4];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int krow = ks * 16 + lk, chunk = (wm * 32 + i * 16) / 8 + lc;   // A operand (m = f, k = batch) from A[batch][f]
            ldmatrix_x4_trans(a_a + swz128(krow, chunk), af[i]);
          }
#pragma unroll
          for (int jp = 0; jp < 2; ++jp) {
            if (jp * 16 < ncols) {
              uint32_t bf[4];
              const int krow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), chunk = (nb + jp * 16) / 8 + (lane >> 4);
              ldmatrix_x4_trans(d_a + krow * (2 * L) + ((chunk ^ (krow & 7)) << 4), bf);   // B operand (k = batch, n = l) from D[batch][l]
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                mma_16816(acc[i][2 * jp], af[i], bf[0], bf[1]);
                mma_16816(acc[i][2 * jp + 1], af[i], bf[2], bf[3]);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j * 8 < ncols) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int f = wm * 32 + i * 16 + g + 8 * h;
                *reinterpret_cast<float2*>(G + (f0 + f) * L + nb + j * 8 + 2 * t) = make_float2(acc[i][j][2 * h], acc[i][j][2 * h + 1]);
              }
            }
      }
    }
  }
  cp_async_wait<0>();
}

static bool linear_tc_ok(int M, int L, int F) {
  return M >= 1 && M <= kLtMaxM && L >= 64 && L <= 256 && L % 64 == 0 && F >= 4096 && F % 64 == 0;   // L % 64: the XOR swizzle of the [64][L] bf16 tile
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
  VAE_REQUIRE(bytes <= 227 * 1024, "linear_tc: shared memory budget (%zu bytes)", bytes);
  VAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return VAE_OK;
}

}  // namespace vae

using namespace vae;

/* 1 when the tensor-core Linear kernels serve (batch M, latent L, features F): M <= 64, 64 <= L <= 256 a multiple of
 * 64, F >= 4096 a multiple of 64.  (L <= 16 is served by the vae_skinny_* kernels, everything else by vae_gemm.) */
extern "C" int vae_linear_tc_supported(int M, int L, int F) { return linear_tc_ok(M, L, F) ? 1 : 0; }

/* out[M][L] (fp32) = A[M][F] (bf16) . W + bias; w_layout 0: W[L][F] (fc_mu / fc_var, models.py:137,141),
 * 1: W[F][L] (latent gradient of decoder_input, train.py:650).  Split over F across the SMs, fp32 atomics.
 * replaces: aten::addmm / aten::mm */
extern "C" int vae_linear_reduce_tc(const void* A, const float* W, const float* bias, float* out, int M, int L, int F,
                                    int w_layout, vae_stream_t stream) {
  VAE_REQUIRE(A && W && out && (w_layout == 0 || w_layout == 1), "linear_reduce_tc: bad argument");
  VAE_REQUIRE(linear_tc_ok(M, L, F), "linear_reduce_tc: shape not served");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(out, sizeof(float) * (size_t)M * L, st));
  const int tiles = F / kRedKT;
  const int w_bytes = w_layout == 0 ? L * 40 * 4 : kRedKT * (L + 4) * 4;
  const size_t smem = (size_t)kRedStages * (w_bytes + 64 * 64);
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  if (w_layout == 0) {
    int rc = set_smem(linear_reduce_tc_kernel<0>, smem);
    if (rc != VAE_OK) return rc;
    linear_reduce_tc_kernel<0><<<grid, kLtThreads, smem, st>>>((const bf16*)A, W, bias, out, M, L, F, tiles);
  } else {
    int rc = set_smem(linear_reduce_tc_kernel<1>, smem);
    if (rc != VAE_OK) return rc;
    linear_reduce_tc_kernel<1><<<grid, kLtThreads, smem, st>>>((const bf16*)A, W, bias, out, M, L, F, tiles);
  }
  return check_launch("linear_reduce_tc");
}

/* out[M][F] (dt_out) = bias[f] + D0[M][L] . W0 (+ D1 . W1); D fp32; w_layout 0: W[L][F] (input gradient of
 * fc_mu + fc_var, train.py:650), 1: W[F][L] (decoder_input forward, models.py:162).
 * replaces: aten::addmm / aten::mm */
extern "C" int vae_linear_expand_tc(const float* D0, const float* W0, const float* D1, const float* W1, const float* bias,
                                    void* out, int M, int L, int F, int w_layout, int dt_out, vae_stream_t stream) {
  VAE_REQUIRE(D0 && W0 && out && (w_layout == 0 || w_layout == 1) && ((D1 == nullptr) == (W1 == nullptr)),
              "linear_expand_tc: bad argument");
  VAE_REQUIRE(linear_tc_ok(M, L, F), "linear_expand_tc: shape not served");
  cudaStream_t st = as_stream(stream);
  const int nops = D1 ? 2 : 1;
  const int w_bytes = w_layout == 0 ? L * (kExpNT + 4) * 4 : kExpNT * (L + 8) * 4;
  const size_t smem = (size_t)nops * 64 * L * 2 + 2 * (size_t)w_bytes + 64 * (kExpNT + 4) * 4;
  int grid = sm_count();
  if (grid > F / kExpNT) grid = F / kExpNT;
#define VAE_LT_EXPAND(LAY, T)                                                                                          \
  do {                                                                                                                 \
    int rc = set_smem(linear_expand_tc_kernel<LAY, T>, smem);                                                          \
    if (rc != VAE_OK) return rc;                                                                                       \
    linear_expand_tc_kernel<LAY, T><<<grid, kLtThreads, smem, st>>>(D0, W0, D1, W1, bias, (T*)out, M, L, F);           \
  } while (0)
  if (dt_out == VAE_BF16) {
    if (w_layout == 0) VAE_LT_EXPAND(0, bf16); else VAE_LT_EXPAND(1, bf16);
  } else {
    if (w_layout == 0) VAE_LT_EXPAND(0, float); else VAE_LT_EXPAND(1, float);
  }
#undef VAE_LT_EXPAND
  return check_launch("linear_expand_tc");
}

/* G = D[M][L]^T . A[M][F] (fp32, overwritten): g_layout 0: G[L][F] (fc_mu / fc_var weight gradients),
 * 1: G[F][L] (decoder_input weight gradient); D fp32, A bf16.
 * replaces: the weight-gradient aten::mm of the three Linear layers (train.py:650) */
extern "C" int vae_linear_wgrad_tc(const float* D, const void* A, float* G, int M, int L, int F, int g_layout,
                                   vae_stream_t stream) {
  VAE_REQUIRE(D && A && G && (g_layout == 0 || g_layout == 1), "linear_wgrad_tc: bad argument");
  VAE_REQUIRE(linear_tc_ok(M, L, F), "linear_wgrad_tc: shape not served");
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)64 * L * 2 + (size_t)kWgStages * 8192;
  int grid = 2 * sm_count();   // the kernels are store-bound: two waves of short CTAs balance better than one
  if (grid > F / 64) grid = F / 64;
  if (g_layout == 0) {
    int rc = set_smem(linear_wgrad_tc_kernel<0>, smem);
    if (rc != VAE_OK) return rc;
    linear_wgrad_tc_kernel<0><<<grid, kLtThreads, smem, st>>>(D, (const bf16*)A, G, M, L, F);
  } else {
    int rc = set_smem(linear_wgrad_tc_kernel<1>, smem);
    if (rc != VAE_OK) return rc;
    linear_wgrad_tc_kernel<1><<<grid, kLtThreads, smem, st>>>(D, (const bf16*)A, G, M, L, F);
  }
  return check_launch("linear_wgrad_tc");
}
