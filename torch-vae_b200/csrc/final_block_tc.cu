// tcgen05 / TMEM versions of the fused output-block kernels (final_block.cu) for the bench
// shape of the path: bf16, C = 32 channels, W = 128 pixels per image row.
//
// The C -> 1 convolution is a GEMM with N = 1, but written PER PIXEL it becomes tensor-core
// shaped:   P[p][t] = sum_c act[p][c] * w[t][c]          (M = pixels, N = 9 taps, K = C)
//           logit[q] = bias + sum_t P[q + d(t)][t]        (a 9-point shifted sum of P)
// and its data gradient is   da[p][c] = sum_t dl[p - d(t)] * w[t][c]   (M = pixels, N = C, K = 9 taps)
// so every element of the big tensor y is touched by CUDA cores only for the BatchNorm /
// LeakyReLU transform (4-8 instructions per element instead of ~26-30 in the SIMT kernels).
//
// Data flow per image row (128 pixels = one UMMA M tile):
//   TMA        y row [128 px][64 B] -> shared, 64B-swizzled (a ring of 4 rows, prefetched)
//   workers    thread = pixel: read its 64 B, BatchNorm + LeakyReLU, write bf16 act back IN PLACE:
//              the row buffer now IS the K-major (forward) / MN-major (weight gradient) UMMA operand
//   MMA warp   one elected thread issues tcgen05.mma, accumulators in TMEM
//   workers    tcgen05.ld their pixel's row of the accumulator and finish (shifted sum / BN backward)
//
//   warps 0-3 : workers (128 threads)      warp 4 : TMA producer      warp 5 : MMA issuer
#include "tma_util.cuh"

namespace vae {

constexpr int kFtThreads = 192;
constexpr int kFtSlots = 8;      // ring of raw y rows (64 KB in flight per CTA)
constexpr int kRowBytes = 8192;  // 128 pixels x 32 channels x bf16

struct FinalTcParams {
  CUtensorMap tmap_y;
  const float* mean;
  const float* scale;
  const float* beta;
  const float* rstd;
  const float* gamma;
  const float* w_last;   // [C][9] fp32 (OIHW with O == 1)
  const float* b_last;
  const bf16* dl;        // dlogits [B][H][W]
  bf16* logits;          // [B][H][W]
  float* probs;          // [B][H][W] or null
  bf16* dy;              // [B][H][W][C]
  double* dstats;        // [2C + 1]
  float* dw_last;        // [C][9]
  float* dbias;          // [C]
  int B, H, W, RB, bands_per_img, total_bands, training;
  float slope;
};

__device__ __forceinline__ uint32_t swz64(int row, int chunk) {  // byte offset inside a [rows][64 B] 64B-swizzled tile
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ void unpack8(const uint4& raw, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return raw;
}

__device__ __forceinline__ uint64_t make_desc_mn_tc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                    uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// =============================================================================================
// forward: logits = conv(lrelu(bn(y)), w_last) + b_last, probs = sigmoid(logits)
// =============================================================================================
__global__ void __launch_bounds__(kFtThreads, 2)
final_fwd_tc_kernel(const __grid_constant__ FinalTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;                                   // [4][8 KB]
  uint8_t* wimg = ring + kFtSlots * kRowBytes;            // [32 taps][64 B] swizzled, taps 9..31 zero
  float* psm = reinterpret_cast<float*>(wimg + 2048);     // [4][9][130]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(psm + 4 * 9 * 130);
  uint64_t* empty_bar = full_bar + kFtSlots;
  uint64_t* aready_bar = empty_bar + kFtSlots;
  uint64_t* dfull_bar = aready_bar + kFtSlots;   // [2]
  uint64_t* dempty_bar = dfull_bar + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < kFtSlots; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
        mbar_init(&aready_bar[s], 128);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&dfull_bar[a], 1);
        mbar_init(&dempty_bar[a], 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 64);
  }
  if (warp < 4) {
    // weight image: row t (tap), 32 channels; thread -> (tap, 8-channel chunk)
    const int t = threadIdx.x >> 2, c8 = threadIdx.x & 3;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = t < 9 ? P.w_last[(c8 * 8 + k) * 9 + t] : 0.0f;
    *reinterpret_cast<uint4*>(wimg + swz64(t, c8)) = pack8(v);
    for (int i = threadIdx.x; i < 4 * 9 * 130; i += 128) psm[i] = 0.0f;
    fence_proxy_async_smem();
  }
  if (warp == 4 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_y) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int H = P.H, RB = P.RB;

  if (warp == 4) {
    // ============================== producer ==============================
    if (lane == 0) {
      uint32_t n = 0;
      for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
        const int b = band / P.bands_per_img, h0 = (band - b * P.bands_per_img) * RB;
        for (int i = 0; i < RB + 2; ++i) {
          const int h = h0 - 1 + i;
          if (h < 0 || h >= H) continue;
          const uint32_t slot = n % kFtSlots, ph = (n / kFtSlots) & 1u;
          mbar_wait(&empty_bar[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[slot], kRowBytes);
          tma_load_4d(ring + slot * kRowBytes, &P.tmap_y, &full_bar[slot], 0, 0, h, b);
          ++n;
        }
      }
    }
  } else if (warp == 5) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
    const uint32_t ring_a = smem_u32(ring), w_a = smem_u32(wimg);
    uint32_t n = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int h0 = (band % P.bands_per_img) * RB;
      for (int i = 0; i < RB + 2; ++i) {
        const int h = h0 - 1 + i;
        if (h < 0 || h >= H) continue;
        const uint32_t slot = n % kFtSlots, ph = (n / kFtSlots) & 1u, db = n & 1u, dph = (n >> 1) & 1u;
        mbar_wait(&aready_bar[slot], ph);
        mbar_wait(&dempty_bar[db], dph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t ad = make_desc_sw(ring_a + slot * kRowBytes + k * 32, 512, 4);
            const uint64_t bd = make_desc_sw(w_a + k * 32, 512, 4);
            mma_bf16_ss(tmem_base + db * 32, ad, bd, idesc, k > 0 ? 1u : 0u);
          }
          mma_commit(&empty_bar[slot]);
          mma_commit(&dfull_bar[db]);
        }
        __syncwarp();
        ++n;
      }
    }
  } else {
    // =============================== workers ==============================
    const int p = threadIdx.x;  // pixel of the row
    float sc[32], sh[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      sc[c] = P.scale[c];
      sh[c] = fmaf(-P.mean[c], sc[c], P.beta[c]);
    }
    const float slope = P.slope, bias = P.b_last != nullptr ? P.b_last[0] : 0.0f;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t n_t = 0, n_e = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int b = band / P.bands_per_img, h0 = (band - b * P.bands_per_img) * RB;
      auto transform = [&](int i) {
        const int h = h0 - 1 + i;
        if (h < 0 || h >= H) return;
        const uint32_t slot = n_t % kFtSlots, ph = (n_t / kFtSlots) & 1u;
        mbar_wait(&full_bar[slot], ph);
        uint8_t* row = ring + slot * kRowBytes;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4* ptr = reinterpret_cast<uint4*>(row + swz64(p, c4));
          float v[8];
          unpack8(*ptr, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float u = fmaf(v[k], sc[c4 * 8 + k], sh[c4 * 8 + k]);
            v[k] = fmaxf(u, u * slope);  // LeakyReLU for 0 < slope < 1
          }
          *ptr = pack8(v);
        }
        fence_proxy_async_smem();
        mbar_arrive(&aready_bar[slot]);
        ++n_t;
      };
      transform(0);
      for (int i = 0; i < RB + 2; ++i) {
        if (i + 1 < RB + 2) transform(i + 1);
        const int h = h0 - 1 + i;
        float* ps = psm + (i & 3) * (9 * 130) + p + 1;
        if (h >= 0 && h < H) {
          const uint32_t db = n_e & 1u, dph = (n_e >> 1) & 1u;
          mbar_wait(&dfull_bar[db], dph);
          tc_fence_after();
          float v[32];
          tmem_ld_32x32(t_lane + db * 32, v);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&dempty_bar[db]);
          ++n_e;
#pragma unroll
          for (int t = 0; t < 9; ++t) ps[t * 130] = v[t];
        } else {
#pragma unroll
          for (int t = 0; t < 9; ++t) ps[t * 130] = 0.0f;
        }
        named_bar_sync(1, 128);
        if (i >= 2) {
          const int q = h - 1;  // image row whose 3x3 window is now complete
          float acc = bias;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* pr = psm + ((i - 2 + kh) & 3) * (9 * 130) + p;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) acc += pr[(kh * 3 + kw) * 130 + kw];
          }
          const long long o = ((long long)b * H + q) * 128 + p;
          const bf16 lg = __float2bfloat16_rn(acc);
          P.logits[o] = lg;
          if (P.probs != nullptr) P.probs[o] = 1.0f / (1.0f + expf(-__bfloat162float(lg)));
        }
      }
      named_bar_sync(1, 128);  // the next band reuses the P ring from slot 0
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// =============================================================================================
// backward, common pieces
//   dl ring   : 4 rows of dlogits as fp32 with zero borders, [4][130]
//   A1 tile   : dlwin[p][widx] = dl[h + a - 1][p + b - 1], widx = 3a + b  (9 of 32 bf16 per 64-byte row)
//   B1 image  : [c][widx] = w[c][8 - widx]      =>  da = A1 . B1^T  (one K = 16 MMA)
// =============================================================================================
__device__ __forceinline__ void load_dl_row(const bf16* __restrict__ dl, float* dlsm, int b, int h, int H, int p,
                                            int slot) {
  float v = 0.0f;
  if (h >= 0 && h < H) v = __bfloat162float(dl[((long long)b * H + h) * 128 + p]);
  dlsm[slot * 130 + p + 1] = v;
}

// writes the K = 16 part of row p of the A1 tile (chunks 0 and 1); returns the centre value
__device__ __forceinline__ float build_a1_row(const float* dlsm, uint8_t* a1, int i, int p) {
  float w[16];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) w[a * 3 + bb] = dlsm[((i + a) & 3) * 130 + p + bb];
#pragma unroll
  for (int k = 9; k < 16; ++k) w[k] = 0.0f;
  *reinterpret_cast<uint4*>(a1 + swz64(p, 0)) = pack8(w);
  *reinterpret_cast<uint4*>(a1 + swz64(p, 1)) = pack8(w + 8);
  return w[4];
}

struct BwdSmem {
  uint8_t* ring;    // [kFtSlots][8 KB] y rows (TMA), rewritten in place as act by pass 1
  uint8_t* a1;      // [2][8 KB]
  uint8_t* b1;      // [32][64 B]
  float* dlsm;      // [4][130]
  float* cst;       // per-channel constants, [5][32]
  float* red;       // [2][32] block reduction scratch (+ 8 for the bias sum)
  uint64_t* full_bar;   // [4] TMA -> workers
  uint64_t* empty_bar;  // [4] -> producer
  uint64_t* a1ready_bar;  // [2] workers -> MMA
  uint64_t* d1full_bar;   // [2] MMA -> workers
  uint64_t* d1empty_bar;  // [2] workers -> MMA
  uint64_t* a2ready_bar;  // [4] workers -> MMA (pass 1)
  uint64_t* done_bar;     // [1]
  uint32_t* tmem_ptr;
};

constexpr size_t kBwdSmemBytes = 1024 + kFtSlots * kRowBytes + 2 * kRowBytes + 8192 /* pad: garbage M groups stay in bounds */ +
                                 2048 + 4 * 130 * 4 + 5 * 32 * 4 + 72 * 4 + 8 * (3 * kFtSlots + 7) + 16;

__device__ __forceinline__ BwdSmem carve_bwd(uint8_t* smem) {
  BwdSmem s;
  s.ring = smem;
  s.a1 = s.ring + kFtSlots * kRowBytes;
  s.b1 = s.a1 + 2 * kRowBytes + 8192;
  s.dlsm = reinterpret_cast<float*>(s.b1 + 2048);
  s.cst = s.dlsm + 4 * 130;
  s.red = s.cst + 5 * 32;
  s.full_bar = reinterpret_cast<uint64_t*>(s.red + 72);
  s.empty_bar = s.full_bar + kFtSlots;
  s.a1ready_bar = s.empty_bar + kFtSlots;
  s.d1full_bar = s.a1ready_bar + 2;
  s.d1empty_bar = s.d1full_bar + 2;
  s.a2ready_bar = s.d1empty_bar + 2;
  s.done_bar = s.a2ready_bar + kFtSlots;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.done_bar + 1);
  return s;
}

// PASS: 1 = reduce (BN sums + weight / bias gradient of the final conv), 2 = apply (dy)
template <int PASS>
__global__ void __launch_bounds__(kFtThreads, 2)
final_bwd_tc_kernel(const __grid_constant__ FinalTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  BwdSmem S = carve_bwd(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = P.H, RB = P.RB;
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < kFtSlots; ++s) {
        mbar_init(&S.full_bar[s], 1);
        mbar_init(&S.empty_bar[s], PASS == 1 ? 1 : 128);  // pass 1: freed by the MMA commit; pass 2: by the workers
        mbar_init(&S.a2ready_bar[s], 128);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&S.a1ready_bar[a], 128);
        mbar_init(&S.d1full_bar[a], 1);
        mbar_init(&S.d1empty_bar[a], 128);
      }
      mbar_init(S.done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(S.tmem_ptr, 128);
  }
  if (warp < 4) {
    const int tid = threadIdx.x;
    {  // B1 image: row c, chunk j -> widx = 8j .. 8j+7
      const int c = tid >> 2, j = tid & 3;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int widx = j * 8 + k;
        v[k] = widx < 9 ? P.w_last[c * 9 + (8 - widx)] : 0.0f;
      }
      *reinterpret_cast<uint4*>(S.b1 + swz64(c, j)) = pack8(v);
    }
    for (int i = tid; i < 2 * kRowBytes / 16; i += 128) reinterpret_cast<uint4*>(S.a1)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 4 * 130; i += 128) S.dlsm[i] = 0.0f;
    if (tid < 32) {
      const int c = tid;
      const float sc = P.scale[c], mean = P.mean[c];
      S.cst[c] = sc;
      S.cst[32 + c] = fmaf(-mean, sc, P.beta[c]);
      if (PASS == 2) {
        const double inv_m = 1.0 / ((double)P.B * P.H * 128.0);
        const double ga = (double)P.gamma[c] * (double)P.rstd[c];
        const double c1 = P.training ? P.dstats[c] * inv_m : 0.0;
        const double c2 = P.training ? P.dstats[32 + c] * inv_m : 0.0;
        const double k2 = ga * c2 * (double)P.rstd[c];
        S.cst[64 + c] = (float)ga;
        S.cst[96 + c] = (float)(ga * c1 - k2 * (double)mean);   // k1
        S.cst[128 + c] = (float)k2;
      }
    }
    fence_proxy_async_smem();
  }
  if (warp == 4 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_y) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *S.tmem_ptr;

  if (warp == 4) {
    // ============================== producer ==============================
    if (lane == 0) {
      uint32_t n = 0;
      for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
        const int b = band / P.bands_per_img, h0 = (band - b * P.bands_per_img) * RB;
        for (int i = 0; i < RB; ++i, ++n) {
          const uint32_t slot = n % kFtSlots, ph = (n / kFtSlots) & 1u;
          mbar_wait(&S.empty_bar[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&S.full_bar[slot], kRowBytes);
          tma_load_4d(S.ring + slot * kRowBytes, &P.tmap_y, &S.full_bar[slot], 0, 0, h0 + i, b);
        }
      }
    }
  } else if (warp == 5) {
    // ============================= MMA issuer =============================
    const uint32_t idesc1 = make_idesc_bf16(128, 32, 0, 0);   // da: K-major operands
    const uint32_t idesc2 = make_idesc_bf16(128, 32, 1, 1);   // weight gradient: MN-major operands
    const uint32_t ring_a = smem_u32(S.ring), a1_a = smem_u32(S.a1), b1_a = smem_u32(S.b1);
    uint32_t total = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) total += (uint32_t)RB;
    auto issue_da = [&](uint32_t n) {
      const uint32_t buf = n & 1u, bph = (n >> 1) & 1u;
      mbar_wait(&S.a1ready_bar[buf], bph);
      mbar_wait(&S.d1empty_bar[buf], bph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        mma_bf16_ss(tmem_base + buf * 32, make_desc_sw(a1_a + buf * kRowBytes, 512, 4), make_desc_sw(b1_a, 512, 4),
                    idesc1, 0u);
        mma_commit(&S.d1full_bar[buf]);
      }
      __syncwarp();
    };
    if (total > 0) issue_da(0);
    for (uint32_t n = 0; n < total; ++n) {
      if (n + 1 < total) issue_da(n + 1);   // the next row's da runs under this row's elementwise math
      if (PASS == 1) {
        const uint32_t slot = n % kFtSlots, ph = (n / kFtSlots) & 1u, buf = n & 1u;
        mbar_wait(&S.a2ready_bar[slot], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {   // 16 pixels per MMA
            const uint64_t ad = make_desc_mn_tc(ring_a + slot * kRowBytes + ks * 1024, kRowBytes, 512, 4);
            const uint64_t bd = make_desc_mn_tc(a1_a + buf * kRowBytes + ks * 1024, kRowBytes, 512, 4);
            mma_bf16_ss(tmem_base + 64, ad, bd, idesc2, (n > 0 || ks > 0) ? 1u : 0u);
          }
          mma_commit(&S.empty_bar[slot]);
        }
        __syncwarp();
      }
    }
    if (PASS == 1) {
      if (elect_one()) mma_commit(S.done_bar);
      __syncwarp();
    }
  } else {
    // =============================== workers ==============================
    const int p = threadIdx.x;
    const float slope = P.slope;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    float acc1[32], acc2[32];   // pass 1: sum du, sum du*y ; pass 2: sum dy (acc1 only)
#pragma unroll
    for (int c = 0; c < 32; ++c) acc1[c] = acc2[c] = 0.0f;
    float sdl = 0.0f;
    uint32_t n = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int b = band / P.bands_per_img, h0 = (band - b * P.bands_per_img) * RB;
      // dl rows j = 0 .. RB+1  <->  image rows h0-1 .. h0+RB ; slot j & 3
      load_dl_row(P.dl, S.dlsm, b, h0 - 1, H, p, 0);
      load_dl_row(P.dl, S.dlsm, b, h0, H, p, 1);
      load_dl_row(P.dl, S.dlsm, b, h0 + 1, H, p, 2);
      named_bar_sync(1, 128);
      auto stage_a1 = [&](int i, uint32_t nn) {   // builds the A1 tile of band row i (sequence number nn)
        const uint32_t buf = nn & 1u, bph = (nn >> 1) & 1u;
        // A1[buf] was last read by the MMAs of row nn-2: da (observed through d1full two rows ago) and,
        // in pass 1, the weight-gradient MMAs, observed through the release of that row's ring slot
        if (PASS == 1 && nn >= 2) {
          const uint32_t m = nn - 2;
          mbar_wait(&S.empty_bar[m % kFtSlots], (m / kFtSlots) & 1u);
        }
        (void)bph;
        const float centre = build_a1_row(S.dlsm, S.a1 + buf * kRowBytes, i, p);
        if (PASS == 1) sdl += centre;
        fence_proxy_async_smem();
        mbar_arrive(&S.a1ready_bar[buf]);
      };
      stage_a1(0, n);
      for (int i = 0; i < RB; ++i, ++n) {
        // prefetch the next dl row and stage the next A1 tile so its MMA overlaps this row's math
        if (i + 1 < RB) {
          load_dl_row(P.dl, S.dlsm, b, h0 + i + 2, H, p, (i + 3) & 3);
          named_bar_sync(1, 128);
          stage_a1(i + 1, n + 1);
        }
        const uint32_t slot = n % kFtSlots, ph = (n / kFtSlots) & 1u, buf = n & 1u, bph = (n >> 1) & 1u;
        mbar_wait(&S.full_bar[slot], ph);
        mbar_wait(&S.d1full_bar[buf], bph);
        tc_fence_after();
        float da[32];
        tmem_ld_32x32(t_lane + buf * 32, da);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&S.d1empty_bar[buf]);
        uint8_t* row = S.ring + slot * kRowBytes;
        const long long opix = (((long long)b * H + h0 + i) * 128 + p) * 32;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4* ptr = reinterpret_cast<uint4*>(row + swz64(p, c4));
          float y[8], o[8];
          unpack8(*ptr, y);
          const float4 s0 = *reinterpret_cast<const float4*>(S.cst + c4 * 8), s1 = *reinterpret_cast<const float4*>(S.cst + c4 * 8 + 4);
          const float4 h0v = *reinterpret_cast<const float4*>(S.cst + 32 + c4 * 8), h1v = *reinterpret_cast<const float4*>(S.cst + 36 + c4 * 8);
          const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float shv[8] = {h0v.x, h0v.y, h0v.z, h0v.w, h1v.x, h1v.y, h1v.z, h1v.w};
          if (PASS == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = c4 * 8 + k;
              const float u = fmaf(y[k], scv[k], shv[k]);
              const float du = u > 0.0f ? da[c] : da[c] * slope;
              acc1[c] += du;
              acc2[c] = fmaf(du, y[k], acc2[c]);
              o[k] = fmaxf(u, u * slope);
            }
            *ptr = pack8(o);   // act, in place: MN-major A operand of the weight-gradient MMA
          } else {
            const float4 g0 = *reinterpret_cast<const float4*>(S.cst + 64 + c4 * 8), g1 = *reinterpret_cast<const float4*>(S.cst + 68 + c4 * 8);
            const float4 a0 = *reinterpret_cast<const float4*>(S.cst + 96 + c4 * 8), a1v = *reinterpret_cast<const float4*>(S.cst + 100 + c4 * 8);
            const float4 b0 = *reinterpret_cast<const float4*>(S.cst + 128 + c4 * 8), b1v = *reinterpret_cast<const float4*>(S.cst + 132 + c4 * 8);
            const float gav[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float k1v[8] = {a0.x, a0.y, a0.z, a0.w, a1v.x, a1v.y, a1v.z, a1v.w};
            const float k2v[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = c4 * 8 + k;
              const float u = fmaf(y[k], scv[k], shv[k]);
              const float du = u > 0.0f ? da[c] : da[c] * slope;
              const float r = fmaf(gav[k], du, -fmaf(k2v[k], y[k], k1v[k]));   // ga*(du - c1 - xhat*c2)
              acc1[c] += r;
              o[k] = r;
            }
            *reinterpret_cast<uint4*>(P.dy + opix + c4 * 8) = pack8(o);
          }
        }
        if (PASS == 1) {
          fence_proxy_async_smem();
          mbar_arrive(&S.a2ready_bar[slot]);
        } else {
          mbar_arrive(&S.empty_bar[slot]);
        }
      }
      named_bar_sync(1, 128);  // the next band restarts the dl ring
    }
    // ------------------------------ reductions ------------------------------
    warp_transpose_reduce32(acc1, lane);   // lane c now holds this warp's sum of channel c
    if (PASS == 1) warp_transpose_reduce32(acc2, lane);
    if (warp == 0) {
      S.red[lane] = 0.0f;
      S.red[32 + lane] = 0.0f;
      if (lane < 8) S.red[64 + lane] = 0.0f;
    }
    named_bar_sync(1, 128);
    atomicAdd(&S.red[lane], acc1[0]);
    if (PASS == 1) {
      atomicAdd(&S.red[32 + lane], acc2[0]);
      const float d = warp_sum(sdl);
      if (lane == 0) S.red[64 + warp] = d;
    }
    named_bar_sync(1, 128);
    if (warp == 0) {
      const int c = lane;
      if (PASS == 1) {
        const double s1 = (double)S.red[c], s2y = (double)S.red[32 + c];
        atomicAdd(&P.dstats[c], s1);
        atomicAdd(&P.dstats[32 + c], (double)P.rstd[c] * (s2y - (double)P.mean[c] * s1));   // sum du * xhat
        if (c == 0) atomicAdd(&P.dstats[64], (double)S.red[64] + (double)S.red[65] + (double)S.red[66] + (double)S.red[67]);
        // weight gradient of the final conv: D2 lane c, column widx -> dw[c][8 - widx]
        mbar_wait(S.done_bar, 0);
        tc_fence_after();
        float v[32];
        tmem_ld_32x32(tmem_base + 64, v);
        tmem_ld_wait();
#pragma unroll
        for (int widx = 0; widx < 9; ++widx) atomicAdd(&P.dw_last[c * 9 + (8 - widx)], v[widx]);
      } else {
        atomicAdd(&P.dbias[c], S.red[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// =============================================================================================
// host side
// =============================================================================================
static bool g_final_tc_enabled = true;

static bool final_tc_ok(int B, int H, int W, int C, int dtype) {
  return g_final_tc_enabled && dtype == VAE_BF16 && C == 32 && W == 128 && B > 0 && H >= 16 && H % 16 == 0 &&
         get_encode_fn() != nullptr;
}

static int final_tc_setup(FinalTcParams& P, const void* y, int B, int H, int W) {
  cuuint64_t gdim[4] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {64, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
  cuuint32_t box[4] = {32, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = get_encode_fn()(&P.tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(y), gdim, gstr, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VAE_REQUIRE(cr == CUDA_SUCCESS, "final_block_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  P.B = B; P.H = H; P.W = W;
  P.RB = H % 32 == 0 ? 32 : 16;
  P.bands_per_img = H / P.RB;
  P.total_bands = B * P.bands_per_img;
  return VAE_OK;
}

}  // namespace vae

using namespace vae;

extern "C" int vae_set_final_tc(int enable) {
  int old = g_final_tc_enabled ? 1 : 0;
  g_final_tc_enabled = enable != 0;
  return old;
}

/* 1 when the tcgen05 versions of the output-block kernels serve this shape (bf16, C = 32, W = 128) */
extern "C" int vae_final_tc_supported(int B, int H, int W, int C, int dtype) {
  return final_tc_ok(B, H, W, C, dtype) ? 1 : 0;
}

/* tcgen05 version of vae_final_conv_fwd (same contract, bf16 only) */
extern "C" int vae_final_conv_fwd_tc(const void* y, const float* mean, const float* scale, const float* beta,
                                     const float* w_last, const float* b_last, void* logits, float* probs, int B,
                                     int H, int W, int C, float slope, vae_stream_t stream) {
  VAE_REQUIRE(y && mean && scale && beta && w_last && logits, "final_conv_fwd_tc: null argument");
  VAE_REQUIRE(final_tc_ok(B, H, W, C, VAE_BF16), "final_conv_fwd_tc: shape not served");
  VAE_REQUIRE(slope > 0.0f && slope < 1.0f, "final_conv_fwd_tc: LeakyReLU slope must be in (0, 1)");
  FinalTcParams P{};
  int rc = final_tc_setup(P, y, B, H, W);
  if (rc != VAE_OK) return rc;
  P.mean = mean; P.scale = scale; P.beta = beta; P.w_last = w_last; P.b_last = b_last;
  P.logits = (bf16*)logits; P.probs = probs; P.slope = slope;
  const size_t smem = 1024 + kFtSlots * kRowBytes + 2048 + 4 * 9 * 130 * 4 + 8 * (3 * kFtSlots + 4) + 16;
  static bool attr = false;
  if (!attr) {
    VAE_CUDA(cudaFuncSetAttribute(final_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int grid = 2 * sm_count();
  if (grid > P.total_bands) grid = P.total_bands;
  final_fwd_tc_kernel<<<grid, kFtThreads, smem, as_stream(stream)>>>(P);
  return check_launch("final_fwd_tc");
}

static int final_bwd_tc_launch(int pass, FinalTcParams& P, cudaStream_t st) {
  static bool attr1 = false, attr2 = false;
  if (pass == 1 && !attr1) {
    VAE_CUDA(cudaFuncSetAttribute(final_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes));
    attr1 = true;
  }
  if (pass == 2 && !attr2) {
    VAE_CUDA(cudaFuncSetAttribute(final_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes));
    attr2 = true;
  }
  int grid = 2 * sm_count();
  if (grid > P.total_bands) grid = P.total_bands;
  if (pass == 1) final_bwd_tc_kernel<1><<<grid, kFtThreads, kBwdSmemBytes, st>>>(P);
  else final_bwd_tc_kernel<2><<<grid, kFtThreads, kBwdSmemBytes, st>>>(P);
  return check_launch(pass == 1 ? "final_bwd_reduce_tc" : "final_bwd_apply_tc");
}

/* tcgen05 version of vae_final_bwd_reduce (same contract, bf16 only) */
extern "C" int vae_final_bwd_reduce_tc(const void* dlogits, const void* y, const float* mean, const float* scale,
                                       const float* beta, const float* rstd, const float* w_last, double* dstats,
                                       float* dw_last, int B, int H, int W, int C, float slope, vae_stream_t stream) {
  VAE_REQUIRE(dlogits && y && mean && scale && beta && rstd && w_last && dstats && dw_last,
              "final_bwd_reduce_tc: null argument");
  VAE_REQUIRE(final_tc_ok(B, H, W, C, VAE_BF16), "final_bwd_reduce_tc: shape not served");
  VAE_REQUIRE(slope > 0.0f && slope < 1.0f, "final_bwd_reduce_tc: LeakyReLU slope must be in (0, 1)");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(cudaMemsetAsync(dstats, 0, sizeof(double) * (2 * C + 1), st));
  VAE_CUDA(cudaMemsetAsync(dw_last, 0, sizeof(float) * 9 * C, st));
  FinalTcParams P{};
  int rc = final_tc_setup(P, y, B, H, W);
  if (rc != VAE_OK) return rc;
  P.mean = mean; P.scale = scale; P.beta = beta; P.rstd = rstd; P.w_last = w_last;
  P.dl = (const bf16*)dlogits; P.dstats = dstats; P.dw_last = dw_last; P.slope = slope;
  return final_bwd_tc_launch(1, P, st);
}

/* tcgen05 version of vae_final_bwd_apply (same contract, bf16 only) */
extern "C" int vae_final_bwd_apply_tc(const void* dlogits, const void* y, const float* mean, const float* scale,
                                      const float* beta, const float* rstd, const float* gamma, const float* w_last,
                                      const double* dstats, void* dy, float* dgamma, float* dbeta, float* dbias,
                                      float* db_last, int B, int H, int W, int C, float slope, int training,
                                      vae_stream_t stream);

namespace vae {
__global__ void final_small_grads_kernel(const double* __restrict__ dstats, float* dgamma, float* dbeta, float* db_last,
                                         int C) {
  const int c = threadIdx.x;
  if (c < C) {
    dbeta[c] = (float)dstats[c];
    dgamma[c] = (float)dstats[C + c];
  }
  if (c == 0) db_last[0] = (float)dstats[2 * C];
}
}  // namespace vae

extern "C" int vae_final_bwd_apply_tc(const void* dlogits, const void* y, const float* mean, const float* scale,
                                      const float* beta, const float* rstd, const float* gamma, const float* w_last,
                                      const double* dstats, void* dy, float* dgamma, float* dbeta, float* dbias,
                                      float* db_last, int B, int H, int W, int C, float slope, int training,
                                      vae_stream_t stream) {
  VAE_REQUIRE(dlogits && y && mean && scale && beta && rstd && gamma && w_last && dstats && dy && dgamma && dbeta &&
                  dbias && db_last,
              "final_bwd_apply_tc: null argument");
  VAE_REQUIRE(final_tc_ok(B, H, W, C, VAE_BF16), "final_bwd_apply_tc: shape not served");
  VAE_REQUIRE(slope > 0.0f && slope < 1.0f, "final_bwd_apply_tc: LeakyReLU slope must be in (0, 1)");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C, st));
  final_small_grads_kernel<<<1, 64, 0, st>>>(dstats, dgamma, dbeta, db_last, C);
  FinalTcParams P{};
  int rc = final_tc_setup(P, y, B, H, W);
  if (rc != VAE_OK) return rc;
  P.mean = mean; P.scale = scale; P.beta = beta; P.rstd = rstd; P.gamma = gamma; P.w_last = w_last;
  P.dl = (const bf16*)dlogits; P.dstats = const_cast<double*>(dstats); P.dy = (bf16*)dy; P.dbias = dbias;
  P.slope = slope; P.training = training;
  return final_bwd_tc_launch(2, P, st);
}
