// tcgen05 / TMEM versions of the fused output-block kernels (final_block.cu) for the bench
// shape of the path: bf16, C = 32 channels, W = 128 pixels per image row.
//
// The C -> 1 convolution is a GEMM with N = 1, but written PER PIXEL it becomes tensor-core
// shaped:   P[p][t] = sum_c act[p][c] * w[t][c]          (M = pixels, N = 9 taps, K = C)
//           logit[q] = bias + sum_t P[q + d(t)][t]        (a 9-point shifted sum of P)
// and its data gradient is   da[p][c] = sum_t dl[p - d(t)] * w[t][c]   (M = pixels, N = C, K = 9 taps)
// so every element of the big tensor y is touched by CUDA cores only for the BatchNorm /
// LeakyReLU transform (5-8 instructions per element instead of ~26-30 in the SIMT kernels).
//
// Data flow per image row (128 pixels = one UMMA M tile):
//   TMA        y row [128 px][64 B] -> shared, 64B-swizzled (ring of rows, prefetched)
//   workers    two threads per pixel (16 channels each): read, BatchNorm + LeakyReLU, write bf16 act
//              back IN PLACE: the row buffer now IS the K-major (forward) / MN-major (weight
//              gradient) UMMA operand
//   MMA warp   one elected thread issues tcgen05.mma, accumulators in TMEM
//   workers    tcgen05.ld their pixel's row of the accumulator and finish (shifted sum / BN backward)
//
//   warps 0-7 : workers (256 threads)      warp 8 : TMA producer      warp 9 : MMA issuer
#include <cstdlib>

#include "tma_util.cuh"

namespace vae {

constexpr int kFtThreads = 320;
constexpr int kFtWorkers = 256;
constexpr int kFtSlots = 6;      // ring of raw y rows
constexpr int kRowBytes = 8192;  // 128 pixels x 32 channels x bf16
constexpr int kA1Bufs = 3;       // dlogits-window tiles in flight (backward)

struct FinalTcParams {
  CUtensorMap tmap_y;
  CUtensorMap tmap_dy;   // pass 2: same geometry as y, used for the TMA store of dy rows
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
  double* ds1;           // [32] sum du of this launch's channels        (dstats[0 .. C) of the caller, + channel offset)
  double* ds2;           // [32] sum du * xhat                            (dstats[C .. 2C), + channel offset)
  double* dsdl;          // sum of dlogits (dstats[2C]); null for every channel group but the first
  float* dw_last;        // [C][9]
  const float* partial_in;   // forward, channel groups after the first: fp32 partial logits [B][H][W] to continue from
  float* partial_out;        // forward, channel groups before the last: write the fp32 partial instead of logits / probs
  // a launch covers ONE group of 32 channels (tensor maps and per-channel pointers are offset by the caller) and tiles
  // of 128 pixels: tile cb of an image row starts at column cb * wstep - halo.  halo = 1 (forward, W > 128): tiles
  // overlap by two pixels and only their inner 126 produce logits; halo = 0: the tile is the whole row (W = 128) or,
  // in the backward kernels, an aligned block whose dlogits window is read from global memory across its edges.
  int B, H, W, RB, bands_per_img, total_bands, training, tpr, wstep, halo;
  float slope;
};

// ---- shared-memory access by 32-bit shared address (keeps the compiler off generic LD/ST) ----
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void load8f(uint32_t a, float* v) {  // 8 consecutive floats
  const float4 x = lds_f4(a), y = lds_f4(a + 16);
  v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
}

__device__ __forceinline__ uint32_t swz64(int row, int chunk) {  // byte offset inside a [rows][64 B] 64B-swizzled tile
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ void unpack8(const uint4& raw, float* v) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
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

// shared -> global tensor store (bulk async-group completion)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 32 lanes x 16 columns of fp32
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// =============================================================================================
// forward: logits = conv(lrelu(bn(y)), w_last) + b_last, probs = sigmoid(logits)
// =============================================================================================
constexpr size_t kFwdSmemBytes = 1024 + kFtSlots * kRowBytes + 2048 + 4 * 9 * 130 * 4 + 64 * 4 + 8 * (3 * kFtSlots + 4) + 16;

// WIDE = false: one 128-pixel tile per image row (W == 128): no column arithmetic, no edge masks
// PARTIAL = true: several groups of 32 channels accumulate the logits through an fp32 partial buffer
template <bool WIDE, bool PARTIAL>
__global__ void __launch_bounds__(kFtThreads, 3)
final_fwd_tc_kernel(const __grid_constant__ FinalTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* ring = smem;                                   // [slots][8 KB]
  uint8_t* wimg = ring + kFtSlots * kRowBytes;            // [32 taps][64 B] swizzled, taps 9..31 zero
  float* psm = reinterpret_cast<float*>(wimg + 2048);     // [4][9][130]
  float* cst = psm + 4 * 9 * 130;                         // scale[32], shift[32]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(cst + 64);
  uint64_t* empty_bar = full_bar + kFtSlots;
  uint64_t* aready_bar = empty_bar + kFtSlots;
  uint64_t* dfull_bar = aready_bar + kFtSlots;   // [2]
  uint64_t* dempty_bar = dfull_bar + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 9) {
    if (lane == 0) {
      for (int s = 0; s < kFtSlots; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
        mbar_init(&aready_bar[s], kFtWorkers);
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
  if (warp < 8) {
    const int tid = threadIdx.x;
    if (tid < 128) {  // weight image: row t (tap), 32 channels; thread -> (tap, 8-channel chunk)
      const int t = tid >> 2, c8 = tid & 3;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = t < 9 ? P.w_last[(c8 * 8 + k) * 9 + t] : 0.0f;
      *reinterpret_cast<uint4*>(wimg + swz64(t, c8)) = pack8(v);
    } else if (tid < 160) {
      const int c = tid - 128;
      const float sc = P.scale[c];
      cst[c] = sc;
      cst[32 + c] = fmaf(-P.mean[c], sc, P.beta[c]);
    }
    for (int i = tid; i < 4 * 9 * 130; i += kFtWorkers) psm[i] = 0.0f;
    fence_proxy_async_smem();
  }
  if (warp == 8 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_y) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int H = P.H, RB = P.RB;

  if (warp == 8) {
    // ============================== producer ==============================
    if (lane == 0) {
      uint32_t n = 0;
      for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
        const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
        for (int i = 0; i < RB + 2; ++i) {
          const int h = h0 - 1 + i;
          if (h < 0 || h >= H) continue;
          const uint32_t slot = n % kFtSlots, ph = (n / kFtSlots) & 1u;
          mbar_wait(&empty_bar[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[slot], kRowBytes);
          tma_load_4d(ring + slot * kRowBytes, &P.tmap_y, &full_bar[slot], 0, w0, h, b);
          ++n;
        }
      }
    }
  } else if (warp == 9) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
    const uint32_t ring_a = smem_u32(ring), w_a = smem_u32(wimg);
    uint32_t n = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int h0 = ((WIDE ? band / P.tpr : band) % P.bands_per_img) * RB;
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
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;  // pixel of the row, channel half
    const float slope = P.slope, bias = (P.b_last != nullptr && (!PARTIAL || P.partial_in == nullptr)) ? P.b_last[0] : 0.0f;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t ring_a = smem_u32(ring), psm_a = smem_u32(psm), cst_a = smem_u32(cst);
    const uint32_t off0 = swz64(p, 2 * half), off1 = swz64(p, 2 * half + 1);
    uint32_t n_t = 0, n_e = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
      // a halo tile (W > 128) may reach one pixel past either edge of the image: TMA zero-fills y there, but the conv's
      // zero padding applies to the ACTIVATION, so those pixels are forced to 0 instead of lrelu(shift)
      const bool col_ok = !WIDE || (unsigned)(w0 + p) < (unsigned)P.W;
      const bool out_ok = !WIDE || (col_ok && (P.halo == 0 || (p >= 1 && p <= 126)));
      auto transform = [&](int i) {
        const int h = h0 - 1 + i;
        if (h < 0 || h >= H) return;
        const uint32_t slot = n_t % kFtSlots, ph = (n_t / kFtSlots) & 1u;
        mbar_wait(&full_bar[slot], ph);
        const uint32_t row = ring_a + slot * kRowBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t a = row + (j == 0 ? off0 : off1);
          float v[8], sc[8], sh[8];
          unpack8(lds128(a), v);
          load8f(cst_a + (2 * half + j) * 32, sc);
          load8f(cst_a + 128 + (2 * half + j) * 32, sh);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float u = fmaf(v[k], sc[k], sh[k]);
            v[k] = col_ok ? fmaxf(u, u * slope) : 0.0f;  // LeakyReLU for 0 < slope < 1; zero padding outside the image
          }
          sts128(a, pack8(v));
        }
        fence_proxy_async_smem();
        mbar_arrive(&aready_bar[slot]);
        ++n_t;
      };
      transform(0);
      for (int i = 0; i < RB + 2; ++i) {
        if (i + 1 < RB + 2) transform(i + 1);
        const int h = h0 - 1 + i;
        const bool valid = h >= 0 && h < H;
        if (half == 0) {  // P row -> ring of fp32 rows [4][9][130]
          const uint32_t ps = psm_a + (uint32_t)(((i & 3) * (9 * 130) + p + 1) * 4);
          if (valid) {
            const uint32_t db = n_e & 1u, dph = (n_e >> 1) & 1u;
            mbar_wait(&dfull_bar[db], dph);
            tc_fence_after();
            float v[16];
            tmem_ld_32x16(t_lane + db * 32, v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&dempty_bar[db]);
#pragma unroll
            for (int t = 0; t < 9; ++t) sts_f(ps + t * 520, v[t]);
          } else {
#pragma unroll
            for (int t = 0; t < 9; ++t) sts_f(ps + t * 520, 0.0f);
          }
        }
        if (valid) ++n_e;
        named_bar_sync(1, kFtWorkers);
        if (half == 1 && i >= 2) {
          const int q = h - 1;  // image row whose 3x3 window is now complete
          float acc = bias;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t pr = psm_a + (uint32_t)(((((i - 2 + kh) & 3) * (9 * 130)) + p) * 4);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) acc += lds_f(pr + ((kh * 3 + kw) * 130 + kw) * 4);
          }
          if (out_ok) {
            const long long o = ((long long)b * H + q) * (WIDE ? P.W : 128) + w0 + p;
            if (PARTIAL && P.partial_in != nullptr) acc += P.partial_in[o];   // the other channel groups' share (bias is in the first)
            if (PARTIAL && P.partial_out != nullptr) {
              P.partial_out[o] = acc;
            } else {
              const bf16 lg = __float2bfloat16_rn(acc);
              P.logits[o] = lg;
              if (P.probs != nullptr) P.probs[o] = 1.0f / (1.0f + expf(-__bfloat162float(lg)));
            }
          }
        }
      }
      named_bar_sync(1, kFtWorkers);  // the next band reuses the P ring from slot 0
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
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
struct BwdSmem {
  uint8_t* ring;    // [kFtSlots][8 KB] y rows (TMA), rewritten in place as act by pass 1
  uint8_t* a1;      // [kA1Bufs][8 KB]
  uint8_t* b1;      // [32][64 B]
  float* dlsm;      // [4][130]
  float* cst;       // per-channel constants, [5][32]
  float* red;       // [2][32] block reduction scratch (+ 8 for the bias sum)
  int* coord;       // [slots][3] (h, b, w0) of the row held by a slot (pass 2 store)
  uint64_t* full_bar;     // TMA -> workers
  uint64_t* empty_bar;    // -> producer
  uint64_t* a1ready_bar;  // [2] workers -> MMA
  uint64_t* d1full_bar;   // [2] MMA -> workers
  uint64_t* d1empty_bar;  // [2] workers -> MMA
  uint64_t* a2ready_bar;  // workers -> MMA (pass 1)
  uint64_t* done_bar;     // [1]
  uint32_t* tmem_ptr;
};

constexpr size_t bwd_smem_bytes(int slots) {
  return 1024 + (size_t)slots * kRowBytes + kA1Bufs * kRowBytes + 8192 /* pad: garbage M groups stay in bounds */ +
         2048 + 4 * 130 * 4 + 5 * 32 * 4 + 72 * 4 + 4 * kFtSlots * 4 + 8 * (3 * kFtSlots + 8) + 16;
}
constexpr size_t kBwdSmemBytes = bwd_smem_bytes(kFtSlots);

__device__ __forceinline__ BwdSmem carve_bwd(uint8_t* smem, int slots) {   // barrier arrays stay sized for kFtSlots
  BwdSmem s;
  s.ring = smem;
  s.a1 = s.ring + slots * kRowBytes;
  s.b1 = s.a1 + kA1Bufs * kRowBytes + 8192;
  s.dlsm = reinterpret_cast<float*>(s.b1 + 2048);
  s.cst = s.dlsm + 4 * 130;
  s.red = s.cst + 5 * 32;
  s.coord = reinterpret_cast<int*>(s.red + 72);
  s.full_bar = reinterpret_cast<uint64_t*>(s.coord + 4 * kFtSlots);
  s.empty_bar = s.full_bar + kFtSlots;
  s.a1ready_bar = s.empty_bar + kFtSlots;
  s.d1full_bar = s.a1ready_bar + kA1Bufs;
  s.d1empty_bar = s.d1full_bar + 2;
  s.a2ready_bar = s.d1empty_bar + 2;
  s.done_bar = s.a2ready_bar + kFtSlots;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.done_bar + 1);
  return s;
}

// PASS: 1 = reduce (BN sums + weight / bias gradient of the final conv), 2 = apply (dy)
// SLOTS y rows in the ring, CTAS resident CTAs per SM (pass 2 is bound by issue latency, not by bytes in
// flight: three CTAs with four slots each keep the same 12 rows per SM in flight with 50 % more warps)
template <int PASS, int SLOTS, int CTAS, bool WIDE>
__global__ void __launch_bounds__(kFtThreads, CTAS)
final_bwd_tc_kernel(const __grid_constant__ FinalTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // shared-space base at C = 32 (apply pass 232 -> 209 us); at C = 64 (WIDE) that form measured slower (590 -> 611 us)
  uint8_t* smem;
  if constexpr (WIDE) smem = align1024(smem_raw); else smem = align1024_shared(smem_raw);
  BwdSmem S = carve_bwd(smem, SLOTS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = P.H, RB = P.RB;
  if (warp == 9) {
    if (lane == 0) {
      for (int s = 0; s < SLOTS; ++s) {
        mbar_init(&S.full_bar[s], 1);
        mbar_init(&S.empty_bar[s], PASS == 1 ? 1 : kFtWorkers);  // pass 1: freed by the MMA commit; pass 2: by the workers
        mbar_init(&S.a2ready_bar[s], kFtWorkers);
      }
      for (int a = 0; a < kA1Bufs; ++a) mbar_init(&S.a1ready_bar[a], 128);
      for (int a = 0; a < 2; ++a) {
        mbar_init(&S.d1full_bar[a], 1);
        mbar_init(&S.d1empty_bar[a], kFtWorkers);
      }
      mbar_init(S.done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(S.tmem_ptr, 128);
  }
  if (warp < 8) {
    const int tid = threadIdx.x;
    if (tid < 128) {  // B1 image: row c, chunk j -> widx = 8j .. 8j+7
      const int c = tid >> 2, j = tid & 3;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int widx = j * 8 + k;
        v[k] = widx < 9 ? P.w_last[c * 9 + (8 - widx)] : 0.0f;
      }
      *reinterpret_cast<uint4*>(S.b1 + swz64(c, j)) = pack8(v);
    } else if (tid < 160) {
      const int c = tid - 128;
      const float sc = P.scale[c], mean = P.mean[c];
      S.cst[c] = sc;
      S.cst[32 + c] = fmaf(-mean, sc, P.beta[c]);
      if (PASS == 2) {
        const double inv_m = 1.0 / ((double)P.B * P.H * (double)P.W);
        const double ga = (double)P.gamma[c] * (double)P.rstd[c];
        const double c1 = P.training ? P.ds1[c] * inv_m : 0.0;
        const double c2 = P.training ? P.ds2[c] * inv_m : 0.0;
        const double k2 = ga * c2 * (double)P.rstd[c];
        S.cst[64 + c] = (float)ga;
        S.cst[96 + c] = (float)(ga * c1 - k2 * (double)mean);   // k1
        S.cst[128 + c] = (float)k2;
      }
    }
    for (int i = tid; i < kA1Bufs * kRowBytes / 16; i += kFtWorkers) reinterpret_cast<uint4*>(S.a1)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 4 * 130; i += kFtWorkers) S.dlsm[i] = 0.0f;
    if (tid < 72) S.red[tid] = 0.0f;
    fence_proxy_async_smem();
  }
  if (warp == 8 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_y) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *S.tmem_ptr;

  if (warp == 8) {
    // ============================== producer ==============================
    if (lane == 0) {
      uint32_t n = 0;
      if (PASS == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_dy) : "memory");
      // pass 2: the workers overwrite a y row IN PLACE with dy; this thread then sends the row to
      // global memory with one TMA tensor store (coalesced by the copy engine) and only afterwards
      // recycles the slot for a new load
      auto store_row = [&](uint32_t m) {
        const uint32_t slot = m % SLOTS;
        mbar_wait(&S.a2ready_bar[slot], (m / SLOTS) & 1u);
        tma_store_4d(&P.tmap_dy, S.ring + slot * kRowBytes, 0, S.coord[3 * slot + 2], S.coord[3 * slot], S.coord[3 * slot + 1]);
        bulk_commit_group();
      };
      for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
        const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
        for (int i = 0; i < RB; ++i, ++n) {
          const uint32_t slot = n % SLOTS, ph = (n / SLOTS) & 1u;
          if (PASS == 1) {
            mbar_wait(&S.empty_bar[slot], ph ^ 1u);
          } else if (n >= SLOTS) {
            store_row(n - SLOTS);
            bulk_wait_group_read0();  // the store has finished reading the slot
          }
          if (PASS == 2) {
            S.coord[3 * slot] = h0 + i;
            S.coord[3 * slot + 1] = b;
            S.coord[3 * slot + 2] = w0;
          }
          mbar_arrive_expect_tx(&S.full_bar[slot], kRowBytes);
          tma_load_4d(S.ring + slot * kRowBytes, &P.tmap_y, &S.full_bar[slot], 0, w0, h0 + i, b);
        }
      }
      if (PASS == 2) {
        for (uint32_t m = n > SLOTS ? n - SLOTS : 0; m < n; ++m) store_row(m);
        bulk_wait_group0();
      }
    }
  } else if (warp == 9) {
    // ============================= MMA issuer =============================
    const uint32_t idesc1 = make_idesc_bf16(128, 32, 0, 0);   // da: K-major operands
    const uint32_t idesc2 = make_idesc_bf16(128, 32, 1, 1);   // weight gradient: MN-major operands
    const uint32_t ring_a = smem_u32(S.ring), a1_a = smem_u32(S.a1), b1_a = smem_u32(S.b1);
    uint32_t total = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) total += (uint32_t)RB;
    auto issue_da = [&](uint32_t n) {
      const uint32_t buf = n & 1u, bph = (n >> 1) & 1u;          // TMEM accumulator of da
      const uint32_t ab = n % kA1Bufs, aph = (n / kA1Bufs) & 1u;  // A1 tile
      mbar_wait(&S.a1ready_bar[ab], aph);
      mbar_wait(&S.d1empty_bar[buf], bph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        mma_bf16_ss(tmem_base + buf * 32, make_desc_sw(a1_a + ab * kRowBytes, 512, 4), make_desc_sw(b1_a, 512, 4),
                    idesc1, 0u);
        mma_commit(&S.d1full_bar[buf]);
      }
      __syncwarp();
    };
    if (total > 0) issue_da(0);
    for (uint32_t n = 0; n < total; ++n) {
      if (n + 1 < total) issue_da(n + 1);   // the next row's da runs under this row's elementwise math
      if (PASS == 1) {
        const uint32_t slot = n % SLOTS, ph = (n / SLOTS) & 1u, ab = n % kA1Bufs;
        mbar_wait(&S.a2ready_bar[slot], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {   // 16 pixels per MMA
            const uint64_t ad = make_desc_mn_tc(ring_a + slot * kRowBytes + ks * 1024, kRowBytes, 512, 4);
            const uint64_t bd = make_desc_mn_tc(a1_a + ab * kRowBytes + ks * 1024, kRowBytes, 512, 4);
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
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    const float slope = P.slope;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 16);
    const uint32_t ring_a = smem_u32(S.ring), a1_a = smem_u32(S.a1), dl_a = smem_u32(S.dlsm), cst_a = smem_u32(S.cst);
    const uint32_t off0 = swz64(p, 2 * half), off1 = swz64(p, 2 * half + 1);
    const uint32_t a1off0 = swz64(p, 0), a1off1 = swz64(p, 1);
    float acc[32];   // pass 1: [0,16) sum du, [16,32) sum du*y for this thread's 16 channels
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
    float sdl = 0.0f;
    uint32_t n = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
      // dlogits window of a pixel: rows j = 0 .. RB+1 of the band <-> image rows h0-1 .. h0+RB, columns
      // p-1, p, p+1.  Each half-0 thread keeps a sliding 3 x 3 window in registers and loads the three
      // values of the next row one iteration ahead (L1-resident after the first touch), so there is no
      // shared-memory exchange and no block barrier on the per-row path.
      auto fetch3 = [&](int h, float* v) {
        if (h >= 0 && h < H) {
          const bf16* row = P.dl + ((long long)b * H + h) * (WIDE ? P.W : 128) + w0;
          v[0] = w0 + p > 0 ? __bfloat162float(row[p - 1]) : 0.0f;
          v[1] = __bfloat162float(row[p]);
          v[2] = w0 + p < (WIDE ? P.W : 128) - 1 ? __bfloat162float(row[p + 1]) : 0.0f;
        } else {
          v[0] = v[1] = v[2] = 0.0f;
        }
      };
      // the look-ahead row stays RAW (bf16 bits) until it enters the window one iteration later: a
      // conversion next to the load would make the thread wait for the load at once (ncu: 15 % of all
      // stall samples of the kernel sat on that conversion)
      auto fetch3_raw = [&](int h, unsigned short* v) {
        v[0] = v[1] = v[2] = 0;
        if (h >= 0 && h < H) {
          const unsigned short* row = reinterpret_cast<const unsigned short*>(P.dl) + ((long long)b * H + h) * (WIDE ? P.W : 128) + w0;
          if (w0 + p > 0) v[0] = __ldg(row + p - 1);
          v[1] = __ldg(row + p);
          if (w0 + p < (WIDE ? P.W : 128) - 1) v[2] = __ldg(row + p + 1);
        }
      };
      float r0[3], r1[3], r2[3];
      unsigned short pf[3];   // 16-bit type: widening (and with it any use of the load) happens at the shift below
      if (half == 0) {
        fetch3(h0 - 1, r0);
        fetch3(h0, r1);
        fetch3(h0 + 1, r2);
        fetch3_raw(h0 + 2, pf);
      }
      auto stage_a1 = [&](uint32_t nn) {   // half 0: builds the A1 tile (sequence number nn) from the window r0..r2
        const uint32_t buf = nn % kA1Bufs;
        // A1[buf] was last read by the MMAs of row nn - kA1Bufs: da (observed through d1full by this thread,
        // in program order) and, in pass 1, the weight-gradient MMAs, observed through the release of that
        // row's ring slot.  With three tiles the wait is for work issued two rows ago, off the critical path.
        if (PASS == 1 && nn >= (uint32_t)kA1Bufs) {
          const uint32_t m = nn - kA1Bufs;
          mbar_wait(&S.empty_bar[m % SLOTS], (m / SLOTS) & 1u);
        }
        float w[16];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          w[k] = r0[k];
          w[3 + k] = r1[k];
          w[6 + k] = r2[k];
        }
#pragma unroll
        for (int k = 9; k < 16; ++k) w[k] = 0.0f;
        const uint32_t base = a1_a + buf * kRowBytes;
        sts128(base + a1off0, pack8(w));
        sts128(base + a1off1, pack8(w + 8));
        if (PASS == 1) sdl += w[4];
        fence_proxy_async_smem();
        mbar_arrive(&S.a1ready_bar[buf]);
      };
      if (half == 0) stage_a1(n);
      for (int i = 0; i < RB; ++i, ++n) {
        // stage the next row's A1 tile first so that its MMA overlaps this row's elementwise math
        if (i + 1 < RB && half == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            r0[k] = r1[k];
            r1[k] = r2[k];
            r2[k] = __uint_as_float((uint32_t)pf[k] << 16);
          }
          fetch3_raw(h0 + i + 3, pf);
          stage_a1(n + 1);
        }
        const uint32_t slot = n % SLOTS, ph = (n / SLOTS) & 1u, buf = n & 1u, bph = (n >> 1) & 1u;
        mbar_wait(&S.full_bar[slot], ph);
        mbar_wait(&S.d1full_bar[buf], bph);
        tc_fence_after();
        float da[16];
        tmem_ld_32x16(t_lane + buf * 32, da);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&S.d1empty_bar[buf]);
        const uint32_t row = ring_a + slot * kRowBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t a = row + (j == 0 ? off0 : off1);
          const uint32_t cb = cst_a + (uint32_t)((2 * half + j) * 32);
          float y[8], o[8], scv[8], shv[8];
          unpack8(lds128(a), y);
          load8f(cb, scv);
          load8f(cb + 128, shv);
          if (PASS == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = j * 8 + k;
              const float u = fmaf(y[k], scv[k], shv[k]);
              const float du = u > 0.0f ? da[c] : da[c] * slope;
              acc[c] += du;
              acc[16 + c] = fmaf(du, y[k], acc[16 + c]);
              o[k] = fmaxf(u, u * slope);
            }
            sts128(a, pack8(o));   // act, in place: MN-major A operand of the weight-gradient MMA
          } else {
            float gav[8], k1v[8], k2v[8];
            load8f(cb + 256, gav);
            load8f(cb + 384, k1v);
            load8f(cb + 512, k2v);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = j * 8 + k;
              const float u = fmaf(y[k], scv[k], shv[k]);
              const float du = u > 0.0f ? da[c] : da[c] * slope;
              o[k] = fmaf(gav[k], du, -fmaf(k2v[k], y[k], k1v[k]));   // ga*(du - c1 - xhat*c2)
            }
            sts128(a, pack8(o));   // dy, in place; the producer thread stores the row with TMA
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&S.a2ready_bar[slot]);   // pass 1: act operand ready; pass 2: dy row ready to store
      }
    }
    // ------------------------------ reductions (pass 1) ------------------------------
    if (PASS == 1) {
      warp_transpose_reduce32(acc, lane);   // lane l: l < 16 -> sum du of channel 16*half + l ; else sum du*y of channel 16*half + l - 16
      atomicAdd(&S.red[(lane >> 4) * 32 + half * 16 + (lane & 15)], acc[0]);
      const float d = warp_sum(sdl);
      if (lane == 0 && half == 0) S.red[64 + warp] = d;
      named_bar_sync(1, kFtWorkers);
      if (warp == 0) {
        const int c = lane;
        const double s1 = (double)S.red[c], s2y = (double)S.red[32 + c];
        atomicAdd(&P.ds1[c], s1);
        atomicAdd(&P.ds2[c], (double)P.rstd[c] * (s2y - (double)P.mean[c] * s1));   // sum du * xhat
        if (c == 0 && P.dsdl != nullptr) atomicAdd(P.dsdl, (double)S.red[64] + (double)S.red[65] + (double)S.red[66] + (double)S.red[67]);
        // weight gradient of the final conv: D2 lane c, column widx -> dw[c][8 - widx]
        mbar_wait(S.done_bar, 0);
        tc_fence_after();
        float v[16];
        tmem_ld_32x16(tmem_base + 64, v);
        tmem_ld_wait();
#pragma unroll
        for (int widx = 0; widx < 9; ++widx) atomicAdd(&P.dw_last[c * 9 + (8 - widx)], v[widx]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}


// ---- LeakyReLU branch of u = fmaf(y, scale, shift) as a bf16 COMPARISON ----------------------------------
// y is a stored bf16 value, so the predicate g(y) = [fmaf(y, scale, shift) > 0] (the forward kernel's own
// expression; one rounding, which preserves the sign) is a step function of y: for scale > 0 there is a
// largest bf16 value T with g(T) false and g(y) <=> y > T; for scale < 0 the same holds for -y.  T is found
// from -shift/scale and then corrected by walking the bf16 grid with g itself, so the comparison agrees with
// the forward's branch for EVERY bf16 input.  One HSET2 per channel pair replaces unpack + FMA + compare.
__device__ __forceinline__ uint32_t bf16_bits_step(uint32_t b, bool up) {   // neighbour on the bf16 grid (finite range)
  const bool neg = (b & 0x8000u) != 0;
  const uint32_t mag = b & 0x7FFFu;
  if (mag == 0) return up ? 0x0001u : 0x8001u;
  const bool grow = (up != neg);   // moving away from zero
  return (uint32_t)((neg ? 0x8000u : 0u) | (grow ? mag + 1 : mag - 1));
}
__device__ __forceinline__ float bf16_bits_to_f(uint32_t b) { return __uint_as_float(b << 16); }

// -> {threshold bits (bf16), sign-flip mask (0x8000 when scale < 0)}
__device__ __forceinline__ void lrelu_threshold(float scale, float shift, uint32_t* thr, uint32_t* sgn) {
  if (scale == 0.0f || !(scale == scale) || !(shift == shift)) {   // constant branch (NaN parameters: "false")
    *thr = (shift > 0.0f) ? 0xFF80u : 0x7F80u;   // y > -inf: always; y > +inf: never
    *sgn = 0u;
    return;
  }
  const float s = fabsf(scale);
  *sgn = scale < 0.0f ? 0x8000u : 0u;
  auto g = [&](uint32_t bits) { return fmaf(bf16_bits_to_f(bits), s, shift) > 0.0f; };   // predicate on y' = +-y
  const uint32_t kMaxPos = 0x7F7Fu, kMaxNeg = 0xFF7Fu;
  if (g(kMaxNeg)) { *thr = 0xFF80u; return; }    // true on the whole finite range
  if (!g(kMaxPos)) { *thr = 0x7F80u; return; }   // false on the whole finite range
  const double t = -(double)shift / (double)s;   // finite here, inside the bf16 range
  uint32_t T = __float_as_uint((float)t) >> 16;  // truncation: a nearby grid point; corrected below
  if ((T & 0x7F80u) == 0x7F80u) T = (T & 0x8000u) | 0x7F7Fu;
  for (int it = 0; it < 300 && g(T); ++it) T = bf16_bits_step(T, false);                        // ensure g(T) false
  for (int it = 0; it < 300 && T != kMaxPos && !g(bf16_bits_step(T, true)); ++it) T = bf16_bits_step(T, true);
  *thr = T;
}

// =============================================================================================
// backward pass 1, all-GEMM form.  With b = [u > 0] (u = scale*y + shift), q = slope + (1-slope)*b:
//     du  = q * da,   da[p][c] = sum_w dlwin[p][w] * wf[c][w]        (wf[c][w] = w_last[c][8 - w])
//   sum_p du[p][c]        = sum_w wf[c][w] * G1[c][w],    G1 = q^T . dlwin
//   sum_p du[p][c]*y[p][c]= sum_w wf[c][w] * G2[c][w],    G2 = (q*y)^T . dlwin
//   dW_last[c][8 - w]     = sum_p lrelu(u)[p][c] * dlwin[p][w] = scale[c]*G2[c][w] + shift[c]*G1[c][w]
// and q^T.dlwin = slope * colsum(dlwin) + (1-slope) * b^T.dlwin,  (q*y)^T.dlwin = slope * y^T.dlwin + (1-slope) * (b*y)^T.dlwin.
// So the whole pass is ONE accumulating GEMM per image row,  [y | b | b*y]^T (96 x 128 pixels) . dlwin (128 x 9),
// with exact bf16 operands (y is the stored tensor, b is 0/1, b*y is y or 0, dlwin is the stored dlogits):
// the workers only derive b and b*y (6 instructions per element, no per-thread accumulators, no da).
// Shared memory: three rings at a common stride L -- [y rows (TMA)] [b rows] [b*y rows] -- so that one
// MN-major descriptor with LBO = L addresses the three 32-channel groups of a row as M = 96 (+32 unused):
// 8 MMAs per row (an M = 128, K = 16 MMA costs ~58 cycles whatever N <= 64 is: scripts/micro/mma_cost.cu).
// =============================================================================================
constexpr int kP1Slots = 3;
constexpr int kP1Ring = kP1Slots * kRowBytes;   // L
constexpr size_t kP1SmemBytes = 1024 + 3 * kP1Ring + kA1Bufs * kRowBytes + 8192 + 64 * 4 + (3 * 32 * 9 + 16) * 4 +
                                8 * (3 * kP1Slots + kA1Bufs + 1) + 16 + 32 * 4;

template <bool WIDE>
__global__ void __launch_bounds__(kFtThreads, 2)
final_bwd_reduce_gemm_kernel(const __grid_constant__ FinalTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* yring = smem;                         // [slots][8 KB]   (b ring at +L, b*y ring at +2L)
  uint8_t* a1 = smem + 3 * kP1Ring;              // [kA1Bufs][8 KB] dlwin tiles; + 8 KB pad (unused M group of the last slot)
  float* cst = reinterpret_cast<float*>(a1 + kA1Bufs * kRowBytes + 8192);   // scale[32], shift[32]
  float* gsm = cst + 64;                         // [3][32][9] accumulators, [16] column sums of dlwin
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gsm + 3 * 32 * 9 + 16);
  uint64_t* empty_bar = full_bar + kP1Slots;
  uint64_t* opready_bar = empty_bar + kP1Slots;
  uint64_t* a1ready_bar = opready_bar + kP1Slots;
  uint64_t* done_bar = a1ready_bar + kA1Bufs;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  uint16_t* thr_s = reinterpret_cast<uint16_t*>(tmem_ptr + 4);   // [32] branch thresholds, [32] sign-flip masks
  uint16_t* sgn_s = thr_s + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = P.H, RB = P.RB;
  if (warp == 9) {
    if (lane == 0) {
      for (int s = 0; s < kP1Slots; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
        mbar_init(&opready_bar[s], kFtWorkers);
      }
      for (int a = 0; a < kA1Bufs; ++a) mbar_init(&a1ready_bar[a], 128);
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 32);
  }
  if (warp < 8) {
    const int tid = threadIdx.x;
    if (tid < 32) {
      const float sc = P.scale[tid];
      const float sh = fmaf(-P.mean[tid], sc, P.beta[tid]);
      cst[tid] = sc;
      cst[32 + tid] = sh;
      uint32_t thr, sgn;
      lrelu_threshold(sc, sh, &thr, &sgn);
      thr_s[tid] = (uint16_t)thr;
      sgn_s[tid] = (uint16_t)sgn;
    }
    for (int i = tid; i < kA1Bufs * kRowBytes / 16; i += kFtWorkers) reinterpret_cast<uint4*>(a1)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 16; i += kFtWorkers) gsm[3 * 32 * 9 + i] = 0.0f;
    fence_proxy_async_smem();
  }
  if (warp == 8 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_y) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  uint32_t total = 0;
  for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) total += (uint32_t)RB;

  if (warp == 8) {
    // ============================== producer ==============================
    if (lane == 0) {
      uint32_t n = 0;
      for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
        const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
        for (int i = 0; i < RB; ++i, ++n) {
          const uint32_t slot = n % kP1Slots, ph = (n / kP1Slots) & 1u;
          mbar_wait(&empty_bar[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[slot], kRowBytes);
          tma_load_4d(yring + slot * kRowBytes, &P.tmap_y, &full_bar[slot], 0, w0, h0 + i, b);
        }
      }
    }
  } else if (warp == 9) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(128, 32, 1, 1);   // both operands MN-major
    const uint32_t y_a = smem_u32(yring), a1_a = smem_u32(a1);
    for (uint32_t n = 0; n < total; ++n) {
      const uint32_t slot = n % kP1Slots, ph = (n / kP1Slots) & 1u, ab = n % kA1Bufs, aph = (n / kA1Bufs) & 1u;
      mbar_wait(&a1ready_bar[ab], aph);
      mbar_wait(&opready_bar[slot], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 pixels per MMA
          const uint64_t ad = make_desc_mn_tc(y_a + slot * kRowBytes + ks * 1024, kP1Ring, 512, 4);
          const uint64_t bd = make_desc_mn_tc(a1_a + ab * kRowBytes + ks * 1024, kRowBytes, 512, 4);
          mma_bf16_ss(tmem_base, ad, bd, idesc, (n > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(&empty_bar[slot]);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(done_bar);
    __syncwarp();
  } else {
    // =============================== workers ==============================
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    const uint32_t y_a = smem_u32(yring), a1_a = smem_u32(a1);
    const uint32_t off0 = swz64(p, 2 * half), off1 = swz64(p, 2 * half + 1);
    const uint32_t a1off0 = swz64(p, 0), a1off1 = swz64(p, 1);
    // this thread's 16 channels (two 16-byte chunks) as 8 bf16x2 words: thresholds and sign-flip masks
    uint32_t thr2[8], sgn2[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = half * 16 + 2 * w;
      thr2[w] = (uint32_t)thr_s[c] | ((uint32_t)thr_s[c + 1] << 16);
      sgn2[w] = (uint32_t)sgn_s[c] | ((uint32_t)sgn_s[c + 1] << 16);
    }
    float colsum[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) colsum[k] = 0.0f;
    uint32_t n = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
      auto fetch3 = [&](int h, float* v) {
        if (h >= 0 && h < H) {
          const bf16* row = P.dl + ((long long)b * H + h) * (WIDE ? P.W : 128) + w0;
          v[0] = w0 + p > 0 ? __bfloat162float(row[p - 1]) : 0.0f;
          v[1] = __bfloat162float(row[p]);
          v[2] = w0 + p < (WIDE ? P.W : 128) - 1 ? __bfloat162float(row[p + 1]) : 0.0f;
        } else {
          v[0] = v[1] = v[2] = 0.0f;
        }
      };
      // the look-ahead row stays RAW (bf16 bits) until it enters the window one iteration later: a
      // conversion next to the load would make the thread wait for the load at once (ncu: 15 % of all
      // stall samples of the kernel sat on that conversion)
      auto fetch3_raw = [&](int h, unsigned short* v) {
        v[0] = v[1] = v[2] = 0;
        if (h >= 0 && h < H) {
          const unsigned short* row = reinterpret_cast<const unsigned short*>(P.dl) + ((long long)b * H + h) * (WIDE ? P.W : 128) + w0;
          if (w0 + p > 0) v[0] = __ldg(row + p - 1);
          v[1] = __ldg(row + p);
          if (w0 + p < (WIDE ? P.W : 128) - 1) v[2] = __ldg(row + p + 1);
        }
      };
      float r0[3], r1[3], r2[3];
      unsigned short pf[3];   // 16-bit type: widening (and with it any use of the load) happens at the shift below
      if (half == 0) {
        fetch3(h0 - 1, r0);
        fetch3(h0, r1);
        fetch3(h0 + 1, r2);
        fetch3_raw(h0 + 2, pf);
      }
      auto stage_a1 = [&](uint32_t nn) {   // half 0: dlwin tile of row nn from the register window
        const uint32_t buf = nn % kA1Bufs;
        if (nn >= (uint32_t)kA1Bufs) {     // the MMAs of row nn - kA1Bufs have finished reading this tile
          const uint32_t m = nn - kA1Bufs;
          mbar_wait(&empty_bar[m % kP1Slots], (m / kP1Slots) & 1u);
        }
        float w[16];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          w[k] = r0[k];
          w[3 + k] = r1[k];
          w[6 + k] = r2[k];
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) colsum[k] += w[k];
#pragma unroll
        for (int k = 9; k < 16; ++k) w[k] = 0.0f;
        const uint32_t base = a1_a + buf * kRowBytes;
        sts128(base + a1off0, pack8(w));
        sts128(base + a1off1, pack8(w + 8));
        // no fence / arrive here: the caller publishes the tile together with the row operands (ONE proxy fence per
        // row and thread; the fences and barrier round trips, not the arithmetic, set the per-row latency of a warp)
      };
      if (half == 0) {
        stage_a1(n);
        fence_proxy_async_smem();
        mbar_arrive(&a1ready_bar[n % kA1Bufs]);
      }
      for (int i = 0; i < RB; ++i, ++n) {
        if (i + 1 < RB && half == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            r0[k] = r1[k];
            r1[k] = r2[k];
            r2[k] = __uint_as_float((uint32_t)pf[k] << 16);
          }
          fetch3_raw(h0 + i + 3, pf);
          stage_a1(n + 1);
        }
        const uint32_t slot = n % kP1Slots, ph = (n / kP1Slots) & 1u;
        mbar_wait(&full_bar[slot], ph);
        const uint32_t row = y_a + slot * kRowBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t a = row + (j == 0 ? off0 : off1);
          const uint4 raw = lds128(a);
          const uint32_t yw[4] = {raw.x, raw.y, raw.z, raw.w};
          uint32_t bw[4], byw[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            // b = [u > 0] as bf16 1.0 / 0.0 for two channels at once; b * y is exact (y or +-0)
            const uint32_t yy = yw[w] ^ sgn2[j * 4 + w];
            asm("set.gt.bf16x2.bf16x2 %0, %1, %2;" : "=r"(bw[w]) : "r"(yy), "r"(thr2[j * 4 + w]));
            asm("mul.bf16x2 %0, %1, %2;" : "=r"(byw[w]) : "r"(bw[w]), "r"(yw[w]));
          }
          sts128(a + kP1Ring, make_uint4(bw[0], bw[1], bw[2], bw[3]));
          sts128(a + 2 * kP1Ring, make_uint4(byw[0], byw[1], byw[2], byw[3]));
        }
        fence_proxy_async_smem();
        mbar_arrive(&opready_bar[slot]);
        if (i + 1 < RB && half == 0) mbar_arrive(&a1ready_bar[(n + 1) % kA1Bufs]);   // staged above, fenced just now
      }
    }
    // ------------------------------ reductions ------------------------------
    if (half == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float v = warp_sum(colsum[k]);
        if (lane == 0) atomicAdd(&gsm[3 * 32 * 9 + k], v);
      }
    }
    if (warp < 3 && total > 0) {   // TMEM lanes 32w..32w+31 = rows of [y | b | b*y]^T . dlwin
      mbar_wait(done_bar, 0);
      tc_fence_after();
      float v[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 9; ++k) gsm[(warp * 32 + lane) * 9 + k] = v[k];
    }
    named_bar_sync(1, kFtWorkers);
    if (warp == 0 && total > 0) {
      const int c = lane;
      const float sl = P.slope, om = 1.0f - P.slope, sc = cst[c], sh = cst[32 + c];
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float g1 = sl * gsm[3 * 32 * 9 + k] + om * gsm[(32 + c) * 9 + k];          // q^T . dlwin
        const float g2 = sl * gsm[c * 9 + k] + om * gsm[(64 + c) * 9 + k];               // (q*y)^T . dlwin
        const float wf = P.w_last[c * 9 + (8 - k)];
        s1 += (double)wf * (double)g1;
        s2 += (double)wf * (double)g2;
        atomicAdd(&P.dw_last[c * 9 + (8 - k)], fmaf(sc, g2, sh * g1));
      }
      atomicAdd(&P.ds1[c], s1);
      atomicAdd(&P.ds2[c], (double)P.rstd[c] * (s2 - (double)P.mean[c] * s1));   // sum du * xhat
      if (c == 0 && P.dsdl != nullptr) atomicAdd(P.dsdl, (double)gsm[3 * 32 * 9 + 4]);                // centre tap: sum of dlogits
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// =============================================================================================
// backward pass 1, deep-ring form of the all-GEMM kernel above.  The three-slot ring of [y | b | b*y] rows kept only
// 48 KB of y in flight per SM (two CTAs) and a slot's turn-around (TMA -> workers -> 8 MMAs -> commit -> reload) is
// ~2.5 us, so the kernel sat at 2.8 TB/s.  Here the LANDING ring (raw y rows, 14 x 8 KB, freed by the workers as soon
// as they have read a row) is decoupled from the OPERAND ring (3 x [y | b | b*y], freed by the MMA commit): one CTA
// per SM keeps 112 KB of y in flight; the workers copy y next to b and b*y (a third 16-byte store per chunk).
// MEASURED (B200, B = 512): 288 us against 197 us for the two-CTA kernel above -- bytes in flight are not the limit;
// the eight worker warps of ONE CTA need ~1 200 cycles per row (fences, barrier round trips, dependent shared-memory
// accesses at ~0.1 IPC per warp), and one CTA per SM halves the number of rows being worked on.  Kept as an opt-in
// experiment (VAE_FINAL_DEEP=1); DESIGN.md section 6.1.
// =============================================================================================
constexpr int kDLSlots = 14;                    // landing ring: raw y rows in flight (TMA)
constexpr int kDOSlots = 3;                     // operand ring: [y | b | b*y] rows the MMAs read (== kA1Bufs)
constexpr int kDOSlot = 3 * kRowBytes;          // one operand slot: three 32-channel groups, 8 KB apart (LBO)
constexpr size_t kDeepSmemBytes = 1024 + kDLSlots * kRowBytes + kDOSlots * kDOSlot + 8192 + kA1Bufs * kRowBytes + 64 * 4 +
                                  (3 * 32 * 9 + 16) * 4 + 8 * (2 * kDLSlots + 2 * kDOSlots + kA1Bufs + 1) + 16 + 32 * 4;
static_assert(kA1Bufs == kDOSlots, "the dlwin tile of a row is recycled together with its operand slot");

__global__ void __launch_bounds__(kFtThreads, 1)
final_bwd_reduce_deep_kernel(const __grid_constant__ FinalTcParams P) {
  constexpr bool WIDE = true;   // opt-in experiment: always the general form
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* yring = smem;                         // [kDLSlots][8 KB] landing ring (TMA)
  uint8_t* oring = smem + kDLSlots * kRowBytes;  // [kDOSlots][y | b | b*y] operand ring; + 8 KB pad (unused M group of the last slot)
  uint8_t* a1 = oring + kDOSlots * kDOSlot + 8192;   // [kA1Bufs][8 KB] dlwin tiles
  float* cst = reinterpret_cast<float*>(a1 + kA1Bufs * kRowBytes);   // scale[32], shift[32]
  float* gsm = cst + 64;                         // [3][32][9] accumulators, [16] column sums of dlwin
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gsm + 3 * 32 * 9 + 16);
  uint64_t* lempty_bar = full_bar + kDLSlots;      // workers -> producer: the landing slot has been read
  uint64_t* opready_bar = lempty_bar + kDLSlots;   // workers -> MMA: operand slot written
  uint64_t* empty_bar = opready_bar + kDOSlots;    // MMA commit -> workers: operand slot (and its dlwin tile) free
  uint64_t* a1ready_bar = empty_bar + kDOSlots;
  uint64_t* done_bar = a1ready_bar + kA1Bufs;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  uint16_t* thr_s = reinterpret_cast<uint16_t*>(tmem_ptr + 4);   // [32] branch thresholds, [32] sign-flip masks
  uint16_t* sgn_s = thr_s + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = P.H, RB = P.RB;
  if (warp == 9) {
    if (lane == 0) {
      for (int s = 0; s < kDLSlots; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&lempty_bar[s], kFtWorkers);
      }
      for (int s = 0; s < kDOSlots; ++s) {
        mbar_init(&empty_bar[s], 1);
        mbar_init(&opready_bar[s], kFtWorkers);
      }
      for (int a = 0; a < kA1Bufs; ++a) mbar_init(&a1ready_bar[a], 128);
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 32);
  }
  if (warp < 8) {
    const int tid = threadIdx.x;
    if (tid < 32) {
      const float sc = P.scale[tid];
      const float sh = fmaf(-P.mean[tid], sc, P.beta[tid]);
      cst[tid] = sc;
      cst[32 + tid] = sh;
      uint32_t thr, sgn;
      lrelu_threshold(sc, sh, &thr, &sgn);
      thr_s[tid] = (uint16_t)thr;
      sgn_s[tid] = (uint16_t)sgn;
    }
    for (int i = tid; i < kA1Bufs * kRowBytes / 16; i += kFtWorkers) reinterpret_cast<uint4*>(a1)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < (kDOSlots * kDOSlot + 8192) / 16; i += kFtWorkers) reinterpret_cast<uint4*>(oring)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 16; i += kFtWorkers) gsm[3 * 32 * 9 + i] = 0.0f;
    fence_proxy_async_smem();
  }
  if (warp == 8 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_y) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  uint32_t total = 0;
  for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) total += (uint32_t)RB;

  if (warp == 8) {
    // ============================== producer ==============================
    if (lane == 0) {
      uint32_t n = 0;
      for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
        const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
        for (int i = 0; i < RB; ++i, ++n) {
          const uint32_t slot = n % kDLSlots, ph = (n / kDLSlots) & 1u;
          mbar_wait(&lempty_bar[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[slot], kRowBytes);
          tma_load_4d(yring + slot * kRowBytes, &P.tmap_y, &full_bar[slot], 0, w0, h0 + i, b);
        }
      }
    }
  } else if (warp == 9) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(128, 32, 1, 1);   // both operands MN-major
    const uint32_t y_a = smem_u32(oring), a1_a = smem_u32(a1);
    for (uint32_t n = 0; n < total; ++n) {
      const uint32_t slot = n % kDOSlots, ph = (n / kDOSlots) & 1u, ab = n % kA1Bufs, aph = (n / kA1Bufs) & 1u;
      mbar_wait(&a1ready_bar[ab], aph);
      mbar_wait(&opready_bar[slot], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 pixels per MMA
          const uint64_t ad = make_desc_mn_tc(y_a + slot * kDOSlot + ks * 1024, kRowBytes, 512, 4);
          const uint64_t bd = make_desc_mn_tc(a1_a + ab * kRowBytes + ks * 1024, kRowBytes, 512, 4);
          mma_bf16_ss(tmem_base, ad, bd, idesc, (n > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(&empty_bar[slot]);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(done_bar);
    __syncwarp();
  } else {
    // =============================== workers ==============================
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    const uint32_t y_a = smem_u32(yring), o_a = smem_u32(oring), a1_a = smem_u32(a1);
    const uint32_t off0 = swz64(p, 2 * half), off1 = swz64(p, 2 * half + 1);
    const uint32_t a1off0 = swz64(p, 0), a1off1 = swz64(p, 1);
    // this thread's 16 channels (two 16-byte chunks) as 8 bf16x2 words: thresholds and sign-flip masks
    uint32_t thr2[8], sgn2[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = half * 16 + 2 * w;
      thr2[w] = (uint32_t)thr_s[c] | ((uint32_t)thr_s[c + 1] << 16);
      sgn2[w] = (uint32_t)sgn_s[c] | ((uint32_t)sgn_s[c + 1] << 16);
    }
    float colsum[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) colsum[k] = 0.0f;
    uint32_t n = 0;
    for (int band = blockIdx.x; band < P.total_bands; band += gridDim.x) {
      const int cb = WIDE ? band % P.tpr : 0, rbnd = WIDE ? band / P.tpr : band;
        const int b = rbnd / P.bands_per_img, h0 = (rbnd - b * P.bands_per_img) * RB, w0 = WIDE ? cb * P.wstep - P.halo : 0;
        (void)w0;
      auto fetch3 = [&](int h, float* v) {
        if (h >= 0 && h < H) {
          const bf16* row = P.dl + ((long long)b * H + h) * (WIDE ? P.W : 128) + w0;
          v[0] = w0 + p > 0 ? __bfloat162float(row[p - 1]) : 0.0f;
          v[1] = __bfloat162float(row[p]);
          v[2] = w0 + p < (WIDE ? P.W : 128) - 1 ? __bfloat162float(row[p + 1]) : 0.0f;
        } else {
          v[0] = v[1] = v[2] = 0.0f;
        }
      };
      // the look-ahead row stays RAW (bf16 bits) until it enters the window one iteration later: a
      // conversion next to the load would make the thread wait for the load at once (ncu: 15 % of all
      // stall samples of the kernel sat on that conversion)
      auto fetch3_raw = [&](int h, unsigned short* v) {
        v[0] = v[1] = v[2] = 0;
        if (h >= 0 && h < H) {
          const unsigned short* row = reinterpret_cast<const unsigned short*>(P.dl) + ((long long)b * H + h) * (WIDE ? P.W : 128) + w0;
          if (w0 + p > 0) v[0] = __ldg(row + p - 1);
          v[1] = __ldg(row + p);
          if (w0 + p < (WIDE ? P.W : 128) - 1) v[2] = __ldg(row + p + 1);
        }
      };
      float r0[3], r1[3], r2[3];
      unsigned short pf[3];   // 16-bit type: widening (and with it any use of the load) happens at the shift below
      if (half == 0) {
        fetch3(h0 - 1, r0);
        fetch3(h0, r1);
        fetch3(h0 + 1, r2);
        fetch3_raw(h0 + 2, pf);
      }
      auto stage_a1 = [&](uint32_t nn) {   // half 0: dlwin tile of row nn from the register window
        const uint32_t buf = nn % kA1Bufs;
        if (nn >= (uint32_t)kA1Bufs) {     // the MMAs of row nn - kA1Bufs have finished reading this tile
          const uint32_t m = nn - kA1Bufs;
          mbar_wait(&empty_bar[m % kDOSlots], (m / kDOSlots) & 1u);
        }
        float w[16];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          w[k] = r0[k];
          w[3 + k] = r1[k];
          w[6 + k] = r2[k];
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) colsum[k] += w[k];
#pragma unroll
        for (int k = 9; k < 16; ++k) w[k] = 0.0f;
        const uint32_t base = a1_a + buf * kRowBytes;
        sts128(base + a1off0, pack8(w));
        sts128(base + a1off1, pack8(w + 8));
        fence_proxy_async_smem();
        mbar_arrive(&a1ready_bar[buf]);
      };
      if (half == 0) stage_a1(n);
      for (int i = 0; i < RB; ++i, ++n) {
        if (i + 1 < RB && half == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            r0[k] = r1[k];
            r1[k] = r2[k];
            r2[k] = __uint_as_float((uint32_t)pf[k] << 16);
          }
          fetch3_raw(h0 + i + 3, pf);
          stage_a1(n + 1);
        }
        const uint32_t slot = n % kDLSlots, ph = (n / kDLSlots) & 1u, os = n % kDOSlots;
        mbar_wait(&full_bar[slot], ph);
        if (n >= (uint32_t)kDOSlots) {   // the MMAs of row n - kDOSlots have finished reading this operand slot
          const uint32_t m = n - kDOSlots;
          mbar_wait(&empty_bar[os], (m / kDOSlots) & 1u);
        }
        const uint32_t row = y_a + slot * kRowBytes, orow = o_a + os * kDOSlot;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t off = j == 0 ? off0 : off1;
          const uint4 raw = lds128(row + off);
          const uint32_t a = orow + off;
          const uint32_t yw[4] = {raw.x, raw.y, raw.z, raw.w};
          uint32_t bw[4], byw[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            // b = [u > 0] as bf16 1.0 / 0.0 for two channels at once; b * y is exact (y or +-0)
            const uint32_t yy = yw[w] ^ sgn2[j * 4 + w];
            asm("set.gt.bf16x2.bf16x2 %0, %1, %2;" : "=r"(bw[w]) : "r"(yy), "r"(thr2[j * 4 + w]));
            asm("mul.bf16x2 %0, %1, %2;" : "=r"(byw[w]) : "r"(bw[w]), "r"(yw[w]));
          }
          sts128(a, raw);
          sts128(a + kRowBytes, make_uint4(bw[0], bw[1], bw[2], bw[3]));
          sts128(a + 2 * kRowBytes, make_uint4(byw[0], byw[1], byw[2], byw[3]));
        }
        mbar_arrive(&lempty_bar[slot]);   // the landing slot may be refilled
        fence_proxy_async_smem();
        mbar_arrive(&opready_bar[os]);
      }
    }
    // ------------------------------ reductions ------------------------------
    if (half == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float v = warp_sum(colsum[k]);
        if (lane == 0) atomicAdd(&gsm[3 * 32 * 9 + k], v);
      }
    }
    if (warp < 3 && total > 0) {   // TMEM lanes 32w..32w+31 = rows of [y | b | b*y]^T . dlwin
      mbar_wait(done_bar, 0);
      tc_fence_after();
      float v[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 9; ++k) gsm[(warp * 32 + lane) * 9 + k] = v[k];
    }
    named_bar_sync(1, kFtWorkers);
    if (warp == 0 && total > 0) {
      const int c = lane;
      const float sl = P.slope, om = 1.0f - P.slope, sc = cst[c], sh = cst[32 + c];
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float g1 = sl * gsm[3 * 32 * 9 + k] + om * gsm[(32 + c) * 9 + k];          // q^T . dlwin
        const float g2 = sl * gsm[c * 9 + k] + om * gsm[(64 + c) * 9 + k];               // (q*y)^T . dlwin
        const float wf = P.w_last[c * 9 + (8 - k)];
        s1 += (double)wf * (double)g1;
        s2 += (double)wf * (double)g2;
        atomicAdd(&P.dw_last[c * 9 + (8 - k)], fmaf(sc, g2, sh * g1));
      }
      atomicAdd(&P.ds1[c], s1);
      atomicAdd(&P.ds2[c], (double)P.rstd[c] * (s2 - (double)P.mean[c] * s1));   // sum du * xhat
      if (c == 0 && P.dsdl != nullptr) atomicAdd(P.dsdl, (double)gsm[3 * 32 * 9 + 4]);                // centre tap: sum of dlogits
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// =============================================================================================
// host side
// =============================================================================================
static bool g_final_tc_enabled = true;
static bool g_final_gemm_reduce = true;   // pass 1 as one accumulating GEMM (final_bwd_reduce_gemm_kernel)
static int g_final_deep_ring = -1;        // pass 1 with the decoupled landing / operand rings (env VAE_FINAL_DEEP=1; measured slower: default 0)

static bool final_tc_ok(int B, int H, int W, int C, int dtype) {
  // bf16; channel groups of 32 (C = 32: one launch per kernel, C = 64: two); rows of whole 128-pixel tiles
  return g_final_tc_enabled && dtype == VAE_BF16 && (C == 32 || C == 64) && W >= 128 && W % 128 == 0 && B > 0 &&
         H >= 16 && H % 16 == 0 && (long long)B * H * W < (1LL << 31) && get_encode_fn() != nullptr;
}

// tensor map + tiling of ONE channel group `grp` (32 channels) of an NHWC tensor with C channels per pixel.
// halo: forward kernel with W > 128 (overlapping tiles of 126 output pixels); else aligned 128-pixel tiles.
static int final_tc_setup(FinalTcParams& P, const void* y, int B, int H, int W, int C, int grp, bool halo) {
  const cuuint64_t pix = (cuuint64_t)C * 2;   // bytes per pixel
  cuuint64_t gdim[4] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {pix, (cuuint64_t)W * pix, (cuuint64_t)H * W * pix};
  cuuint32_t box[4] = {32, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* base = const_cast<uint8_t*>(static_cast<const uint8_t*>(y)) + (size_t)grp * 64;
  CUresult cr = get_encode_fn()(&P.tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VAE_REQUIRE(cr == CUDA_SUCCESS, "final_block_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  P.B = B; P.H = H; P.W = W;
  P.RB = H % 32 == 0 ? 32 : 16;
  P.bands_per_img = H / P.RB;
  if (halo) {
    P.halo = 1; P.wstep = 126; P.tpr = (W + 125) / 126;
  } else {
    P.halo = 0; P.wstep = 128; P.tpr = W / 128;
  }
  P.total_bands = B * P.bands_per_img * P.tpr;
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
  const int groups = C / 32;
  VAE_REQUIRE(groups == 1 || probs != nullptr, "final_conv_fwd_tc: C = 64 needs the fp32 probs buffer (it carries the partial sums)");
  const size_t smem = kFwdSmemBytes;
  static bool attr = false;
  if (!attr) {
    VAE_CUDA(cudaFuncSetAttribute(final_fwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VAE_CUDA(cudaFuncSetAttribute(final_fwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VAE_CUDA(cudaFuncSetAttribute(final_fwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VAE_CUDA(cudaFuncSetAttribute(final_fwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  for (int g = 0; g < groups; ++g) {   // one launch per group of 32 channels; partial logits travel in `probs` (fp32)
    FinalTcParams P{};
    int rc = final_tc_setup(P, y, B, H, W, C, g, W > 128);
    if (rc != VAE_OK) return rc;
    P.mean = mean + 32 * g; P.scale = scale + 32 * g; P.beta = beta + 32 * g; P.w_last = w_last + 32 * 9 * g; P.b_last = b_last;
    P.logits = (bf16*)logits; P.probs = probs; P.slope = slope;
    P.partial_in = g > 0 ? probs : nullptr;
    P.partial_out = g + 1 < groups ? probs : nullptr;
    int grid = 3 * sm_count();
    if (grid > P.total_bands) grid = P.total_bands;
    if (W > 128 && groups > 1) final_fwd_tc_kernel<true, true><<<grid, kFtThreads, smem, as_stream(stream)>>>(P);
    else if (W > 128) final_fwd_tc_kernel<true, false><<<grid, kFtThreads, smem, as_stream(stream)>>>(P);
    else if (groups > 1) final_fwd_tc_kernel<false, true><<<grid, kFtThreads, smem, as_stream(stream)>>>(P);
    else final_fwd_tc_kernel<false, false><<<grid, kFtThreads, smem, as_stream(stream)>>>(P);
    rc = check_launch("final_fwd_tc");
    if (rc != VAE_OK) return rc;
  }
  return VAE_OK;
}

template <bool WIDE>
static int final_bwd_tc_launch_t(int pass, FinalTcParams& P, cudaStream_t st) {
  // pass 2: <4 slots, 3 CTAs per SM> measured 254 us against 297 us for <6 slots, 2 CTAs> on the B = 512 tensor
  constexpr int kP2Slots = 4, kP2Ctas = 3;
  static bool attr1 = false, attr2 = false;
  if (pass == 1 && !attr1) {
    VAE_CUDA(cudaFuncSetAttribute(final_bwd_tc_kernel<1, kFtSlots, 2, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes));
    attr1 = true;
  }
  if (pass == 2 && !attr2) {
    VAE_CUDA(cudaFuncSetAttribute(final_bwd_tc_kernel<2, kP2Slots, kP2Ctas, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)bwd_smem_bytes(kP2Slots)));
    attr2 = true;
  }
  int grid = (pass == 2 ? kP2Ctas : 2) * sm_count();
  if (grid > P.total_bands) grid = P.total_bands;
  if (pass == 1 && g_final_gemm_reduce) {
    if (g_final_deep_ring < 0) {
      const char* e = getenv("VAE_FINAL_DEEP");
      g_final_deep_ring = e ? atoi(e) : 0;
    }
    if (g_final_deep_ring) {
      static bool attr4 = false;
      if (!attr4) {
        VAE_CUDA(cudaFuncSetAttribute(final_bwd_reduce_deep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDeepSmemBytes));
        attr4 = true;
      }
      int g1 = sm_count();
      if (g1 > P.total_bands) g1 = P.total_bands;
      final_bwd_reduce_deep_kernel<<<g1, kFtThreads, kDeepSmemBytes, st>>>(P);
      return check_launch("final_bwd_reduce_deep");
    }
    static bool attr3 = false;
    if (!attr3) {
      VAE_CUDA(cudaFuncSetAttribute(final_bwd_reduce_gemm_kernel<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kP1SmemBytes));
      attr3 = true;
    }
    final_bwd_reduce_gemm_kernel<WIDE><<<grid, kFtThreads, kP1SmemBytes, st>>>(P);
    return check_launch("final_bwd_reduce_gemm");
  }
  if (pass == 1) final_bwd_tc_kernel<1, kFtSlots, 2, WIDE><<<grid, kFtThreads, kBwdSmemBytes, st>>>(P);
  else final_bwd_tc_kernel<2, kP2Slots, kP2Ctas, WIDE><<<grid, kFtThreads, bwd_smem_bytes(kP2Slots), st>>>(P);
  return check_launch(pass == 1 ? "final_bwd_reduce_tc" : "final_bwd_apply_tc");
}

static int final_bwd_tc_launch(int pass, FinalTcParams& P, cudaStream_t st) {
  return P.W > 128 ? final_bwd_tc_launch_t<true>(pass, P, st) : final_bwd_tc_launch_t<false>(pass, P, st);
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
  VAE_CUDA(zero_scratch(dstats, sizeof(double) * (2 * C + 1), st));
  VAE_CUDA(zero_scratch(dw_last, sizeof(float) * 9 * C, st));
  for (int g = 0; g < C / 32; ++g) {   // the reductions are per channel: one launch per group of 32
    FinalTcParams P{};
    int rc = final_tc_setup(P, y, B, H, W, C, g, false);
    if (rc != VAE_OK) return rc;
    P.mean = mean + 32 * g; P.scale = scale + 32 * g; P.beta = beta + 32 * g; P.rstd = rstd + 32 * g; P.w_last = w_last + 32 * 9 * g;
    P.dl = (const bf16*)dlogits; P.dw_last = dw_last + 32 * 9 * g; P.slope = slope;
    P.ds1 = dstats + 32 * g; P.ds2 = dstats + C + 32 * g; P.dsdl = g == 0 ? dstats + 2 * C : nullptr;
    rc = final_bwd_tc_launch(1, P, st);
    if (rc != VAE_OK) return rc;
  }
  return VAE_OK;
}

/* tcgen05 version of vae_final_bwd_apply (same contract, bf16 only) */
extern "C" int vae_final_bwd_apply_tc(const void* dlogits, const void* y, const float* mean, const float* scale,
                                      const float* beta, const float* rstd, const float* gamma, const float* w_last,
                                      const double* dstats, void* dy, float* dgamma, float* dbeta, float* dbias,
                                      float* db_last, int B, int H, int W, int C, float slope, int training,
                                      vae_stream_t stream);

namespace vae {
// dbias = sum over pixels of dy: identically zero through a train-mode BatchNorm (dy has its batch mean
// removed; the reference's value is pure rounding noise), gamma*rstd*sum(du) in eval mode.
__global__ void final_small_grads_kernel(const double* __restrict__ dstats, const float* __restrict__ gamma,
                                         const float* __restrict__ rstd, float* dgamma, float* dbeta, float* dbias,
                                         float* db_last, int C, int training) {
  const int c = threadIdx.x;
  if (c < C) {
    dbeta[c] = (float)dstats[c];
    dgamma[c] = (float)dstats[C + c];
    dbias[c] = training ? 0.0f : (float)((double)gamma[c] * (double)rstd[c] * dstats[c]);
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
  final_small_grads_kernel<<<1, 64, 0, st>>>(dstats, gamma, rstd, dgamma, dbeta, dbias, db_last, C, training);
  for (int g = 0; g < C / 32; ++g) {
    FinalTcParams P{};
    int rc = final_tc_setup(P, y, B, H, W, C, g, false);
    if (rc != VAE_OK) return rc;
    P.mean = mean + 32 * g; P.scale = scale + 32 * g; P.beta = beta + 32 * g; P.rstd = rstd + 32 * g; P.gamma = gamma + 32 * g;
    P.w_last = w_last + 32 * 9 * g;
    {
      FinalTcParams Q{};
      rc = final_tc_setup(Q, dy, B, H, W, C, g, false);
      if (rc != VAE_OK) return rc;
      P.tmap_dy = Q.tmap_y;
    }
    P.dl = (const bf16*)dlogits; P.dy = (bf16*)dy;
    P.ds1 = const_cast<double*>(dstats) + 32 * g; P.ds2 = const_cast<double*>(dstats) + C + 32 * g; P.dsdl = nullptr;
    P.slope = slope; P.training = training;
    rc = final_bwd_tc_launch(2, P, st);
    if (rc != VAE_OK) return rc;
  }
  return VAE_OK;
}
