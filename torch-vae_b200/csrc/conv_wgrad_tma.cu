// TMA-fed tcgen05 / TMEM weight-gradient kernel (bf16 operands, fp32 accumulate), sm_100a.
//
//   dW[co][ci][kh][kw] = sum_j  Y[j][co] * X[pix(j, kh, kw)][ci]        j over B x Ho x Wo
//
// with Y the LOW-resolution tensor [B,Ho,Wo,Co] and X the HIGH-resolution one [B,Hi,Wi,Ci]
// (Conv2d: Y = dL/dy, X = layer input; ConvTranspose2d: Y = layer input, X = dL/d(output)).
//
// The reduction runs over PIXELS, the slow index of both NHWC tensors, so both operands are
// MN-major: a shared-memory "region" is simply [KP pixels][64 or 128 bytes of channels] -- the
// bytes an NHWC row holds -- swizzled by the TMA unit.  The same strided-box tensor map that
// feeds the forward conv performs the im2col gather: a k-block of KP pixels is made of whole
// output rows, so for one filter tap the pixels X[s*j + (kh-1, kw-1)] form a box {C, W step s,
// H step s, N}; padding is zero-filled by the TMA unit.  One thread issues all loads.
//
//   D[(tap, ci)  x  co] (TMEM)  +=  A[(tap, ci) x 16 pixels] * B[co x 16 pixels]^T
//
// A CTA owns a group of TG = 3 taps (one kernel row kh), a tile of <= 128 output channels and a
// contiguous range of pixel k-blocks (split-K); partials are added with fp32 atomics into a
// [tap][co][ci] workspace (consecutive lanes -> consecutive addresses) that a small second
// kernel transposes into the reference's OIHW layout.
//   warp 0    : TMA producer (one thread)
//   warp 1    : TMEM allocation + single-thread tcgen05.mma issue
//   warps 2-5 : epilogue (TMEM lane = (tap, ci) row)
#include "tma_util.cuh"

namespace vae {

constexpr int kWtThreads = 192;
constexpr int kWtMaxStages = 8;

struct WgradTmaParams {
  CUtensorMap tmap_x, tmap_y;
  float* ws;                 // [9][Co][Ci] fp32 partial sums
  int Ci, Co, s, Ho, Wo;
  int KP;                    // pixels per k-block (64 or 128)
  int TG, n_tap_groups;      // taps per CTA
  int co_tile, n_co_tiles;
  int cx, gpt;               // channels per X region (32 or 64), regions per tap
  int cy, ypr;               // channels per Y region, Y regions per co tile
  int x_region, y_region;    // bytes
  int stage_bytes, stages;
  int mtiles, tmem_cols;
  int kblocks, kb_per_split;
  int layout_x, layout_y;    // UMMA swizzle codes (4 = 64B, 2 = 128B)
  int ctas_per_sm;
  int rows_w, rows_h, rows_n;  // a k-block = rows_n images x rows_h rows x rows_w pixels of the Ho x Wo grid
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                 uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

template <int KSTEPS, int MTILES>
__device__ __forceinline__ void issue_stage(uint32_t tmem_base, uint32_t hi_x, uint32_t hi_y, uint32_t lo_x, uint32_t lo_y,
                                            uint32_t kstep_x, uint32_t kstep_y, uint32_t mt_step, uint32_t co_tile,
                                            uint32_t idesc, uint32_t accum_first) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const uint64_t bd = ((uint64_t)hi_y << 32) | (lo_y + (uint32_t)ks * kstep_y);
#pragma unroll
    for (int mt = 0; mt < MTILES; ++mt) {
      const uint64_t ad = ((uint64_t)hi_x << 32) | (lo_x + (uint32_t)ks * kstep_x + (uint32_t)mt * mt_step);
      mma_bf16_ss(tmem_base + (uint32_t)mt * co_tile, ad, bd, idesc, ks > 0 ? 1u : accum_first);
    }
  }
}

__global__ void __launch_bounds__(kWtThreads, 2)
umma_wgrad_tma_kernel(const __grid_constant__ WgradTmaParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* tail = smem + P.stages * P.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kWtMaxStages;
  uint64_t* done_bar = empty_bar + kWtMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int combo = blockIdx.x, split = blockIdx.y;
  const int tg = combo % P.n_tap_groups, cot = combo / P.n_tap_groups;
  const int co_base = cot * P.co_tile;
  const int kb_begin = split * P.kb_per_split;
  const int kb_end = min(P.kblocks, kb_begin + P.kb_per_split);
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kWtMaxStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, (uint32_t)P.tmem_cols);
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_y) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int x_bytes = P.TG * P.gpt * P.x_region;

  if (warp == 0) {
    // ============================== producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      const uint32_t Wo = (uint32_t)P.Wo, Ho = (uint32_t)P.Ho;
      const uint32_t tx = (uint32_t)(x_bytes + P.ypr * P.y_region);
      const int s = P.s, TG = P.TG, gpt = P.gpt, ypr = P.ypr, cx = P.cx, cy = P.cy;
      const int x_region = P.x_region, y_region = P.y_region;
      // taps of this CTA: consecutive taps of one kernel row (TG = 3) or a single tap (TG = 1)
      const int tap0 = tg * TG;
      const int kh0 = tap0 / 3, kw0 = tap0 - kh0 * 3;
      // position of the first k-block (the only divisions of this thread); later k-blocks advance it
      // incrementally -- an integer division costs this lone, latency-bound thread ~150 cycles
      int jw0, jh0, b0;
      {
        const uint32_t m0 = (uint32_t)kb_begin * (uint32_t)P.KP;
        jw0 = (int)(m0 % Wo);
        const uint32_t t2 = m0 / Wo;
        jh0 = (int)(t2 % Ho);
        b0 = (int)(t2 / Ho);
      }
      const int rows_w = P.rows_w, rows_h = P.rows_h, rows_n = P.rows_n, iWo = P.Wo, iHo = P.Ho;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int xw = jw0 * s + kw0 - 1, xh = jh0 * s + kh0 - 1;
        mbar_wait(&empty_bar[stage], ph ^ 1u);
        uint8_t* st = smem + stage * P.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], tx);
        uint8_t* dstp = st;
        for (int tl = 0; tl < TG; ++tl)
          for (int g = 0; g < gpt; ++g, dstp += x_region)
            tma_load_4d(dstp, &P.tmap_x, &full_bar[stage], g * cx, xw + tl, xh, b0);
        dstp = st + x_bytes;
        for (int g = 0; g < ypr; ++g, dstp += y_region)
          tma_load_4d(dstp, &P.tmap_y, &full_bar[stage], co_base + g * cy, jw0, jh0, b0);
        if (++stage == P.stages) { stage = 0; ph ^= 1u; }
        jw0 += rows_w;
        if (jw0 >= iWo) {
          jw0 = 0;
          jh0 += rows_h;
          if (jh0 >= iHo) {
            jh0 = 0;
            b0 += rows_n;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)P.co_tile, 1, 1);
    const uint32_t rowx = (uint32_t)P.cx * 2u, rowy = (uint32_t)P.cy * 2u;   // bytes per pixel row of a region
    // descriptor words (see make_desc_mn): hi = SBO | version | swizzle, lo = address>>4 | LBO<<16.
    // Shared-memory addresses are < 256 KB, so address>>4 never carries out of its 14-bit field and a
    // descriptor is advanced with one 32-bit add.
    const uint32_t hi_x = (((8u * rowx) >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)P.layout_x << 29);
    const uint32_t hi_y = (((8u * rowy) >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)P.layout_y << 29);
    const uint32_t lbo_x = (((uint32_t)P.x_region >> 4) & 0x3FFFu) << 16;
    const uint32_t lbo_y = (((uint32_t)P.y_region >> 4) & 0x3FFFu) << 16;
    const uint32_t kstep_x = (16u * rowx) >> 4, kstep_y = (16u * rowy) >> 4;
    const uint32_t mt_step = (uint32_t)((128 / P.cx) * P.x_region) >> 4;  // X regions of one 128-row M tile
    const int ksteps = P.KP / 16, mtiles = P.mtiles;
    const uint32_t co_tile = (uint32_t)P.co_tile;
    const uint32_t smem0 = smem_u32(smem), stage_bytes = (uint32_t)P.stage_bytes, xb = (uint32_t)x_bytes;
    int stage = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full_bar[stage], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xa = smem0 + (uint32_t)stage * stage_bytes;
        const uint32_t lo_x = (xa >> 4) | lbo_x;
        const uint32_t lo_y = ((xa + xb) >> 4) | lbo_y;
        const uint32_t accum = kb > 0 ? 1u : 0u;
        // straight-line issue: the single issuing thread is latency-bound, so the (k-step, M-tile) loops are
        // compile-time unrolled and every descriptor is base + constant * index (no loop-carried chain)
#define VAE_WT_ISSUE(KS_, MT_) \
        issue_stage<KS_, MT_>(tmem_base, hi_x, hi_y, lo_x, lo_y, kstep_x, kstep_y, mt_step, co_tile, idesc, accum)
        if (ksteps == 8) {
          if (mtiles == 1) VAE_WT_ISSUE(8, 1);
          else if (mtiles == 2) VAE_WT_ISSUE(8, 2);
          else if (mtiles == 3) VAE_WT_ISSUE(8, 3);
          else VAE_WT_ISSUE(8, 4);
        } else {
          if (mtiles == 1) VAE_WT_ISSUE(4, 1);
          else if (mtiles == 2) VAE_WT_ISSUE(4, 2);
          else if (mtiles == 3) VAE_WT_ISSUE(4, 3);
          else VAE_WT_ISSUE(4, 4);
        }
#undef VAE_WT_ISSUE
        mma_commit(&empty_bar[stage]);
        if (kb == nkb - 1) mma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == P.stages) { stage = 0; ph ^= 1u; }
    }
  } else if (nkb > 0) {
    // ============================== epilogue ==============================
    const int lg = warp & 3;  // TMEM lane group
    mbar_wait_sleep(done_bar, 0, 1000);  // idle for the whole main loop: do not compete with warps 0 / 1 for issue slots
    tc_fence_after();
    const int rows = P.TG * P.Ci;
    for (int mt = 0; mt < P.mtiles; ++mt) {
      const int row = mt * 128 + lg * 32 + lane;
      const bool valid = row < rows;
      const int tl = valid ? row / P.Ci : 0;
      const int ci = row - tl * P.Ci;
      const int tap = tg * P.TG + tl;
      float* dst = P.ws + ((long long)tap * P.Co + co_base) * P.Ci + ci;
      const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mt * P.co_tile);
      for (int c0 = 0; c0 < P.co_tile; c0 += 32) {
        float v[32];
        tmem_ld_32x32(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + (long long)(c0 + i) * P.Ci, v[i]);
        }
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

// ws[tap][co][ci] -> dW[co][ci][tap]  (OIHW)
__global__ void wgrad_finish_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Co, int Ci) {
  const int n = Co * Ci;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = ws[(long long)t * n + i];
#pragma unroll
    for (int t = 0; t < 9; ++t) dw[(long long)i * 9 + t] = v[t];
  }
}

static bool g_wgrad_tma_enabled = true;

// KP pixels = rows_n x rows_h x rows_w whole rows of the Ho x Wo grid
static bool kblock_rows(int Ho, int Wo, int KP, int* rw, int* rh, int* rn) {
  if (Wo >= KP) {
    if (Wo % KP) return false;
    *rw = KP; *rh = 1; *rn = 1;
    return true;
  }
  if (KP % Wo) return false;
  const int rows = KP / Wo;
  if (rows <= Ho) {
    if (Ho % rows) return false;
    *rw = Wo; *rh = rows; *rn = 1;
    return true;
  }
  if (rows % Ho) return false;
  *rw = Wo; *rh = Ho; *rn = rows / Ho;
  return true;
}

struct WtPlan {
  WgradTmaParams P;
  int rw, rh, rn;
};

static bool wgrad_tma_plan(WtPlan& pl, int B, int Hi, int Wi, int Ci, int Co, int s) {
  WgradTmaParams& P = pl.P;
  if (!g_wgrad_tma_enabled || B <= 0 || (s != 1 && s != 2) || Hi % s || Wi % s) return false;
  if (!(Ci == 32 || Ci == 64 || Ci == 128 || Ci == 256) || Co % 32 != 0 || Co < 32) return false;
  if (Co > 128 && Co % 128 != 0) return false;
  if (Co < 128 && !(Co == 32 || Co == 64)) return false;
  if (get_encode_fn() == nullptr) return false;
  P.Ci = Ci; P.Co = Co; P.s = s; P.Ho = Hi / s; P.Wo = Wi / s;
  P.cx = Ci == 32 ? 32 : 64; P.gpt = Ci / P.cx;
  P.co_tile = Co < 128 ? Co : 128; P.n_co_tiles = Co / P.co_tile;
  P.cy = P.co_tile == 32 ? 32 : 64; P.ypr = P.co_tile / P.cy;
  P.layout_x = P.cx == 32 ? 4 : 2; P.layout_y = P.cy == 32 ? 4 : 2;
  for (int tg_try = 0; tg_try < 2; ++tg_try) {
    P.TG = tg_try == 0 ? 3 : 1;
    P.n_tap_groups = 9 / P.TG;
    P.mtiles = (P.TG * Ci + 127) / 128;
    if (P.mtiles * P.co_tile > 512) continue;
    // Preference: configurations that leave room for TWO CTAs per SM (the per-CTA chain TMA gather -> MMA ->
    // commit -> slot reuse is latency-bound; two interleaved chains hide it: 234 -> 142 us on the 32->32
    // layer at 128x128), larger k-blocks first; then one CTA per SM.
    for (int attempt = 0; attempt < 4; ++attempt) {
      const int want_ctas = attempt < 2 ? 2 : 1;
      P.KP = (attempt & 1) == 0 ? 128 : 64;
      if (!kblock_rows(P.Ho, P.Wo, P.KP, &pl.rw, &pl.rh, &pl.rn)) continue;
      if (pl.rw * s > 256 || pl.rh * s > 256 || pl.rn > 256) continue;
      P.x_region = P.KP * P.cx * 2; P.y_region = P.KP * P.cy * 2;
      // the last M tile may address up to (mtiles*128/cx) X regions: keep them inside the stage
      const int x_regions = P.TG * P.gpt, x_addr = P.mtiles * (128 / P.cx);
      const int pad_regions = x_addr > x_regions ? x_addr - x_regions : 0;
      int y_bytes = P.ypr * P.y_region;
      if (pad_regions * P.x_region > y_bytes) y_bytes = pad_regions * P.x_region;
      P.stage_bytes = (x_regions * P.x_region + y_bytes + 1023) / 1024 * 1024;
      P.ctas_per_sm = want_ctas;
      P.stages = ((want_ctas == 2 ? 111 : 200) * 1024 - 2048) / P.stage_bytes;
      if (P.stages > kWtMaxStages) P.stages = kWtMaxStages;
      if (P.stages < 3) continue;
      if (want_ctas == 2 && 2 * P.mtiles * P.co_tile > 512) continue;  // TMEM columns of two resident CTAs
      int cols = 32;
      while (cols < P.mtiles * P.co_tile) cols *= 2;
      P.tmem_cols = cols;
      const long long J = (long long)B * P.Ho * P.Wo;
      if (J >= (1LL << 31)) return false;
      P.kblocks = (int)((J + P.KP - 1) / P.KP);
      P.rows_w = pl.rw; P.rows_h = pl.rh; P.rows_n = pl.rn;
      return true;
    }
  }
  return false;
}

}  // namespace vae

using namespace vae;

extern "C" int vae_set_wgrad_tma(int enable) {
  int old = g_wgrad_tma_enabled ? 1 : 0;
  g_wgrad_tma_enabled = enable != 0;
  return old;
}

/* 1 if vae_conv_wgrad_tma serves this layer */
extern "C" int vae_wgrad_tma_eligible(int B, int Hi, int Wi, int Ci, int Co, int stride) {
  WtPlan pl{};
  return wgrad_tma_plan(pl, B, Hi, Wi, Ci, Co, stride) ? 1 : 0;
}

/* dw[Co][Ci][3][3] fp32 (overwritten) from bf16 NHWC x [B,Hi,Wi,Ci] (high-res side) and
 * dy [B,Hi/s,Wi/s,Co] (low-res side); ws: caller-provided fp32 workspace of 9*Co*Ci elements.
 * replaces: the weight-gradient half of aten::convolution_backward (train.py:650) */
extern "C" int vae_conv_wgrad_tma(const void* x, const void* dy, float* dw, float* ws, int B, int Hi, int Wi, int Ci,
                                  int Co, int stride, vae_stream_t stream) {
  VAE_REQUIRE(x && dy && dw && ws, "conv_wgrad_tma: null argument");
  WtPlan pl{};
  VAE_REQUIRE(wgrad_tma_plan(pl, B, Hi, Wi, Ci, Co, stride), "conv_wgrad_tma: shape not served");
  WgradTmaParams& P = pl.P;
  EncodeTiledFn encode = get_encode_fn();
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Ci, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)Ci * 2, (cuuint64_t)Wi * Ci * 2, (cuuint64_t)Hi * Wi * Ci * 2};
    cuuint32_t box[4] = {(cuuint32_t)P.cx, (cuuint32_t)(pl.rw * stride), (cuuint32_t)(pl.rh * stride), (cuuint32_t)pl.rn};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult cr = encode(&P.tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         P.cx == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VAE_REQUIRE(cr == CUDA_SUCCESS, "conv_wgrad_tma: tensor map (x) failed (%d)", (int)cr);
  }
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Co, (cuuint64_t)P.Wo, (cuuint64_t)P.Ho, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)Co * 2, (cuuint64_t)P.Wo * Co * 2, (cuuint64_t)P.Ho * P.Wo * Co * 2};
    cuuint32_t box[4] = {(cuuint32_t)P.cy, (cuuint32_t)pl.rw, (cuuint32_t)pl.rh, (cuuint32_t)pl.rn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&P.tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         P.cy == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VAE_REQUIRE(cr == CUDA_SUCCESS, "conv_wgrad_tma: tensor map (y) failed (%d)", (int)cr);
  }
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(ws, sizeof(float) * 9 * (size_t)Ci * Co, st));
  P.ws = ws;
  const int combos = P.n_tap_groups * P.n_co_tiles;
  int splits = P.ctas_per_sm * sm_count() / combos;
  if (splits < 1) splits = 1;
  if (splits > P.kblocks) splits = P.kblocks;
  P.kb_per_split = (P.kblocks + splits - 1) / splits;
  splits = (P.kblocks + P.kb_per_split - 1) / P.kb_per_split;
  const size_t smem = 1024 + (size_t)P.stages * P.stage_bytes + 8 * (2 * kWtMaxStages + 1) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    VAE_CUDA(cudaFuncSetAttribute(umma_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid(combos, splits);
  umma_wgrad_tma_kernel<<<grid, kWtThreads, smem, st>>>(P);
  int rc = check_launch("umma_wgrad_tma");
  if (rc != 0) return rc;
  const int n = Co * Ci;
  wgrad_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws, dw, Co, Ci);
  return check_launch("wgrad_finish");
}
