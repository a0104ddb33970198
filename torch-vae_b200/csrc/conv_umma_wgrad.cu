// tcgen05 / TMEM weight-gradient kernel (bf16 operands, fp32 accumulate).
//
//   dW[co][ci][kh][kw] = sum_j  Y[j][co] * X[pix(j, kh, kw)][ci]
//
// with Y the LOW-resolution tensor [J = B*Ho*Wo][Co] and X the HIGH-resolution
// one (conv input side) -- for Conv2d layers (Y = dL/dy, X = layer input) and for
// ConvTranspose2d layers (Y = layer input, X = dL/d(output)) alike.
//
// As a GEMM the reduction runs over PIXELS, which is the slow index of both NHWC
// operands, so both are fed to the tensor core MN-major (a_major = b_major = 1):
// the shared-memory image of a stage is simply [64 pixels][128 B of channels],
// 128B-swizzled -- the same bytes the NHWC rows hold -- and is filled by 16-byte
// cp.async copies with zero-fill for the 3x3 padding (no registers, deep prefetch).
//
//   D[128 (co) x N] (TMEM) += A[128 x 16 pixels] * B[N x 16 pixels]^T,
//   N = the flat (tap, ci) index of a group of TG taps (TG*Ci <= 512 TMEM columns)
//
// grid.x splits the pixel range (split-K); every CTA reduces its pixels into TMEM
// and adds its partial into the fp32 OIHW gradient with red.global.add.f32.
//   warps 0-7: cp.async producers (pixel row = t % 64, every 4th 16-byte chunk each);
//              warps 0-3 then run the epilogue (one TMEM lane = one co each)
//   warp  8  : TMEM allocation + single-thread tcgen05.mma issue
#include "common.cuh"
#include "umma.cuh"

namespace vae {

using namespace umma;

struct WgradParams {
  const bf16* X;
  const bf16* Y;
  float* dW;
  int B, Hi, Wi, Ci, Co, s, Ho, Wo;
  long long J;
  long long pix_per_split;  // multiple of 64
  int TG;                   // taps per CTA
  int n_tap_groups;         // 9 / TG
  int n_groups;             // ceil(TG * Ci / 64) column groups of the B operand
  int cpt;                  // Ci / 8
  int stages;
  int stage_bytes;          // 16 KB (A) + n_groups * 8 KB (B)
  int tmem_cols;
};

__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kWgProducers = 256;            // 8 producer warps
constexpr int kWgThreads = kWgProducers + 32;  // + the MMA warp
constexpr int kAStage = 2 * 8192;  // two 64-channel groups x 64 pixel rows x 128 B

__global__ void __launch_bounds__(kWgThreads, 1)
umma_wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* tail = smem + P.stages * P.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* done_bar = empty_bar + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.y / P.n_tap_groups;
  const int tg = blockIdx.y - m_tile * P.n_tap_groups;
  const int co_base = m_tile * 128;
  const int co_tile = min(128, P.Co - co_base);
  const int CA = co_tile / 8;      // real 16-byte chunks per A row
  const int CB = P.TG * P.cpt;     // real 16-byte chunks per B row
  const long long j_begin = (long long)blockIdx.x * P.pix_per_split;
  long long j_end = j_begin + P.pix_per_split;
  if (j_end > P.J) j_end = P.J;
  const int nkb = (int)((j_end - j_begin + 63) / 64);

  // zero every stage once: chunks that are never copied (channel padding) must read as 0
  {
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const int n16 = P.stages * P.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kWgThreads) reinterpret_cast<uint4*>(smem)[i] = z;
    fence_proxy_async_smem();
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < P.stages; ++s) {
        mbar_init(&full_bar[s], kWgProducers);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, (uint32_t)P.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 8) {
    // ============================ producers ============================
    const int p = threadIdx.x & 63, part = threadIdx.x >> 6;  // pixel row, chunk residue mod 4
    const uint32_t row_off = (uint32_t)p * 128u;
    const uint32_t xr = (uint32_t)(p & 7);
    int stage = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const long long j = j_begin + (long long)kb * 64 + p;
      const bool valid = j < j_end;
      int b = 0, ho = 0, wo = 0;
      if (valid) {  // J < 2^31 (checked on the host): 32-bit divisions
        const uint32_t ju = (uint32_t)j;
        wo = (int)(ju % (uint32_t)P.Wo);
        const uint32_t t2 = ju / (uint32_t)P.Wo;
        ho = (int)(t2 % (uint32_t)P.Ho);
        b = (int)(t2 / (uint32_t)P.Ho);
      }
      mbar_wait(&empty_bar[stage], ph ^ 1u);
      const uint32_t a_base = smem_u32(smem + stage * P.stage_bytes);
      const uint32_t b_base = a_base + kAStage;
      // A: Y[j][co_base + 8*ca ...]
      const bf16* ysrc = P.Y + (valid ? j : 0) * P.Co + co_base;
      for (int ca = part; ca < CA; ca += 4)
        cp_async_16(a_base + (uint32_t)(ca >> 3) * 8192u + row_off + ((((uint32_t)ca & 7u) ^ xr) << 4),
                    ysrc + ca * 8, valid ? 16u : 0u);
      // B: X[pix(j, tap)][8*c8 ...] for the TG taps of this CTA
      for (int tl = 0; tl < P.TG; ++tl) {
        const int tap = tg * P.TG + tl;
        const int kh = tap / 3, kw = tap - kh * 3;
        const int ih = ho * P.s + kh - 1, iw = wo * P.s + kw - 1;
        const bool inb = valid && ih >= 0 && ih < P.Hi && iw >= 0 && iw < P.Wi;
        const bf16* xsrc = inb ? P.X + (((long long)b * P.Hi + ih) * P.Wi + iw) * P.Ci : P.X;
        const uint32_t nbytes = inb ? 16u : 0u;
        for (int c8 = part; c8 < P.cpt; c8 += 4) {
          const uint32_t cb = (uint32_t)(tl * P.cpt + c8);
          cp_async_16(b_base + (cb >> 3) * 8192u + row_off + (((cb & 7u) ^ xr) << 4), xsrc + c8 * 8, nbytes);
        }
      }
      cp_async_arrive_noinc(&full_bar[stage]);  // arrives when this thread's copies have landed
      if (++stage == P.stages) { stage = 0; ph ^= 1u; }
    }
    // ============================= epilogue =============================
    const int row = warp * 32 + lane;  // co within the tile == TMEM lane (warps 0-3 only)
    const int co = co_base + row;
    const int ncols = warp < 4 ? P.TG * P.Ci : 0;
    if (warp < 4) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      float v[32];
      tmem_ld_32x32(t_row + (uint32_t)c0, v);
      tmem_ld_wait();
      if (row < co_tile) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int q = c0 + i;
          if (q < ncols) {
            const int tl = q / P.Ci, ci = q - tl * P.Ci;
            atomicAdd(&P.dW[((long long)co * P.Ci + ci) * 9 + tg * P.TG + tl], v[i]);
          }
        }
      }
    }
  } else {
    // ============================ MMA issuer ============================
    int stage = 0;
    uint32_t ph = 0;
    const int n_chunks = (P.n_groups + 3) / 4;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full_bar[stage], ph);
      fence_proxy_async_smem();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_base = smem_u32(smem + stage * P.stage_bytes);
        const uint32_t b_base = a_base + kAStage;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // 16 pixels per MMA
          const uint64_t ad = make_smem_desc(a_base + ks * 2048, 8192, 1024);
          for (int nc = 0; nc < n_chunks; ++nc) {
            const int g_this = min(4, P.n_groups - 4 * nc);
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)g_this * 64u, 1, 1);
            const uint64_t bd = make_smem_desc(b_base + nc * 4 * 8192 + ks * 2048, 8192, 1024);
            mma_bf16_ss(tmem_base + (uint32_t)nc * 256u, ad, bd, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          }
        }
        mma_commit(&empty_bar[stage]);
        if (kb == nkb - 1) mma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == P.stages) { stage = 0; ph ^= 1u; }
    }
    if (nkb == 0 && elect_one()) mbar_arrive(done_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

static bool g_umma_wgrad_enabled = true;

// returns false when the shape is not served
static bool wgrad_plan(WgradParams& P) {
  if (P.Ci % 32 != 0 || P.Co % 8 != 0 || P.Ci > 256 || P.Co < 8) return false;
  P.cpt = P.Ci / 8;
  P.TG = 9 * P.Ci <= 512 ? 9 : (3 * P.Ci <= 512 ? 3 : 1);
  P.n_tap_groups = 9 / P.TG;
  P.n_groups = (P.TG * P.Ci + 63) / 64;
  P.stage_bytes = kAStage + P.n_groups * 8192;
  P.stages = 200 * 1024 / P.stage_bytes;
  if (P.stages > 4) P.stages = 4;
  if (P.stages < 2) return false;
  int cols = 32;
  while (cols < P.n_groups * 64) cols *= 2;
  P.tmem_cols = cols;
  return true;
}

}  // namespace vae

using namespace vae;

extern "C" int vae_set_umma_wgrad(int enable) {
  int old = g_umma_wgrad_enabled ? 1 : 0;
  g_umma_wgrad_enabled = enable != 0;
  return old;
}

/* 1 if vae_conv_wgrad_umma serves this layer shape */
extern "C" int vae_umma_wgrad_supported(int Ci, int Co) {
  WgradParams P{};
  P.Ci = Ci; P.Co = Co;
  return (g_umma_wgrad_enabled && wgrad_plan(P)) ? 1 : 0;
}

/* dw[Co][Ci][3][3] fp32 (overwritten) from bf16 NHWC x [B,Hi,Wi,Ci] (high-res side)
 * and dy [B,Hi/s,Wi/s,Co] (low-res side) on the tcgen05 path.
 * replaces: the weight-gradient half of aten::convolution_backward (train.py:650) */
extern "C" int vae_conv_wgrad_umma(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Ci,
                                   int Co, int stride, vae_stream_t stream) {
  VAE_REQUIRE(x && dy && dw && B > 0 && (stride == 1 || stride == 2) && Hi % stride == 0 && Wi % stride == 0,
              "conv_wgrad_umma: bad argument");
  WgradParams P{};
  P.X = (const bf16*)x; P.Y = (const bf16*)dy; P.dW = dw;
  P.B = B; P.Hi = Hi; P.Wi = Wi; P.Ci = Ci; P.Co = Co; P.s = stride;
  P.Ho = Hi / stride; P.Wo = Wi / stride;
  P.J = (long long)B * P.Ho * P.Wo;
  VAE_REQUIRE(wgrad_plan(P), "conv_wgrad_umma: shape (Ci=%d, Co=%d) not served by the tcgen05 path", Ci, Co);
  VAE_REQUIRE(P.J < (1LL << 31), "conv_wgrad_umma: too many pixels for 32-bit indexing");
  cudaStream_t st = as_stream(stream);
  VAE_CUDA(zero_scratch(dw, sizeof(float) * 9 * (size_t)Ci * Co, st));
  const int m_tiles = (Co + 127) / 128;
  const int combos = m_tiles * P.n_tap_groups;
  long long kblocks = (P.J + 63) / 64;
  int splits = (2 * sm_count() + combos - 1) / combos;
  if (splits > kblocks) splits = (int)kblocks;
  if (splits < 1) splits = 1;
  long long kb_per = (kblocks + splits - 1) / splits;
  P.pix_per_split = kb_per * 64;
  splits = (int)((kblocks + kb_per - 1) / kb_per);
  size_t smem = 1024 + (size_t)P.stages * P.stage_bytes + 128;
  static bool attr_set = false;
  if (!attr_set) {
    VAE_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid(splits, combos);
  umma_wgrad_kernel<<<grid, kWgThreads, smem, st>>>(P);
  return check_launch("umma_wgrad");
}
