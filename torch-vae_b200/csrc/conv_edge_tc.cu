// Edge layer on tensor cores: Conv2d(1 -> N, k3, s2, p1) forward of the first encoder block (models.py:45,
// enc.conv0) for N = 32 / 64 output channels, fp32 piano-roll in, bf16 NHWC out, BatchNorm statistics in the epilogue.
//
// With ONE input channel the conv has K = 9: per output pixel  y[p][c] = bias[c] + sum_t xwin[p][t] * w[c][t], which is
// a GEMM with M = pixels, N = channels, K = 9 (padded to 16) -- the same "window tile" trick as the output block
// (final_block_tc.cu): a thread builds its pixel's 3x3 window as one 32-byte K-major row of the A operand (bf16; the
// rolls are {0, 1}, exact), ONE tcgen05.mma (M = 128, N, K = 16) produces the 128 x N tile in TMEM, and the same
// thread reads its pixel's accumulator row back for bias / statistics / store.  The CUDA-core kernel this replaces
// (conv_edge.cu: 9 FMAs per output element) was issue-bound at 72 us for the B = 512 tensor (ideal 26 us).
//
//   warps 0-7 : two groups of 128 workers; a group owns its tiles end to end (window -> A tile -> wait -> epilogue) and
//               builds the NEXT tile's window before it waits for the current accumulator
//   warp 8    : MMA issuer (one thread), round-robin over the two groups
#include "tma_util.cuh"

namespace vae {

constexpr int kEtThreads = 288;
constexpr int kEtTile = 128;

struct EdgeTcParams {
  const float* x;       // [B][Hi][Wi] fp32
  const bf16* wq;       // [N][9] bf16 (the CUDA-core pack of a 1-channel conv)
  const float* bias;    // [N] or null
  bf16* out;            // [B][Ho][Wo][N]
  double* stats;        // [2N] or null
  int B, Hi, Wi, Ho, Wo, N;
  int tiles;            // B * Ho * Wo / 128
};

__device__ __forceinline__ uint32_t et_swz64(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

template <int N>
__global__ void __launch_bounds__(kEtThreads, 2)
edge_cin1_tc_kernel(const __grid_constant__ EdgeTcParams P) {
  // shared: A tiles [2 groups][2 buffers][128 rows][64 B] (SW64, K-major, 16 of 32 bf16 used), B image [N][64 B]
  __shared__ __align__(1024) uint8_t a_s[2 * 2 * 8192];
  __shared__ __align__(1024) uint8_t b_s[64 * 64];
  __shared__ uint64_t aready_bar[4], dfull_bar[4];
  __shared__ uint32_t tmem_ptr_s;
  __shared__ float s_stats[2 * N], s_bias[N];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) {
        mbar_init(&aready_bar[i], 128);
        mbar_init(&dfull_bar[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_ptr_s, 4 * N);   // two groups x two accumulators of N columns (128 or 256: a power of two)
  }
  for (int i = threadIdx.x; i < 2 * N; i += kEtThreads) s_stats[i] = 0.0f;
  for (int i = threadIdx.x; i < N; i += kEtThreads) s_bias[i] = P.bias != nullptr ? P.bias[i] : 0.0f;
  // weight image: row c, 16-byte chunks 0 / 1 = taps 0..7 / 8..15 (tap 8 + zeros); chunks 2 / 3 unused
  for (int i = threadIdx.x; i < N * 2; i += kEtThreads) {
    const int c = i >> 1, half = i & 1;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    const unsigned short* src = reinterpret_cast<const unsigned short*>(P.wq) + c * 9;
    if (half == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = (uint32_t)src[2 * k] | ((uint32_t)src[2 * k + 1] << 16);
    } else {
      w[0] = (uint32_t)src[8];
    }
    *reinterpret_cast<uint4*>(b_s + et_swz64(c, half)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int i = threadIdx.x; i < 2 * 2 * 8192 / 16; i += kEtThreads) reinterpret_cast<uint4*>(a_s)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr_s;
  // tiles are dealt to (CTA, group) pairs: unit u = 2 * blockIdx.x + group handles tiles u, u + 2 * gridDim.x, ...
  const int units = 2 * gridDim.x;

  if (warp == 8) {
    // ============================= MMA issuer =============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, (uint32_t)N, 0, 0);
      const uint32_t a_a = smem_u32(a_s), b_a = smem_u32(b_s);
      const uint64_t bd = make_desc_sw(b_a, 512, 4);
      int n[2] = {0, 0};
      bool more = true;
      while (more) {
        more = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const long long tile = (long long)(2 * blockIdx.x + g) + (long long)n[g] * units;
          if (tile >= P.tiles) continue;
          more = true;
          const int buf = n[g] & 1;
          mbar_wait(&aready_bar[g * 2 + buf], (uint32_t)(n[g] >> 1) & 1u);
          tc_fence_after();
          mma_bf16_ss(tmem_base + (uint32_t)((g * 2 + buf) * N), make_desc_sw(a_a + (uint32_t)((g * 2 + buf) * 8192), 512, 4), bd, idesc, 0u);
          mma_commit(&dfull_bar[g * 2 + buf]);
          ++n[g];
        }
      }
    }
  } else {
    // =============================== workers ==============================
    const int g = warp >> 2, p = threadIdx.x & 127;
    const uint32_t a_a = smem_u32(a_s) + (uint32_t)(g * 2 * 8192);
    const uint32_t off0 = et_swz64(p, 0), off1 = et_swz64(p, 1);
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 2 * N);
    const int Wo = P.Wo, Ho = P.Ho, Wi = P.Wi, Hi = P.Hi;
    float run1[32], run2[32];   // N == 32: per-channel running sums in registers
#pragma unroll
    for (int i = 0; i < 32; ++i) run1[i] = run2[i] = 0.0f;

    auto build = [&](long long tile, int buf) {   // this pixel's 3x3 window -> its K-major row of the A tile
      const long long m = tile * kEtTile + p;
      const int ow = (int)(m % Wo);
      const long long t2 = m / Wo;
      const int oh = (int)(t2 % Ho);
      const long long b = t2 / Ho;
      const float* img = P.x + b * (long long)Hi * Wi;
      float w[9];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int ih = 2 * oh - 1 + a;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int iw = 2 * ow - 1 + c;
          w[a * 3 + c] = (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) ? __ldg(img + (long long)ih * Wi + iw) : 0.0f;
        }
      }
      uint4 lo, hi = make_uint4(0u, 0u, 0u, 0u);
      __nv_bfloat162 h2;
      h2 = __floats2bfloat162_rn(w[0], w[1]); lo.x = *reinterpret_cast<uint32_t*>(&h2);
      h2 = __floats2bfloat162_rn(w[2], w[3]); lo.y = *reinterpret_cast<uint32_t*>(&h2);
      h2 = __floats2bfloat162_rn(w[4], w[5]); lo.z = *reinterpret_cast<uint32_t*>(&h2);
      h2 = __floats2bfloat162_rn(w[6], w[7]); lo.w = *reinterpret_cast<uint32_t*>(&h2);
      h2 = __floats2bfloat162_rn(w[8], 0.0f); hi.x = *reinterpret_cast<uint32_t*>(&h2);
      const uint32_t base = a_a + (uint32_t)(buf * 8192);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + off0), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + off1), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
      fence_proxy_async_smem();
      mbar_arrive(&aready_bar[g * 2 + buf]);
    };

    const long long first = 2 * blockIdx.x + g;
    int n = 0;
    if (first < P.tiles) build(first, 0);
    for (long long tile = first; tile < P.tiles; tile += units, ++n) {
      const int buf = n & 1;
      // the other A buffer was last read by the MMA of tile n - 1, whose accumulator this thread consumed in the
      // previous iteration (dfull observed): it is free
      if (tile + units < P.tiles) build(tile + units, buf ^ 1);
      mbar_wait(&dfull_bar[g * 2 + buf], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      bf16* dst = P.out + (tile * kEtTile + p) * N;
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld_32x32(t_lane + (uint32_t)(buf * N + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + i]);
          v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[i + 2 * k], v[i + 2 * k + 1]);
          *reinterpret_cast<uint4*>(dst + c0 + i) = pk;
        }
        if (P.stats != nullptr) {
          if (N == 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              run1[i] += v[i];
              run2[i] = fmaf(v[i], v[i], run2[i]);
            }
          } else {
            float sq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
            warp_transpose_reduce32(v, lane);
            warp_transpose_reduce32(sq, lane);
            atomicAdd(&s_stats[c0 + lane], v[0]);
            atomicAdd(&s_stats[N + c0 + lane], sq[0]);
          }
        }
      }
      tc_fence_before();   // the accumulator has been read: orders the tcgen05.ld before the next MMA into this buffer
    }
    if (P.stats != nullptr) {
      if (N == 32) {
        warp_transpose_reduce32(run1, lane);
        warp_transpose_reduce32(run2, lane);
        atomicAdd(&s_stats[lane], run1[0]);
        atomicAdd(&s_stats[N + lane], run2[0]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight worker warps only
      for (int i = threadIdx.x; i < 2 * N; i += 256) atomicAdd(&P.stats[i], (double)s_stats[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 4 * N);
  }
}

// MEASURED (B200): 98.9 us against 83.0 us for the CUDA-core strip kernel at B = 512 (160 vs 141 us at the wide /
// long shape): one thread per pixel pays ~220 instructions per tile (nine bounds-checked loads, fences, barrier round
// trips, a 128-byte epilogue) and the 16 worker warps per SM run latency-bound at ~0.5 IPC.  Off by default; the
// strip kernel re-uses every loaded input across 4 pixels x 8 channels.  DESIGN.md section 6.1.
static bool g_edge_tc_enabled = false;

// returns 1 if the shape is not served (the caller falls back to the CUDA-core kernel), 0 / negative otherwise
int edge_cin1_tc(const GatherGeom& g, const void* in, const void* wq, const float* bias, void* out, double* stats,
                 int dt_in, int dt_out, cudaStream_t st) {
  const int N = g.Nout;
  if (!g_edge_tc_enabled || g.Cin != 1 || !(N == 32 || N == 64) || g.in_mul != 2 || g.out_mul != 1 || g.nphase != 1 ||
      dt_in != VAE_F32 || dt_out != VAE_BF16)
    return 1;
  const long long M = (long long)g.B * g.Hj * g.Wj;
  if (M % kEtTile != 0 || M / kEtTile >= (1LL << 30)) return 1;
  EdgeTcParams P{};
  P.x = (const float*)in; P.wq = (const bf16*)wq; P.bias = bias; P.out = (bf16*)out; P.stats = stats;
  P.B = g.B; P.Hi = g.Hin; P.Wi = g.Win; P.Ho = g.Hj; P.Wo = g.Wj; P.N = N;
  P.tiles = (int)(M / kEtTile);
  int grid = 2 * sm_count();
  if (grid > (P.tiles + 1) / 2) grid = (P.tiles + 1) / 2;
  if (N == 32) edge_cin1_tc_kernel<32><<<grid, kEtThreads, 0, st>>>(P);
  else edge_cin1_tc_kernel<64><<<grid, kEtThreads, 0, st>>>(P);
  return check_launch("edge_cin1_tc");
}

}  // namespace vae

/* Switch for the tensor-core edge-layer forward kernel (tests compare it with the CUDA-core kernel); returns the
 * previous setting. */
extern "C" int vae_set_edge_tc(int enable) {
  int old = vae::g_edge_tc_enabled ? 1 : 0;
  vae::g_edge_tc_enabled = enable != 0;
  return old;
}
