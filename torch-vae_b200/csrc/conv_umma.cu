// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 x bf16 -> fp32).
//
// One persistent, warp-specialised kernel runs every "gather-GEMM" of the hot
// path -- Conv2d(k3,s2,p1) forward, ConvTranspose2d(k3,s2,p1,op1) forward (the
// data-gradient of the equivalent conv, as 4 output-parity phases with 1/2/2/4
// taps) and both data gradients (geometry: conv_geom.cuh):
//
//   D[128 pixels x N channels] (TMEM, fp32) += A[128 x 64] (smem) * B[N x 64]^T (smem)
//
//   warps 0-7  A producers: per k-block every GEMM row (output pixel) needs 64 input
//              channels (8 x 16 B, one tap of the 3x3 window).  Thread t copies chunk
//              column t%8 of rows t/8 + 32*i with cp.async (zero-fill for padding)
//              straight into the 128B-swizzled K-major tile UMMA expects; the 8
//              lanes of a row fetch 128 contiguous bytes.  Up to D+1 k-blocks per thread are in flight (a ring that
//              runs across tile boundaries), so ~80 KB per SM is always
//              outstanding.  When a k-block has landed the thread optionally
//              applies the PREVIOUS layer's BatchNorm+LeakyReLU to its own row in
//              place (padding stays zero), fences the async proxy and arrives.
//              The weight tile of the k-block is a pre-swizzled contiguous image
//              in global memory and arrives by ONE bulk-TMA copy (cp.async.bulk)
//              signalling the same mbarrier.
//   warp  12   MMA issuer: one elected thread issues 4 x tcgen05.mma (K=16 each)
//              per k-block and commits to the stage's "empty" mbarrier; after the
//              last k-block it commits to the accumulator's "full" mbarrier.
//   warps 8-11 epilogue: tcgen05.ld the fp32 accumulator (double-buffered in TMEM,
//              so the epilogue of tile i overlaps the main loop of tile i+1), add
//              bias, accumulate per-channel sum / sum-of-squares for the following
//              train-mode BatchNorm (warp transpose-reduce -> smem -> one fp64
//              atomic per channel per CTA at the very end), convert to bf16, store.
//
// Grid = #SMs (persistent, static round-robin tile schedule), 1 CTA / SM.
#include "common.cuh"
#include "conv_geom.cuh"
#include "umma.cuh"

namespace vae {

using namespace umma;

constexpr int kMaxStages = 6;
constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KB
constexpr int kProducers = 256;                   // 8 producer warps
constexpr int kThreads = kProducers + 128 + 32;   // + 4 epilogue warps + 1 MMA warp
constexpr int kProducerWarps = kProducers / 32;
constexpr int kWarpDepth = 1;  // k-blocks a producer warp keeps in flight beyond the oldest
constexpr int kRowRing = 8;    // shared ring of tile row tables (> ring stages: tiles that can be in flight)

struct __align__(8) RowInfo {
  uint32_t off;   // element offset of in[b][in_mul*jh][in_mul*jw][0], the row's window centre
  uint32_t mask;  // bit t: tap t of the phase's tap list falls inside the image (0: row beyond M)
};

struct UmmaGatherParams {
  GatherGeom g;
  int nkb[4];        // k-blocks per phase
  int b_off[4];      // element offset of the phase's first weight image
  int tiles_per_phase;
  int total_tiles;
  int cpt;           // 16-byte chunks per tap = Cin / 8
  int N;             // output channels (multiple of 16, <= 256)
  int b_stage_bytes; // N * 128 rounded up to 1024
  int tmem_cols;     // power of two >= 2 * N
  int stages;        // smem ring depth (4..6)
  int depth;         // k-blocks a producer thread keeps in flight beyond the oldest (D)
  const bf16* in;
  const bf16* wpk;
  const float* bias;
  bf16* out;
  double* stats;
  const float* pro_scale;  // optional fused BatchNorm+LeakyReLU on the gathered operand:
  const float* pro_mean;   //   leaky_relu((x - mean) * scale + beta), ATen operation order
  const float* pro_beta;
  float slope;
  long long* trace;  // optional [6][512] clock64 samples of CTA 0 (debug / tuning aid)
};

// --------------------------------------------------------------------------
// weight images: for phase p, k-block kb: [N rows][64 k] bf16, 128B-swizzled,
// element (n, e) = W(n, tap(q), c) with chunk q = kb*8 + e/8 -> tap = q / cpt,
// c = (q % cpt)*8 + e%8; zero beyond the phase's taps.
// mode 0 (fprop):  W(n, t, c) = w[co=n][ci=c][t]     (Nout = Co, Cin_gemm = Ci)
// mode 1 (dgrad):  W(n, t, c) = w[co=c][ci=n][t]     (Nout = Ci, Cin_gemm = Co)
// --------------------------------------------------------------------------
__global__ void pack_umma_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out,
                                        const __grid_constant__ UmmaGatherParams P, int Co, int Ci,
                                        int mode) {
  const int phase = blockIdx.y;
  const TapTable& tt = P.g.phase[phase];
  const long long n_elem = (long long)P.nkb[phase] * P.N * 64;
  bf16* dst = out + P.b_off[phase];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem;
       i += (long long)gridDim.x * blockDim.x) {
    int e = (int)(i & 63);
    int n = (int)((i >> 6) % P.N);
    int kb = (int)((i >> 6) / P.N);
    int q = kb * 8 + (e >> 3);
    int tap_i = q / P.cpt;
    float v = 0.0f;
    if (tap_i < tt.ntaps) {
      int c = (q - tap_i * P.cpt) * 8 + (e & 7);
      int t = tt.wt[tap_i];
      int co = mode == 0 ? n : c, ci = mode == 0 ? c : n;
      v = w[((long long)co * Ci + ci) * 9 + t];
    }
    long long off = (long long)kb * P.N * 64 + (long long)n * 64 + ((((e >> 3) ^ (n & 7)) << 3) | (e & 7));
    dst[off] = __float2bfloat16_rn(v);
  }
}

// --------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ uint4 bn_lrelu_8(uint4 raw, const float* __restrict__ sc, const float* __restrict__ mn,
                                            const float* __restrict__ bt, int c0, float slope) {
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    const int c = c0 + 2 * i;
    float u0 = fmaf(f.x - mn[c], sc[c], bt[c]);
    float u1 = fmaf(f.y - mn[c + 1], sc[c + 1], bt[c + 1]);
    u0 = u0 > 0.0f ? u0 : u0 * slope;
    u1 = u1 > 0.0f ? u1 : u1 * slope;
    h[i] = __floats2bfloat162_rn(u0, u1);
  }
  return raw;
}

__device__ __forceinline__ void cp_async_16_zfill(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#define VAE_TRACE(role, idx)                                                              \
  do {                                                                                   \
    if (P.trace != nullptr && blockIdx.x == 0 && lane == 0 && (idx) < 512)                \
      P.trace[(role) * 512 + (idx)] = clock64();                                          \
  } while (0)

__global__ void __launch_bounds__(kThreads, 1)
umma_gather_gemm_kernel(const __grid_constant__ UmmaGatherParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB][B b_stage_bytes] | barriers | tmem ptr | stats | prologue vectors
  uint8_t* smem = align1024(smem_raw);
  const int stage_bytes = kABytes + P.b_stage_bytes;
  const int kStages = P.stages;
  uint8_t* tail = smem + kStages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile uint32_t* s_consumed = tmem_ptr + 1;  // k-blocks whose MMAs have completed (written by the MMA warp)
  float* s_stats = reinterpret_cast<float*>(tmem_ptr + 4);  // [2][N]
  float* s_pro = s_stats + 2 * 256;                         // [3][Cin] (scale, mean, beta)
  RowInfo* s_rows = reinterpret_cast<RowInfo*>(s_pro + 3 * 2048);  // [kRowRing][128]
  float* s_bias = reinterpret_cast<float*>(s_rows + kRowRing * kTileM);  // [256]
  uint32_t* s_ready = reinterpret_cast<uint32_t*>(s_bias + 256);       // [kRowRing]: tile index + 1 whose table is valid

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatherGeom& g = P.g;
  const int N = P.N;

  for (int i = threadIdx.x; i < 2 * N; i += kThreads) s_stats[i] = 0.0f;
  for (int i = threadIdx.x; i < 256; i += kThreads) s_bias[i] = (P.bias != nullptr && i < N) ? P.bias[i] : 0.0f;
  if (threadIdx.x < kRowRing) s_ready[threadIdx.x] = 0u;
  if (P.pro_scale != nullptr)
    for (int i = threadIdx.x; i < g.Cin; i += kThreads) {
      s_pro[i] = P.pro_scale[i];
      s_pro[g.Cin + i] = P.pro_mean[i];
      s_pro[2 * g.Cin + i] = P.pro_beta[i];
    }
  if (warp == 12) {
    if (lane == 0) {
      tmem_ptr[1] = 0u;
      for (int s = 0; s < kMaxStages; ++s) {
        mbar_init(&full_bar[s], 32 + 1);  // the 32 lanes of the owning producer warp + the expect_tx arrive
        mbar_init(&empty_bar[s], 1);       // one tcgen05.commit
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const long long M_total = (long long)g.B * g.Hj * g.Wj;

  if (warp < 8) {
    // =========================== A/B producers ===========================
    // The CTA's k-blocks form one global sequence (the order the MMA warp consumes
    // them).  Producer warp w owns k-blocks seq == w (mod 8) and produces each of them
    // completely by itself: eight independent instruction streams, and the window of
    // ring slots that is open at any moment always maps to distinct warps, so ~6
    // k-blocks (96 KB) are in flight per SM.  Lane l copies 16-byte chunk column
    // j = l % 8 of rows (l / 8) + 4*i, i = 0..31: the 8 lanes of a row fetch 128
    // contiguous bytes.  Row geometry of a tile is computed once, by the warp that owns
    // the tile's first k-block, into a shared ring; the others wait for its ready flag.
    const int pw = warp;
    const int j = lane & 7, rsub = lane >> 3;
    const bool fuse = P.pro_scale != nullptr;
    const int G = (int)gridDim.x;
    const int n_my = (P.total_tiles - (int)blockIdx.x + G - 1) / G;  // tiles of this CTA
    const uint32_t M32 = (uint32_t)M_total;

    struct Cursor {
      int seq, i, kb, nkb, phase;
    };
    auto tile_info = [&](Cursor& c) {
      if (c.i >= n_my) return;
      const int tile = (int)blockIdx.x + c.i * G;
      c.phase = tile / P.tiles_per_phase;
      c.nkb = P.nkb[c.phase];
    };
    auto advance = [&](Cursor& c) {
      c.seq += kProducerWarps;
      c.kb += kProducerWarps;
      while (c.i < n_my && c.kb >= c.nkb) {
        c.kb -= c.nkb;
        ++c.i;
        tile_info(c);
      }
    };
    auto build_rows = [&](const Cursor& c) {  // 4 rows per lane
      const TapTable& tt = g.phase[c.phase];
      const int tile = (int)blockIdx.x + c.i * G;
      const int mt = tile - c.phase * P.tiles_per_phase;
      RowInfo* tab = s_rows + (c.i & (kRowRing - 1)) * kTileM;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = lane + 32 * k;
        const uint32_t m = (uint32_t)mt * kTileM + row;
        RowInfo ri;
        ri.off = 0u; ri.mask = 0u;
        if (m < M32) {
          const uint32_t jw = m % (uint32_t)g.Wj, t2 = m / (uint32_t)g.Wj;
          const uint32_t jh = t2 % (uint32_t)g.Hj, b = t2 / (uint32_t)g.Hj;
          const int bh = (int)jh * g.in_mul, bw = (int)jw * g.in_mul;
          ri.off = ((b * (uint32_t)g.Hin + (uint32_t)bh) * (uint32_t)g.Win + (uint32_t)bw) * (uint32_t)g.Cin;
          for (int ti = 0; ti < tt.ntaps; ++ti) {
            const int ih = bh + tt.dh[ti], iw = bw + tt.dw[ti];
            if (ih >= 0 && ih < g.Hin && iw >= 0 && iw < g.Win) ri.mask |= 1u << ti;
          }
        }
        tab[row] = ri;
      }
      __syncwarp();
      if (lane == 0) st_release_shared(&s_ready[c.i & (kRowRing - 1)], (uint32_t)(c.i + 1));
    };
    // geometry of this lane's chunk column in the cursor's k-block
    struct ChunkGeom { uint32_t tapbit; long long delta; int c8; };
    auto chunk_geom = [&](const Cursor& c) {
      const TapTable& tt = g.phase[c.phase];
      const int q = c.kb * 8 + j;
      const int tap_i = q / P.cpt;
      ChunkGeom cg;
      cg.c8 = q - tap_i * P.cpt;
      cg.tapbit = 0u;
      cg.delta = 0;
      if (tap_i < tt.ntaps) {
        cg.tapbit = 1u << tap_i;
        cg.delta = ((long long)tt.dh[tap_i] * g.Win + tt.dw[tap_i]) * g.Cin + cg.c8 * 8;
      }
      return cg;
    };
    // byte offset of (row rsub + 4*i, chunk j) inside a swizzled A stage
    auto dst_of = [&](int i) -> uint32_t {
      const uint32_t row = (uint32_t)(rsub + 4 * i);
      return row * 128u + ((((uint32_t)j) ^ (row & 7u)) << 4);
    };
    // a k-block may be issued when (a) its ring slot has been consumed -- producer warps
    // run out of order with respect to each other, so a 1-bit mbarrier parity could
    // alias; the MMA warp publishes an exact completed-k-block count instead -- and
    // (b) its tile's row table exists.  Never BLOCK here: the slot may only free up
    // after the MMA warp has consumed what this warp still holds unpublished.
    auto can_issue = [&](const Cursor& c) -> bool {
      if (c.seq >= kStages &&
          ld_acquire_shared(const_cast<const uint32_t*>(s_consumed)) < (uint32_t)(c.seq - kStages + 1))
        return false;
      return c.kb == 0 || ld_acquire_shared(&s_ready[c.i & (kRowRing - 1)]) == (uint32_t)(c.i + 1);
    };
    auto issue = [&](Cursor& c) {
      VAE_TRACE(0, c.seq);
      if (c.kb == 0) build_rows(c);
      const int stage = c.seq % kStages;
      uint8_t* a_st = smem + stage * stage_bytes;
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(N * 128));
        bulk_g2s(a_st + kABytes, P.wpk + P.b_off[c.phase] + (long long)c.kb * N * 64, (uint32_t)(N * 128),
                 &full_bar[stage]);
      }
      const ChunkGeom cg = chunk_geom(c);
      const uint32_t a_dst = smem_u32(a_st);
      const RowInfo* rows = s_rows + (c.i & (kRowRing - 1)) * kTileM + rsub;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const RowInfo ri = rows[4 * i];
        const bool ok = (ri.mask & cg.tapbit) != 0u;
        const bf16* src = ok ? P.in + ((long long)ri.off + cg.delta) : P.in;
        cp_async_16_zfill(a_dst + dst_of(i), src, ok ? 16u : 0u);
      }
      if (fuse) {
        cp_async_commit();
      } else {
        // no transform needed: let the copies themselves arrive on the barrier when they
        // land (no wait, no fence in this warp; the MMA warp fences the proxy once)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
      }
      VAE_TRACE(1, c.seq);
    };
    auto publish = [&](Cursor& c) {
      const int stage = c.seq % kStages;
      if (fuse) {  // previous layer's BatchNorm + LeakyReLU, in place on this lane's own chunks
        const ChunkGeom cg = chunk_geom(c);
        uint8_t* a_st = smem + stage * stage_bytes;
        const RowInfo* rows = s_rows + (c.i & (kRowRing - 1)) * kTileM + rsub;
        float sc[8], mn[8], bt[8];  // the lane's 8 channels are the same for all of its rows
        {
          const float4* ps = reinterpret_cast<const float4*>(s_pro + cg.c8 * 8);
          const float4* pm = reinterpret_cast<const float4*>(s_pro + g.Cin + cg.c8 * 8);
          const float4* pb = reinterpret_cast<const float4*>(s_pro + 2 * g.Cin + cg.c8 * 8);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 a = ps[h], b = pm[h], d = pb[h];
            sc[4 * h] = a.x; sc[4 * h + 1] = a.y; sc[4 * h + 2] = a.z; sc[4 * h + 3] = a.w;
            mn[4 * h] = b.x; mn[4 * h + 1] = b.y; mn[4 * h + 2] = b.z; mn[4 * h + 3] = b.w;
            bt[4 * h] = d.x; bt[4 * h + 1] = d.y; bt[4 * h + 2] = d.z; bt[4 * h + 3] = d.w;
          }
        }
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          if (rows[4 * i].mask & cg.tapbit) {
            uint4* p4 = reinterpret_cast<uint4*>(a_st + dst_of(i));
            *p4 = bn_lrelu_8(*p4, sc, mn, bt, 0, P.slope);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[stage]);
    };

    Cursor head;
    head.seq = pw; head.i = 0; head.kb = pw; head.nkb = 0; head.phase = 0;
    tile_info(head);
    while (head.i < n_my && head.kb >= head.nkb) {  // position on k-block number pw
      head.kb -= head.nkb;
      ++head.i;
      tile_info(head);
    }
    Cursor tail = head;
    int in_flight = 0;
    for (;;) {
      const bool head_live = head.i < n_my;
      if (head_live && in_flight <= kWarpDepth && can_issue(head)) {
        issue(head);
        advance(head);
        if (fuse) ++in_flight;
      } else if (in_flight > 0) {
        if (in_flight == kWarpDepth + 1) cp_async_wait<kWarpDepth>();
        else cp_async_wait<0>();
        publish(tail);
        advance(tail);
        --in_flight;
      } else if (!head_live) {
        break;
      }
    }
  } else if (warp == 12) {
    // ============================ MMA issuer =============================
    const uint32_t idesc = make_idesc_bf16(kTileM, (uint32_t)N, 0, 0);
    int stage = 0;
    uint32_t ph = 0;
    int prev_stage = -1;
    uint32_t prev_ph = 0, seq = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int phase = tile / P.tiles_per_phase;
      const int nkb = P.nkb[phase];
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], ph);
        VAE_TRACE(2, seq);
        fence_proxy_async_smem();  // cp.async (generic proxy) writes -> visible to the tensor core
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
            mma_bf16_ss(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(&empty_bar[stage]);              // completes when these MMAs have read the stage
          if (kb == nkb - 1) mma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        VAE_TRACE(3, seq);
        // retire the PREVIOUS k-block (its MMAs overlap the ones just issued) and publish the
        // exact count of consumed k-blocks to the producer warps
        if (prev_stage >= 0) {
          mbar_wait(&empty_bar[prev_stage], prev_ph);
          if (lane == 0) st_release_shared(const_cast<uint32_t*>(s_consumed), seq);
        }
        prev_stage = stage; prev_ph = ph;
        ++seq;
        if (++stage == kStages) { stage = 0; ph ^= 1u; }
      }
    }
    if (prev_stage >= 0) {
      mbar_wait(&empty_bar[prev_stage], prev_ph);
      if (lane == 0) st_release_shared(const_cast<uint32_t*>(s_consumed), seq);
    }
  } else {
    // ============================= epilogue ==============================
    const int ew = warp - 8;             // TMEM lane group = warp % 4
    const int r = ew * 32 + lane;
    float run1[32], run2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) run1[i] = run2[i] = 0.0f;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int phase = tile / P.tiles_per_phase;
      const int mt = tile - phase * P.tiles_per_phase;
      const TapTable& tt = g.phase[phase];
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const long long m = (long long)mt * kTileM + r;
      const bool valid = m < M_total;
      bf16* dst = nullptr;
      if (valid) {
        int jw = (int)(m % g.Wj);
        long long t2 = m / g.Wj;
        int jh = (int)(t2 % g.Hj);
        int b = (int)(t2 / g.Hj);
        int oh = jh * g.out_mul + tt.ph, ow = jw * g.out_mul + tt.pw;
        dst = P.out + (((long long)b * g.Hout + oh) * g.Wout + ow) * g.Nout;
      }
      mbar_wait(&tfull_bar[acc], acc_ph);
      if (ew == 0) VAE_TRACE(4, it);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * N);
      for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld_32x32(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= N) {  // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tempty_bar[acc]);
          if (ew == 0) VAE_TRACE(5, it);
        }
        const int ncol = min(32, N - c0);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[(c0 + i) & 255]);
          v[i] = (valid && i < ncol) ? v[i] + b4.x : 0.0f;
          v[i + 1] = (valid && i + 1 < ncol) ? v[i + 1] + b4.y : 0.0f;
          v[i + 2] = (valid && i + 2 < ncol) ? v[i + 2] + b4.z : 0.0f;
          v[i + 3] = (valid && i + 3 < ncol) ? v[i + 3] + b4.w : 0.0f;
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
          if (N <= 32) {
            // narrow layers have thousands of tiles per CTA: keep the column sums of this
            // thread's rows in registers and reduce across the warp once, after the loop
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
            if (lane < ncol) {
              atomicAdd(&s_stats[c0 + lane], v[0]);
              atomicAdd(&s_stats[N + c0 + lane], sq[0]);
            }
          }
        }
      }
    }
    if (P.stats != nullptr && N <= 32) {
      warp_transpose_reduce32(run1, lane);
      warp_transpose_reduce32(run2, lane);
      if (lane < N) {
        atomicAdd(&s_stats[lane], run1[0]);
        atomicAdd(&s_stats[N + lane], run2[0]);
      }
    }
    if (P.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
      for (int i = threadIdx.x - kProducers; i < 2 * N; i += 128) atomicAdd(&P.stats[i], (double)s_stats[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// ============================================================================
// host side
// ============================================================================
static bool g_umma_enabled = true;
static long long* g_trace = nullptr;
long long* umma_trace_ptr() { return g_trace; }

static bool umma_shape_ok(int Cin_gemm, int Nout) {
  return g_umma_enabled && Cin_gemm % 32 == 0 && Nout % 16 == 0 && Nout >= 16 && Nout <= 256 && Cin_gemm <= 2048;
}

static void fill_params(UmmaGatherParams& P, const GatherGeom& g) {
  P.g = g;
  P.cpt = g.Cin / 8;
  P.N = g.Nout;
  P.b_stage_bytes = (g.Nout * 128 + 1023) / 1024 * 1024;
  int cols = 32;
  while (cols < 2 * g.Nout) cols *= 2;
  P.tmem_cols = cols;
  P.stages = (170 * 1024) / (kABytes + P.b_stage_bytes);
  if (P.stages > kMaxStages) P.stages = kMaxStages;
  P.depth = P.stages - 2;
  if (P.depth > 4) P.depth = 4;
  if (P.depth < 2) P.depth = 2;
  long long M = (long long)g.B * g.Hj * g.Wj;
  P.tiles_per_phase = (int)((M + kTileM - 1) / kTileM);
  P.total_tiles = P.tiles_per_phase * g.nphase;
  int off = 0;
  for (int p = 0; p < 4; ++p) {
    if (p < g.nphase) {
      P.nkb[p] = (g.phase[p].ntaps * P.cpt + 7) / 8;
      P.b_off[p] = off;
      off += P.nkb[p] * g.Nout * 64;
    } else {
      P.nkb[p] = 0;
      P.b_off[p] = off;
    }
  }
}

static long long packed_elems(const UmmaGatherParams& P) {
  long long n = 0;
  for (int p = 0; p < P.g.nphase; ++p) n += (long long)P.nkb[p] * P.N * 64;
  return n;
}

static int launch_umma(UmmaGatherParams& P, cudaStream_t st) {
  size_t smem = 1024 + (size_t)P.stages * (kABytes + P.b_stage_bytes) + 256 + 16 + 2 * 256 * 4 + 3 * 2048 * 4 +
                kRowRing * kTileM * sizeof(RowInfo) + 256 * 4 + kRowRing * 4;
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(umma_gather_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      set_error("umma: smem attribute: %s", cudaGetErrorString(e));
      return VAE_ERR_CUDA;
    }
    attr_bytes = 227 * 1024;
  }
  if (smem > 227 * 1024) {
    set_error("umma: tile needs %zu bytes of shared memory", smem);
    return VAE_ERR_UNSUPPORTED;
  }
  int grid = sm_count();
  if (grid > P.total_tiles) grid = P.total_tiles;
  umma_gather_gemm_kernel<<<grid, kThreads, smem, st>>>(P);
  return check_launch("umma_gather_gemm");
}

}  // namespace vae

using namespace vae;

extern "C" int vae_has_umma(void) { return 1; }

/* tuning aid: device buffer of 6*512 int64 that CTA 0 of vae_conv_umma fills with clock64()
 * samples per k-block / tile (NULL switches it off) */
extern "C" int vae_umma_set_trace(void* buf) {
  g_trace = (long long*)buf;
  return VAE_OK;
}

extern "C" int vae_set_umma(int enable) {
  int old = g_umma_enabled ? 1 : 0;
  g_umma_enabled = enable != 0;
  return old;
}

/* number of bf16 elements vae_umma_pack_weight writes (0 if the shape is not
 * served by the tcgen05 path) */
extern "C" long long vae_umma_packed_elems(int Ci, int Co, int stride, int mode) {
  GatherGeom g = mode == 0 ? geom_fprop(1, 2 * stride, 2 * stride, Ci, Co, stride)
                           : geom_dgrad(1, 2 * stride, 2 * stride, Ci, Co, stride);
  if (!umma_shape_ok(g.Cin, g.Nout)) return 0;
  UmmaGatherParams P{};
  fill_params(P, g);
  return packed_elems(P);
}

extern "C" int vae_umma_pack_weight(const float* w_oihw, void* out, int Ci, int Co, int stride, int mode,
                                    vae_stream_t stream) {
  VAE_REQUIRE(w_oihw && out && (mode == 0 || mode == 1) && (stride == 1 || stride == 2), "umma_pack_weight: bad argument");
  GatherGeom g = mode == 0 ? geom_fprop(1, 2 * stride, 2 * stride, Ci, Co, stride)
                           : geom_dgrad(1, 2 * stride, 2 * stride, Ci, Co, stride);
  VAE_REQUIRE(umma_shape_ok(g.Cin, g.Nout), "umma_pack_weight: shape not served by the tcgen05 path");
  UmmaGatherParams P{};
  fill_params(P, g);
  dim3 grid(64, g.nphase);
  pack_umma_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(w_oihw, (bf16*)out, P, Co, Ci, mode);
  return check_launch("pack_umma_weight");
}

/* Conv forward (mode 0) / data gradient == ConvTranspose forward (mode 1) on the
 * tcgen05 path.  `in`, `out` NHWC bf16; wpk from vae_umma_pack_weight(same Ci, Co,
 * stride, mode).  Hi, Wi are the HIGH-resolution (conv input side) extents as in
 * vae_conv_fprop / vae_conv_dgrad.  pro_scale/pro_mean/pro_beta (nullable, fp32
 * [C_in_gemm]) fuse the previous layer's BatchNorm + LeakyReLU into the operand load. */
extern "C" int vae_conv_umma(const void* in, const void* wpk, const float* bias, void* out, double* stats,
                             const float* pro_scale, const float* pro_mean, const float* pro_beta, float slope,
                             int B, int Hi, int Wi, int Ci, int Co, int stride, int mode, vae_stream_t stream) {
  VAE_REQUIRE(in && wpk && out && B > 0 && (stride == 1 || stride == 2) && Hi % stride == 0 && Wi % stride == 0 &&
                  (mode == 0 || mode == 1) && ((pro_scale == nullptr) == (pro_mean == nullptr)) &&
                  ((pro_scale == nullptr) == (pro_beta == nullptr)),
              "conv_umma: bad argument");
  GatherGeom g = mode == 0 ? geom_fprop(B, Hi, Wi, Ci, Co, stride) : geom_dgrad(B, Hi, Wi, Ci, Co, stride);
  VAE_REQUIRE(umma_shape_ok(g.Cin, g.Nout), "conv_umma: shape (Cin=%d, N=%d) not served by the tcgen05 path", g.Cin, g.Nout);
  VAE_REQUIRE((long long)g.B * g.Hin * g.Win * g.Cin < (1LL << 32) && (long long)g.B * g.Hj * g.Wj < (1LL << 31),
              "conv_umma: tensor too large for 32-bit row offsets");
  cudaStream_t st = as_stream(stream);
  if (stats) VAE_CUDA(zero_scratch(stats, sizeof(double) * 2 * g.Nout, st));
  UmmaGatherParams P{};
  fill_params(P, g);
  P.in = (const bf16*)in; P.wpk = (const bf16*)wpk; P.bias = bias; P.out = (bf16*)out; P.stats = stats;
  P.pro_scale = pro_scale; P.pro_mean = pro_mean; P.pro_beta = pro_beta; P.slope = slope;
  P.trace = g_trace;
  return launch_umma(P, st);
}

// hooks used by conv_simt.cu's generic entry points: the packed-weight tcgen05
// path has its own entry point (vae_conv_umma), so these never take over.
namespace vae {
int umma_conv_fprop(const void*, const void*, const float*, void*, double*, int, int, int, int, int,
                    int, cudaStream_t, bool* handled) { *handled = false; return VAE_OK; }
int umma_conv_dgrad(const void*, const void*, const float*, void*, double*, int, int, int, int, int,
                    int, cudaStream_t, bool* handled) { *handled = false; return VAE_OK; }
int umma_conv_wgrad(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t,
                    bool* handled) { *handled = false; return VAE_OK; }
}  // namespace vae
