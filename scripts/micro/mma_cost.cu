// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of N, operand majors
// and swizzle, operands resident in shared memory (contents irrelevant).  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_cost mma_cost.cu && ./mma_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../torch-vae_b200/csrc/umma.cuh"
using namespace vae::umma;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) k(int N, int a_mn, int b_mn, int layout, int reps, long long* out, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[8];
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 128 * 1024 / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&tptr, 256);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x < 32) {
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)N, (uint32_t)a_mn, (uint32_t)b_mn);
    const uint32_t row = layout == 4 ? 64u : 128u;  // bytes per row
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        // K-major: advance 32 B per k-step inside the swizzle atom; MN-major: 16 rows per k-step
        const uint32_t ka = a_mn ? (uint32_t)(r & 7) * 16u * row : (uint32_t)(r & (row / 32 - 1)) * 32u;
        const uint32_t kb = b_mn ? (uint32_t)(r & 7) * 16u * row : (uint32_t)(r & (row / 32 - 1)) * 32u;
        mma_bf16_ss(tm, desc(a0 + ka, 8192, 8 * row, layout), desc(b0 + kb, 8192, 8 * row, layout), idesc, r > 0);
        if (commit_every > 0 && (r % commit_every) == commit_every - 1) mma_commit(&dummy[(r / commit_every) & 7]);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    const long long tt = __shfl_sync(0xffffffffu, t0, 0);
    (void)tt;
    if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 512;
  for (int ce : {0, 8, 2, 1})
  for (int layout : {4})
    for (int am = 1; am < 2; ++am)
      for (int bm = 1; bm < 2; ++bm)
        for (int N : {32, 256}) {
          long long h = 0;
          printf("commit every %d MMAs: ", ce);
          for (int it = 0; it < 2; ++it) {
            k<<<148, 128, 130 * 1024>>>(N, am, bm, layout, reps, d, ce);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          }
          printf("swizzle %s  A %s  B %s  N %3d : %7.1f cycles / MMA\n", layout == 4 ? " 64B" : "128B", am ? "MN" : "K ",
                 bm ? "MN" : "K ", N, (double)h / reps);
        }
  return 0;
}
