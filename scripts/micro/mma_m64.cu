// Micro-benchmark + layout probe: tcgen05.mma kind::f16 with M = 64 (cta_group::1), MN-major 64B-swizzled
// operands as in the output-block backward kernel.  (1) cycles per MMA for M in {64,128}, N in {16,32};
// (2) which TMEM lane holds accumulator row m when M = 64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_m64 mma_m64.cu && ./mma_m64
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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

__global__ void __launch_bounds__(128, 1) k(int M, int N, int reps, long long* out, float* probe) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // A: MN-major, [16 k][64 B] per 32-channel group, groups 8 KB apart (LBO); k = 0 row holds m + 1
  if (threadIdx.x < 128) {
    const int m = threadIdx.x;
    __nv_bfloat16* a = (__nv_bfloat16*)(smem + (m / 32) * 8192);
    a[m % 32] = __float2bfloat16((float)(m + 1));
  }
  if (threadIdx.x == 0) ((__nv_bfloat16*)(smem + 64 * 1024))[0] = __float2bfloat16(1.0f);   // B[k=0][n=0] = 1
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&tptr, 64);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x < 32) {
    const uint32_t idesc = make_idesc_bf16((uint32_t)M, (uint32_t)N, 1, 1);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r)
        mma_bf16_ss(tm, desc(a0 + (uint32_t)(r & 7) * 1024u * (reps > 1), 8192, 512, 4), desc(b0 + (uint32_t)(r & 7) * 1024u * (reps > 1), 8192, 512, 4), idesc, r > 0);
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (probe != nullptr && blockIdx.x == 0) {
    float v[32];
    tmem_ld_32x32(tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    tmem_ld_wait();
    probe[threadIdx.x] = v[0];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  float* p; cudaMalloc(&p, 128 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  for (int M : {64, 128})
    for (int N : {16, 32, 64}) {
      long long h = 0;
      for (int it = 0; it < 2; ++it) {
        k<<<148, 128, 100 * 1024>>>(M, N, 512, d, nullptr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      }
      printf("M %3d N %3d MN-major SW64: %7.1f cycles / MMA\n", M, N, (double)h / 512);
    }
  for (int M : {64, 128}) {
    cudaMemset(p, 0, 512);
    k<<<1, 128, 100 * 1024>>>(M, 16, 1, d, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float h[128];
    cudaMemcpy(h, p, 512, cudaMemcpyDeviceToHost);
    printf("M %d: TMEM lane -> accumulator row + 1 (0 = unused):\n", M);
    for (int i = 0; i < 128; ++i) printf("%3.0f%s", h[i], (i % 32 == 31) ? "\n" : " ");
  }
  return 0;
}
