// TMA helpers shared by the tensor-map fed tcgen05 kernels: the 4-D tiled load, swizzle-aware
// shared-memory descriptors, the driver entry point for cuTensorMapEncodeTiled and the
// "whole output rows" tile decomposition.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_geom.cuh"
#include "umma.cuh"

namespace vae {

using namespace umma;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 16-byte read-only global load that stays where it is written: as `asm volatile` it is ordered against the
// (volatile) mbarrier waits, so a load issued before a wait really flies during the wait (a plain __ldg is sunk
// by the compiler to its first use, behind the wait)
__device__ __forceinline__ uint4 ldg_nc_v4_pinned(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 lds_f4_addr(uint32_t a) {   // LDS.128 by 32-bit shared address (no generic LD)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// the same for data that never changes after the kernel's set-up barrier (bias, BatchNorm vectors): not volatile, so the
// compiler may hoist it above waits
__device__ __forceinline__ float4 lds_f4_ro(uint32_t a) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// 16-byte store to GLOBAL memory (a pointer rebuilt from integers / shuffles would otherwise be stored through generically)
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// shared-memory matrix descriptor with a selectable swizzle (2 = 128B, 4 = 64B)
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tile decomposition 128 = rows_n * rows_h * rows_w made of whole output rows
static inline bool tile_rows(const GatherGeom& g, int* rw, int* rh, int* rn) {
  if (g.Wj >= 128) {
    if (g.Wj % 128) return false;
    *rw = 128; *rh = 1; *rn = 1;
    return true;
  }
  if (128 % g.Wj) return false;
  int rows = 128 / g.Wj;
  if (rows <= g.Hj) {
    if (g.Hj % rows) return false;
    *rw = g.Wj; *rh = rows; *rn = 1;
    return true;
  }
  if (rows % g.Hj) return false;
  *rw = g.Wj; *rh = g.Hj; *rn = rows / g.Hj;
  return true;
}


}  // namespace vae
