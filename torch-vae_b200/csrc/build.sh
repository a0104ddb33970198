#!/usr/bin/env bash
# Build libvae_b200.so (the C-ABI library) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libvae_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "${HERE}/obj"
pids=()
for f in runtime adamw elementwise elbo conv_simt conv_edge final_block final_block_tc heads gemm_simt conv_umma conv_umma_tma conv_umma_rw conv_umma_wgrad conv_wgrad_tma; do
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/obj/${f}.o" > "${HERE}/obj/${f}.log" 2>&1 &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
if [ "$fail" -ne 0 ]; then cat "${HERE}"/obj/*.log | grep -E "error|Error" -A3 | head -80; exit 1; fi
"${NVCC}" -shared -o "${OUT}" "${HERE}"/obj/*.o -lcudart
echo "built ${OUT}"
