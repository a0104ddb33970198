#!/usr/bin/env bash
# Build libvae_b200.so (the C-ABI library) for sm_100a, in-tree.  Incremental: a source is recompiled
# when it, a header next to it, or the public header is newer than its object (FORCE=1 rebuilds all).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libvae_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v -Wno-deprecated-gpu-targets)
mkdir -p "${HERE}/obj"
newest_hdr=0
for h in "${HERE}"/*.cuh "${HERE}/../../include/vae_b200.h" "${HERE}/build.sh"; do
  t=$(stat -c %Y "$h"); [ "$t" -gt "$newest_hdr" ] && newest_hdr=$t
done
pids=(); names=()
for src in "${HERE}"/*.cu; do
  f="$(basename "${src}" .cu)"
  obj="${HERE}/obj/${f}.o"
  if [ "${FORCE:-0}" = "1" ] || [ ! -f "$obj" ] || [ "$(stat -c %Y "$src")" -gt "$(stat -c %Y "$obj")" ] || [ "$newest_hdr" -gt "$(stat -c %Y "$obj")" ]; then
    "${NVCC}" "${FLAGS[@]}" -c "${src}" -o "${obj}" > "${HERE}/obj/${f}.log" 2>&1 &
    pids+=($!); names+=("$f")
  fi
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then fail=1; echo "== ${names[$i]}.cu failed"; grep -E "error|Error" -A3 "${HERE}/obj/${names[$i]}.log" | head -40; rm -f "${HERE}/obj/${names[$i]}.o"; fi
done
[ "$fail" -ne 0 ] && exit 1
# objects whose source was deleted must not be linked
for obj in "${HERE}"/obj/*.o; do [ -f "${HERE}/$(basename "$obj" .o).cu" ] || rm -f "$obj"; done
"${NVCC}" -shared -o "${OUT}" "${HERE}"/obj/*.o -lcudart
echo "built ${OUT} (${#pids[@]} recompiled: ${names[*]:-none})"
