#!/usr/bin/env bash
# Builds libcdc_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libcdc_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default)
OBJ="${HERE}/../build"
mkdir -p "${OBJ}"
pids=()
for f in conv_tc conv_kf elementwise attention entropy runner; do
  if [ ! -f "${OBJ}/${f}.o" ] || [ "${HERE}/${f}.cu" -nt "${OBJ}/${f}.o" ] || [ -n "$(find "${HERE}" "${HERE}/../../include" -name '*.cuh' -newer "${OBJ}/${f}.o" -o -name '*.h' -newer "${OBJ}/${f}.o" 2>/dev/null | head -1)" ]; then
    "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${OBJ}/${f}.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${OUT}" "${OBJ}"/conv_tc.o "${OBJ}"/conv_kf.o "${OBJ}"/elementwise.o "${OBJ}"/attention.o "${OBJ}"/entropy.o "${OBJ}"/runner.o
echo "built ${OUT}"
