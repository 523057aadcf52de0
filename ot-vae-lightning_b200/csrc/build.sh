#!/usr/bin/env bash
# Builds libotk.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
cd "$HERE"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       -Xptxas -v -rdc=false ${OTK_EXTRA_NVCC_FLAGS:-})
SRCS=$(ls *.cu)
mkdir -p build
pids=()
for f in $SRCS; do
  o="build/${f%.cu}.o"
  if [[ ! -f "$o" || "$f" -nt "$o" || -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" -print -quit)" || ../../include/otk.h -nt "$o" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "build/${f%.cu}.log" 2>&1 || { cat "build/${f%.cu}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o libotk.so build/*.o -lcudart_static -lpthread -ldl -lrt
echo "built $HERE/libotk.so"
