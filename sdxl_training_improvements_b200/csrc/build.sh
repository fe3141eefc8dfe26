#!/bin/bash
# Build libsdxl_b200.so for sm_100a (in-tree, so it travels with the gpurun snapshot).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr ${B2_NVCC_EXTRA}"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/obj/$(basename "${f%.cu}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/tc.cuh" -nt "$o" ] || [ "$HERE/../../include/sdxl_b200.h" -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT/libsdxl_b200.so" "$HERE"/obj/*.o -lcudart_static -ldl -lpthread -lrt
echo "built $OUT/libsdxl_b200.so"
