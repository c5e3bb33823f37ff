#!/bin/bash
# Builds pb_chime5_b200/csrc/libgss.so for sm_100a (in-tree; the .so travels to the GPU box).
# Usage: build.sh [fast]   -- "fast" restricts the CACGMM channel variants (developer loop).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
EXTRA=""
if [ "$1" = "fast" ]; then EXTRA="-DGSS_DP_LIST=GSS_CASE(4)GSS_CASE(8)GSS_CASE(24)"; fi
mkdir -p build
pids=()
for f in *.cu; do
  o=build/${f%.cu}.o
  dep=""; case "$f" in cacgmm_part*) dep=cacgmm.cu;; esac
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$dep" -a "$dep" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o")" ] || [ ../../include/gss.h -nt "$o" ] || [ -n "$EXTRA" -a ! -f build/.fast ] || [ -z "$EXTRA" -a -f build/.fast ]; then
    ( $NVCC $FLAGS $EXTRA -c "$f" -o "$o" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
if [ -n "$EXTRA" ]; then touch build/.fast; else rm -f build/.fast; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libgss.so build/*.o -lcudart
echo "built $(pwd)/libgss.so"
