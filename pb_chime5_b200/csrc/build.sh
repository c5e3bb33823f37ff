#!/bin/bash
# Builds, in-tree (the .so files travel to the GPU box), for sm_100a:
#   libgss.so      the product: the C ABI of include/gss.h, nothing else
#   libgss_dev.so  libgss.so + the developer / measurement entry points of include/gss_dev.h
#                  (gss_debug_*: loaded by tests/, tools/ and bench.py's side measurements only)
# Usage: build.sh [fast]   -- "fast" restricts the CACGMM channel variants (developer loop).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
EXTRA=""
if [ "$1" = "fast" ]; then EXTRA="-DGSS_DP_LIST=GSS_CASE(4)GSS_CASE(8)GSS_CASE(24)"; fi
DEV_ONLY="mstep_i8.cu probe.cu"          # sources that exist only in libgss_dev.so
mkdir -p build
pids=()
stale() {  # stale <src> <obj> [extra dependency]
  [ ! -f "$2" ] || [ "$1" -nt "$2" ] || [ -n "$3" -a "$3" -nt "$2" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$2")" ] \
    || [ ../../include/gss.h -nt "$2" ] || [ ../../include/gss_dev.h -nt "$2" ] \
    || [ -n "$EXTRA" -a ! -f build/.fast ] || [ -z "$EXTRA" -a -f build/.fast ]
}
for f in *.cu; do
  o=build/${f%.cu}.o
  dep=""; case "$f" in cacgmm_part*) dep=cacgmm.cu;; esac
  if stale "$f" "$o" "$dep"; then ( $NVCC $FLAGS $EXTRA -c "$f" -o "$o" ) & pids+=($!); fi
done
# wpe.cu carries gss_debug_wpe_gram behind GSS_DEV_API (it needs the file-local launch helpers)
if stale wpe.cu build/dev_wpe.o; then ( $NVCC $FLAGS $EXTRA -DGSS_DEV_API -c wpe.cu -o build/dev_wpe.o ) & pids+=($!); fi
for p in "${pids[@]}"; do wait $p; done
if [ -n "$EXTRA" ]; then touch build/.fast; else rm -f build/.fast; fi
PROD=""; DEV=""
for f in *.cu; do
  o=build/${f%.cu}.o
  case " $DEV_ONLY " in *" $f "*) DEV="$DEV $o";; *) PROD="$PROD $o"; [ "$f" = wpe.cu ] && DEV="$DEV build/dev_wpe.o" || DEV="$DEV $o";; esac
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libgss.so $PROD -lcudart
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libgss_dev.so $DEV -lcudart
echo "built $(pwd)/libgss.so $(pwd)/libgss_dev.so"
