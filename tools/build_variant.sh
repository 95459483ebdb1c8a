#!/bin/bash
# Developer tool: build a variant of the library whose z-marching gather kernels (csrc/march_*.cu) are
# compiled with extra macros; everything else is linked from the default build's objects.
# usage: tools/build_variant.sh NAME -DBMQ_MARCH_BY=8 -DBMQ_MARCH_MINBLOCKS=3 ...
# result: gpufluidsimulation_b200/lib/variants/NAME.so (load it with BMQ_LIB=...; tools/ab_variants.sh times all)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/gpufluidsimulation_b200/csrc
out=$root/gpufluidsimulation_b200/lib/variants
tmp=/tmp/bmq_variant_$name
mkdir -p "$out" "$tmp"
make -s -C "$src" >/dev/null
arch="-gencode arch=compute_100a,code=sm_100a"
pids=()
for f in march_advect march_error march_cumulate march_apply; do
  nvcc -std=c++17 -O3 $arch -lineinfo -Xcompiler -fPIC "$@" -c "$src/$f.cu" -o "$tmp/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
others=$(ls "$src"/build/*.o | grep -v '/march_')
nvcc $arch -shared -o "$out/$name.so" $others "$tmp"/march_*.o -lcudart
echo "$out/$name.so"
