#!/bin/bash
# A/B builds: scripts/build_variant.sh NAME [extra nvcc flags...] -> build/ab/NAME.so (matching.cu / dem.cu / api.cu rebuilt
# with the flags, the other objects reused).  build/ is git-ignored but travels to the GPU box.  Use with
# scripts/lib_bench.py build/ab/NAME.so
set -e
cd "$(dirname "$0")/../graphqec_b200/csrc"
name=$1; shift
out=../../build/ab; mkdir -p $out/$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for f in matching dem api; do $NV "$@" -c $f.cu -o $out/$name/$f.o & done
wait
$NV -shared -o $out/$name.so $out/$name/matching.o $out/$name/dem.o $out/$name/api.o bp.o circuit_dem.o -lcudart
echo built $out/$name.so
