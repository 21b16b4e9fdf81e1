#!/bin/bash
# builds nexus_b200/variants/lib_<tag>.so with extra -D flags (kernel-variant experiments; select with NEXUS_B200_LIB)
TAG=$1; shift
D=nexus_b200/csrc; O=$D/build/var_$TAG; mkdir -p $O nexus_b200/variants
for f in context bvh_builder scene render shade; do
  X=; [ $f = shade ] && X=--use_fast_math
  nvcc $X -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" -c $D/$f.cu -o $O/$f.o 2> $O/$f.log &
done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o nexus_b200/variants/lib_$TAG.so $O/*.o
grep -E "trace_(closest|any)_kernelILb0" -A2 $O/render.log | grep -E "Used" 
