#!/bin/bash
# build a tuning variant of the library: tools/build_variant.sh <tag> [-D flags...]
# only the 3-D mixed hyper_J2 / small_J2 combos are recompiled; output calibr8_b200/lib/variants/libc8b200_<tag>.so
set -e
tag=$1; shift
cd "$(dirname "$0")/../calibr8_b200/csrc"
out=../lib/variants; mkdir -p $out/obj_$tag
for f in combo_3d_mixed_hyper_j2 combo_3d_mixed_small_j2; do
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" \
    -I../../include -I../host -c $f.cu -o $out/obj_$tag/$f.o 2> $out/obj_$tag/$f.ptxas.log &
done
wait
others=$(ls ../lib/obj/*.o | grep -v "combo_3d_mixed_hyper_j2\|combo_3d_mixed_small_j2")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libc8b200_$tag.so $out/obj_$tag/*.o $others -lcudart
grep -A2 "k_forward_jacobian" $out/obj_$tag/*.ptxas.log | grep "registers\|spill" 
