#!/bin/bash
# Build an experiment copy of libdpsttc.so with extra -D flags for ONE translation unit:
#   tools/build_variant.sh <name> <file.cu> "<-DFOO=1 -DBAR=2>"   →  dps_ttc_b200/build_variants/libdpsttc_<name>.so
# Run with DPSTTC_LIB=<that path> (dps_ttc_b200/_lib.py honours it).  Needs a prior `make` (links the other objects).
set -e
name=$1; src=$2; flags=$3
cd "$(dirname "$0")/../dps_ttc_b200/csrc"
mkdir -p ../build_variants build
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static $flags -c -o build/${src%.cu}_$name.o $src
objs=""
for o in api update operator project inpaint blur_separable blur_fused blur_sparse resize resize_fused phase resample; do
  if [ "$o.cu" == "$src" ]; then objs="$objs build/${o}_$name.o"; else objs="$objs build/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC -o ../build_variants/libdpsttc_$name.so $objs
echo ../build_variants/libdpsttc_$name.so
