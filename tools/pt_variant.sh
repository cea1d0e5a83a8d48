#!/bin/bash
# developer aid: build a variant of the library that differs only in rb_pathtracer.cu's compile flags
#   tools/pt_variant.sh NAME [-DFLAG ...]   ->  rendu_b200/librendu_b200_variant_NAME.so  (other objects from csrc/build)
set -e
cd "$(dirname "$0")/../rendu_b200/csrc"
name=$1; shift
mkdir -p build_v
/usr/local/cuda/bin/nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-O2 -Xptxas -v -c rb_pathtracer.cu -o build_v/rb_pathtracer_$name.o 2> build_v/rb_pathtracer_$name.log
objs=$(ls build/*.o | grep -v rb_pathtracer.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../librendu_b200_variant_$name.so build_v/rb_pathtracer_$name.o $objs -cudart static
grep -A2 "rb_k_pt_shade" build_v/rb_pathtracer_$name.log | grep -E "stack|Used" | tr '\n' ' '; echo
