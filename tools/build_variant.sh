#!/bin/bash
# development: build a variant of libobca.so with extra flags into build/libobca_<name>.so (fast build: <2,true> only)
# usage: tools/build_variant.sh name [-DOBCA_...=..] [FMAD=true]
name=$1; shift
fmad=true
args=()
for a in "$@"; do if [ "$a" = "FMAD=true" ]; then fmad=true; else args+=("$a"); fi; done
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=$fmad -diag-suppress 177 -Xcompiler -fPIC -shared -DOBCA_FAST_BUILD "${args[@]}" \
  -o build/libobca_$name.so obca_b200/csrc/obca_lib.cu
