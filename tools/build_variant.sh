#!/bin/bash
# development: build a variant of libobca.so with extra -D flags into build/libobca_<name>.so (fast build: <2,true> only)
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -diag-suppress 177 -Xcompiler -fPIC -shared -DOBCA_FAST_BUILD "$@" \
  -o build/libobca_$name.so obca_b200/csrc/obca_lib.cu
