#!/bin/bash
# development: default / rounds-only / tail-only timings for every build/libobca_*.so
for so in build/libobca_*.so; do
  echo "== $so"
  OBCA_SO=$PWD/$so python tools/gpu_one.py ${B:-4096} 3 2>&1 | tail -1
  OBCA_SO=$PWD/$so OBCA_MODE=2 OBCA_TAIL_THRESH=0 python tools/gpu_one.py ${B:-4096} 3 2>&1 | tail -1
  OBCA_SO=$PWD/$so OBCA_MODE=1 python tools/gpu_one.py ${B:-4096} 3 2>&1 | tail -1
done
