#!/bin/bash
# development: time every build/libobca_*.so variant: full-load rounds (hand-over at B/2), rounds only, default schedule
for so in build/libobca_*.so; do
  echo "== $so"
  OBCA_SO=$PWD/$so python tools/gpu_one.py ${B:-4096} 2 2048 t 2>&1 | tail -1
  OBCA_SO=$PWD/$so python tools/gpu_one.py ${B:-4096} 2 0 t 2>&1 | tail -1
  OBCA_SO=$PWD/$so python tools/gpu_one.py ${B:-4096} 0 2>&1 | tail -1
done
