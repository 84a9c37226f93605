#!/bin/bash
# development: time every build/libobca_*.so variant (rounds only + default schedule)
for so in build/libobca_*.so; do
  echo "== $so"
  OBCA_SO=$PWD/$so OBCA_MODE=2 OBCA_TAIL_THRESH=${THRESH:-0} python tools/gpu_one.py ${B:-4096} 3 2>/dev/null | tail -1
  OBCA_SO=$PWD/$so python tools/gpu_one.py ${B:-4096} 3 2>/dev/null | tail -1
done
