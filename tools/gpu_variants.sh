#!/bin/bash
# development: time every build/libobca_*.so variant (rounds only, per-kernel event times)
for so in build/libobca_*.so; do
  echo "== $so"
  OBCA_SO=$PWD/$so OBCA_MODE=2 OBCA_TAIL_THRESH=${THRESH:-0} OBCA_PHASE_TIMING=1 python tools/gpu_one.py ${B:-4096} 2 2>&1 | tail -2
done
