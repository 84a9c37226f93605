#!/bin/bash
for so in build/libobca_*.so; do echo "== $so"; OBCA_SO=$PWD/$so python tools/gpu_quad.py ${B:-2048} 2>&1 | grep -v "^  phase" ; done
