#!/bin/bash
# compute-sanitizer over every kernel of the library: memcheck, racecheck, synccheck and initcheck on small
# batches in every schedule (OBCA_MODE 1 = persistent kernel, 2 = rounds only, 2 + hand-over = rounds then tail kernel).
# Run on the GPU box:   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'      logs: gpurun_out/sanitize_*.log, summary on stdout
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
out=gpurun_out
mkdir -p $out
run() {   # tool tag env... -- args
  tool=$1; tag=$2; shift 2
  log=$out/sanitize_${tool}_${tag}.log
  env "$@" $CS --tool $tool --print-limit 20 python tools/sanitize_case.py $CASE > $log 2>&1
  echo "$tool $tag: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -n 1) | $(grep -E '^(parking|dist|quad) mode' $log | cut -c1-160)"
}
for tool in memcheck racecheck; do
  CASE="parking 4 10"; run $tool sd_handover OBCA_MODE=2 OBCA_TAIL_THRESH=2      # rounds (k_pk_eval / k_pk_sweep / k_pk_step) then k_pk_tail
  CASE="dist 2 10";    run $tool d_tail OBCA_MODE=1
  CASE="quad 2 8";     run $tool quad OBCA_MODE=0
done
for tool in synccheck initcheck; do
  CASE="parking 4 10"; run $tool sd_handover OBCA_MODE=2 OBCA_TAIL_THRESH=2
done
