CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for n in 40 100; do
    log=gpurun_out/sanitize_r02b_${tool}_quad_N$n.log
    QUAD_N=$n timeout 600 $CS --tool $tool --print-limit 20 python tools/sanitize_case.py quad 2 6 > $log 2>&1
    echo "$tool quad N=$n: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -n 1) | $(grep -E '^quad mode' $log | cut -c1-140)"
  done
done
