python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests7.log 2>&1; echo tests rc=$?; tail -n 3 gpurun_out/r2_gpu_tests7.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke3.log 2>&1; echo smoke rc=$?
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo bench rc=$?; cut -c1-200 gpurun_out/r2_bench5.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref5.json 2> gpurun_out/r2_bench_ref5.err; echo ref rc=$?
OBCA_GRAPH=0 ncu --set full --import-source on --clock-control none -k regex:k_pk_ -s 6 -c 3 -o gpurun_out/r2f_rounds python tools/gpu_one.py 4096 0 > gpurun_out/r2f_ncu1.log 2>&1; echo ncu1 rc=$?
OBCA_GRAPH=0 ncu --set full --clock-control none -k regex:k_pk_tail -c 1 -o gpurun_out/r2f_tail python tools/gpu_one.py 4096 0 > gpurun_out/r2f_ncu2.log 2>&1; echo ncu2 rc=$?
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  OBCA_MODE=2 OBCA_TAIL_THRESH=2 timeout 400 $CS --tool $tool --print-limit 10 python tools/sanitize_case.py parking 4 8 > gpurun_out/sanitize_r02c_${tool}_parking.log 2>&1
  echo "$tool parking: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_r02c_${tool}_parking.log | tail -n 1)"
done
ls -la gpurun_out/r2f_*.ncu-rep
