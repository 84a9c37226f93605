python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests2.log 2>&1; tail -n 12 gpurun_out/r2_gpu_tests2.log
tools/gpu_variants.sh > gpurun_out/r2_var4.log 2>&1; cat gpurun_out/r2_var4.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -n 3 gpurun_out/r2_smoke.log
