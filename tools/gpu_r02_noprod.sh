#!/bin/bash
# GPU box: fixed-order Gaussian without a producer warp (256 threads, 64 registers): parity, fuzz, A/B, sanitizer.
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_size.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02n_tests.txt
bash tools/ab_lib2.sh tools/ab_libs/lib_g64.so tools/ab_libs/lib_noprod.so 2>&1 | tee gpurun_out/r02n_ab.txt
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done | tee gpurun_out/r02n_sanitizer.txt
