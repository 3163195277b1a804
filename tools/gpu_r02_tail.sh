#!/bin/bash
# GPU box: parity tests of the fused Gaussian after the tail / store rewrite, then A/B against the previous build.
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_size.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02f_tests.txt
cat gpurun_out/r02f_tests.txt
bash tools/ab_lib2.sh tools/ab_libs/lib_head.so tools/ab_libs/lib_tail.so 2>&1 | tee gpurun_out/r02f_ab_tail.txt
