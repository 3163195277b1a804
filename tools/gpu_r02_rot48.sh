#!/bin/bash
# GPU box: R = 48 rotation filter with per-message tap ranges: parity at 1000 x 1000 x 48, then cfg-5 A/B.
python -m pytest tests/test_gpu_bench_size.py tests/test_gpu_parity.py -x -q -m gpu -k "cfg5 or rot" 2>&1 | tail -4 | tee gpurun_out/r02j_tests.txt
export BENCH_ARGS="--workload cfg5 --steps 3"
bash tools/ab_lib2.sh tools/ab_libs/lib_nofill.so tools/ab_libs/lib_rot48.so 2>&1 | tee gpurun_out/r02j_ab.txt
