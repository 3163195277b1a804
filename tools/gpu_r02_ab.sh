#!/bin/bash
# GPU box: parity suites on the current build, then tools/ab_lib2.sh over the given builds.   usage: gpu_r02_ab.sh <tag> lib...
tag=$1; shift
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_size.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${tag}_tests.txt
cat gpurun_out/${tag}_tests.txt
bash tools/ab_lib2.sh "$@" 2>&1 | tee gpurun_out/${tag}_ab.txt
