#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_size.py tests/test_gpu_cli.py -x -q -k "not pcp and not fast_math_argmax" 2>&1 | tail -3
echo "== sanitizer"; bash tools/sanitize_box.sh 2>&1 | tail -4
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe"
ab() { name=$1; shift; echo "== A/B $name"; env "$@" timeout 300 $B ${EXTRA:-} 2> $O/ab_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', d['value'], d['e2e']['value'], 'gauss', k['gauss_xy']['ms_per_image'] if 'gauss_xy' in k else None, 'msg', d['roofline']['ms_per_message'])" | tee -a $O/r02_ab8.txt; }
ab table X=1
ab dynamic PSINFER_GAUSS_DYNAMIC=1
EXTRA="--streams 1" ab table_s1 X=1
EXTRA="--streams 1" ab dynamic_s1 PSINFER_GAUSS_DYNAMIC=1
EXTRA="--workload cfg4 --steps 4" ab cfg4_table X=1
