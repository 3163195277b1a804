#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== full suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/r02_suite_full.txt
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['launches_per_image'], d['roofline']['frac'], d['roofline']['ms_per_message'], d['parity_check']['parity']['marginal_cells_differing'], d['other_mode']['value'])"; tail -3 $O/r02_bench_1gpu.err
for st in 1 4; do python bench.py --workload cfg4 --streams $st --steps 4 --warmup 3 --no-cpu-baseline --no-mode-probe 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 s$st', d['value'], d['e2e']['value'])"; done
python bench.py --streams 1 --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 s1', d['value'], d['e2e']['value'])"
PSINFER_GAUSS_DYNAMIC=1 python bench.py --streams 1 --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 s1 dynamic', d['value'], d['e2e']['value'])"
