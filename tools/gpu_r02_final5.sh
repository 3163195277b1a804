#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== full suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $O/r02_suite_full.txt
echo "== sanitizer"; bash tools/sanitize_box.sh 2>&1 | tail -4 | tee $O/r02_sanitizer_lines.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['launches_per_image'], d['roofline']['frac'], d['roofline']['ms_per_message'], d['parity_check']['parity']['marginal_cells_differing'], d['other_mode']['value'], d['clocks'])"; tail -3 $O/r02_bench_1gpu.err
echo "== profile"; timeout 1500 bash tools/profile_box_r02.sh r02e 2>&1 | tail -3
