#!/bin/bash
# Last confirmation of the shipped tree: full GPU suite, smoke, default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== full suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $O/r02_suite_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; python -c "import json; d=json.load(open('$O/r02_bench_default.json')); print(d['value'], d['e2e']['value'], d['launches_per_image'], d['roofline']['frac'], d['parity_check']['parity']['marginal_cells_differing'], d['other_mode']['value'], d['clocks'], d['cpu_baseline']['value'])"; tail -3 $O/r02_bench_default.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-400
