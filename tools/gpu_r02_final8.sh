#!/bin/bash
# Final 1-GPU evidence of the shipped build (64-register fixed-order Gaussian without a producer warp).
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== full suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $O/r02_suite_full.txt
echo "== initcheck"; timeout 400 compute-sanitizer --tool initcheck python tools/sanitize_small.py > $O/sanitize_initcheck.log 2>&1; echo "initcheck: $(grep -E 'ERROR SUMMARY' $O/sanitize_initcheck.log | tail -1)" | tee $O/r02p_initcheck.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['launches_per_image'], d['roofline']['frac'], d['roofline']['ms_per_message'], d['parity_check']['parity']['marginal_cells_differing'], d['other_mode']['value'], d['clocks'])"; tail -3 $O/r02_bench_1gpu.err
echo "== cfg4"; timeout 600 python bench.py --workload cfg4 --steps 6 --warmup 3 > $O/bench_r02_cfg4_1gpu.json 2> $O/bench_r02_cfg4.err; python -c "import json; d=json.load(open('$O/bench_r02_cfg4_1gpu.json')); print('cfg4', d['value'], d['e2e']['value'], d['other_mode']['value'])"
echo "== cfg5"; timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 3 > $O/bench_r02_cfg5_1gpu.json 2> $O/bench_r02_cfg5.err; python -c "import json; d=json.load(open('$O/bench_r02_cfg5_1gpu.json')); print('cfg5', d['value'], d['e2e']['value'], d['other_mode']['value'])"
echo "== profile"; timeout 1200 bash tools/profile_box_r02.sh r02p 2>&1 | tail -3
