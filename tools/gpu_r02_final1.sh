#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== sanitizer"; bash tools/sanitize_box.sh 2>&1 | tail -4
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_reference_arm.json 2>/dev/null; tail -c 900 $O/r02_bench_reference_arm.json
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_1gpu.json')); print(d['value'], d['e2e'], d['launches_per_image'], {k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic','traffic_over_algorithmic','ms_per_message')}, d['parity_check'], d['cpu_baseline']['value'], d['other_mode'])"; tail -3 $O/r02_bench_1gpu.err
