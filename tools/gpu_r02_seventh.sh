#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe"
ab() { name=$1; shift; echo "== A/B $name"; env "$@" timeout 300 $B ${EXTRA:-} 2> $O/ab_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', d['value'], d['e2e']['value'], 'gauss', k['gauss_xy']['ms_per_image'] if 'gauss_xy' in k else None, 'msg', d['roofline']['ms_per_message'], 'lpi', d['launches_per_image'])" | tee -a $O/r02_ab6.txt; }
for mb in 1 2 3; do
  EXTRA="--streams 8" ab maxbatch$mb PSINFER_MAX_BATCH=$mb
done
EXTRA="--streams 4" ab maxbatch2_s4 PSINFER_MAX_BATCH=2
EXTRA="--streams 3" ab maxbatch3_s3 PSINFER_MAX_BATCH=3
EXTRA="--streams 8" ab nobatch PSINFER_NO_BATCH=1
echo "== cfg4"; timeout 600 python bench.py --workload cfg4 --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_r02_cfg4_1gpu.json 2> $O/bench_r02_cfg4.err; python -c "import json; d=json.load(open('$O/bench_r02_cfg4_1gpu.json')); print('cfg4', d['value'], d['e2e']['value'], d['launches_per_image'], d.get('other_mode'))"; tail -3 $O/bench_r02_cfg4.err
echo "== cfg5"; timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_r02_cfg5_1gpu.json 2> $O/bench_r02_cfg5.err; python -c "import json; d=json.load(open('$O/bench_r02_cfg5_1gpu.json')); print('cfg5', d['value'], d['e2e']['value'], d['launches_per_image'], d.get('other_mode'))"; tail -3 $O/bench_r02_cfg5.err
echo "== cli"; timeout 900 python tools/cli_throughput.py --images 256 --distinct 16 --out $O/r02_cli_throughput.json 2>&1 | tail -6
