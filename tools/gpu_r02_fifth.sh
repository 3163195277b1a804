#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== quick parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_size.py -x -q -k "test_message_matches_oracle or test_infer_matches_oracle or cfg2_whole or pos_gaussian or disc_ps or work_lists_change" 2>&1 | tail -4
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe"
ab() { name=$1; shift; echo "== A/B $name"; env "$@" timeout 300 $B ${EXTRA:-} 2> $O/ab_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', d['value'], d['e2e']['value'], 'gauss', k['gauss_xy']['ms_per_image'] if 'gauss_xy' in k else None, 'msg', d['roofline']['ms_per_message'])" | tee -a $O/r02_ab4.txt; }
EXTRA="--streams 8" ab prio_hint100k X=1
EXTRA="--streams 8" ab noprio_hint100k PSINFER_GAUSS_PRIO=0
EXTRA="--streams 8" ab prio_hint2000 PSINFER_MBAR_HINT=2000
EXTRA="--streams 8" ab prio_hint300 PSINFER_MBAR_HINT=300
EXTRA="--streams 8" ab prio_ns2_pad0 PSINFER_GAUSS_STAGES=2 PSINFER_GAUSS_SMEM=0
EXTRA="--streams 8" ab prio_bps1 PSINFER_GAUSS_BPS=1
EXTRA="--streams 4" ab prio_s4 X=1
EXTRA="--streams 16 --images 32" ab prio_s16 X=1
EXTRA="--streams 8 --fast-math" ab fast_prio X=1
for i in 1 2 3; do echo "== cli test $i"; timeout 600 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -3; done
