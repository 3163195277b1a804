#!/bin/bash
# First GPU call of round 2: fused-kernel diagnostic, parity tests, bench + A/B knobs.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== debug fused"; timeout 300 python tools/gpu_debug_fused.py 2>&1 | tail -20 | tee $O/r02_debug_fused.txt
echo "== quick parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "test_message_matches_oracle or test_infer_matches_oracle or golden or reference_code" 2>&1 | tail -8 | tee $O/r02_quick.txt
echo "== old suite"; timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_bench_size.py 2>&1 | tail -25 | tee $O/r02_suite.txt
echo "== bench default"; timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_r02a.json 2> $O/bench_r02a.err; tail -c 2500 $O/bench_r02a.json; tail -3 $O/bench_r02a.err
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe"
ab() { name=$1; shift; echo "== A/B $name"; env "$@" timeout 300 $B ${EXTRA:-} 2> $O/ab_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['value'], d['e2e']['value'], d.get('launches_per_image'), d['roofline']['ms_per_message'])" | tee -a $O/r02_ab.txt; }
ab default X=1
ab nobatch PSINFER_NO_BATCH=1
ab nograph PSINFER_NO_GRAPH=1
ab nosnake PSINFER_NO_SNAKE=1
ab maxbatch3 PSINFER_MAX_BATCH=3
EXTRA="--streams 4" ab streams4 X=1
EXTRA="--streams 2" ab streams2 X=1
EXTRA="--streams 16 --images 32" ab streams16 X=1
EXTRA="--fast-math" ab fastmath X=1
echo "== bench-size tests"; timeout 1500 python -m pytest tests/test_gpu_bench_size.py -q 2>&1 | tail -15 | tee $O/r02_bench_size.txt
