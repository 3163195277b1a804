#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== quick parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_size.py -x -q -k "test_message_matches_oracle or test_infer_matches_oracle or cfg2_whole or 22_part" 2>&1 | tail -4
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe"
ab() { name=$1; shift; echo "== A/B $name"; env "$@" timeout 300 $B ${EXTRA:-} 2> $O/ab_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', d['value'], d['e2e']['value'], 'gauss', k['gauss_xy']['ms_per_image'] if 'gauss_xy' in k else None, 'msg', d['roofline']['ms_per_message'])" | tee -a $O/r02_ab3.txt; }
for st in 4 8; do
  EXTRA="--streams $st" ab bps2_ns1_pad72_s$st X=1
  EXTRA="--streams $st" ab bps2_ns2_pad0_s$st PSINFER_GAUSS_STAGES=2 PSINFER_GAUSS_SMEM=0
  EXTRA="--streams $st" ab bps1_ns1_pad116_s$st PSINFER_GAUSS_BPS=1
  EXTRA="--streams $st" ab bps2_ns1_pad0_s$st PSINFER_GAUSS_SMEM=0
done
EXTRA="--streams 8" ab bps2_ns1_pad90_s8 PSINFER_GAUSS_SMEM=92160
EXTRA="--streams 6" ab bps2_ns1_pad72_s6 X=1
EXTRA="--streams 12 --images 36" ab bps2_ns1_pad72_s12 X=1
EXTRA="--streams 8 --fast-math" ab fast_bps2_ns1_pad72_s8 X=1
echo "== ncu gauss"; PSINFER_NO_GRAPH=1 ncu --set full --clock-control none -k regex:k_gauss_xy -s 8 -c 2 -f -o $O/prof_r02b python bench.py --ncu --images 1 --streams 1 --steps 1 --warmup 2 > $O/prof_r02b.log 2>&1
ncu -i $O/prof_r02b.ncu-rep --page raw --csv > $O/prof_r02b_raw.csv 2>/dev/null; python tools/ncu_summary.py raw $O/prof_r02b_raw.csv | head -44; rm -f $O/prof_r02b.ncu-rep
echo "== cli test"; timeout 600 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -30
