#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== misc debug"; timeout 300 python tools/gpu_debug_misc.py 2>&1 | tail -30 | tee $O/r02_debug_misc.txt
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe"
ab() { name=$1; shift; echo "== A/B $name"; env "$@" timeout 300 $B ${EXTRA:-} 2> $O/ab_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', d['value'], d['e2e']['value'], 'gauss', k['gauss_xy']['ms_per_image'] if 'gauss_xy' in k else None, 'msg', d['roofline']['ms_per_message'])" | tee -a $O/r02_ab2.txt; }
for st in 2 4 8; do
  EXTRA="--streams $st" ab bps1_pad116_s$st X=1
  EXTRA="--streams $st" ab bps1_pad0_s$st PSINFER_GAUSS_SMEM=0
  EXTRA="--streams $st" ab bps2_s$st PSINFER_GAUSS_BPS=2
done
EXTRA="--streams 3" ab bps1_pad116_s3 X=1
EXTRA="--streams 1" ab bps1_pad116_s1 X=1
EXTRA="--streams 1" ab bps2_s1 PSINFER_GAUSS_BPS=2
EXTRA="--streams 8" ab bps1_pad150_s8 PSINFER_GAUSS_SMEM=153600
echo "== profile"; timeout 1500 bash tools/profile_box_r02.sh r02a 2>&1 | tail -45
