#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for wl in cfg2r8 cfg2; do for st in 1 8; do
python bench.py --workload $wl --streams $st --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$wl s$st', d['value'], {n: v['ms_per_image'] for n,v in k.items()})"
done; done
echo "== cli"; timeout 900 python tools/cli_throughput.py --images 256 --distinct 16 --contexts 12 --out $O/r02_cli_throughput_c12.json 2>&1 | tail -4
