#!/bin/bash
# Multi-GPU refresh of the cfg-2 bench line (run under gpurun --gpus N).
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== cfg2 x$N"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_${N}gpu.json 2> $O/r02_bench_${N}gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_${N}gpu.json')); print('cfg2', d['n_gpus'], d['value'], d['e2e']['value'], d['clocks'])"; tail -2 $O/r02_bench_${N}gpu.err
