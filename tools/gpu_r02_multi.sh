#!/bin/bash
# Multi-GPU evidence (run under gpurun --gpus N): bench at N ranks for the three workloads, CLI over N GPUs.
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== cfg2 x$N"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_${N}gpu.json 2> $O/r02_bench_${N}gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_${N}gpu.json')); print('cfg2', d['n_gpus'], d['value'], d['e2e']['value'], d['clocks'])"; tail -2 $O/r02_bench_${N}gpu.err
echo "== cfg4 x$N"; timeout 600 $TR bench.py --gpus $N --workload cfg4 --steps 6 --warmup 3 > $O/r02_bench_cfg4_${N}gpu.json 2> $O/r02_bench_cfg4_${N}gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_cfg4_${N}gpu.json')); print('cfg4', d['n_gpus'], d['value'], d['e2e']['value'], d['clocks'])"; tail -2 $O/r02_bench_cfg4_${N}gpu.err
echo "== cfg5 x$N"; timeout 900 $TR bench.py --gpus $N --workload cfg5 --steps 3 --warmup 3 > $O/r02_bench_cfg5_${N}gpu.json 2> $O/r02_bench_cfg5_${N}gpu.err; python -c "import json; d=json.load(open('$O/r02_bench_cfg5_${N}gpu.json')); print('cfg5', d['n_gpus'], d['value'], d['e2e']['value'], d['clocks'])"; tail -2 $O/r02_bench_cfg5_${N}gpu.err
echo "== cli x$N"; timeout 900 python tools/cli_throughput.py --images 1024 --distinct 16 --contexts 2 --out $O/r02_cli_throughput_${N}gpu.json 2>&1 | tail -5
