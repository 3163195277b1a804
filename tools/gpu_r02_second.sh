#!/bin/bash
# Second GPU call of round 2: ncu evidence, the other workloads, the C++ CLI.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== profile"; timeout 1500 bash tools/profile_box_r02.sh r02a 2>&1 | tail -45
echo "== cfg4"; timeout 600 python bench.py --workload cfg4 --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_r02_cfg4_1gpu.json 2> $O/bench_r02_cfg4.err; tail -c 1500 $O/bench_r02_cfg4_1gpu.json; tail -3 $O/bench_r02_cfg4.err
echo "== cfg5"; timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_r02_cfg5_1gpu.json 2> $O/bench_r02_cfg5.err; tail -c 1500 $O/bench_r02_cfg5_1gpu.json; tail -3 $O/bench_r02_cfg5.err
echo "== cli"; timeout 900 python tools/cli_throughput.py --images 256 --distinct 16 --out $O/r02_cli_throughput.json 2>&1 | tail -8
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_r02_reference_arm.json 2>/dev/null; tail -c 800 $O/bench_r02_reference_arm.json
