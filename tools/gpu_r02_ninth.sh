#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
nproc
for c in 4 12; do timeout 900 python tools/cli_throughput.py --images 128 --distinct 16 --contexts $c --out $O/r02_cli_throughput_c$c.json 2>&1 | tail -3; done
