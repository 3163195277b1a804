#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== profile"; timeout 1500 bash tools/profile_box_r02.sh r02d 2>&1 | tail -42
echo "== cli 1024"; timeout 900 python tools/cli_throughput.py --images 1024 --distinct 16 --contexts 12 --out $O/r02_cli_throughput_1gpu.json 2>&1 | tail -3
