#!/bin/bash
# Runs ON THE GPU BOX: one `--set full` capture WITH sources of the n-th launch of each named kernel (eager launches, one
# image in flight), exported as the source-page CSV (per-SASS-instruction executed counts and stall samples).
# usage: tools/profile_src_kernel.sh <tag> <kernel regex> [skip] [count]   -> gpurun_out/src_<tag>.csv.gz, prof_<tag>_raw.csv
set -u
tag=$1; rx=$2; skip=${3:-8}; count=${4:-1}
export PSINFER_NO_GRAPH=1
B="python bench.py --ncu --images 1 --streams 1 --steps 1 --warmup 2"
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $count -f -o gpurun_out/prof_${tag} $B > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv > gpurun_out/src_${tag}.csv 2>gpurun_out/src_${tag}.err
ls -la gpurun_out/src_${tag}.csv
gzip -f gpurun_out/src_${tag}.csv
rm -f gpurun_out/prof_${tag}.ncu-rep
tail -3 gpurun_out/prof_${tag}.log
