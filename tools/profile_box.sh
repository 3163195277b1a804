#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): ncu launch list + one `--set full` capture of a slice of one image.
# usage: tools/profile_box.sh <tag> [skip] [count]      outputs under gpurun_out/
set -u
tag=${1:-rXX}; skip=${2:-200}; count=${3:-40}
B="python bench.py --ncu --images 1 --streams 1 --steps 2 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 160 --csv \
    --log-file gpurun_out/launches_${tag}.csv $B > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -s ${skip} -c ${count} -f -o gpurun_out/prof_${tag} \
    $B > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv > gpurun_out/src_${tag}.csv 2>/dev/null
gzip -f gpurun_out/src_${tag}.csv
# the .ncu-rep itself is too big for the 64 MiB return channel once sources are imported: keep the CSV exports
rm -f gpurun_out/prof_${tag}.ncu-rep
ls -la gpurun_out/ | grep ${tag}
