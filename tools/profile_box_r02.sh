#!/bin/bash
# Runs ON THE GPU BOX: ncu launch list of three images (one ctx, eager launches), then `--set full` over the launches of
# the LAST image, exports and the per-message DRAM traffic.   usage: tools/profile_box_r02.sh <tag>   -> gpurun_out/
set -u
tag=${1:-r02a}
export PSINFER_NO_GRAPH=1
B="python bench.py --ncu --images 1 --streams 1 --steps 1 --warmup 2"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv $B > gpurun_out/launches_${tag}.log 2>&1
read N PER < <(python tools/ncu_summary.py lastimage gpurun_out/launches_${tag}.csv)
echo "launches total $N, last image $PER"
python tools/ncu_summary.py launches gpurun_out/launches_${tag}.csv > gpurun_out/${tag}_launch_summary.txt
ncu --set full --clock-control none --import-source on -s $((N-PER)) -c ${PER} -f -o gpurun_out/prof_${tag} $B > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv > gpurun_out/src_${tag}.csv 2>/dev/null
gzip -f gpurun_out/src_${tag}.csv
python tools/ncu_summary.py raw gpurun_out/prof_${tag}_raw.csv > gpurun_out/${tag}_ncu_full_summary.txt
python tools/ncu_summary.py traffic gpurun_out/prof_${tag}_raw.csv 18 > gpurun_out/${tag}_traffic.json
rm -f gpurun_out/prof_${tag}.ncu-rep
tail -30 gpurun_out/${tag}_launch_summary.txt; cat gpurun_out/${tag}_traffic.json
