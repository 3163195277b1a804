#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== full suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $O/r02_suite_full.txt
echo "== ncu gauss src"; PSINFER_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_gauss_xy -s 8 -c 1 -f -o $O/prof_r02c python bench.py --ncu --images 1 --streams 1 --steps 1 --warmup 2 > $O/prof_r02c.log 2>&1
ncu -i $O/prof_r02c.ncu-rep --page source --csv > $O/src_r02c.csv 2>/dev/null; ncu -i $O/prof_r02c.ncu-rep --page raw --csv > $O/prof_r02c_raw.csv 2>/dev/null; rm -f $O/prof_r02c.ncu-rep
python tools/ncu_src_stalls.py $O/src_r02c.csv k_gauss_xy 2>&1 | head -32
gzip -f $O/src_r02c.csv
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-mode-probe"
cp partapp_b200/libpsinfer.so /tmp/lib_default.so
for lib in /tmp/lib_default.so tools/ab_libs/lib_minb4.so; do
  [ -f $lib ] || continue
  cp $lib partapp_b200/libpsinfer.so
  $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$lib', d['value'], d['e2e']['value'], 'gauss', k['gauss_xy']['ms_per_image'])" | tee -a $O/r02_ab5.txt
done
cp /tmp/lib_default.so partapp_b200/libpsinfer.so
