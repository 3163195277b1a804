#!/bin/bash
# A/B: run the bench against alternative builds of libpsinfer (tools/ab_lib.sh libA.so libB.so ...)
cp partapp_b200/libpsinfer.so /tmp/libpsinfer_base.so
for lib in "$@"; do
  cp "$lib" partapp_b200/libpsinfer.so
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-mode-probe ${BENCH_ARGS:-} 2>/dev/null | tail -1 > /tmp/ab.json
  python -c "
import json
d=json.load(open('/tmp/ab.json')); k=d['roofline']['kernel_ms_per_image']
print('$lib', 'value', d['value'], 'e2e', d['e2e']['value'], {n: k[n] for n in ('warp_back','warp_bilinear','warp_direct','rotconv','epilogue','conv_cols','conv_rows') if n in k})"
done
cp /tmp/libpsinfer_base.so partapp_b200/libpsinfer.so
