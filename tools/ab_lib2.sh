#!/bin/bash
# A/B on the GPU box: the bench against alternative builds of libpsinfer.   usage: tools/ab_lib2.sh libA.so libB.so ...
# Each build runs twice, interleaved, so that drift shows.  BENCH_ARGS adds bench flags.
cp partapp_b200/libpsinfer.so /tmp/libpsinfer_base.so
for rep in 1 2; do
for lib in "$@"; do
  cp "$lib" partapp_b200/libpsinfer.so
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>/dev/null | tail -1 > /tmp/ab.json
  python -c "
import json
d=json.load(open('/tmp/ab.json')); k=d['roofline']['kernels']
print('$lib', 'value', d['value'], 'e2e', d['e2e']['value'], 'fast', (d.get('other_mode') or {}).get('value'), {n: round(k[n]['ms_per_image'],4) for n in k})"
done
done
cp /tmp/libpsinfer_base.so partapp_b200/libpsinfer.so
