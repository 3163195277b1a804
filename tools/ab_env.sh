#!/bin/bash
# Runs ON THE GPU BOX: the bench under alternative environment settings (tools/ab_env.sh "VAR=1" "VAR=2 OTHER=3" ...)
for e in "$@"; do
  env $e python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-mode-probe ${BENCH_ARGS:-} 2>/dev/null | tail -1 > /tmp/ab.json
  python -c "
import json
d=json.load(open('/tmp/ab.json')); k=d['roofline']['kernel_ms_per_image']
print('$e', 'value', d['value'], 'e2e', d['e2e']['value'], {n: k[n] for n in ('warp_back','warp_bilinear','warp_direct','rotconv','epilogue','conv_cols','conv_rows') if n in k})"
done
