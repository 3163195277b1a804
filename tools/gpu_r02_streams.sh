#!/bin/bash
# GPU box: images in flight per GPU with the shipped build.
for s in 6 8 12 16; do
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-mode-probe --streams $s 2>/dev/null | tail -1 > /tmp/ab.json
  python -c "
import json
d=json.load(open('/tmp/ab.json')); print('streams $s', 'value', d['value'], 'e2e', d['e2e']['value'])"
done | tee gpurun_out/r02o_streams.txt
