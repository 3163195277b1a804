#!/bin/bash
# usage: tools/scan_streams.sh  -- throughput vs number of concurrent contexts/streams per GPU
for s in 1 2 3 4; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --images 8 --streams $s 2>&1 | tail -1 > /tmp/b_$s.json
  python -c "
import json
d=json.load(open('/tmp/b_$s.json')); print('streams', $s, 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
done
