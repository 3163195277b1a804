#!/bin/bash
# Runs ON THE GPU BOX: A/B over (library build, environment) pairs: tools/ab_lib_env.sh "lib.so VAR=1" "lib2.so VAR=2" ...
cp partapp_b200/libpsinfer.so /tmp/libpsinfer_keep.so
for pair in "$@"; do
  set -- $pair
  lib=$1; shift
  cp "$lib" partapp_b200/libpsinfer.so
  env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-mode-probe ${BENCH_ARGS:-} 2>/dev/null | tail -1 > /tmp/ab.json
  python -c "
import json
d=json.load(open('/tmp/ab.json')); print('$pair', 'value', d['value'], 'e2e', d['e2e']['value'])"
done
cp /tmp/libpsinfer_keep.so partapp_b200/libpsinfer.so
