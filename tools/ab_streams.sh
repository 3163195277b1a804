#!/bin/bash
# Runs ON THE GPU BOX: GPU parity tests, the default bench, then a sweep over images in flight / conv blocks per SM.
python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/ab_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_image"], d["roofline"]["fp32_pipe"]["kernels"], d["other_mode"])
PY
for v in ${SWEEP:-4:4:0 4:4:1 8:8:0 8:8:1 6:6:0}; do
  set -- ${v//:/ }
  PSINFER_CONV_BLOCKS=$3 python bench.py --no-cpu-baseline --no-mode-probe --images $1 --streams $2 --steps 10 2>/dev/null > /tmp/ab.json
  python -c "
import json; d=json.load(open('/tmp/ab.json')); print('images/streams/convblocks', '$v', d['value'], d['e2e']['value'])"
done
