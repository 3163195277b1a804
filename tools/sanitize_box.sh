#!/bin/bash
# Runs ON THE GPU BOX: compute-sanitizer (memcheck, racecheck, initcheck) over tools/sanitize_small.py
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
