#!/bin/bash
# scratch driver for one gpurun call: compute-sanitizer over the new integer kernels
O=gpurun_out
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 70 $CS --tool $tool --print-limit 20 python profiles/sanitize_graph_device.py > $O/sanitizer_${tool}_graph_device_r02.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all checks passed|Error|error' $O/sanitizer_${tool}_graph_device_r02.log | head -5 | tr '\n' ' ')"
done
