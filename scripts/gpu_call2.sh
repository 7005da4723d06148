#!/bin/bash
mkdir -p gpurun_out
for part in sem p1 p3 perf; do
  timeout 120 python scripts/tc_probe.py $part > gpurun_out/tc_probe_$part.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_$part.log
  tail -30 gpurun_out/tc_probe_$part.log
done
nvidia-smi --query-gpu=name,temperature.gpu --format=csv
