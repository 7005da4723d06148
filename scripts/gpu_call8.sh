#!/bin/bash
mkdir -p gpurun_out
for part in p3 v3 perf; do
  timeout 300 python scripts/tc_probe.py $part > gpurun_out/tc_probe_$part.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_$part.log
  tail -64 gpurun_out/tc_probe_$part.log
done
