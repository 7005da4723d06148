#!/bin/bash
mkdir -p gpurun_out
for part in p3 v3 v4 perf4; do
  timeout 300 python scripts/tc_probe.py $part > gpurun_out/tc_probe_$part.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_$part.log
done
grep -h "FAILED\|exit" gpurun_out/tc_probe_p3.log gpurun_out/tc_probe_v3.log gpurun_out/tc_probe_v4.log
grep -v "bn256 " gpurun_out/tc_probe_perf4.log | tail -32
bash scripts/gpu_call6.sh
