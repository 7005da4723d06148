#!/bin/bash
mkdir -p gpurun_out
for part in p3 perf; do
  timeout 300 python scripts/tc_probe.py $part > gpurun_out/tc_probe_$part.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_$part.log
  tail -70 gpurun_out/tc_probe_$part.log
done
timeout 300 python bench.py --ops-only > gpurun_out/ops.log 2>&1; echo "exit $?" >> gpurun_out/ops.log; tail -20 gpurun_out/ops.log
bash scripts/gpu_tests_bench.sh
