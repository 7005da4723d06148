#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --ops-only > gpurun_out/ops.log 2>&1; echo "exit $?" >> gpurun_out/ops.log; tail -20 gpurun_out/ops.log
STEPS=10 bash scripts/gpu_tests_bench.sh
