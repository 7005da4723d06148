#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --profile --no-ops --no-cpu > gpurun_out/profile_step.log 2>&1; tail -45 gpurun_out/profile_step.log
STEPS=10 bash scripts/gpu_tests_bench.sh 2>&1 | cut -c1-900
