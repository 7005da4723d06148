#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --profile --no-ops --no-cpu > gpurun_out/profile_step.log 2>&1; grep -v "ms/step=   0.0[0-4]" gpurun_out/profile_step.log | tail -60
