#!/bin/bash
mkdir -p gpurun_out
python scripts/launch_gap_probe.py 2>&1 | tail -4
DLGRAD_PDL=0 python scripts/launch_gap_probe.py 2>&1 | tail -4
STEPS=10 bash scripts/gpu_tests_bench.sh 2>&1 | grep "smoke\|passed\|failed\|steps_per_s" | cut -c1-330
DLGRAD_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-ops --no-cpu 2>/dev/null | tail -1 | cut -c1-200
