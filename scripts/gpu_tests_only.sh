#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/test_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/test_gpu.log
