#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/tc_probe.py sk > gpurun_out/tc_probe_sk.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_sk.log
cat gpurun_out/tc_probe_sk.log | tail -50
