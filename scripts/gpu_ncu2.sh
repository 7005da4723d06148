#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/ncu_targets.py"
$CMD > gpurun_out/ncu_targets_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 9 -c 9 -o gpurun_out/prof_targets $CMD > gpurun_out/ncu_targets.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/ncu_targets.log; ls -la gpurun_out/prof_targets.ncu-rep
