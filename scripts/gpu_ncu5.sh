#!/bin/bash
# ncu --set full over the dominant kernels (scripts/ncu_targets.py), after the same command ran clean without ncu
mkdir -p gpurun_out
CMD="python scripts/ncu_targets.py"
$CMD > gpurun_out/ncu_targets_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_targets_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -c 30 -f -o gpurun_out/prof_targets $CMD > gpurun_out/ncu_targets.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_targets.log; ls -la gpurun_out/prof_targets.ncu-rep
