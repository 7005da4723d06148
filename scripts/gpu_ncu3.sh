#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/ncu_attn.py"
$CMD > gpurun_out/ncu_attn_plain.log 2>&1 && \
ncu --set full --clock-control none -s 2 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_attn.log
