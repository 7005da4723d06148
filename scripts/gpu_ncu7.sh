#!/bin/bash
# ncu --set full on the fused attention kernels (after the same command ran clean)
mkdir -p gpurun_out
CMD="python scripts/attn_bwd_trace.py"
timeout 120 $CMD || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 1 -f -o gpurun_out/attn_bwd $CMD > gpurun_out/ncu_attn_bwd.log 2>&1
tail -3 gpurun_out/ncu_attn_bwd.log; ls -la gpurun_out/attn_bwd.ncu-rep
