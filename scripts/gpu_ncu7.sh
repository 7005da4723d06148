#!/bin/bash
# ncu --set full on the fused attention kernels and the tile-skipping GEMMs (after the same command ran clean)
mkdir -p gpurun_out
CMD="python scripts/ncu_attn.py"
timeout 120 $CMD || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_|mm_tcts_sp|mask_bits|rowdot" -c 16 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log; ls -la gpurun_out/prof_attn.ncu-rep
