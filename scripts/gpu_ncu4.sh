#!/bin/bash
# one --set full capture (with SASS-level sampling) of a single matmul kernel variant
mkdir -p gpurun_out
ARGS="${MM_ARGS:-4096 4096 4096 128 v4}"
timeout 120 python scripts/mm_one.py $ARGS || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_ -s 2 -c 1 -f -o gpurun_out/mm_one python scripts/mm_one.py $ARGS > gpurun_out/ncu_mm_one.log 2>&1
tail -5 gpurun_out/ncu_mm_one.log
ls -la gpurun_out/mm_one.ncu-rep
