#!/bin/bash
mkdir -p gpurun_out
python scripts/flash_bench.py > gpurun_out/r02_flash_bench.txt 2>&1; cat gpurun_out/r02_flash_bench.txt
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash -c 3 -f -o /tmp/r02_ncu_flash python scripts/flash_bench.py > gpurun_out/r02_ncu_flash.log 2>&1; ncu -i /tmp/r02_ncu_flash.ncu-rep --page raw --csv > gpurun_out/r02_ncu_flash_raw.csv 2>/dev/null
echo "ncu exit $?"; tail -3 gpurun_out/r02_ncu_flash.log; ls -la gpurun_out/r02_ncu_flash_raw.csv
