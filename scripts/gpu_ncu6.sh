#!/bin/bash
# ncu launch list of a short bench run (per-launch device time), after the same command ran clean
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-ops --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-200
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"; tail -2 gpurun_out/ncu_list.log | cut -c1-200; wc -l gpurun_out/launches.csv
