#!/bin/bash
# ncu evidence: (1) per-launch device time of a short bench run, (2) full counters for the tcgen05 matmul
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-ops --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mm_tc -s 40 -c 3 -o gpurun_out/prof_mm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_list.log; tail -3 gpurun_out/ncu_full.log; wc -l gpurun_out/launches.csv; ls -la gpurun_out/*.ncu-rep
