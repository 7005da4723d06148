#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config c4 --batch 16 --steps 20 --warmup 5 --no-ops > gpurun_out/bench_c4.log 2> gpurun_out/bench_c4.err; echo "exit $?"; tail -1 gpurun_out/bench_c4.log | cut -c1-1500
timeout 900 python bench.py --config c4 --batch 16 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c4_ref.log 2> gpurun_out/bench_c4_ref.err; echo "exit $?"; tail -1 gpurun_out/bench_c4_ref.log | cut -c1-900
