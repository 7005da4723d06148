#!/bin/bash
# what the driver does at round end: GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/test_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/test_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
