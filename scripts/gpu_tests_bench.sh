#!/bin/bash
# parity tests + smoke + bench on one B200
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/test_gpu.log
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -2 gpurun_out/smoke.log; grep -E "^FAILED|passed|failed" gpurun_out/test_gpu.log | tail -25; cut -c1-1800 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
