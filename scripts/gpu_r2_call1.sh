#!/bin/bash
# round 2, call 1: the whole GPU suite with the per-element parity report, smoke, bench (+profile, +reference arm)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/r2_env.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r2_env.txt
DLGRAD_PARITY_REPORT=gpurun_out/r2_parity_report.jsonl timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/r2_pytest_gpu.txt 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/r2_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2_smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/r2_bench.err
timeout 300 python bench.py --profile > gpurun_out/r2_profile.txt 2>&1; echo "profile exit $?"; head -40 gpurun_out/r2_profile.txt
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref exit $?"
cut -c1-900 gpurun_out/r2_bench_ref.json
