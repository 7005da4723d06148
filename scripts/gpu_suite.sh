#!/bin/bash
# flash attention bring-up: the flash tests alone first (short timeout: a protocol bug traps after ~2 s per wait), then the rest
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_r2.py -q -m gpu -p no:cacheprovider -k "flash" -x > gpurun_out/r02_flash_tests.txt 2>&1
echo "flash tests exit $?"; tail -40 gpurun_out/r02_flash_tests.txt
if [ "$1" = "all" ]; then
  DLGRAD_PARITY_REPORT=gpurun_out/r02_parity_report.jsonl timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1
  echo "pytest exit $?"; tail -15 gpurun_out/r02_pytest_gpu.txt
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-small --no-ops > gpurun_out/r02_bench_flash.json 2> gpurun_out/r02_bench_flash.err; echo "bench exit $?"; tail -c 400 gpurun_out/r02_bench_flash.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_flash.json').read().strip().splitlines()[-1])
print('steps/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for s in d['roofline']['by_shape']: print(s)
"
  timeout 300 python bench.py --profile > gpurun_out/r02_profile_flash.txt 2>&1; echo "profile exit $?"; head -45 gpurun_out/r02_profile_flash.txt
fi
