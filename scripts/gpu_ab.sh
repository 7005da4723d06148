#!/bin/bash
# same-box A/B of environment switches on the bench step:  bash scripts/gpu_ab.sh "A=1" "DLGRAD_X=1" ...
mkdir -p gpurun_out
i=0
for v in "$@" "$1"; do
  i=$((i+1))
  env $v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_ab_$i.json 2> gpurun_out/r02_ab_$i.err
  tail -1 gpurun_out/r02_ab_$i.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]'.ljust(28), 'value ms', round(d['ms_per_step'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'launches', d['gpu_launches']//d['steps'], 'loss', d['last_loss'])" || tail -3 gpurun_out/r02_ab_$i.err
done
