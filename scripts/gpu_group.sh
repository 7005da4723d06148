#!/bin/bash
# grouped sibling-projection launches: parity tests, the whole GPU suite, then same-box A/B of the bench step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_r2.py -q -m gpu -p no:cacheprovider -x -k "grouped or adds_ride" > gpurun_out/r02_group_tests.txt 2>&1
echo "group tests exit $?"; tail -25 gpurun_out/r02_group_tests.txt
if [ "$1" = "all" ]; then
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r02_pytest_gpu.txt
fi
nvidia-smi --query-gpu=utilization.gpu,memory.used,clocks.sm --format=csv,noheader
one() {
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_group_$name.json 2> gpurun_out/r02_group_$name.err
  tail -1 gpurun_out/r02_group_$name.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); n=d['n_gpus']
print('$name'.ljust(14), 'value ms', round(d['ms_per_step'],3), 'e2e ms', round(1e3*n/d['e2e']['value'],3), 'launches', d['gpu_launches']//d['steps'], 'loss', d['last_loss'])
for s in d['roofline']['by_shape']:
    if 'group' in str(s) or '768, 768' in str(s): print('   ', s)
" || tail -5 gpurun_out/r02_group_$name.err
}
one grouped A=1
one single DLGRAD_GROUP_LINEAR=0
one grouped_b A=1
one single_b DLGRAD_GROUP_LINEAR=0
one grouped_c A=1
one single_c DLGRAD_GROUP_LINEAR=0
