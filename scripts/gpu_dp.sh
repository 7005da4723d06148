#!/bin/bash
# data-parallel check on N GPUs (final code): the multi-rank tests (both exchange schedules), then bench.py at N with each
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_cuda_parity.py -q -m gpu -p no:cacheprovider -k "fused_dp or dp_adam" > gpurun_out/r02_dp_tests_$N.txt 2>&1
echo "dp tests exit $?"; tail -3 gpurun_out/r02_dp_tests_$N.txt
for ov in 0 1; do
  DLGRAD_DP_OVERLAP=$ov NCCL_DEBUG=WARN timeout 300 $RUN --master-port 2953$ov scripts/dp_fused_check.py > gpurun_out/r02_dp_fused_check_${N}gpu_overlap$ov.txt 2> gpurun_out/r02_dp_fused_check_${N}gpu_overlap$ov.err
  echo "check overlap=$ov exit $?"; tail -2 gpurun_out/r02_dp_fused_check_${N}gpu_overlap$ov.txt
done
port=29541
for name in monolithic overlapped monolithic_b; do
  port=$((port+1)); ov=0; [ "$name" = "overlapped" ] && ov=1
  DLGRAD_DP_OVERLAP=$ov timeout 400 $RUN --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_bench_dp${N}_$name.json 2> gpurun_out/r02_bench_dp${N}_$name.err
  tail -1 gpurun_out/r02_bench_dp${N}_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); n=d['n_gpus']; print('$name'.ljust(14), 'n_gpus', n, 'value ms', round(d['ms_per_step'],3), 'steps/s', round(d['value'],2), 'e2e ms', round(1e3*n/d['e2e']['value'],3), 'last_loss', repr(d['last_loss']), 'replicas_identical', d.get('replicas_identical'), d['clocks'].get('sm_mhz'))" || tail -8 gpurun_out/r02_bench_dp${N}_$name.err
done
if [ "$N" = "2" ]; then
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_bench_1gpu_samebox.json 2> gpurun_out/r02_bench_1gpu_samebox.err
  tail -1 gpurun_out/r02_bench_1gpu_samebox.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1 gpu same box', round(d['ms_per_step'],3), 'ms/step', round(d['value'],2), 'steps/s', 'e2e ms', round(1e3/d['e2e']['value'],3))"
fi
