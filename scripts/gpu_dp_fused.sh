#!/bin/bash
# fused data-parallel step on N GPUs: correctness check, then a same-box A/B against ncclAllReduce + local Adam
#   gpurun --gpus 2 -- 'bash scripts/gpu_dp_fused.sh 2'
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 120 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k fused_dp 2>&1 | tail -3
NCCL_DEBUG=WARN timeout 240 $RUN --master-port 29533 scripts/dp_fused_check.py > gpurun_out/dp_fused_check_$N.txt 2> gpurun_out/dp_fused_check_$N.err
echo "check exit $?"; cat gpurun_out/dp_fused_check_$N.txt; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/dp_fused_check_$N.err | tail -12
port=29540
for v in 1 0 1 0; do
  port=$((port + 1))
  DLGRAD_DP_FUSED=$v timeout 300 $RUN --master-port $port bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_dp${N}_fused$v.json 2> gpurun_out/bench_dp${N}_fused$v.err
  echo "exit $?"
  tail -1 gpurun_out/bench_dp${N}_fused$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('DLGRAD_DP_FUSED=$v', 'n_gpus', d['n_gpus'], round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s', d['clocks'].get('sm_mhz'))" || tail -8 gpurun_out/bench_dp${N}_fused$v.err
done
