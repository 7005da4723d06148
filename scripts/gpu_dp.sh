#!/bin/bash
# data-parallel check on N GPUs: the 2-GPU fused-step test, the bit-level check of the fused exchange, and bench.py at N
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_cuda_parity.py -q -m gpu -p no:cacheprovider -k fused_dp 2>&1 | tail -3
NCCL_DEBUG=WARN timeout 300 $RUN --master-port 29533 scripts/dp_fused_check.py > gpurun_out/r02_dp_fused_check_$N.txt 2> gpurun_out/r02_dp_fused_check_$N.err
echo "check exit $?"; tail -6 gpurun_out/r02_dp_fused_check_$N.txt
timeout 400 $RUN --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_dp$N.json 2> gpurun_out/r02_bench_dp$N.err
echo "bench exit $?"
tail -1 gpurun_out/r02_bench_dp$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n_gpus', d['n_gpus'], round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s', 'replicas_identical', d.get('replicas_identical'), d['clocks'].get('sm_mhz'))" || tail -8 gpurun_out/r02_bench_dp$N.err
