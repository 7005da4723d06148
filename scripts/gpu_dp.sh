#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_dp$N.log 2> gpurun_out/bench_dp$N.err
echo "exit $?"; cut -c1-1200 gpurun_out/bench_dp$N.log; tail -15 gpurun_out/bench_dp$N.err
