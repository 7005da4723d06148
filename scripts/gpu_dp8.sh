#!/bin/bash
# 8-GPU box: bench.py with the monolithic and the overlapped exchange, same box, alternating
mkdir -p gpurun_out
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29541
run() {
  name=$1; shift
  port=$((port+1))
  env "$@" timeout 300 $RUN --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_dp${N}_$name.json 2> gpurun_out/r02_dp${N}_$name.err
  tail -1 gpurun_out/r02_dp${N}_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); n=d['n_gpus']; print('$name'.ljust(22), 'value ms', round(d['ms_per_step'],3), 'steps/s', round(d['value'],2), 'e2e ms', round(1e3*n/d['e2e']['value'],3), 'last_loss', repr(d['last_loss']), 'replicas_identical', d.get('replicas_identical'), d['clocks'])" || tail -8 gpurun_out/r02_dp${N}_$name.err
}
run mono DLGRAD_DP_OVERLAP=0 DLGRAD_BENCH_STEP_TIMES=1
grep "rank 0.*region A per-step" gpurun_out/r02_dp${N}_mono.err
run over DLGRAD_DP_OVERLAP=1 DLGRAD_BENCH_STEP_TIMES=1
grep "rank 0.*region A per-step" gpurun_out/r02_dp${N}_over.err

