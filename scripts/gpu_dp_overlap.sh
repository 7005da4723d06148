#!/bin/bash
# overlapped data-parallel exchange on N GPUs: bench.py at N with the overlapped and the monolithic schedule, with and without
# the per-launch roofline events of region A, same box, alternating; last_loss must agree bit for bit across all of them
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29541
run() {
  name=$1; shift
  port=$((port+1))
  env "$@" timeout 300 $RUN --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_dp${N}_$name.json 2> gpurun_out/r02_dp${N}_$name.err
  tail -1 gpurun_out/r02_dp${N}_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); n=d['n_gpus']; print('$name'.ljust(22), 'value ms', round(d['ms_per_step'],3), 'e2e ms', round(1e3*n/d['e2e']['value'],3), 'last_loss', repr(d['last_loss']), 'replicas_identical', d.get('replicas_identical'), d['clocks'].get('sm_mhz'))" || tail -8 gpurun_out/r02_dp${N}_$name.err
}
run mono DLGRAD_DP_OVERLAP=0 DLGRAD_BENCH_STEP_TIMES=1
grep "rank 0.*per-step" gpurun_out/r02_dp${N}_mono.err
run over DLGRAD_BENCH_STEP_TIMES=1
grep "rank 0.*per-step" gpurun_out/r02_dp${N}_over.err
run mono_nofreeze DLGRAD_DP_OVERLAP=0 DLGRAD_BENCH_STEP_TIMES=1 DLGRAD_BENCH_GC_FREEZE=0
grep "rank 0.*region A per-step" gpurun_out/r02_dp${N}_mono_nofreeze.err
run mono_b DLGRAD_DP_OVERLAP=0
run over_b A=1
N1="timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-small --no-ops"
one() {
  name=$1; shift
  env "$@" $N1 > gpurun_out/r02_dp1_$name.json 2> gpurun_out/r02_dp1_$name.err
  tail -1 gpurun_out/r02_dp1_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); n=d['n_gpus']; print('1gpu $name'.ljust(22), 'value ms', round(d['ms_per_step'],3), 'e2e ms', round(1e3*n/d['e2e']['value'],3))"
}
if [ "$N" = "2" ]; then
one plain A=1
one plain_b A=1
fi
