#!/bin/bash
# same-box A/B of the two schedules of distributed.Adam on N GPUs (bash scripts/gpu_dp_overlap.sh N [extra env for every run]):
# monolithic exchange kernel vs parameter buckets exchanged on the comm stream during backward, with per-step device times;
# last_loss must agree bit for bit.  DLGRAD_BENCH_GC_FREEZE=0 as extra env reproduces the collector pause of
# profiles/r02_dp_overlap_and_gc_stall.md.
N=${1:-2}; shift
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29541
run() {
  name=$1; shift
  port=$((port+1))
  env "$@" DLGRAD_BENCH_STEP_TIMES=1 timeout 300 $RUN --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_dp${N}_$name.json 2> gpurun_out/r02_dp${N}_$name.err
  tail -1 gpurun_out/r02_dp${N}_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); n=d['n_gpus']; print('$name'.ljust(12), 'value ms', round(d['ms_per_step'],3), 'steps/s', round(d['value'],2), 'e2e ms', round(1e3*n/d['e2e']['value'],3), 'last_loss', repr(d['last_loss']), 'replicas_identical', d.get('replicas_identical'), d['clocks'])" || tail -8 gpurun_out/r02_dp${N}_$name.err
  grep "rank 0.*region A per-step" gpurun_out/r02_dp${N}_$name.err
}
run mono DLGRAD_DP_OVERLAP=0 "$@"
run over DLGRAD_DP_OVERLAP=1 "$@"
run mono_b DLGRAD_DP_OVERLAP=0 "$@"
