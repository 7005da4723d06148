#!/bin/bash
# overlapped data-parallel exchange on N GPUs: single-GPU schedule tests, the bit-level multi-rank check, and bench.py at N
# with the overlapped schedule and with the monolithic one (same box, alternating), last_loss must agree bit for bit
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_cuda_parity.py -q -m gpu -p no:cacheprovider -k "dp_adam or fused_dp" -x > gpurun_out/r02_dp_tests.txt 2>&1
echo "dp tests exit $?"; tail -5 gpurun_out/r02_dp_tests.txt
NCCL_DEBUG=WARN timeout 300 $RUN --master-port 29533 scripts/dp_fused_check.py > gpurun_out/r02_dp_overlap_check_$N.txt 2> gpurun_out/r02_dp_overlap_check_$N.err
echo "check exit $?"; tail -6 gpurun_out/r02_dp_overlap_check_$N.txt; tail -5 gpurun_out/r02_dp_overlap_check_$N.err
port=29541
for rep in 1 2; do
  for ov in 1 0; do
    port=$((port+1))
    DLGRAD_DP_OVERLAP=$ov timeout 300 $RUN --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_bench_dp${N}_ov${ov}_$rep.json 2> gpurun_out/r02_bench_dp${N}_ov${ov}_$rep.err
    echo "bench overlap=$ov rep $rep exit $?"
    tail -1 gpurun_out/r02_bench_dp${N}_ov${ov}_$rep.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  n_gpus', d['n_gpus'], round(d['ms_per_step'],3), 'ms/step', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'last_loss', repr(d['last_loss']), 'replicas_identical', d.get('replicas_identical'), d['clocks'].get('sm_mhz'))" || tail -8 gpurun_out/r02_bench_dp${N}_ov${ov}_$rep.err
  done
done
if [ "$N" = "2" ]; then
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_bench_1gpu_samebox.json 2> gpurun_out/r02_bench_1gpu_samebox.err
  tail -1 gpurun_out/r02_bench_1gpu_samebox.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  1 gpu same box', round(d['ms_per_step'],3), 'ms/step', round(d['value'],2), 'steps/s')"
fi
