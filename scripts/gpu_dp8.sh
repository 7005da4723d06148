mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_dp8.json 2> gpurun_out/r02_bench_dp8.err
echo "bench exit $?"
tail -1 gpurun_out/r02_bench_dp8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n_gpus', d['n_gpus'], round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'replicas_identical', d.get('replicas_identical'), d['clocks'].get('sm_mhz'))" || tail -8 gpurun_out/r02_bench_dp8.err
timeout 300 $RUN --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_dp8_b.json 2> gpurun_out/r02_bench_dp8_b.err
tail -1 gpurun_out/r02_bench_dp8_b.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n_gpus', d['n_gpus'], round(d['ms_per_step'],2), 'ms/step', round(d['value'],2))"
