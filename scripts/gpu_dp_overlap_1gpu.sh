#!/bin/bash
# what do the comm-stream kernels of the overlapped exchange cost the backward next to them?  One GPU, world == 1.
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-small --no-ops > gpurun_out/r02_ov1gpu_$name.json 2> gpurun_out/r02_ov1gpu_$name.err
  tail -1 gpurun_out/r02_ov1gpu_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],3), 'ms/step', 'launches', d['gpu_launches']//d['steps'], 'last_loss', repr(d['last_loss']), d['clocks'].get('sm_mhz'))" || tail -5 gpurun_out/r02_ov1gpu_$name.err
}
timeout 300 python -m pytest tests/test_cuda_parity.py -q -m gpu -p no:cacheprovider -k "dp_adam or fused_dp" -x 2>&1 | tail -3
run plain_adam A=1
run overlap_lowprio DLGRAD_DP_FUSED=force
run overlap_highprio DLGRAD_DP_FUSED=force DLGRAD_COMM_PRIORITY=high
run monolithic DLGRAD_DP_FUSED=force DLGRAD_DP_OVERLAP=0
run overlap_lowprio_b DLGRAD_DP_FUSED=force
run overlap_t256 DLGRAD_DP_FUSED=force DLGRAD_DP_THREADS=256
run overlap_bucket50 DLGRAD_DP_FUSED=force DLGRAD_DP_BUCKET_MB=50
run plain_adam_b A=1
