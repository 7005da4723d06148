#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --ops-only > gpurun_out/ops.log 2>&1; echo "exit $?" >> gpurun_out/ops.log; tail -20 gpurun_out/ops.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-ops --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], d["clocks"])
r=d["roofline"]; print(r["achieved"], r["frac"], r["matmul_share_of_step"])
for t in r["by_shape"]: print(t)
PY
