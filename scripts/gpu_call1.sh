#!/bin/bash
# first GPU call: environment facts, smoke, parity tests, short bench, TF32 peak probe
mkdir -p gpurun_out
{ lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-300; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv; nproc; } > gpurun_out/env.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/test_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python - > gpurun_out/tf32_peak.txt 2>&1 <<'PY'
import torch, time
torch.backends.cuda.matmul.allow_tf32 = True
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
for _ in range(3): (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("tf32 8192^3 burst TFLOP/s", 2 * 8192**3 / (best * 1e-3) / 1e12)
t0 = time.time(); n = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(10): c = a @ b
    n += 10; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print("tf32 8192^3 sustained TFLOP/s", n * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
torch.backends.cuda.matmul.allow_tf32 = False
for _ in range(2): (a @ b)
torch.cuda.synchronize(); e0.record()
for _ in range(5): c = a @ b
e1.record(); torch.cuda.synchronize()
print("fp32 (no tf32) 8192^3 TFLOP/s", 5 * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
PY
tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/test_gpu.log; cat gpurun_out/bench.log | cut -c1-1500; tail -5 gpurun_out/bench.err; cat gpurun_out/tf32_peak.txt
