"""One attention block at the benchmark's size (B=8, h=12, T=1024, D=64), forward + backward, timed with CUDA events; the
target of the ncu captures of the flash kernels (profiles/r02_ncu_flash_*.md)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200 import Tensor
from dlgrad_b200.runtime.cuda import attention, profile, shim

B, H, T, D = int(os.environ.get("B", "8")), 12, int(os.environ.get("T", "1024")), int(os.environ.get("D", "64"))
rng = np.random.default_rng(0)
split = os.environ.get("LAYOUT", "split") == "split"          # q, k, v stored (B, T, h, D) like the projections produce them
shape = (B, T, H, D) if split else (B, H, T, D)
q, k, v = ((rng.random(shape, dtype=np.float32) - 0.5).astype(np.float32) for _ in range(3))
mask = Tensor(np.triu(np.ones((T, T), dtype=np.float32), 1), device="cuda")
w = Tensor(np.ones(shape, dtype=np.float32), device="cuda")
scale = float(1 / np.sqrt(D))
reps = int(os.environ.get("REPS", "5"))
for r in range(reps):
    Q0, K0, V0 = (Tensor(a, device="cuda", requires_grad=True) for a in (q, k, v))
    Q, K, V = ((t.transpose(1, 2) for t in (Q0, K0, V0)) if split else (Q0, K0, V0))
    profile.matmul_timer = profile.KernelTimer()
    y = ((Q @ K.transpose(2, 3)) * scale).masked_fill(mask, float("-inf")).softmax(dim=3) @ V
    if split:
        y = y.transpose(1, 2)
    (y * w).sum().backward()
    shim.synchronize()
    mt = profile.matmul_timer.collect()
    profile.matmul_timer = None
for tag, val in sorted(mt["by_tag"].items(), key=lambda kv: -kv[1][1]):
    print(tag, f"{1e3 * val[1] / val[0]:.1f} us")
