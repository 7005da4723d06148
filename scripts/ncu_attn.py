"""A few launches of the fused attention kernels and the tile-skipping GEMMs at C5 size, for `ncu --set full`
(run once plain, then under ncu: scripts/gpu_ncu7.sh)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from dlgrad_b200 import Tensor  # noqa: E402
from dlgrad_b200.runtime.cuda import shim  # noqa: E402

rng = np.random.default_rng(0)
B, H, T, D = 8, 12, 1024, 64
q, k, v = (Tensor((rng.random((B, H, T, D), dtype=np.float32) - 0.5), device="cuda", requires_grad=True) for _ in range(3))
mask = Tensor(np.triu(np.ones((T, T), dtype=np.float32), 1), device="cuda")
w = Tensor(rng.random((B, H, T, D), dtype=np.float32), device="cuda")
for _ in range(2):
    att = ((q @ k.transpose(2, 3)) * 0.125).masked_fill(mask, float("-inf")).softmax(dim=3)
    y = att @ v
    (y * w).sum().backward()
    q.grad = k.grad = v.grad = None
shim.synchronize()
print("ok")
