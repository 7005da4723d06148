"""ncu target: the HBM-bound kernels added late in round 1, at C5 sizes on one GPU —
dp_adam (world 1: flat-bucket Adam, the single-rank form of the fused data-parallel step), pack_multi (all gradients
into the flat bucket in one launch), uniform_philox (4096^2 draws), and the split-K weight-gradient GEMM with its
in-kernel finish."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dlgrad_b200 import Tensor, distributed  # noqa: E402
from dlgrad_b200.runtime.cuda import shim  # noqa: E402

cfg, gpt, params, opt = bench.build_gpt(bench.C5, "cuda", fused_dp=True)       # distributed.Adam at world 1
rng = np.random.default_rng(0)
for p in params:
    p.grad = Tensor(rng.standard_normal(p.shape).astype(np.float32), device="cuda")
for _ in range(3):
    opt.step()
for _ in range(3):
    u = Tensor.uniform((4096, 4096), device="cuda")
x = Tensor(rng.standard_normal((8192, 768)).astype(np.float32), device="cuda", requires_grad=True)
W = Tensor(rng.standard_normal((768, 768)).astype(np.float32), device="cuda", requires_grad=True)
for _ in range(3):
    W.grad = None
    x.linear(W, None).sum().backward()
shim.synchronize()
print("ok", float(u.numpy().mean()))
