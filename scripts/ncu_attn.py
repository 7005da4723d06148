"""Attention-shaped batched GEMMs for an ncu --set full capture."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200 import Tensor
from dlgrad_b200.runtime.cuda import shim
rng = np.random.default_rng(0)
T = lambda *s: Tensor(rng.random(s, dtype=np.float32), device="cuda")   # noqa: E731
q, kT, v, att = T(8, 12, 1024, 64), T(8, 12, 64, 1024), T(8, 12, 1024, 64), T(8, 12, 1024, 1024)
for _ in range(2):
    s = q @ kT          # QK^T-like: (1024 x 1024 x 64) x 96, output-bound
    y = att @ v         # AV-like:   (1024 x 64 x 1024) x 96, input-bound
shim.synchronize()
print("ok")
