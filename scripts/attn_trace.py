"""Per-key-tile timestamps of the fused attention-scores kernel (CTA 0, first 32 tile passes).
    DLGRAD_TC_TRACE=1 python scripts/attn_trace.py"""
import os
import sys
os.environ["DLGRAD_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from dlgrad_b200 import Tensor  # noqa: E402
from dlgrad_b200.runtime.cuda import shim, transfer  # noqa: E402

shim.init()
rng = np.random.default_rng(0)
B, H, T, D = 8, 12, 1024, 64
q = Tensor((rng.random((B, H, T, D), dtype=np.float32) - 0.5), device="cuda")
k = Tensor((rng.random((B, H, T, D), dtype=np.float32) - 0.5), device="cuda")
mask = Tensor(np.triu(np.ones((T, T), dtype=np.float32), 1), device="cuda")
for _ in range(3):
    out = ((q @ k.transpose(2, 3)) * 0.125).masked_fill(mask, float("-inf")).softmax(dim=3)
shim.synchronize()
if shim.COMPILE_ONLY:
    sys.exit(0)
host = transfer.download(out.data.ptr, 320).view(np.int64).reshape(5, 32)
t0 = host[4, 0]
print("  u  tma(first kb)  mma start  mma issued  rows start  rows done | mma busy  rows busy  d(mma start)")
for i in range(32):
    t = host[:, i] - t0
    d = host[0, i] - host[0, i - 1] if i else 0
    print(f"{i:3d} {t[4]:12d} {t[0]:10d} {t[1]:11d} {t[2]:11d} {t[3]:10d} | {t[1]-t[0]:8d} {t[3]-t[2]:10d} {d:10d}")
