"""Timing (and, with DLGRAD_TC_TRACE=1, per-key-tile timestamps) of the fused attention-backward kernel at C5 size.
    python scripts/attn_bwd_trace.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from dlgrad_b200 import Tensor  # noqa: E402
from dlgrad_b200.runtime.cuda import attention, profile, shim, transfer  # noqa: E402

shim.init()
rng = np.random.default_rng(0)
B, H, T, D = 8, 12, 1024, 64
mk = lambda *s: Tensor((rng.random(s, dtype=np.float32) - 0.5), device="cuda").data  # noqa: E731
dO, v, O = mk(B, H, T, D), mk(B, H, T, D), mk(B, H, T, D)
P = Tensor(rng.random((B, H, T, T), dtype=np.float32), device="cuda").data
mask = Tensor(np.triu(np.ones((T, T), dtype=np.float32), 1), device="cuda").data
for _ in range(3):
    out = attention.dscores(dO, v, O, P, mask, 0.125)
shim.synchronize()
e0 = profile.Event().record()
for _ in range(10):
    out = attention.dscores(dO, v, O, P, mask, 0.125)
e1 = profile.Event().record()
e1.sync()
print(f"attn_dscores C5: {e1.ms_since(e0) * 100:.1f} us per launch (incl. rowdot + mask_bits)")
if os.environ.get("DLGRAD_TC_TRACE") == "1" and not shim.COMPILE_ONLY:
    host = transfer.download(out, 320).view(np.int64).reshape(5, 32)
    t0 = host[4, 0]
    print("  u  tma(first kb)  mma start  mma issued  rows start  rows done | mma busy  rows busy  d(mma start)")
    for i in range(20):
        t = host[:, i] - t0
        d = host[0, i] - host[0, i - 1] if i else 0
        print(f"{i:3d} {t[4]:12d} {t[0]:10d} {t[1]:11d} {t[2]:11d} {t[3]:10d} | {t[1]-t[0]:8d} {t[3]-t[2]:10d} {d:10d}")
