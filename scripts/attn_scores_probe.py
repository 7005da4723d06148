"""Fused attention-scores kernel vs the eager chain (GEMM -> scale+mask+softmax kernel): values and time.

    python scripts/attn_scores_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dlgrad_b200 import Tensor  # noqa: E402
from dlgrad_b200.runtime.cuda import attention, profile, shim  # noqa: E402

shim.init()
rng = np.random.default_rng(0)


def run(B, H, T, D, mask_np, fill, fused, reps=0):
    attention.ENABLED = fused
    q = Tensor((rng.random((B, H, T, D), dtype=np.float32) - 0.5), device="cuda")
    k = Tensor((rng.random((B, H, T, D), dtype=np.float32) - 0.5), device="cuda")
    mask = Tensor(mask_np.copy(), device="cuda")
    scale = 1.0 / np.sqrt(D)

    def f():
        return ((q @ k.transpose(2, 3)) * scale).masked_fill(mask, fill).softmax(dim=3)
    out = f()
    res = out.numpy()
    ms = None
    if reps:
        for _ in range(3):
            out = f()
        shim.synchronize()
        e0 = profile.Event().record()
        for _ in range(reps):
            out = f()
        e1 = profile.Event().record()
        e1.sync()
        ms = e1.ms_since(e0) / reps
    qn, kn = q.numpy().astype(np.float64), k.numpy().astype(np.float64)
    s = np.einsum("bhqd,bhkd->bhqk", qn, kn) * scale
    s = np.where(mask_np > 0, fill, s)
    with np.errstate(invalid="ignore"):
        e = np.exp(s - s.max(axis=3, keepdims=True))
        ref = e / e.sum(axis=3, keepdims=True)
    return res, ref, ms


for name, (B, H, T, D) in {"small": (2, 4, 256, 64), "D=32": (2, 8, 256, 32), "D=128": (1, 4, 512, 128), "C5": (8, 12, 1024, 64)}.items():
    causal = np.triu(np.ones((T, T), dtype=np.float32), 1)
    for mname, mk, fill in (("causal -inf", causal, float("-inf")), ("random -1e9", (rng.random((T, T)) > 0.6).astype(np.float32), -1e9),
                            ("random -inf + dead rows", np.maximum((rng.random((T, T)) > 0.5), (np.arange(T)[:, None] % 37 == 0)).astype(np.float32), float("-inf"))):
        rng = np.random.default_rng(1)
        got_f, ref, ms_f = run(B, H, T, D, mk, fill, True, reps=10 if name == "C5" else 0)
        rng = np.random.default_rng(1)
        got_e, _, ms_e = run(B, H, T, D, mk, fill, False, reps=10 if name == "C5" else 0)
        if shim.COMPILE_ONLY:
            continue
        nan_ok = bool((np.isnan(got_f) == np.isnan(ref)).all())
        ok = np.isfinite(ref)
        err_f = float(np.abs(got_f[ok] - ref[ok]).max())
        err_e = float(np.abs(got_e[ok] - ref[ok]).max())
        t = f"  fused {ms_f * 1e3:7.1f} us  eager {ms_e * 1e3:7.1f} us" if ms_f else ""
        print(f"{name:6s} {mname:26s} max|err| fused {err_f:.2e} eager {err_e:.2e}  NaN pattern ok: {nan_ok}{t}", flush=True)
