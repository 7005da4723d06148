"""Runs a parity case on a given device through dlgrad_b200's public Tensor API and returns numpy."""
from __future__ import annotations

import numpy as np

from dlgrad_b200 import Tensor

# cases whose forward AND backward are pure selection / data movement / exact IEEE single ops
# (division is NOT here: under -ffast-math gcc turns the reference's x / y into x * (1/y) with an
# approximate reciprocal refinement, so the reference itself is ~1 ulp off IEEE division)
EXACT_FWD = ("add_", "sub_", "mul_", "neg", "relu", "leaky_relu", "clamp", "permute_", "max_")
EXACT_BWD = ("neg", "relu", "permute_")

# tolerances of BASELINE.json's north star: 1e-5 relative for elementwise / reduce / softmax,
# 1e-4 relative for matmul.  "relative" is taken against the largest magnitude of the expected
# tensor (atol = tol * max|expected|) plus the same rtol per element, so that sums that cancel to
# ~0 are not asked for more digits than fp32 carries.
TOL_ELEMENTWISE = 1e-5
TOL_MATMUL = 1e-4


def tol_for(name: str) -> float:
    return TOL_MATMUL if name.startswith("matmul") else TOL_ELEMENTWISE


def is_exact_fwd(name: str) -> bool:
    return name.startswith(EXACT_FWD)


def is_exact_bwd(name: str) -> bool:
    return name.startswith(EXACT_BWD)


def run_case(name, fn, inputs, diff, device, w=None) -> dict:
    ts = [Tensor(np.array(a, dtype=np.float32), requires_grad=diff, device=device) for a in inputs]
    y = fn(*ts)
    res = {"out": y.numpy().reshape(y.shape)}
    if diff:
        if y.shape:
            if w is None:
                w = np.random.default_rng(99).random(y.shape, dtype=np.float32).astype(np.float32)
            loss = (y * Tensor(np.array(w, dtype=np.float32), device=device)).sum()
        else:
            loss = y
        loss.backward()
        for i, t in enumerate(ts):
            if t.grad is not None:
                res[f"grad{i}"] = t.grad.numpy().reshape(t.grad.shape)
    return res


def rel_check(check, actual, expected, tol, what):
    scale = float(np.max(np.abs(expected))) if np.size(expected) else 1.0
    scale = scale if np.isfinite(scale) and scale > 0 else 1.0
    check(actual, expected, rtol=tol, atol=tol * scale, what=what)
