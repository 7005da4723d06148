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


# forward values of these cases are pointwise functions of their inputs (no cancelling sum): the stated tolerance is
# asserted PER ELEMENT on every entry above 1e-3 x max, not only against the largest magnitude
ELEMENTWISE_FWD = ("add_", "sub_", "mul_", "div_", "exp", "log", "pow", "rsqrt", "sqrt", "sigmoid", "softmax",
                   "leaky_relu", "clamp", "neg", "relu", "max_")


def is_elementwise_fwd(name: str) -> bool:
    return name.startswith(ELEMENTWISE_FWD)


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


# Per-element criterion: the max-norm bound above lets a small entry (a softmax tail, a small
# gradient) be wholly wrong.  Every comparison therefore ALSO measures, for the entries with |expected| > 1e-3 * max,
# the worst per-element relative error, and for the rest the worst error in ulps of the largest magnitude; both figures
# are logged for every comparison (DLGRAD_PARITY_REPORT=<path>, profiles/r02_parity_report_summary.md).
#   * `elementwise=True` (forward values of pointwise / softmax work: no cancelling sum) ASSERTS the per-element
#     relative error against the same stated tolerance — measured on B200: worst case 0.04 x tolerance;
#   * dot products get the componentwise bound of a dot product instead (`gemm_check`: |err_ij| <= tol * (|A| |B|)_ij —
#     fp32 does not carry more digits than that in a sum that cancels);
#   * gradients of whole models are sums of cancelling terms of unknown operands: reported, max-norm asserted.
PER_ELEMENT_FLOOR = 1e-3
REPORT: list = []


def per_element_stats(actual, expected) -> dict:
    a = np.asarray(actual, dtype=np.float64).reshape(-1)
    e = np.asarray(expected, dtype=np.float64).reshape(-1)
    fin = np.isfinite(e) & np.isfinite(a)
    if a.size != e.size or not fin.any():
        return {"n": int(e.size), "big": 0, "worst_rel": 0.0, "rest_ulp": 0.0}
    a, e = a[fin], e[fin]
    scale = float(np.max(np.abs(e)))
    if not scale > 0:
        return {"n": int(e.size), "big": 0, "worst_rel": 0.0, "rest_ulp": float(np.max(np.abs(a)) > 0)}
    big = np.abs(e) > PER_ELEMENT_FLOOR * scale
    worst = float(np.max(np.abs(a[big] - e[big]) / np.abs(e[big]))) if big.any() else 0.0
    ulp = float(np.spacing(np.float32(scale)))
    rest = float(np.max(np.abs(a[~big] - e[~big])) / ulp) if (~big).any() else 0.0
    return {"n": int(e.size), "big": int(big.sum()), "worst_rel": worst, "rest_ulp": rest}


def rel_check(check, actual, expected, tol, what, elementwise: bool = False):
    from tests.conftest import COMPILE_ONLY
    scale = float(np.max(np.abs(expected))) if np.size(expected) else 1.0
    scale = scale if np.isfinite(scale) and scale > 0 else 1.0
    check(actual, expected, rtol=tol, atol=tol * scale, what=what)
    if COMPILE_ONLY:
        return
    st = per_element_stats(actual, expected)
    st.update(what=what, tol=tol, elementwise=elementwise)
    REPORT.append(st)
    if elementwise:
        # 4 ulp of slack at the element itself: an entry at the floor of a 1e-5 comparison is only ~80 ulp wide
        assert st["worst_rel"] <= tol + 4 * 6e-8, (
            f"{what}: per-element relative error {st['worst_rel']:.3g} on entries above {PER_ELEMENT_FLOOR:g} x max "
            f"exceeds {tol:g}")


def gemm_check(check, actual, A, B, tol, what, bias=None):
    """`actual` against A @ B (+ bias) in float64 with the COMPONENTWISE bound of a dot product:
    |actual_ij - exact_ij| <= tol * (sum_k |A_ik| |B_kj| + |bias_j|) for every single entry (Higham, Accuracy and
    Stability of Numerical Algorithms, §3.5), in addition to the max-norm comparison of rel_check."""
    from tests.conftest import COMPILE_ONLY
    A64, B64 = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    exact = A64 @ B64 + (0.0 if bias is None else np.asarray(bias, dtype=np.float64))
    rel_check(check, actual, exact, tol, what)
    if COMPILE_ONLY:
        return
    bound = np.abs(A64) @ np.abs(B64) + (0.0 if bias is None else np.abs(np.asarray(bias, dtype=np.float64)))
    err = np.abs(np.asarray(actual, dtype=np.float64).reshape(exact.shape) - exact)
    worst = float(np.max(err / np.maximum(bound, 1e-300)))
    REPORT.append({"what": what + " [componentwise]", "tol": tol, "worst_rel": worst, "n": int(exact.size), "big": int(exact.size),
                   "rest_ulp": 0.0, "elementwise": True})
    assert worst <= tol, f"{what}: componentwise error {worst:.3g} x (|A||B|)_ij exceeds {tol:g}"
