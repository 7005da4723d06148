"""Cross-check of the oracle's C restatement against the reference's OWN generated kernels
(oracle/_ref/*.so, compiled by oracle/build.py from /root/reference/dlgrad/codegen/cpu_kernel.py with
the reference's flags).  Skipped where oracle/_ref is absent.

Exact ops must agree bit for bit.  Sums may differ in the last bits because gcc vectorises the
reference's rank-specific loop nest and the restatement's (outer, R, inner) nest with different
partial-sum orders under -ffast-math; the bound asserted here (1e-5 of the largest magnitude) is the
same one the CUDA path is held to.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import build

REF = build.REF_DIR
pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or not os.listdir(REF),
                                reason="oracle/_ref not built (needs /root/reference)")
fp = C.POINTER(C.c_float)


def ia(*v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def P(a):
    return a.ctypes.data_as(fp)


def ref(stem):
    return C.CDLL(os.path.join(REF, f"{stem}.so"))


@pytest.fixture(scope="module")
def orc():
    return C.CDLL(build.build_oracle())


def strides(shape):
    s, acc = [], 1
    for d in reversed(shape):
        s.append(acc)
        acc *= d
    return list(reversed(s))


def pad4(v, fill):
    return [fill] * (4 - len(v)) + list(v)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "eq", "gt", "gte"])
@pytest.mark.parametrize("xs,ys", [((2, 3, 4, 5), (2, 3, 4, 5)), ((2, 3, 4, 5), (1, 3, 1, 5)), ((64, 48), (64, 1)),
                                   ((64, 48), (1, 48)), ((7,), (7,))])
def test_binary_bit_exact(orc, op, xs, ys):
    rng = np.random.default_rng(0)
    x = (rng.random(xs, dtype=np.float32) * 4 - 2).astype(np.float32)
    y = (rng.random(ys, dtype=np.float32) + 0.5).astype(np.float32)
    if op in ("eq", "gt", "gte"):
        x, y = np.round(x), np.round(y)
    nd = len(xs)
    out_shape = np.broadcast_shapes(xs, ys)
    bx = [0 if xs[i] == 1 and out_shape[i] > 1 else strides(xs)[i] for i in range(nd)]
    by = [0 if ys[i] == 1 and out_shape[i] > 1 else strides(ys)[i] for i in range(nd)]
    a = np.empty(out_shape, dtype=np.float32)
    b = np.empty(out_shape, dtype=np.float32)
    getattr(ref(f"{op}_{nd}d"), f"{op}_{nd}d")(P(x), P(y), P(a), ia(*out_shape), ia(*bx), ia(*by))
    getattr(orc, f"orc_{op}")(P(x), P(y), P(b), ia(*pad4(out_shape, 1)), ia(*pad4(bx, 0)), ia(*pad4(by, 0)))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("fn", ["neg", "exp", "log", "sqrt", "rsqrt"])
def test_unary(orc, fn):
    rng = np.random.default_rng(1)
    x = (rng.random(4099, dtype=np.float32) * 3 + 0.05).astype(np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    getattr(ref(f"c_{fn}"), f"c_{fn}")(P(x), P(a), C.c_int(x.size))
    getattr(orc, f"orc_{fn}")(P(x), P(b), C.c_int(x.size))
    # same source expression and flags except -march (x86-64-v3 for the travelling _ref objects,
    # native for the oracle): the vector-libm variant chosen may differ by an ulp
    np.testing.assert_allclose(a, b, rtol=3e-7)


@pytest.mark.parametrize("val", [2, 3, -1, 0, 1])
def test_pow(orc, val):
    x = (np.random.default_rng(2).random(1000, dtype=np.float32) + 0.25).astype(np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    ref("c_pow").c_pow(P(x), P(a), C.c_int(x.size), C.c_int(val))
    orc.orc_pow(P(x), P(b), C.c_int(x.size), C.c_int(val))
    np.testing.assert_allclose(a, b, rtol=3e-7)


@pytest.mark.parametrize("op", ["sum", "max", "mean"])
@pytest.mark.parametrize("shape,dim", [((2, 3, 4, 5), 0), ((2, 3, 4, 5), 2), ((2, 2, 16, 128), 3), ((64, 48), 0),
                                       ((64, 48), 1), ((513, 300), 0), ((513, 300), 1), ((5, 33, 36), 1)])
def test_reduce_axis(orc, op, shape, dim):
    x = (np.random.default_rng(3).random(shape, dtype=np.float32) * 4 - 2).astype(np.float32)
    nd = len(shape)
    oshape = [s for i, s in enumerate(shape) if i != dim]
    a, b = np.empty(oshape, dtype=np.float32), np.empty(oshape, dtype=np.float32)
    getattr(ref(f"reduce_{op}_{nd}d_axis{dim}"), f"reduce_{op}_{nd}d_axis{dim}")(
        P(x), P(a), ia(*shape), ia(*strides(shape)), ia(*strides(oshape)))
    outer, R, inner = int(np.prod(shape[:dim])), shape[dim], int(np.prod(shape[dim + 1:]))
    getattr(orc, f"orc_reduce_{op}")(P(x), P(b), C.c_int(outer), C.c_int(R), C.c_int(inner))
    if op == "max":
        assert np.array_equal(a, b)
    else:
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-5 * float(np.abs(a).max()))


@pytest.mark.parametrize("op", ["sum", "max", "mean"])
def test_reduce_all(orc, op):
    x = (np.random.default_rng(4).random(100003, dtype=np.float32)).astype(np.float32)
    a, b = np.empty(1, dtype=np.float32), np.empty(1, dtype=np.float32)
    getattr(ref(f"reduce_{op}"), f"reduce_{op}")(P(x), P(a), C.c_int(x.size))
    getattr(orc, f"orc_reduce_all_{op}")(P(x), P(b), C.c_int(x.size))
    np.testing.assert_allclose(a, b, rtol=1e-5 if op != "max" else 0)


def test_matmul_2d_3d_4d(orc):
    rng = np.random.default_rng(5)
    x = (rng.random((33, 20), dtype=np.float32) - 0.5).astype(np.float32)
    y = (rng.random((20, 17), dtype=np.float32) - 0.5).astype(np.float32)
    a = np.zeros((33, 17), dtype=np.float32)
    b = np.zeros((33, 17), dtype=np.float32)
    ref("matmul_2d").matmul_2d(P(x), P(y), P(a), ia(33, 20), ia(20, 1), ia(20, 17), ia(17, 1))
    orc.orc_matmul(P(x), P(y), P(b), *[C.c_int(v) for v in (1, 1, 33, 20, 17, 0, 0, 0, 0)])
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-6 * float(np.abs(a).max()))
    # 3-D with the second operand broadcast over the batch (stride 0), as Linear layers do
    x3 = (rng.random((4, 33, 20), dtype=np.float32) - 0.5).astype(np.float32)
    a = np.zeros((4, 33, 17), dtype=np.float32)
    b = np.zeros((4, 33, 17), dtype=np.float32)
    ref("matmul_3d").matmul_3d(P(x3), P(y), P(a), C.c_int(4), C.c_int(33), C.c_int(20), C.c_int(17),
                               ia(660, 20, 1), ia(0, 17, 1), ia(561, 17, 1))
    orc.orc_matmul(P(x3), P(y), P(b), *[C.c_int(v) for v in (1, 4, 33, 20, 17, 0, 660, 0, 0)])
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-6 * float(np.abs(a).max()))
    x4 = (rng.random((2, 3, 16, 8), dtype=np.float32) - 0.5).astype(np.float32)
    y4 = (rng.random((2, 3, 8, 16), dtype=np.float32) - 0.5).astype(np.float32)
    a = np.zeros((2, 3, 16, 16), dtype=np.float32)
    b = np.zeros((2, 3, 16, 16), dtype=np.float32)
    ref("matmul_4d").matmul_4d(P(x4), P(y4), P(a), ia(2, 3, 16, 8, 16), ia(384, 128, 8, 1), ia(384, 128, 16, 1),
                               ia(768, 256, 16, 1))
    orc.orc_matmul(P(x4), P(y4), P(b), *[C.c_int(v) for v in (2, 3, 16, 8, 16, 384, 128, 384, 128)])
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-6 * float(np.abs(a).max()))


def test_permute_where_masked_fill_argmax_bit_exact(orc):
    rng = np.random.default_rng(6)
    x = rng.random((2, 5, 12, 36), dtype=np.float32)
    order = (0, 2, 1, 3)
    oshape = [x.shape[i] for i in order]
    pstr = [strides(x.shape)[i] for i in order]
    a, b = np.empty(oshape, dtype=np.float32), np.empty(oshape, dtype=np.float32)
    ref("permute_4d").permute_4d(P(x), P(a), ia(*oshape), ia(*pstr), ia(*strides(oshape)))
    orc.orc_permute(P(x), P(b), ia(*oshape), ia(*pstr))
    assert np.array_equal(a, b) and np.array_equal(a, x.transpose(order))
    cond = (rng.random((5, 6)) > 0.5).astype(np.float32)
    xv, yv = rng.random((5, 6), dtype=np.float32), rng.random((1,), dtype=np.float32)
    a, b = np.empty((5, 6), dtype=np.float32), np.empty((5, 6), dtype=np.float32)
    ref("where_2d").where_2d(P(cond), P(xv), P(yv), P(a), ia(5, 6), ia(6, 1), ia(6, 1), ia(0, 0), ia(6, 1))
    orc.orc_where(P(cond), P(xv), P(yv), P(b), ia(1, 1, 5, 6), ia(0, 0, 6, 1), ia(0, 0, 6, 1), ia(0, 0, 0, 0))
    assert np.array_equal(a, b)
    big = rng.random((2, 3, 6, 6), dtype=np.float32)
    mask = np.triu(np.ones((6, 6), dtype=np.float32), 1)
    a, b = np.empty_like(big), np.empty_like(big)
    ref("masked_fill_4d").masked_fill_4d(P(big), P(mask), P(a), ia(2, 3, 6, 6), ia(108, 36, 6, 1), ia(0, 0, 6, 1),
                                         ia(108, 36, 6, 1), C.c_float(float("-inf")))
    orc.orc_masked_fill(P(big), P(mask), P(b), ia(2, 3, 6, 6), ia(108, 36, 6, 1), ia(0, 0, 6, 1), C.c_float(float("-inf")))
    assert np.array_equal(a, b) and np.isneginf(a).sum() == 2 * 3 * 15
    t = rng.integers(0, 4, size=(37, 11)).astype(np.float32)
    for d, oshape in ((0, (11,)), (1, (37,))):
        a, b = np.empty(oshape, dtype=np.float32), np.empty(oshape, dtype=np.float32)
        getattr(ref(f"argmax_2d_axis{d}"), f"argmax_2d_axis{d}")(P(t), P(a), ia(37, 11), ia(11, 1), ia(1))
        outer, R, inner = (1, 37, 11) if d == 0 else (37, 11, 1)
        orc.orc_argmax(P(t), P(b), C.c_int(outer), C.c_int(R), C.c_int(inner))
        assert np.array_equal(a, b) and np.array_equal(a, t.argmax(d).astype(np.float32))


def test_ce_and_embedding_bit_exact(orc):
    rng = np.random.default_rng(7)
    x = rng.random((32, 10), dtype=np.float32)
    t = rng.integers(0, 10, size=32).astype(np.float32)
    a, b = np.empty(32, dtype=np.float32), np.empty(32, dtype=np.float32)
    ref("ce_forward").ce_forward(P(x), P(t), P(a), C.c_int(32), ia(10, 1))
    orc.orc_ce_forward(P(x), P(t), P(b), C.c_int(32), C.c_int(10))
    assert np.array_equal(a, b)
    xa, xb = x.copy(), x.copy()
    ref("ce_backward").ce_backward(P(xa), P(t), ia(32, 10), ia(10, 1))
    orc.orc_ce_backward(P(xb), P(t), C.c_int(32), C.c_int(10))
    assert np.array_equal(xa, xb)
    W = rng.random((11, 8), dtype=np.float32)
    idx = rng.integers(0, 11, size=54).astype(np.float32)
    a, b = np.empty((54, 8), dtype=np.float32), np.empty((54, 8), dtype=np.float32)
    ref("embedding").embedding(P(W), P(idx), P(a), C.c_int(54), C.c_int(8))
    orc.orc_embedding(P(W), P(idx), P(b), C.c_int(54), C.c_int(8))
    assert np.array_equal(a, b)
    g = rng.random((54, 8), dtype=np.float32)
    a, b = np.zeros((11, 8), dtype=np.float32), np.zeros((11, 8), dtype=np.float32)
    ref("embedding_backward").embedding_backward(P(a), P(g), P(idx), C.c_int(54), C.c_int(8))
    orc.orc_embedding_backward(P(b), P(g), P(idx), C.c_int(54), C.c_int(8))
    assert np.array_equal(a, b)
