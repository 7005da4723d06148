"""Oracle runtime: (op, Device.CPU) functions over oracle/kernels.c.  TEST INFRASTRUCTURE ONLY.

`install()` registers these with dlgrad_b200's dispatcher so that the SAME Tensor / ops / nn / optim
layer can be executed on the CPU restatement of the reference's kernels and compared with
Device.CUDA on identical inputs.  Fused ops of the CUDA backend (softmax, cross-entropy gradient,
optimizer steps) are registered here as the reference's compositions of primitive kernels, op for
op (citations per function).  Nothing in dlgrad_b200/ imports this module; without install() a
Device.CPU dispatch raises.

Each function mirrors the argument handling of its counterpart in dlgrad/runtime/cpu.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from dlgrad_b200.buffer import Buffer
from dlgrad_b200.device import Device
from dlgrad_b200.dispatch import dispatcher
from dlgrad_b200.helpers import (
    BinaryOps,
    BufferOps,
    CustomOps,
    UnaryOps,
    broadcast_shapes_2,
    calculate_stride,
    ffi,
    get_broadcast_strides_2,
    prod_,
)
from oracle import build

CPU = Device.CPU
_lib = None
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_rng = np.random.default_rng(0)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build.build_oracle())
    return _lib


# ---- the reference's OWN generated kernels (oracle/_ref/*.so, built by oracle/build.py from
# /root/reference/dlgrad/codegen/cpu_kernel.py) are used in place of the restatement when they are present and
# USE_REF is on; per op the call below mirrors the argument packing of dlgrad/runtime/cpu.py.  `kernel_origin`
# records which implementation actually served each op (bench.py reports cpu_baseline.kind from it).
USE_REF = os.environ.get("DLGRAD_ORACLE_REF", "1") != "0"
_ref_libs: dict = {}
kernel_origin = {"reference": 0, "port": 0}


def _ref(stem: str):
    if not USE_REF:
        return None
    lib_ = _ref_libs.get(stem, False)
    if lib_ is False:
        path = os.path.join(build.REF_DIR, f"{stem}.so")
        lib_ = C.CDLL(path) if os.path.exists(path) else None
        _ref_libs[stem] = lib_
    return lib_


def _ints(vals) -> "C.Array":
    vals = [int(v) for v in vals]
    return (C.c_int * max(1, len(vals)))(*vals)


def _count(which: str) -> None:
    kernel_origin[which] += 1


def _np(b: Buffer) -> np.ndarray:
    """numpy view of a host Buffer's storage (flat)."""
    if b.scalar is not None:
        return np.array([b.scalar], dtype=np.float32)
    if b.metadata.device is not CPU:
        raise RuntimeError("oracle backend received a device buffer")
    return np.frombuffer(ffi.buffer(b.ptr, b.metadata.numel * 4), dtype=np.float32)


def _ret(arr: np.ndarray):
    """cdata `float *` that keeps `arr` alive — the oracle's version of ffi.gc(malloc, free)."""
    return ffi.from_buffer("float *", arr, require_writable=True)


def _f(a: np.ndarray):
    return a.ctypes.data_as(_fp)


def _i4(vals) -> "C.Array":
    vals = list(vals)
    vals = [1] * (4 - len(vals)) + vals if len(vals) <= 4 else vals
    return (C.c_int * len(vals))(*vals)


def _s4(strides) -> "C.Array":
    strides = list(strides)
    return (C.c_int * 4)(*([0] * (4 - len(strides)) + strides))


# ---- binary: CPU._binary_op, dlgrad/runtime/cpu.py:257-291 ----------------------------------
def _binary(name: str, x: Buffer, y: Buffer, out_shape=None):
    xs, ys = x.metadata.shape, y.metadata.shape
    oshape = tuple(out_shape) if out_shape is not None else broadcast_shapes_2(xs, ys)
    out = np.empty(prod_(oshape), dtype=np.float32)
    sx = get_broadcast_strides_2(oshape, xs, calculate_stride(xs))
    sy = get_broadcast_strides_2(oshape, ys, calculate_stride(ys))
    xa, ya = _np(x), _np(y)
    nd = len(oshape)
    r = _ref(f"{name}_{nd}d") if (name != "lte" and 0 <= nd <= 4) else (_ref("cmp") if (name == "lte" and nd == 2) else None)
    if r is not None:
        _count("reference")
        if name == "lte":      # cmp_2d, cpu.py:835-862
            r.cmp(_f(xa), _f(ya), _f(out), _ints(oshape), _ints(sx), _ints(sy), _ints(calculate_stride(oshape)))
        else:                  # arithmetic(op, ndim), cpu.py:257-291
            getattr(r, f"{name}_{nd}d")(_f(xa), _f(ya), _f(out), _ints(oshape), _ints(sx), _ints(sy))
        return _ret(out)
    _count("port")
    getattr(lib(), f"orc_{name}")(_f(xa), _f(ya), _f(out), _i4(oshape), _s4(sx), _s4(sy))
    return _ret(out)


def _reg_binary():
    for op, name in ((BinaryOps.ADD, "add"), (BinaryOps.SUB, "sub"), (BinaryOps.MUL, "mul"),
                     (BinaryOps.DIV, "div"), (BinaryOps.EQT, "eq")):
        dispatcher.register(op, CPU)(lambda x, y, _n=name: _binary(_n, x, y))
    # gt/gte take a python scalar or Buffer (cpu.py:784-791)
    dispatcher.register(BinaryOps.GT, CPU)(lambda x, y: _binary("gt", x, Buffer._as_buffer(y)))
    dispatcher.register(BinaryOps.GTE, CPU)(lambda x, y: _binary("gte", x, Buffer._as_buffer(y)))
    # cmp: cpu.py:835-862
    dispatcher.register(BinaryOps.CMP, CPU)(
        lambda x, y, out_shape, mode="<=": _binary("lte", x, y, out_shape))


# ---- unary: cpu.py:314-502 ---------------------------------------------------------------------
def _unary(name: str, x: Buffer, *extra):
    xa = _np(x)
    out = np.empty(xa.size, dtype=np.float32)
    if name == "clamp":
        has_min, lo, has_max, hi = (e.value for e in extra)
        r = _ref(f"clamp_{int(bool(has_min))}{int(bool(has_max))}")
        if r is not None:      # clamp(has_min, has_max), cpu.py:431-462
            _count("reference")
            fn = getattr(r, f"clamp_{'min' if has_min else ''}_{'max' if has_max else ''}")
            args = [C.c_float(lo)] * bool(has_min) + [C.c_float(hi)] * bool(has_max)
            fn(_f(xa), _f(out), C.c_int(xa.size), *args)
            return _ret(out)
    else:
        r = _ref("relu") if name == "relu" else _ref(f"c_{name}")
        if r is not None:      # utils(func) / relu, cpu.py:314-502,764
            _count("reference")
            getattr(r, "relu" if name == "relu" else f"c_{name}")(_f(xa), _f(out), C.c_int(xa.size), *extra)
            return _ret(out)
    _count("port")
    getattr(lib(), f"orc_{name}")(_f(xa), _f(out), C.c_int(xa.size), *extra)
    return _ret(out)


def _reg_unary():
    for op, name in ((UnaryOps.NEG, "neg"), (UnaryOps.EXP, "exp"), (UnaryOps.LOG, "log"),
                     (UnaryOps.SQRT, "sqrt"), (UnaryOps.RSQRT, "rsqrt"), (UnaryOps.RELU, "relu")):
        dispatcher.register(op, CPU)(lambda x, _n=name: _unary(_n, x))
    dispatcher.register(UnaryOps.POW, CPU)(lambda x, val: _unary("pow", x, C.c_int(int(val))))   # cpu.py:500
    dispatcher.register(UnaryOps.CLAMP, CPU)(
        lambda x, min_val, max_val: _unary("clamp", x, C.c_int(min_val is not None),
                                           C.c_float(float(min_val or 0.0)), C.c_int(max_val is not None),
                                           C.c_float(float(max_val or 0.0))))


# ---- select: cpu.py:794-831, 865-903 -----------------------------------------------------------
def _where(cond: Buffer, x: Buffer, y: Buffer):
    oshape = cond.metadata.shape
    out = np.empty(prod_(oshape), dtype=np.float32)
    st = [get_broadcast_strides_2(oshape, b.metadata.shape, calculate_stride(b.metadata.shape)) for b in (cond, x, y)]
    ca, xa, ya = _np(cond), _np(x), _np(y)
    nd = len(oshape)
    r = _ref(f"where_{nd}d") if 1 <= nd <= 4 else None
    if r is not None:          # where(ndim), cpu.py:865-903
        _count("reference")
        getattr(r, f"where_{nd}d")(_f(ca), _f(xa), _f(ya), _f(out), _ints(oshape), _ints(st[0]), _ints(st[1]),
                                  _ints(st[2]), _ints(calculate_stride(oshape)))
        return _ret(out)
    _count("port")
    lib().orc_where(_f(ca), _f(xa), _f(ya), _f(out), _i4(oshape), _s4(st[0]), _s4(st[1]), _s4(st[2]))
    return _ret(out)


def _masked_fill(x: Buffer, mask: Buffer, val: float, out_shape: tuple):
    oshape = tuple(out_shape)
    out = np.empty(prod_(oshape), dtype=np.float32)
    sx = get_broadcast_strides_2(oshape, x.metadata.shape, calculate_stride(x.metadata.shape))
    sm = get_broadcast_strides_2(oshape, mask.metadata.shape, calculate_stride(mask.metadata.shape))
    xa, ma = _np(x), _np(mask)
    nd = len(oshape)
    r = _ref(f"masked_fill_{nd}d") if 2 <= nd <= 4 else None
    if r is not None:          # masked_fill(ndim), cpu.py:794-831
        _count("reference")
        getattr(r, f"masked_fill_{nd}d")(_f(xa), _f(ma), _f(out), _ints(oshape), _ints(sx), _ints(sm),
                                        _ints(calculate_stride(oshape)), C.c_float(val))
        return _ret(out)
    _count("port")
    lib().orc_masked_fill(_f(xa), _f(ma), _f(out), _i4(oshape), _s4(sx), _s4(sm), C.c_float(val))
    return _ret(out)


# ---- reductions: cpu.py:605-761, 911-959 --------------------------------------------------------
def _reduce(name: str, x: Buffer, dim):
    xa = _np(x)
    shape = x.metadata.shape
    if dim is None or dim == -1:
        out = np.empty(1, dtype=np.float32)
        r = _ref(f"reduce_{name}")
        if r is not None:      # reduce(op), cpu.py:612-630
            _count("reference")
            getattr(r, f"reduce_{name}")(_f(xa), _f(out), C.c_int(xa.size))
            return _ret(out)
        _count("port")
        getattr(lib(), f"orc_reduce_all_{name}")(_f(xa), _f(out), C.c_int(xa.size))
        return _ret(out)
    if dim < 0:
        dim += len(shape)
    outer, R, inner = prod_(shape[:dim]), shape[dim], prod_(shape[dim + 1:])
    out = np.empty(outer * inner, dtype=np.float32)
    nd = len(shape)
    r = _ref(f"reduce_{name}_{nd}d_axis{dim}") if 1 <= nd <= 4 else None
    if r is not None:          # reduce_source(op, ndim, dim), cpu.py:632-658
        _count("reference")
        oshape = [s_ for i, s_ in enumerate(shape) if i != dim]
        getattr(r, f"reduce_{name}_{nd}d_axis{dim}")(_f(xa), _f(out), _ints(shape), _ints(calculate_stride(shape)),
                                                   _ints(calculate_stride(oshape)))
        return _ret(out)
    _count("port")
    getattr(lib(), f"orc_reduce_{name}")(_f(xa), _f(out), C.c_int(outer), C.c_int(R), C.c_int(inner))
    return _ret(out)


def _argmax(x: Buffer, dim):
    xa = _np(x)
    shape = x.metadata.shape
    if dim is None:
        outer, R, inner = 1, xa.size, 1
    else:
        if dim < 0:              # cpu.py:921-922: -1 becomes the LAST axis (Buffer labels it (1,1))
            dim += len(shape)
        outer, R, inner = prod_(shape[:dim]), shape[dim], prod_(shape[dim + 1:])
    out = np.empty(outer * inner, dtype=np.float32)
    r = _ref(f"argmax_2d_axis{dim}") if (dim is not None and len(shape) == 2) else None
    if r is not None:          # argmax(ndim, dim), cpu.py:911-959
        _count("reference")
        oshape = [s_ for i, s_ in enumerate(shape) if i != dim]
        getattr(r, f"argmax_2d_axis{dim}")(_f(xa), _f(out), _ints(shape), _ints(calculate_stride(shape)),
                                         _ints(calculate_stride(oshape)))
        return _ret(out)
    _count("port")
    lib().orc_argmax(_f(xa), _f(out), C.c_int(outer), C.c_int(R), C.c_int(inner))
    return _ret(out)


# ---- permute: cpu.py:334-368 ---------------------------------------------------------------------
def _permute(x: Buffer, order: tuple):
    shape = x.metadata.shape
    in_str = calculate_stride(shape)
    oshape = tuple(shape[i] for i in order)
    pstr = tuple(in_str[i] for i in order)
    xa = _np(x)
    out = np.empty(xa.size, dtype=np.float32)
    nd = len(shape)
    r = _ref(f"permute_{nd}d") if 2 <= nd <= 4 else None
    if r is not None:          # permute(ndim), cpu.py:334-368
        _count("reference")
        getattr(r, f"permute_{nd}d")(_f(xa), _f(out), _ints(oshape), _ints(pstr), _ints(calculate_stride(oshape)))
        return _ret(out)
    _count("port")
    lib().orc_permute(_f(xa), _f(out), _i4(oshape), _s4(pstr))
    return _ret(out)


# ---- matmul: cpu.py:505-602 ----------------------------------------------------------------------
def _matmul(x: Buffer, y: Buffer, trans_x: bool = False, trans_y: bool = False):
    assert not (trans_x or trans_y), "the reference has no transposed matmul; ops.MatMul permutes first on CPU"
    xs, ys = x.metadata.shape, y.metadata.shape
    bx = (1,) * (4 - len(xs)) + tuple(xs)
    by = (1,) * (4 - len(ys)) + tuple(ys)
    B0, B1 = max(bx[0], by[0]), max(bx[1], by[1])
    M, K, N = bx[2], bx[3], by[3]
    sx = calculate_stride(bx)
    sy = calculate_stride(by)
    xb0 = 0 if (bx[0] == 1 and B0 > 1) else sx[0]
    xb1 = 0 if (bx[1] == 1 and B1 > 1) else sx[1]
    yb0 = 0 if (by[0] == 1 and B0 > 1) else sy[0]
    yb1 = 0 if (by[1] == 1 and B1 > 1) else sy[1]
    xa, ya = _np(x), _np(y)
    out = np.zeros(B0 * B1 * M * N, dtype=np.float32)
    nd = len(xs)
    r = _ref(f"matmul_{nd}d") if nd in (2, 3, 4) else None
    if r is not None:          # matmul_{2,3,4}d with the batch strides of a broadcast operand zeroed, cpu.py:505-602
        _count("reference")
        if nd == 2:
            r.matmul_2d(_f(xa), _f(ya), _f(out), _ints(xs), _ints(calculate_stride(xs)), _ints(ys), _ints(calculate_stride(ys)))
        elif nd == 3:
            r.matmul_3d(_f(xa), _f(ya), _f(out), C.c_int(B1), C.c_int(M), C.c_int(K), C.c_int(N),
                        _ints((xb1, K, 1)), _ints((yb1, N, 1)), _ints((M * N, N, 1)))
        else:
            r.matmul_4d(_f(xa), _f(ya), _f(out), _ints((B0, B1, M, K, N)), _ints((xb0, xb1, K, 1)),
                        _ints((yb0, yb1, N, 1)), _ints((B1 * M * N, M * N, N, 1)))
        return _ret(out)
    _count("port")
    lib().orc_matmul(_f(xa), _f(ya), _f(out), *(C.c_int(v) for v in (B0, B1, M, K, N, xb0, xb1, yb0, yb1)))
    return _ret(out)


# ---- cross entropy / embedding: cpu.py:962-1061 ----------------------------------------------------
def _ce_forward(x: Buffer, y: Buffer):
    xa, ta = _np(x), _np(y)
    n = x.metadata.shape[0]
    out = np.empty(n, dtype=np.float32)
    r = _ref("ce_forward")
    if r is not None:
        _count("reference")
        r.ce_forward(_f(xa), _f(ta), _f(out), C.c_int(n), _ints(x.metadata.stride))
        return _ret(out)
    _count("port")
    lib().orc_ce_forward(_f(xa), _f(ta), _f(out), C.c_int(n), C.c_int(x.metadata.stride[0]))
    return _ret(out)


def _ce_backward(x: Buffer, target: Buffer):
    xa, ta = _np(x), _np(target)
    r = _ref("ce_backward")
    if r is not None:
        _count("reference")
        r.ce_backward(_f(xa), _f(ta), _ints(x.metadata.shape), _ints(x.metadata.stride))
        return
    _count("port")
    lib().orc_ce_backward(_f(xa), _f(ta), C.c_int(x.metadata.shape[0]), C.c_int(x.metadata.stride[0]))


def _embedding(x: Buffer, idx: Buffer, out_numel: int, backward: bool = False, upstream_grad: Buffer = None):
    width = x.metadata.shape[-1]
    ia = _np(idx)
    if backward:
        out = np.zeros(out_numel, dtype=np.float32)
        ga = _np(upstream_grad)
        r = _ref("embedding_backward")
        if r is not None:
            _count("reference")
            r.embedding_backward(_f(out), _f(ga), _f(ia), C.c_int(ia.size), C.c_int(width))
            return _ret(out)
        _count("port")
        lib().orc_embedding_backward(_f(out), _f(ga), _f(ia), C.c_int(ia.size), C.c_int(width))
        return _ret(out)
    xa = _np(x)
    out = np.empty(out_numel, dtype=np.float32)
    r = _ref("embedding")
    if r is not None:
        _count("reference")
        r.embedding(_f(xa), _f(ia), _f(out), C.c_int(ia.size), C.c_int(width))
        return _ret(out)
    _count("port")
    lib().orc_embedding(_f(xa), _f(ia), _f(out), C.c_int(ia.size), C.c_int(width))
    return _ret(out)


# ---- creation: cpu.py:194-254 ----------------------------------------------------------------------
def _full(shape: tuple, fill_value: float):
    out = np.empty(prod_(shape), dtype=np.float32)
    r = _ref("full")
    if r is not None:
        _count("reference")
        r.full(_f(out), C.c_int(out.size), C.c_int(int(fill_value)))            # int(): cpu.py:252
        return _ret(out)
    _count("port")
    lib().orc_full(_f(out), C.c_int(out.size), C.c_int(int(fill_value)))
    return _ret(out)


def _arange(shape: tuple):
    out = np.empty(prod_(shape), dtype=np.float32)
    r = _ref("arange")
    if r is not None:
        _count("reference")
        r.arange(_f(out), C.c_int(out.size))
        return _ret(out)
    _count("port")
    lib().orc_arange(_f(out), C.c_int(out.size))
    return _ret(out)


def _uniform(shape: tuple, low: float = 0.0, high: float = 1.0):
    # the reference's PCG32 reseeds from /dev/random per call (cpu_kernel.py:774-874): values are
    # never reproducible, so parity runs inject weights; any U[low, high) stream serves here
    out = (low + (high - low) * _rng.random(prod_(shape), dtype=np.float32)).astype(np.float32)
    return _ret(out)


# ---- fused CUDA ops as the reference's compositions ------------------------------------------------
def _wrap(ptr, shape) -> Buffer:
    return Buffer(ptr, tuple(shape), CPU)


def _softmax(x: Buffer, dim: int, log: bool = False):
    """dlgrad/tensor.py:634-669: max, sub, exp, sum, div | log, sub — five primitive kernels."""
    mx = x.max(dim=dim, keepdim=True)
    sh = x - mx
    e = sh.exp()
    s = e.sum(dim=dim, keepdim=True)
    res = (sh - s.log()) if log else (e / s)
    return res.ptr


def _ce_grad(logsm: Buffer, target: Buffer, upstream: Buffer, n: float):
    """ops.CrossEntropy.backward, dlgrad/ops.py:359-370: exp, in-place -1 at target, * upstream, / N."""
    tmp = logsm.exp()
    _ce_backward(tmp, target)
    return ((tmp * upstream) / float(n)).ptr


def _adam_step(param: Buffer, grad: Buffer, m: Buffer, v: Buffer, beta1, beta2, eps, bc1, bc2, lr, grad_scale=1.0):
    """dlgrad/nn/optim.py:59-70, op for op on the primitive kernels; results written back in place."""
    if grad_scale != 1.0:
        grad = grad * grad_scale
    m_new = (m * beta1) + (grad * (1.0 - beta1))
    v_new = (v * beta2) + ((grad ** 2.0) * (1.0 - beta2))
    m_hat = m_new / bc1
    v_hat = v_new / bc2
    p_new = param - (m_hat / (v_hat.sqrt() + eps)) * lr
    for dst, src in ((param, p_new), (m, m_new), (v, v_new)):
        _np(dst)[:] = _np(src)


def _sgd_step(param: Buffer, grad: Buffer, lr: float, grad_scale: float = 1.0):
    """dlgrad/nn/optim.py:44: param - lr * grad."""
    if grad_scale != 1.0:
        grad = grad * grad_scale
    p_new = param - (Buffer.from_scalar(lr) * grad)
    _np(param)[:] = _np(p_new)


def install() -> None:
    """Register the oracle as the Device.CPU runtime (tests / smoke / cpu_baseline only)."""
    _reg_binary()
    _reg_unary()
    reg = dispatcher.register
    reg(UnaryOps.WHERE, CPU)(_where)
    reg(UnaryOps.MASKED_FILL, CPU)(_masked_fill)
    reg(UnaryOps.SUM, CPU)(lambda x, dim: _reduce("sum", x, dim))
    reg(UnaryOps.MEAN, CPU)(lambda x, dim: _reduce("mean", x, dim))
    reg(UnaryOps.MAX, CPU)(lambda x, dim: _reduce("max", x, dim))
    reg(UnaryOps.ARGMAX, CPU)(_argmax)
    reg(UnaryOps.PERMUTE, CPU)(_permute)
    reg(BinaryOps.MATMUL, CPU)(_matmul)
    reg(CustomOps.CE_FORWARD, CPU)(_ce_forward)
    reg(CustomOps.CE_BACKWARD, CPU)(_ce_backward)
    reg(CustomOps.EMBEDDING, CPU)(_embedding)
    reg(BufferOps.FULL, CPU)(_full)
    reg(BufferOps.ARANGE, CPU)(_arange)
    reg(BufferOps.UNIFORM, CPU)(_uniform)
    reg(UnaryOps.SOFTMAX, CPU)(lambda x, dim: _softmax(x, dim))
    reg(UnaryOps.LOG_SOFTMAX, CPU)(lambda x, dim: _softmax(x, dim, log=True))
    reg(CustomOps.CE_GRAD, CPU)(_ce_grad)
    reg(CustomOps.ADAM_STEP, CPU)(_adam_step)
    reg(CustomOps.SGD_STEP, CPU)(_sgd_step)
    reg(CustomOps.PRINT, CPU)(lambda x: print(_np(x).reshape(x.metadata.shape)))


def uninstall() -> None:
    for key in [k for k in dispatcher._dispatch_table if k[1] is CPU]:
        del dispatcher._dispatch_table[key]


def seed(s: int) -> None:
    global _rng
    _rng = np.random.default_rng(s)
