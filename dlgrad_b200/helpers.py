"""Op enums and shape arithmetic shared by Buffer, the runtimes and the code generator.

The enum member names are the dispatch-table keys of the reference (dlgrad/helpers.py:18-58) and are
part of the drop-in contract; members below the "fused" markers are additions of the CUDA backend.
The shape helpers restate the reference's rules (file:line cited per function) — they are pure
tuple arithmetic with no device dependence.
"""
from __future__ import annotations

import os
from collections.abc import Iterable
from enum import Enum, auto
from math import prod

from cffi import FFI

ffi = FFI()


class UnaryOps(Enum):
    __hash__ = object.__hash__        # identity hash in C: these are dict keys on every dispatch (Enum's is a Python call)
    SUM = auto()
    MAX = auto()
    NEG = auto()
    EXP = auto()
    LOG = auto()
    POW = auto()
    SQRT = auto()
    RSQRT = auto()
    TRANSPOSE = auto()
    RELU = auto()
    ARGMAX = auto()
    WHERE = auto()
    MEAN = auto()
    CLAMP = auto()
    MASKED_FILL = auto()
    PERMUTE = auto()
    # --- fused (CUDA backend) ---
    SOFTMAX = auto()
    LOG_SOFTMAX = auto()
    SOFTMAX_BACKWARD = auto()
    LOG_SOFTMAX_BACKWARD = auto()
    EXPAND = auto()          # broadcast a reduced grad back to the input shape (Sum/Mean/Max backward)


class BufferOps(Enum):
    __hash__ = object.__hash__        # identity hash in C: these are dict keys on every dispatch (Enum's is a Python call)
    CREATE = auto()
    UNIFORM = auto()
    ARANGE = auto()
    FULL = auto()


class BinaryOps(Enum):
    __hash__ = object.__hash__        # identity hash in C: these are dict keys on every dispatch (Enum's is a Python call)
    ADD = auto()
    SUB = auto()
    MUL = auto()
    DIV = auto()
    MATMUL = auto()
    GT = auto()
    GTE = auto()
    EQT = auto()
    CMP = auto()


class CustomOps(Enum):
    __hash__ = object.__hash__        # identity hash in C: these are dict keys on every dispatch (Enum's is a Python call)
    INDEX = auto()
    CE_FORWARD = auto()
    CE_BACKWARD = auto()
    PRINT = auto()
    EMBEDDING = auto()
    # --- fused (CUDA backend) ---
    CE_LOSS = auto()          # log-softmax + gather + (-sum) in one pass, keeps log-softmax for backward
    CE_GRAD = auto()          # (exp(logsm) - onehot) * upstream / N
    ADAM_STEP = auto()
    ADAM_MULTI = auto()       # every parameter tensor in one launch
    SGD_STEP = auto()
    RMSNORM = auto()
    RMSNORM_BACKWARD = auto()
    CAT = auto()              # concatenate two tensors along an axis (device-side KV-cache append)


def prod_(x: Iterable) -> int:
    return prod(x) if x else 1


def check_broadcast(x_shape: tuple, y_shape: tuple) -> bool:
    """Trailing-aligned compatibility test; raises AssertionError like dlgrad/helpers.py:64-83."""
    for a, b in zip(reversed(x_shape), reversed(y_shape)):
        if a != b and a != 1 and b != 1:
            raise AssertionError(
                f"Cannot broadcast {y_shape} to {x_shape}, the dimensions {a} and {b} dont match")
    return True


def calculate_stride(shape: tuple | list) -> tuple:
    """Row-major contiguous strides in elements (dlgrad/helpers.py:176-186)."""
    acc, out = 1, []
    for d in reversed(tuple(shape)):
        out.append(acc)
        acc *= d
    return tuple(reversed(out))


def broadcast_shapes(s1: tuple, s2: tuple) -> tuple[tuple, tuple, tuple]:
    """Matmul rank alignment: left-pad with 1s, broadcast the leading (batch) dims, check K
    (dlgrad/helpers.py:114-135).  Returns (padded1, padded2, out_shape)."""
    n = max(len(s1), len(s2))
    p1 = (1,) * (n - len(s1)) + tuple(s1)
    p2 = (1,) * (n - len(s2)) + tuple(s2)
    batch = []
    for a, b in zip(p1[:-2], p2[:-2]):
        if a != b and 1 not in (a, b):
            raise ValueError(f"Cannot broadcast batch dims: {s1} and {s2}")
        batch.append(max(a, b))
    if p1[-1] != p2[-2]:
        raise ValueError(f"Matrix inner dims must match: {s1} and {s2}")
    return p1, p2, tuple(batch) + (p1[-2], p2[-1])


def get_broadcast_shape(shape1: tuple, shape2: tuple) -> tuple:
    """Full numpy-style broadcast result (dlgrad/helpers.py:137-155)."""
    n = max(len(shape1), len(shape2))
    a = (1,) * (n - len(shape1)) + tuple(shape1)
    b = (1,) * (n - len(shape2)) + tuple(shape2)
    out = []
    for x, y in zip(a, b):
        if x != y and 1 not in (x, y):
            raise ValueError(f"Operands could not be broadcast together with shapes {shape1} and {shape2}")
        out.append(max(x, y))
    return tuple(out)


# the CPU runtime's name for the same rule (dlgrad/helpers.py:310-334)
def broadcast_shapes_2(s1: tuple, s2: tuple) -> tuple:
    try:
        return get_broadcast_shape(s1, s2)
    except ValueError:
        raise ValueError(f"Operands could not be broadcast together with shapes {s1} {s2}") from None


def get_broadcast_strides_2(out_shape: tuple, in_shape: tuple, in_stride: tuple) -> tuple:
    """Strides of `in` viewed at `out_shape`: 0 where a size-1 dim is stretched
    (dlgrad/helpers.py:276-308)."""
    pad = len(out_shape) - len(in_shape)
    shp = (1,) * pad + tuple(in_shape)
    strd = (0,) * pad + tuple(in_stride)
    res = []
    for o, i, s in zip(out_shape, shp, strd):
        if i == o:
            res.append(s)
        elif i == 1:
            res.append(0)
        else:
            raise ValueError(f"Shape mismatch: cannot broadcast {in_shape} to {out_shape}")
    return tuple(res)


def cal_sum_max_out_shape(ndim: int, dim: int, inp_shape: tuple, keepdim: bool = False) -> tuple:
    """Output shape of sum/mean/max.  dim == -1 means "reduce everything"; with keepdim the
    reference returns the *input* shape there (dlgrad/helpers.py:204-213)."""
    if dim == -1:
        return tuple(inp_shape) if keepdim else ()
    t = list(inp_shape)
    if keepdim:
        t[dim] = 1
    else:
        t.pop(dim)
    return tuple(t)


# Kernel cache (cubins).  DLGRAD_CUDA_CACHE overrides; the default lives inside the package so the
# pre-built cubins travel with the repository snapshot to the GPU box.
CACHE_DIR = os.environ.get(
    "DLGRAD_CUDA_CACHE", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_kcache"))
