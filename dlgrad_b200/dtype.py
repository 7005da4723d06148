"""dtypes.  dlgrad computes in float32 only (dlgrad/dtype.py:11-27); indices, labels, masks and
argmax results are float32 holding integers.  The CUDA backend keeps that representation."""
from __future__ import annotations

import enum
from typing import Any, NewType

# a cffi `float *` cdata: host address on CPU, CUdeviceptr on CUDA
CDataPtr = NewType("CDataPtr", Any)
Scalar = float


class DType(enum.Enum):
    """Element types; the value of a member is its size in bytes."""
    __hash__ = object.__hash__        # identity hash in C: a member keys dict lookups on every dispatch

    FLOAT32 = 4

    @property
    def itemsize(self) -> int:
        return self.value

    @staticmethod
    def from_str(d: str) -> "DType":
        member = DType.__members__.get(str(d).strip().upper())
        if member is None:
            raise ValueError(f"Invalid dtype: {d}")
        return member

    @staticmethod
    def get_n_bytes(d: "DType") -> int:
        return d.itemsize
