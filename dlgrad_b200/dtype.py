"""dtypes.  dlgrad computes in float32 only (dlgrad/dtype.py:11-27); indices, labels, masks and
argmax results are float32 holding integers.  The CUDA backend keeps that representation."""
from __future__ import annotations

from enum import Enum, auto
from typing import Any, NewType

CDataPtr = NewType("CDataPtr", Any)   # a cffi `float *` cdata: host address on CPU, CUdeviceptr on CUDA
Scalar = float


class DType(Enum):
    __hash__ = object.__hash__        # identity hash in C: these are dict keys on every dispatch (Enum's is a Python call)
    FLOAT32 = auto()

    @staticmethod
    def from_str(d: str) -> "DType":
        try:
            return DType[d.upper()]
        except KeyError:
            raise ValueError(f"Invalid dtype: {d}") from None

    @staticmethod
    def get_n_bytes(d: "DType") -> int:
        return 4
