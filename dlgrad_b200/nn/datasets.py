"""MNIST idx loader (dlgrad/nn/datasets.py:9-39, C loader dlgrad/codegen/cpu_kernel.py:877-1024).

The reference downloads the four idx files into its cache directory, parses them with a generated C
function (big-endian header, magic check, every byte widened to float, NOT normalised) and wraps the
result in Tensors on the requested device.  Here the parse is host-side numpy and the result is uploaded
once (H2D) when `device` is CUDA.  Files are looked up in `root`, `$DLGRAD_DATA`, then the reference's
cache location `~/.cache/dlgrad/downloads`; only if absent is a download attempted.
"""
from __future__ import annotations

import gzip
import os
import struct

import numpy as np

from dlgrad_b200.device import Device
from dlgrad_b200.tensor import Tensor

BASE_URL = "https://storage.googleapis.com/cvdf-datasets/mnist/"
FILES = [
    ("train-images-idx3-ubyte", True, 2051, (60000, 784)),
    ("train-labels-idx1-ubyte", False, 2049, (60000, 1)),
    ("t10k-images-idx3-ubyte", True, 2051, (10000, 784)),
    ("t10k-labels-idx1-ubyte", False, 2049, (10000, 1)),
]


def read_idx(path: str, images: bool, magic_number: int) -> np.ndarray:
    """Parse one idx file into float32 (N, 784) pixels or (N, 1) labels.  A wrong magic number or a
    truncated file raises MemoryError("Error when loading MNIST data"), the reference's error
    (dlgrad/runtime/cpu.py:162-163)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    try:
        magic, n = struct.unpack(">II", raw[:8])
        if magic != magic_number:
            raise ValueError("magic")
        if images:
            rows, cols = struct.unpack(">II", raw[8:16])
            body = np.frombuffer(raw, dtype=np.uint8, count=n * rows * cols, offset=16)
            return body.astype(np.float32).reshape(n, rows * cols)
        body = np.frombuffer(raw, dtype=np.uint8, count=n, offset=8)
        return body.astype(np.float32).reshape(n, 1)
    except (ValueError, struct.error):
        raise MemoryError("Error when loading MNIST data") from None


def _locate(name: str, root: str | None) -> str:
    dirs = [d for d in (root, os.environ.get("DLGRAD_DATA"),
                        os.path.join(os.path.expanduser("~"), ".cache", "dlgrad", "downloads")) if d]
    for d in dirs:
        for cand in (name, name + ".gz"):
            p = os.path.join(d, cand)
            if os.path.exists(p):
                return p
    dest = os.path.join(dirs[0], name + ".gz")
    os.makedirs(dirs[0], exist_ok=True)
    import urllib.request
    try:
        urllib.request.urlretrieve(BASE_URL + name + ".gz", dest)
    except OSError as exc:
        raise FileNotFoundError(f"{name} not found in {dirs} and the download failed: {exc}") from None
    return dest


def mnist(device: Device | str = Device.CPU, root: str | None = None) -> list[Tensor]:
    """[train_images (60000,784), train_labels (60000,1), test_images (10000,784), test_labels (10000,1)],
    unnormalised float32, requires_grad=False, on `device`."""
    out = []
    for name, is_image, magic, shape in FILES:
        arr = read_idx(_locate(name, root), is_image, magic)
        if arr.shape != shape and root is None:
            raise MemoryError("Error when loading MNIST data")
        out.append(Tensor(arr, requires_grad=False, device=device))
    return out
