"""MNIST idx loader: parse on the host, one H2D per file."""
import gzip
import os
import struct

import numpy as np
import pytest

from tests.conftest import check

from dlgrad_b200.nn import datasets


def _write_idx(tmp_path, n=12, gz=False):
    rng = np.random.default_rng(3)
    px = rng.integers(0, 256, size=(n, 28, 28), dtype=np.uint8)
    lb = rng.integers(0, 10, size=(n,), dtype=np.uint8)
    for prefix in ("train", "t10k"):
        for name, blob in ((f"{prefix}-images-idx3-ubyte", struct.pack(">IIII", 2051, n, 28, 28) + px.tobytes()),
                           (f"{prefix}-labels-idx1-ubyte", struct.pack(">II", 2049, n) + lb.tobytes())):
            p = os.path.join(tmp_path, name + (".gz" if gz else ""))
            with (gzip.open if gz else open)(p, "wb") as f:
                f.write(blob)
    return px, lb


@pytest.mark.parametrize("gz", [False, True])
def test_read_idx(tmp_path, gz):
    px, lb = _write_idx(str(tmp_path), gz=gz)
    ext = ".gz" if gz else ""
    img = datasets.read_idx(os.path.join(tmp_path, "train-images-idx3-ubyte" + ext), True, 2051)
    lab = datasets.read_idx(os.path.join(tmp_path, "train-labels-idx1-ubyte" + ext), False, 2049)
    assert img.dtype == np.float32 and img.shape == (12, 784) and lab.shape == (12, 1)
    np.testing.assert_array_equal(img, px.reshape(12, 784).astype(np.float32))     # widened, not normalised
    np.testing.assert_array_equal(lab[:, 0], lb.astype(np.float32))


def test_bad_magic_and_truncated(tmp_path):
    _write_idx(str(tmp_path))
    with pytest.raises(MemoryError, match="Error when loading MNIST data"):
        datasets.read_idx(os.path.join(tmp_path, "train-images-idx3-ubyte"), True, 2049)
    p = os.path.join(tmp_path, "short")
    with open(p, "wb") as f:
        f.write(struct.pack(">IIII", 2051, 100, 28, 28) + b"\0" * 10)
    with pytest.raises(MemoryError):
        datasets.read_idx(p, True, 2051)


@pytest.mark.gpu
def test_mnist_on_device(tmp_path):
    px, lb = _write_idx(str(tmp_path))
    tensors = datasets.mnist(device="cuda", root=str(tmp_path))
    assert [t.shape for t in tensors] == [(12, 784), (12, 1), (12, 784), (12, 1)]
    assert all(not t.requires_grad for t in tensors)
    check(tensors[0].numpy(), px.reshape(12, 784).astype(np.float32), exact=True, what="train images")
    check(tensors[3].numpy()[:, 0], lb.astype(np.float32), exact=True, what="test labels")
    # usable by the hot path: a normalised batch through a matmul
    x = tensors[0] / 255.0
    ref = px.reshape(12, 784).astype(np.float64) / 255.0
    check((x @ x.T).numpy(), ref @ ref.T, rtol=1e-4, what="normalised batch through a matmul")
