"""pytest configuration.

  -m "not gpu"  runs on the build machine: oracle vs golden vectors, host logic, C-ABI surface.
  -m gpu        runs on a B200: Device.CUDA vs golden vectors and vs the oracle (through the C ABI).

With DLGRAD_CUDA_COMPILE_ONLY=1 the gpu tests are *walked* without a device so that every
shape-specialised kernel they need is compiled into the in-tree cache (see __graft_entry__.build);
comparisons are skipped in that mode and the pending kernels are flushed at session end.
"""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

COMPILE_ONLY = os.environ.get("DLGRAD_CUDA_COMPILE_ONLY", "0") == "1"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if COMPILE_ONLY:
        return
    try:
        from dlgrad_b200.runtime.cuda import shim
        have = shim.device_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("DLGRAD_PARITY_REPORT")
    if path and not COMPILE_ONLY:
        import json
        from tests import runner
        with open(path, "w") as f:
            for row in runner.REPORT:
                f.write(json.dumps(row) + "\n")
    if COMPILE_ONLY:
        from dlgrad_b200.runtime.cuda import compiler
        n = compiler.flush_pending()
        print(f"\n[compile-only] built {n} kernels into the cache")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle registered as the Device.CPU runtime for the whole session."""
    from oracle import backend
    backend.install()
    return backend


@pytest.fixture(scope="session")
def golden_ops():
    return np.load(os.path.join(REPO, "tests", "golden", "ops.npz"))


@pytest.fixture(scope="session")
def golden_models():
    return np.load(os.path.join(REPO, "tests", "golden", "models.npz"))


def check(actual, expected, rtol=0.0, atol=0.0, exact=False, what=""):
    """Comparison helper; a no-op while walking the tests in compile-only mode."""
    if COMPILE_ONLY:
        return
    actual = np.asarray(actual, dtype=np.float32)
    expected = np.asarray(expected, dtype=np.float32)
    assert actual.size == expected.size, f"{what}: size {actual.shape} vs {expected.shape}"
    actual = actual.reshape(expected.shape)
    if exact:
        same = (actual == expected) | (np.isnan(actual) & np.isnan(expected))
        assert same.all(), f"{what}: {(~same).sum()} of {same.size} elements differ (bit-exact required)"
    else:
        np.testing.assert_allclose(actual, expected, rtol=rtol, atol=atol, err_msg=what)
