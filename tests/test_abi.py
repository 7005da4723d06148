"""The C-ABI library loads without a GPU and exports every symbol include/dlgrad_cuda.h declares."""
import os
import re
import subprocess

from tests.conftest import REPO

LIB = os.path.join(REPO, "dlgrad_b200", "libdlgrad_cuda.so")
HEADER = os.path.join(REPO, "include", "dlgrad_cuda.h")


def _declared():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(dlg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exists_and_loads():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    import cffi
    ffi = cffi.FFI()
    ffi.cdef("const char *dlg_last_error(void); int dlg_last_status(void); unsigned long long dlg_launch_count(void);")
    lib = ffi.dlopen(LIB)
    assert lib.dlg_launch_count() == 0
    assert isinstance(ffi.string(lib.dlg_last_error()), bytes)


def test_every_declared_symbol_is_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dlg_[a-z0-9_]+)", out))
    declared = _declared()
    assert len(declared) >= 40
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_header_parses_as_cdef_and_binds():
    from dlgrad_b200.runtime.cuda import shim
    if shim.COMPILE_ONLY:
        return
    for name in _declared():
        assert hasattr(shim.lib, name), name


def test_no_device_fails_loudly():
    """Without a driver the product must raise, not compute somewhere else."""
    from dlgrad_b200.runtime.cuda import shim
    if shim.COMPILE_ONLY or shim.device_available():
        return
    import numpy as np
    import pytest
    from dlgrad_b200 import Tensor
    from dlgrad_b200.dispatch import dispatcher
    from dlgrad_b200.device import Device
    from dlgrad_b200.helpers import BinaryOps
    if dispatcher.has(BinaryOps.ADD, Device.CPU):
        pytest.skip("oracle already installed in this session")
    a = Tensor(np.ones((2, 2), dtype=np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ = a + a


def test_product_never_imports_oracle():
    """No file under dlgrad_b200/ may import or name the oracle package."""
    bad = []
    for root, _dirs, files in os.walk(os.path.join(REPO, "dlgrad_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(root, f))
    assert not bad, bad
