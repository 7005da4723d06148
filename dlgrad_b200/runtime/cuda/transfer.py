"""Host <-> device copies for Buffer: replaces the reference's direct reads/writes of `ptr` as host
memory (dlgrad/tensor.py:112-115,125-138,147-164; dlgrad/buffer.py:67-90)."""
from __future__ import annotations

import numpy as np

from dlgrad_b200.runtime.cuda import shim

_counters = {"h2d_bytes": 0, "d2h_bytes": 0}


def upload(host_ptr, nbytes: int):
    dptr = shim.alloc(nbytes)
    if shim.capture["on"]:
        # host data created inside a captured step (e.g. the position indices of examples/gpt.py:103): frozen into a
        # pinned staging buffer that the graph owns; every replay copies the same bytes again
        pin = shim.lib.dlg_host_alloc(nbytes)
        if pin == shim.NULL:
            raise MemoryError("pinned staging allocation failed: " + shim.last_error())
        shim.ffi.memmove(pin, host_ptr, nbytes)
        shim.capture["keep"].append(pin)
        shim.check(shim.lib.dlg_h2d_async(dptr, pin, nbytes), "dlg_h2d_async")
    else:
        shim.check(shim.lib.dlg_h2d(dptr, host_ptr, nbytes), "dlg_h2d")
    _counters["h2d_bytes"] += nbytes
    return dptr


def upload_numpy(arr: np.ndarray):
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    return upload(shim.ffi.from_buffer("float *", arr, require_writable=False), arr.nbytes)


def download(dev_ptr, numel: int) -> np.ndarray:
    shim.no_capture("reading device memory on the host (numpy(), show(), tobytes())")
    out = np.empty(numel, dtype=np.float32)
    shim.check(shim.lib.dlg_d2h(shim.ffi.from_buffer("float *", out, require_writable=True), dev_ptr, numel * 4),
               "dlg_d2h")
    _counters["d2h_bytes"] += numel * 4
    return out


def write_bytes(dev_ptr, data: bytes) -> None:
    shim.no_capture("copy_from()")
    shim.check(shim.lib.dlg_h2d(dev_ptr, shim.ffi.from_buffer("char *", data), len(data)), "dlg_h2d")
    _counters["h2d_bytes"] += len(data)


def counters() -> dict:
    return dict(_counters)


def reset_counters() -> None:
    _counters["h2d_bytes"] = 0
    _counters["d2h_bytes"] = 0
