"""cffi (ABI-mode) binding of libdlgrad_cuda.so — the same mechanism the reference uses for its
per-kernel CPU objects (`ffi.dlopen`, dlgrad/runtime/cpu.py:64-73).

Nothing here falls back to a host implementation: if the library is missing, or no driver / GPU is
present, the first device request raises RuntimeError.

`DLGRAD_CUDA_COMPILE_ONLY=1` swaps the library for a stand-in that hands out fake addresses and
ignores launches.  It exists so that `__graft_entry__.build()` can walk the workloads on the
GPU-less build machine and pre-compile every shape-specialised kernel into the in-tree cache
(each GPU box starts with a cold cache).  Results are meaningless in that mode.
"""
from __future__ import annotations

import os
import re

from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.path.join(_PKG, "libdlgrad_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "dlgrad_cuda.h")

COMPILE_ONLY = os.environ.get("DLGRAD_CUDA_COMPILE_ONLY", "0") == "1"

ffi = FFI()


def _cdef_from_header() -> str:
    """The header is the single source of truth for the ABI; strip what cffi cannot parse."""
    with open(HEADER_PATH) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    keep = []
    for line in txt.splitlines():
        s = line.strip()
        if s.startswith("#") or s.startswith('extern "C"') or s == "}":
            continue
        keep.append(line)
    return "\n".join(keep)


ffi.cdef(_cdef_from_header())
NULL = ffi.NULL


class _CompileOnlyLib:
    """Stand-in for the shared library on a machine without a GPU (build-time kernel census)."""

    def __init__(self) -> None:
        self._next = 0x7000_0000_0000
        self.launches = 0

    def _addr(self, nbytes: int):
        p = self._next
        self._next += (int(nbytes) + 511) & ~511
        return ffi.cast("void *", p)

    def dlg_init(self, device): return 0
    def dlg_device_count(self): return 1
    def dlg_device_attr(self, attr): return 148
    def dlg_last_error(self): return ffi.new("char[]", b"compile-only")
    def dlg_device_name(self): return ffi.new("char[]", b"compile-only (no device)")
    def dlg_sync(self): return 0
    def dlg_alloc(self, n): return self._addr(n)
    def dlg_calloc(self, n): return self._addr(n)
    def dlg_free(self, p): return None
    def dlg_empty_cache(self): return 0
    def dlg_mem_live(self): return 0
    def dlg_mem_reserved(self): return 0
    def dlg_mem_peak(self): return 0
    def dlg_mem_reset_peak(self): return None
    def dlg_h2d(self, d, s, n): return 0
    def dlg_d2h(self, d, s, n): return 0
    def dlg_d2d(self, d, s, n): return 0
    def dlg_memset32(self, d, w, n): return 0
    def dlg_host_alloc(self, n):
        buf = ffi.new("char[]", int(n) or 1)        # real host memory: callers write through it
        self._host = getattr(self, "_host", []) + [buf]
        return ffi.cast("void *", buf)
    def dlg_host_free(self, p): return None
    def dlg_h2d_async(self, d, s, n): return 0
    def dlg_d2h_async(self, d, s, n): return 0
    def dlg_module_load(self, path, out): out[0] = ffi.cast("void *", 1); return 0
    def dlg_module_get(self, mod, name, out): out[0] = ffi.cast("void *", 1); return 0
    def dlg_launch(self, *a): self.launches += 1; return 0
    def dlg_plan_create(self, fn, gx, gy, gz, bx, by, bz, smem, out_bytes, zero, cl):
        return ffi.cast("void *", int(out_bytes) + 1)
    def dlg_plan_destroy(self, p): return None
    def dlg_plan_run(self, plan, *a):
        self.launches += 1
        n = int(ffi.cast("uintptr_t", plan)) - 1
        return self._addr(n) if n else NULL
    def dlg_plan_run_into(self, plan, out, *a): self.launches += 1; return 0
    def dlg_last_status(self): return 0
    def dlg_launch_count(self): return self.launches
    def dlg_mm_plan_create(self, fn, gx, gy, gz, th, smem, cl, out_bytes, a, b, c):
        return ffi.cast("void *", int(out_bytes) + 1)
    def dlg_mm_plan_run(self, plan, a, b):
        self.launches += 1
        return self._addr(int(ffi.cast("uintptr_t", plan)) - 1)
    def dlg_mm_plan_run_into(self, plan, c, a, b): self.launches += 1; return 0
    def dlg_mm_plan_run_aux(self, plan, a, b, aux0, aux1): return self.dlg_mm_plan_run(plan, a, b)
    def dlg_tc_plan_create(self, fn, grid, threads, smem, ntmaps, maps, nptrs): return ffi.cast("void *", 1)
    def dlg_tc_plan_run(self, plan, bases, ptrs): self.launches += 1; return 0
    def dlg_event_create(self): return ffi.cast("void *", 1)
    def dlg_event_record(self, e): return 0
    def dlg_event_sync(self, e): return 0
    def dlg_event_elapsed_ms(self, a, b): return 1.0
    def dlg_event_destroy(self, e): return None
    def dlg_graph_begin(self): return 0
    def dlg_graph_end(self): return ffi.cast("void *", 1)
    def dlg_graph_launch(self, g): return 0
    def dlg_graph_destroy(self, g): return None
    def dlg_pool_begin(self): return 1
    def dlg_pool_end(self): return 0
    def dlg_pool_destroy(self, p): return 0
    def dlg_nccl_unique_id(self, out): return 0
    def dlg_nccl_init(self, r, w, i): return 0
    def dlg_nccl_allreduce_sum(self, p, n): return 0
    def dlg_nccl_wait(self): return 0
    def dlg_nccl_destroy(self): return 0
    def dlg_nccl_version(self): return 0
    def dlg_ipc_alloc(self, n): return self._addr(n)
    def dlg_ipc_free(self, p): return 0
    def dlg_ipc_export(self, p, h): return 0
    def dlg_ipc_open(self, h): return self._addr(1 << 21)
    def dlg_ipc_close(self, p): return 0
    def dlg_comm_after_compute(self): return 0
    def dlg_compute_after_comm(self): return 0
    def dlg_comm_plan_run_into(self, plan, out, *a): self.launches += 1; return 0
    def dlg_comm_memops_supported(self): return 1
    def dlg_comm_wait_geq32(self, p, v): return 0


def _load():
    if COMPILE_ONLY:
        return _CompileOnlyLib()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or dlgrad_b200/build.py).  dlgrad_b200 has no CPU fallback.")
    return ffi.dlopen(LIB_PATH)


lib = _load()
_state = {"device": None}

# set by dlgrad_b200/graph.py while a step is being captured into a CUDA graph: host round trips are illegal then
# (the stream cannot be synchronised), and host data uploaded inside the step is staged in pinned memory the graph owns
capture = {"on": False, "keep": None, "hooks": None, "rng": None, "refs": None}


def no_capture(what: str) -> None:
    if capture["on"]:
        raise RuntimeError(f"{what} inside a step that is being captured into a CUDA graph: the host cannot wait for "
                           "the device there.  Pass per-step data as graph inputs and read results after the replay.")


def last_error() -> str:
    return ffi.string(lib.dlg_last_error()).decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"CUDA runtime error in {what or 'libdlgrad_cuda'}: {last_error()}")


def init(device: int | None = None) -> int:
    """Bind this process to one GPU (one process per GPU; LOCAL_RANK selects it under a launcher)."""
    if _state["device"] is not None:
        return _state["device"]
    if device is None:
        device = int(os.environ.get("DLGRAD_CUDA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    rc = lib.dlg_init(device)
    if rc != 0:
        raise RuntimeError(
            "Device.CUDA is unavailable and dlgrad_b200 has no CPU fallback: " + last_error())
    _state["device"] = device
    return device


def is_initialized() -> bool:
    return _state["device"] is not None


def device_available() -> bool:
    if COMPILE_ONLY:
        return True
    try:
        return lib.dlg_device_count() > 0
    except Exception:
        return False


def synchronize() -> None:
    if is_initialized():
        no_capture("synchronize()")
        check(lib.dlg_sync(), "dlg_sync")


def alloc(nbytes: int):
    """Fresh device storage owned by the returned cdata (ffi.gc → dlg_free), like CPU.malloc
    (dlgrad/runtime/cpu.py:82-107)."""
    if _state["device"] is None:
        init()
    p = lib.dlg_alloc(nbytes)
    if p == NULL:
        raise MemoryError(f"Failed to allocate {nbytes} bytes of device memory: {last_error()}")
    return own(p, nbytes)


def calloc(nbytes: int):
    """Zero-filled device storage (dlg_calloc), owned like alloc()."""
    if _state["device"] is None:
        init()
    p = lib.dlg_calloc(nbytes)
    if p == NULL:
        raise MemoryError(f"Failed to allocate {nbytes} bytes of device memory: {last_error()}")
    return own(p, nbytes)


_FLOAT_P = ffi.typeof("float *")
_UINTPTR = ffi.typeof("uintptr_t")
_cast, _gc, _free = ffi.cast, ffi.gc, lib.dlg_free


def own(p, nbytes: int = 0):
    """Attach allocator ownership to a pointer returned by dlg_plan_run / dlg_alloc."""
    return _gc(_cast(_FLOAT_P, p), _free, nbytes)


def address(ptr) -> int:
    return int(_cast(_UINTPTR, ptr))
