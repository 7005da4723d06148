"""nvcc driver + on-disk kernel cache for the generated CUDA C.

Same idea as the reference's CPU path — hash the generated source, compile once, cache the object,
load it on later runs (dlgrad/runtime/cpu.py:47-62: write .c, fork the compiler, sha256 key) — with
sm_100a cubins instead of host shared objects:

    source --sha256--> _kcache/<key>.cubin            (one kernel, compiled on first use)
                  \\--> _kcache/pack_<n>.cubin + index  (many kernels, pre-built by build())

The cache directory lives inside the package so pre-built cubins travel with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import shutil
import subprocess
import tempfile

from dlgrad_b200.helpers import CACHE_DIR
from dlgrad_b200.runtime.cuda import shim

NVCC = os.environ.get("DLGRAD_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-O3", "-lineinfo", "-std=c++17",
              "-Xptxas", "-warn-spills"]
_FLAG_TAG = " ".join(NVCC_FLAGS)

KNAME = "DLG_KERNEL_NAME"     # placeholder the code generator puts where the function name goes

_index: dict[str, list[str]] | None = None      # key -> [cubin file name, function name]
_modules: dict[str, object] = {}                # cubin path -> CUmodule handle
_functions: dict[str, object] = {}              # key -> CUfunction handle
_pending: dict[str, tuple[str, str]] = {}       # compile-only mode: key -> (fname, source)


def _index_path() -> str:
    return os.path.join(CACHE_DIR, "index.json")


def _load_index() -> dict:
    global _index
    if _index is None:
        try:
            with open(_index_path()) as f:
                _index = json.load(f)
        except (OSError, ValueError):
            _index = {}
    return _index


def source_key(src: str) -> str:
    return hashlib.sha256((_FLAG_TAG + "\n" + src).encode()).hexdigest()[:32]


def _run_nvcc(cu_path: str, out_path: str) -> None:
    cmd = [NVCC, *NVCC_FLAGS, "-o", out_path, cu_path]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"Compilation failed ({' '.join(cmd)}):\n{res.stderr}\n--- source: {cu_path}")
    if "spill" in res.stderr and "0 bytes spill" not in res.stderr:
        # not fatal, but never silent
        print(f"[dlgrad_b200] nvcc reported spills for {cu_path}:\n{res.stderr}")


def _compile_single(key: str, fname: str, src: str) -> str:
    os.makedirs(CACHE_DIR, exist_ok=True)
    out = os.path.join(CACHE_DIR, f"{key}.cubin")
    tmpdir = tempfile.mkdtemp(prefix="dlgk_")
    try:
        cu = os.path.join(tmpdir, f"{fname}.cu")
        with open(cu, "w") as f:
            f.write(src.replace(KNAME, fname))
        tmp_out = os.path.join(tmpdir, "k.cubin")
        _run_nvcc(cu, tmp_out)
        # several ranks may compile the same missing kernel at once: publish with an atomic rename inside the cache
        staged = f"{out}.{os.getpid()}.tmp"
        shutil.copyfile(tmp_out, staged)
        os.replace(staged, out)
    finally:
        if not os.environ.get("DLGRAD_KEEP_SRC"):
            shutil.rmtree(tmpdir, ignore_errors=True)
    return out


def _module(path: str):
    m = _modules.get(path)
    if m is None:
        out = shim.ffi.new("void **")
        shim.check(shim.lib.dlg_module_load(path.encode(), out), f"dlg_module_load({path})")
        m = _modules[path] = out[0]
    return m


_ENTRY = re.compile(r'(extern "C" __global__ void[^{;]*\{)')
_PDL_PROLOGUE = ('\n    // programmatic dependent launch: let the next kernel start its prologue now; touch global memory'
                 '\n    // only after every earlier kernel on the stream has completed and flushed'
                 '\n    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");'
                 '\n    asm volatile("griddepcontrol.wait;" ::: "memory");')
PDL = os.environ.get("DLGRAD_PDL", "1") != "0"


def with_pdl_prologue(src: str) -> str:
    """Insert the griddepcontrol pair at the top of the kernel (csrc/dlgrad_cuda.cpp launch_ex sets the
    matching launch attribute).  Both are no-ops for a kernel launched without the attribute."""
    if not PDL:
        return src
    out, n = _ENTRY.subn(lambda m: m.group(1) + _PDL_PROLOGUE, src, count=1)
    if n != 1:
        raise RuntimeError("generated kernel without an extern \"C\" __global__ entry point")
    return out


def get_function(base: str, src: str):
    """Return the CUfunction for generated source `src` (function name placeholder = KNAME)."""
    src = with_pdl_prologue(src)
    key = source_key(src)
    fn = _functions.get(key)
    if fn is not None:
        return fn
    fname = f"{base}_{key[:12]}"
    idx = _load_index()
    path = None
    if key in idx:
        cand = os.path.join(CACHE_DIR, idx[key][0])
        if os.path.exists(cand):
            path, fname = cand, idx[key][1]
    if path is None:
        cand = os.path.join(CACHE_DIR, f"{key}.cubin")
        if os.path.exists(cand):
            path = cand
    if path is None:
        if shim.COMPILE_ONLY:
            _pending[key] = (fname, src)
            _functions[key] = shim.ffi.cast("void *", 1)
            return _functions[key]
        path = _compile_single(key, fname, src)
    if shim.COMPILE_ONLY:
        fn = shim.ffi.cast("void *", 1)
    else:
        out = shim.ffi.new("void **")
        shim.check(shim.lib.dlg_module_get(_module(path), fname.encode(), out), f"dlg_module_get({fname})")
        fn = out[0]
    _functions[key] = fn
    return fn


def flush_pending(jobs: int | None = None, pack_size: int = 24) -> int:
    """Compile everything recorded in compile-only mode into pack cubins (nvcc processes in parallel)."""
    global _index
    if not _pending:
        return 0
    os.makedirs(CACHE_DIR, exist_ok=True)
    idx = _load_index()
    items = sorted(_pending.items())
    packs = [items[i:i + pack_size] for i in range(0, len(items), pack_size)]
    jobs = jobs or (os.cpu_count() or 4)
    tmpdir = tempfile.mkdtemp(prefix="dlgp_")
    try:
        todo = []
        for pack in packs:
            h = hashlib.sha256("".join(k for k, _ in pack).encode()).hexdigest()[:16]
            name = f"pack_{h}.cubin"
            cu = os.path.join(tmpdir, f"pack_{h}.cu")
            with open(cu, "w") as f:
                for key, (fname, src) in pack:
                    # each kernel in its own namespace so file-local helpers cannot collide
                    f.write(f"// ---- {fname}\nnamespace ns_{fname} {{\n{src.replace(KNAME, fname)}\n}}\n")
            todo.append((name, cu, os.path.join(tmpdir, name), pack))
        running: list = []
        queue = list(todo)
        while queue or running:
            while queue and len(running) < jobs:
                name, cu, out, pack = queue.pop(0)
                proc = subprocess.Popen([NVCC, *NVCC_FLAGS, "-o", out, cu], stdout=subprocess.PIPE,
                                        stderr=subprocess.PIPE, text=True)
                running.append((proc, name, cu, out, pack))
            proc, name, cu, out, pack = running.pop(0)
            _so, se = proc.communicate()
            if proc.returncode != 0:
                for other in running:
                    other[0].kill()
                raise RuntimeError(f"Compilation failed for {cu}:\n{se}")
            shutil.move(out, os.path.join(CACHE_DIR, name))
            for key, (fname, _src) in pack:
                idx[key] = [name, fname]
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    with open(_index_path() + ".tmp", "w") as f:
        json.dump(idx, f, indent=0, sort_keys=True)
    os.replace(_index_path() + ".tmp", _index_path())
    n = len(_pending)
    _pending.clear()
    return n


def pending_count() -> int:
    return len(_pending)


if shim.COMPILE_ONLY:
    import atexit
    atexit.register(flush_pending)
