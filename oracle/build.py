"""Builds the oracle's shared objects.

  liboracle   oracle/kernels.c, flags = the reference's CFLAGS (dlgrad/runtime/cpu.py:30:
              -shared -fPIC -O3 -march=native -ffast-math, libs -lm -lmvec).  Because of
              -march=native the object is keyed by the host CPU's feature flags and rebuilt on a
              different machine (the GPU box has gcc; this is our own source).
  _ref/       the reference's own kernels: dlgrad/codegen/cpu_kernel.py is imported from
              /root/reference (when mounted), each generator's C text is compiled exactly as the
              reference would (same flags) into oracle/_ref/<name>.so.  Only objects are written —
              no reference source is copied into the repo.  -march=x86-64-v3 replaces -march=native
              there so the objects also run on the GPU box's host CPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"
CFLAGS = ["-shared", "-fPIC", "-O3", "-march=native", "-ffast-math"]
LIBS = ["-lm", "-lmvec"]


def _cpu_tag() -> str:
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha256(flags.encode()).hexdigest()[:10]


def liboracle_path() -> str:
    return os.path.join(BUILD_DIR, f"liboracle_{_cpu_tag()}.so")


def build_oracle(force: bool = False) -> str:
    out = liboracle_path()
    src = os.path.join(HERE, "kernels.c")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ["gcc", *CFLAGS, "-o", out, src, *LIBS]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"oracle build failed: {' '.join(cmd)}\n{res.stderr}")
    return out


# (file stem, generator name, args) — the kernels of SURVEY §2a that are on the hot path
_REF_KERNELS = (
    [(f"{op}_{nd}d", "arithmetic", (op, nd)) for op in ("add", "sub", "mul", "div", "eq", "gt", "gte")
     for nd in (0, 1, 2, 3, 4)]      # 0-d: scalar (op) scalar, e.g. (loss_real + loss_fake) / 2.0 of the GAN script
    + [(f"c_{f}", "utils", (f,)) for f in ("neg", "exp", "log", "pow", "sqrt", "rsqrt")]
    + [(f"reduce_{op}_{nd}d_axis{d}", "reduce_source", (op, nd, d))
       for op in ("sum", "max", "mean") for nd in (1, 2, 3, 4) for d in range(nd)]
    + [(f"reduce_{op}", "reduce", (op,)) for op in ("sum", "max", "mean")]
    + [(f"permute_{nd}d", "permute", (nd,)) for nd in (2, 3, 4)]
    + [(f"where_{nd}d", "where", (nd,)) for nd in (1, 2, 3, 4)]
    + [(f"masked_fill_{nd}d", "masked_fill", (nd,)) for nd in (2, 3, 4)]
    + [(f"argmax_2d_axis{d}", "argmax", (2, d)) for d in (0, 1)]
    + [("matmul_2d", "matmul_2d", ()), ("matmul_3d", "matmul_3d", ()), ("matmul_4d", "matmul_4d", ()),
       ("ce_forward", "ce_forward", ()), ("ce_backward", "ce_backward", ()),
       ("embedding", "embedding", ()), ("embedding_backward", "embedding_backward", ()),
       ("cmp", "cmp_2d", ("<=",)), ("full", "full", ()), ("arange", "arange", ()), ("relu", "relu", ())]
    + [(f"clamp_{int(a)}{int(b)}", "clamp", (a, b)) for a in (False, True) for b in (False, True)]
)


def build_ref(force: bool = False) -> int:
    """Compile the reference's generated kernels (only possible where /root/reference is mounted)."""
    gen_path = os.path.join(REFERENCE_ROOT, "dlgrad", "codegen", "cpu_kernel.py")
    if not os.path.exists(gen_path):
        return 0
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_cpu_kernel", gen_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    os.makedirs(REF_DIR, exist_ok=True)
    flags = [f if f != "-march=native" else "-march=x86-64-v3" for f in CFLAGS]
    built = 0
    with tempfile.TemporaryDirectory(prefix="dlg_ref_") as tmp:
        for stem, gen, args in _REF_KERNELS:
            out = os.path.join(REF_DIR, f"{stem}.so")
            if os.path.exists(out) and not force:
                built += 1
                continue
            code, _cdef = getattr(mod, gen)(*args)
            cpath = os.path.join(tmp, f"{stem}.c")
            with open(cpath, "w") as f:
                f.write(code)
            res = subprocess.run(["gcc", *flags, "-o", out, cpath, *LIBS], capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"reference kernel {stem} failed to compile:\n{res.stderr}")
            built += 1
    return built


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print("reference kernels:", build_ref(force="--force" in sys.argv))
