"""SASS opcode histogram rows for generated kernels (append to profiles/r02_sass_opcode_histogram.md):

    DLGRAD_CUDA_COMPILE_ONLY=1 python scripts/sass_histogram.py

Compiles the three grouped forms of the tensor-core GEMM at the benchmark's sizes (or takes them from the kernel cache) and
counts the instructions of `cuobjdump -sass` by mnemonic prefix."""
import os
import re
import subprocess
import sys
import tempfile

os.environ.setdefault("DLGRAD_CUDA_COMPILE_ONLY", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dlgrad_b200.codegen import cuda_tcgen05 as tc  # noqa: E402
from dlgrad_b200.runtime.cuda import compiler  # noqa: E402

COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMAREDG", "SYNCS", "ELECT", "MUFU.EX2", "MUFU",
        "BAR.SYNC", "LDS", "STS", "LDG", "STG", "FFMA", "FMNMX", "SHFL", "R2UR", "UTCATOMSWS"]


def sass_of(kernel) -> list[str]:
    with tempfile.TemporaryDirectory() as d:
        cu, cubin = os.path.join(d, "k.cu"), os.path.join(d, "k.cubin")
        text = kernel.source.replace(compiler.KNAME, kernel.name)
        open(cu, "w").write(text)
        subprocess.run([compiler.NVCC] + compiler.NVCC_FLAGS + ["-o", cubin, cu], check=True, capture_output=True)
        out = subprocess.run(["cuobjdump", "-sass", cubin], check=True, capture_output=True, text=True).stdout
    ops = []
    for ln in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            ops.append(m.group(1))
    return ops


def row(label: str, kernel) -> str:
    ops = sass_of(kernel)
    counts = [sum(1 for o in ops if o.startswith(c)) for c in COLS]
    return f"| {label} | {len(ops)} | " + " | ".join(str(c) for c in counts) + " |"


if __name__ == "__main__":
    ks = [("mm_tcts_gn3 8192x(3x768)x768 (q / k / v forward, one launch)", tc.matmul_tc_ts(8192, 768, 768, 3, False, False, True, False, group=("n", 3))),
          ("mm_tcts_gk3 8192x768x(3x768) (dq Wq + dk Wk + dv Wv, one launch)", tc.matmul_tc_ts(8192, 768, 2304, 1, False, True, False, False, group=("k", 3))),
          ("mm_tcts_k4_gm3 3 x (768x768x8192 in 4 K-slices) (dWq, dWk, dWv, one launch)", tc.matmul_tc_ts(768, 768, 2048, 12, True, True, False, False, bn=128, splitk=4, group=("m", 3)))]
    for label, k in ks:
        print(row(label, k))
