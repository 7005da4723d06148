"""Host-side cost of one train step with the device stubbed out (DLGRAD_CUDA_COMPILE_ONLY=1): what the Python
autograd layer + dispatcher + plan lookup cost per launch when no kernel runs.

    DLGRAD_CUDA_COMPILE_ONLY=1 python scripts/host_dispatch_probe.py [c4|c5] [--profile]

Prints ms of host time per step, launches per step and microseconds per launch; `--profile` adds a cProfile table
(sorted by own time) of the steady-state steps.  Runs on the GPU-less build machine.
"""
import os
import sys
import time

os.environ.setdefault("DLGRAD_CUDA_COMPILE_ONLY", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from tests import models  # noqa: E402


def main() -> None:
    which = "c5" if "c5" in sys.argv else "c4"
    cfg_kw, batch = (bench.C5, 8) if which == "c5" else (bench.C4, 16)
    import dlgrad_b200 as dl  # noqa: F401
    from dlgrad_b200 import Tensor
    from dlgrad_b200.runtime.cuda import shim
    cfg, gpt, params, opt = bench.build_gpt(cfg_kw, "cuda")
    xs, ys = models.gpt_batch(cfg, batch, np.random.default_rng(7))
    x, y = Tensor(xs, device="cuda"), Tensor(ys, device="cuda")

    def step():
        return bench.train_step(gpt, opt, params, x, y, 1)
    for _ in range(3):
        step()
    n = 10 if which == "c5" else 30
    l0 = shim.lib.dlg_launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    launches = (shim.lib.dlg_launch_count() - l0) / n
    print(f"{which}: host {dt * 1e3:.3f} ms/step, {launches:.0f} launches/step, {dt * 1e6 / launches:.2f} us/launch")
    if "--profile" in sys.argv:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(n):
            step()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(45)


if __name__ == "__main__":
    main()
