"""How much does a kernel launch cost between back-to-back kernels on the compute stream?

Launches N small elementwise kernels (a) eagerly through the plan API and (b) replayed from one CUDA graph,
and a medium one (25 MB, the GPT residual add) the same way.  Per-launch time = total / N, CUDA events.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dlgrad_b200 import Tensor  # noqa: E402
from dlgrad_b200.runtime.cuda import profile, shim  # noqa: E402

shim.init()
rng = np.random.default_rng(0)
N = 300
for label, shape in (("25 MB (GPT residual add)", (8192, 768)), ("100 MB", (8192, 3072)), ("268 MB", (8192, 8192))):
    a = Tensor(rng.random(shape, dtype=np.float32), device="cuda")
    b = Tensor(rng.random(shape, dtype=np.float32), device="cuda")
    for _ in range(5):
        c = a + b
    shim.synchronize()
    # eager
    e0 = profile.Event().record()
    for _ in range(N):
        c = a + b                      # the previous output is released: the allocator recycles two blocks
    e1 = profile.Event().record()
    e1.sync()
    eager_us = e1.ms_since(e0) * 1e3 / N
    # graph replay
    keep = []
    shim.check(shim.lib.dlg_graph_begin(), "begin")
    for _ in range(N):
        keep.append(a + b)
        if len(keep) > 2:
            keep.pop(0)
    g = shim.lib.dlg_graph_end()
    shim.check(shim.lib.dlg_graph_launch(g), "warm")
    shim.synchronize()
    e0 = profile.Event().record()
    shim.check(shim.lib.dlg_graph_launch(g), "launch")
    e1 = profile.Event().record()
    e1.sync()
    graph_us = e1.ms_since(e0) * 1e3 / N
    shim.lib.dlg_graph_destroy(g)
    nbytes = 3 * 4 * shape[0] * shape[1]
    print(f"{label:28s} eager {eager_us:7.2f} us/launch   graph {graph_us:7.2f} us/launch   (pure traffic at 6.4 TB/s: {nbytes / 6.4e6:6.2f} us)", flush=True)
