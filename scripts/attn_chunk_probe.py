"""Does running attention per batch element (12 heads, a 50 MB score tensor that fits L2) beat running it over the
whole (8,12,1024,1024) tensor (403 MB per pass through HBM)?  Same kernels, same values, different order.

    python scripts/attn_chunk_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dlgrad_b200 import Tensor  # noqa: E402
from dlgrad_b200.runtime.cuda import profile, shim  # noqa: E402

shim.init()
rng = np.random.default_rng(0)
B, H, T, D = 8, 12, 1024, 64
q, k, v = (Tensor((rng.random((B, H, T, D), dtype=np.float32) - 0.5), device="cuda") for _ in range(3))
mask = Tensor(np.triu(np.ones((T, T), dtype=np.float32), 1), device="cuda")
scale = 1.0 / 8.0


def attn(qq, kk, vv):
    att = ((qq @ kk.transpose(2, 3)) * scale).masked_fill(mask, float("-inf")).softmax(dim=3)
    return att @ vv


def full():
    return [attn(q, k, v)]


def chunked(n):
    step = B // n
    return [attn(q[i:i + step], k[i:i + step], v[i:i + step]) for i in range(0, B, step)]


def timeit(fn, reps=10):
    """The launches of one call are captured in a CUDA graph and the graph is replayed: Python dispatch
    (~14 us per op) would otherwise hide what the GPU does with 8 x 5 short kernels."""
    for _ in range(2):
        out = fn()
    shim.synchronize()
    shim.check(shim.lib.dlg_graph_begin(), "begin")
    out = fn()
    g = shim.lib.dlg_graph_end()
    shim.check(shim.lib.dlg_graph_launch(g), "warm")
    shim.synchronize()
    e0 = profile.Event().record()
    for _ in range(reps):
        shim.check(shim.lib.dlg_graph_launch(g), "launch")
    e1 = profile.Event().record()
    e1.sync()
    ms = e1.ms_since(e0) / reps
    shim.lib.dlg_graph_destroy(g)
    return ms, out


t_full, o_full = timeit(full)
print(f"whole tensor (96 heads at once):        {t_full * 1e3:8.1f} us per forward attention")
for n in (2, 4, 8):
    t, o = timeit(lambda n=n: chunked(n))
    same = True
    if not shim.COMPILE_ONLY:
        ref = o_full[0].numpy()
        got = np.concatenate([x.numpy() for x in o], axis=0)
        same = bool((ref == got).all())
    print(f"{n} chunks of {96 // n:2d} heads ({403 // n:3d} MB scores each): {t * 1e3:8.1f} us   bit-identical: {same}")
