"""GPU probe: column-reduce variants on 4096x4096 (graph-timed, inputs rotated through > L2)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200.codegen import cuda_kernel as ck
from dlgrad_b200.runtime.cuda import ops, profile, shim, transfer

shim.init()
n, SETS = 4096, 6
rng = np.random.default_rng(0)
ins = [transfer.upload_numpy(rng.random((n, n), dtype=np.float32)) for _ in range(SETS)]
ref = None


def timed(run, rounds=4):
    for i in range(SETS):
        run(i)
    shim.synchronize()
    shim.check(shim.lib.dlg_graph_begin(), "begin")
    keep = [run(i) for _ in range(rounds) for i in range(SETS)]
    g = shim.lib.dlg_graph_end()
    best = 1e9
    for _ in range(4):
        e0 = profile.Event().record(); shim.check(shim.lib.dlg_graph_launch(g), "launch"); e1 = profile.Event().record(); e1.sync()
        best = min(best, e1.ms_since(e0))
    return best * 1e3 / (rounds * SETS)


for slabs in (74, 148, 296, 592, 1184):
    p1 = ops.Plan(ck.reduce_cols_slab("sum", 1, n, n, slabs, False))
    us = timed(lambda i: p1.run(ins[i]))
    print(f"slab pass  slabs={slabs:5d}: {us:7.2f} us  {n*n*4/us/1e3:8.1f} GB/s", flush=True)
    p2 = ops.Plan(ck.reduce_cols("sum", 1, slabs, n, True, 1, scale_R=n))
    part = p1.run(ins[0])
    us2 = timed(lambda i: p2.run(part))
    print(f"   strip pass over {slabs} partial rows: {us2:7.2f} us", flush=True)
for splits in (4, 8):
    p = ops.Plan(ck.reduce_cols("sum", 1, n, n, True, splits, None, True))
    us = timed(lambda i: p.run(ins[i]))
    print(f"cluster strips splits={splits}: {us:7.2f} us  {n*n*4/us/1e3:8.1f} GB/s", flush=True)
for splits in (16, 32, 64, 128):
    p1 = ops.Plan(ck.reduce_cols("sum", 1, n, n, True, splits))
    p2 = ops.Plan(ck.reduce_cols("sum", 1, splits, n, True, 1, scale_R=n))
    us = timed(lambda i: p2.run(p1.run(ins[i])))
    print(f"strips two-pass splits={splits}: {us:7.2f} us  {n*n*4/us/1e3:8.1f} GB/s", flush=True)
p = ops.Plan(ck.reduce_rows("sum", n, n))
us = timed(lambda i: p.run(ins[i]))
print(f"row reduce (axis 1): {us:7.2f} us  {n*n*4/us/1e3:8.1f} GB/s")
