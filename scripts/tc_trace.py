"""Pipeline trace of the persistent 3xTF32 kernel: CTA 0 timestamps (clock64) its first 32 k-blocks.

    DLGRAD_TC_TRACE=1 python scripts/tc_trace.py [M N K bn stages]

Columns (cycles, relative to the producer's first TMA issue):
    tma   producer saw `empty` and issues the loads of k-block i
    full  splitter saw the tile land          split  splitter done (lo written, arrive)
    mma   MMA warp saw `split`                commit MMA warp has issued 12 MMAs + commit
"""
import os
import sys

os.environ.setdefault("DLGRAD_TC_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dlgrad_b200.runtime.cuda import matmul_tc, shim, transfer  # noqa: E402

args = [int(x) for x in sys.argv[1:]]
M, N, K, bn = (args + [4096, 4096, 4096, 128][len(args):])[:4]
stages = (args[4] or None) if len(args) > 4 else None
kernel = {3: "v3", 4: "v4"}.get(args[5], "v2") if len(args) > 5 else "v2"
shim.init()
rng = np.random.default_rng(0)
a = rng.random((1, M, K), dtype=np.float32)
b = rng.random((1, K, N), dtype=np.float32)
plan = matmul_tc.try_build(M, N, K, (1, 1), (0, 0), (0, 0), False, False, True, bn=bn, stages=stages, force=True, kernel=kernel)
da, db = transfer.upload_numpy(a), transfer.upload_numpy(b)
for _ in range(int(os.environ.get("TC_TRACE_ITERS", "3"))):
    out = plan(da, db)
shim.synchronize()
if shim.COMPILE_ONLY:
    sys.exit(0)
host = transfer.download(out, 320).view(np.int64).reshape(5, 32)
t0 = host[0, 0]
print(f"# {kernel} {M}x{N}x{K} bn{bn} stages={stages}")
if os.environ["DLGRAD_TC_TRACE"] == "3":
    cyc, ns = int(host[0, 2] - host[0, 0]), int(host[0, 3] - host[0, 1])
    tiles = -(-M // 128) * -(-N // bn)
    kbs = -(-tiles // 148) * -(-K // 32)
    print(f"CTA0: {cyc} cycles in {ns} ns -> {cyc / ns * 1e3:.0f} MHz; {kbs} k-blocks -> {cyc / kbs:.0f} cycles / {ns / kbs:.0f} ns per k-block")
    sys.exit(0)
if os.environ["DLGRAD_TC_TRACE"] == "2":
    print("  i     top   ready  issued  commit    sync | wait+fence issue12  commit syncwarp  loop(back to top)")
    for i in range(32):
        t = host[:, i] - t0
        nxt = host[0, i + 1] - host[4, i] if i < 31 else 0
        print(f"{i:3d} {t[0]:7d} {t[1]:7d} {t[2]:7d} {t[3]:7d} {t[4]:7d} | {t[1]-t[0]:6d} {t[2]-t[1]:10d} {t[3]-t[2]:8d} {t[4]-t[3]:7d} {nxt:6d}")
    sys.exit(0)
print("  i     tma    full   split     mma  commit | full-tma split-full mma-split  d(tma) d(commit)")
for i in range(32):
    t = host[:, i] - t0
    dt = host[0, i] - host[0, i - 1] if i else 0
    dc = host[4, i] - host[4, i - 1] if i else 0
    print(f"{i:3d} {t[0]:7d} {t[1]:7d} {t[2]:7d} {t[3]:7d} {t[4]:7d} | {t[1]-t[0]:8d} {t[2]-t[1]:10d} {t[3]-t[2]:9d} {dt:7d} {dc:9d}")
