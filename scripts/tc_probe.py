"""GPU probe for the tcgen05 matmul: correctness vs float64 for every operand-major combination, the
TF32 input-conversion semantics of tcgen05 (truncate or round), and timing.  Run under gpurun."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200.runtime.cuda import matmul_tc, profile, shim, transfer  # noqa: E402

shim.init()
rng = np.random.default_rng(0)
results = {}


def run(M, N, K, nb=1, tx=False, ty=False, a_b=False, b_b=False, passes=3, bn=None, stages=None, hi_inplace=False,
        a=None, b=None, time_it=0, kernel=None, split_warps=None, streamk=None):
    batch = (1, nb)
    if a is None:
        a = (rng.random(((1 if a_b else nb), K, M) if tx else ((1 if a_b else nb), M, K), dtype=np.float32) - 0.5).astype(np.float32)
    if b is None:
        b = (rng.random(((1 if b_b else nb), N, K) if ty else ((1 if b_b else nb), K, N), dtype=np.float32) - 0.5).astype(np.float32)
    a_bs = (0, 0) if (a_b or nb == 1) else (0, M * K)
    b_bs = (0, 0) if (b_b or nb == 1) else (0, K * N)
    plan = matmul_tc.try_build(M, N, K, batch, a_bs, b_bs, tx, ty, True, passes=passes, bn=bn, stages=stages,
                               hi_inplace=hi_inplace, force=True, kernel=kernel, split_warps=split_warps, streamk=streamk)
    assert plan is not None, "shape did not qualify"
    da, db = transfer.upload_numpy(a), transfer.upload_numpy(b)
    out = plan(da, db)
    shim.synchronize()
    c = transfer.download(out, nb * M * N).reshape(nb, M, N)
    A = a.astype(np.float64)
    B = b.astype(np.float64)
    A = A.transpose(0, 2, 1) if tx else A
    B = B.transpose(0, 2, 1) if ty else B
    ref = A @ B
    err = float(np.abs(c - ref).max() / np.abs(ref).max())
    ms = None
    if time_it:
        for _ in range(2):
            plan(da, db)
        e0 = profile.Event().record()
        for _ in range(time_it):
            plan(da, db)
        e1 = profile.Event().record()
        e1.sync()
        ms = e1.ms_since(e0) / time_it
    return err, ms, c


def case(name, **kw):
    t0 = time.time()
    try:
        err, ms, _ = run(**kw)
        tf = (2.0 * kw["M"] * kw["N"] * kw["K"] * kw.get("nb", 1) / (ms * 1e-3) / 1e12) if ms else None
        results[name] = {"rel_err": err, "ms": ms, "tflops": tf}
        print(f"{name:38s} rel_err {err:.3e}" + (f"  {ms:.3f} ms  {tf:.1f} TFLOP/s" if ms else ""), flush=True)
    except Exception as exc:  # keep going: every case is independent evidence
        results[name] = {"error": repr(exc)[:300]}
        print(f"{name:38s} FAILED {exc!r}"[:400], flush=True)
    return time.time() - t0


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "sem"):
    # conversion semantics: x = 1 + 2^-11 + 2^-12; truncation -> 1.0, round-to-nearest -> 1 + 2^-10
    x = np.float32(1.0 + 2.0 ** -11 + 2.0 ** -12)
    a = np.full((1, 128, 32), x, dtype=np.float32)
    b = np.zeros((1, 32, 32), dtype=np.float32)
    b[0, 0, :] = 1.0
    try:
        _, _, c = run(128, 32, 32, passes=1, a=a, b=b)
        v = float(c[0, 0, 0])
        results["tf32_input_conversion"] = {"value": v, "mode": "truncate" if v == 1.0 else ("round" if v == 1.0 + 2.0 ** -10 else "other")}
        print("tf32 input conversion:", results["tf32_input_conversion"], flush=True)
    except Exception as exc:
        print("semantics probe failed", repr(exc)[:300], flush=True)
if which in ("all", "p1"):
    case("p1 NN 128x128x128", M=128, N=128, K=128, passes=1)
    case("p1 NN 256x256x64 bn256", M=256, N=256, K=64, passes=1)
    case("p1 NT 128x128x128", M=128, N=128, K=128, ty=True, passes=1)
    case("p1 TN 128x128x128", M=128, N=128, K=128, tx=True, passes=1)
if which in ("all", "p3"):
    case("p3 NN 128x128x128", M=128, N=128, K=128)
    case("p3 NT 128x128x128", M=128, N=128, K=128, ty=True)
    case("p3 TN 128x128x128", M=128, N=128, K=128, tx=True)
    case("p3 NN 200x96x72 ragged", M=200, N=96, K=72)
    case("p3 NN 333x100x44 ragged", M=333, N=100, K=44)
    case("p3 NN 1000x1000x1000", M=1000, N=1000, K=1000)
    case("p3 NN batched 8x1024x96x768 bB", M=1024, N=96, K=768, nb=8, b_b=True)
    case("p3 NN batched 96x1024x64x1024", M=1024, N=64, K=1024, nb=96)
    case("p3 NT batched 96x1024x1024x64", M=1024, N=1024, K=64, nb=96, ty=True)
    case("p3 TN batched 96x1024x64x1024", M=1024, N=64, K=1024, nb=96, tx=True)
    case("p3 NN 4096^3 accuracy U[0,1)", M=4096, N=4096, K=4096,
         a=np.random.default_rng(5).random((1, 4096, 4096), dtype=np.float32), b=np.random.default_rng(6).random((1, 4096, 4096), dtype=np.float32))
    case("p3 TN 768x3072x8192 accuracy U[0,1)", M=768, N=3072, K=8192, tx=True,
         a=np.random.default_rng(5).random((1, 8192, 768), dtype=np.float32), b=np.random.default_rng(6).random((1, 8192, 3072), dtype=np.float32))
    case("p3 NN 512x512x16384 accuracy U[0,1)", M=512, N=512, K=16384,
         a=np.random.default_rng(5).random((1, 512, 16384), dtype=np.float32), b=np.random.default_rng(6).random((1, 16384, 512), dtype=np.float32))
if which in ("all", "v3"):
    case("v3 NN 256x256x128", M=256, N=256, K=128, kernel="v3")
    case("v3 NT 256x256x128", M=256, N=256, K=128, ty=True, kernel="v3")
    case("v3 TN 256x256x128", M=256, N=256, K=128, tx=True, kernel="v3")
    case("v3 NN 1000x1000x1000", M=1000, N=1000, K=1000, kernel="v3")
    case("v3 NN 333x200x44 ragged", M=333, N=200, K=44, kernel="v3")
    case("v3 NN batched 8x1024x96x768 bB bn128", M=1024, N=96, K=768, nb=8, b_b=True, kernel="v3", bn=128)
    case("v3 NN 4096^3 accuracy U[0,1)", M=4096, N=4096, K=4096, kernel="v3",
         a=np.random.default_rng(5).random((1, 4096, 4096), dtype=np.float32), b=np.random.default_rng(6).random((1, 4096, 4096), dtype=np.float32))
if which in ("all", "v4"):
    for bn in (128, 64, 192, 256):
        case(f"v4 NN 256x256x128 bn{bn}", M=256, N=256, K=128, kernel="v4", bn=bn)
        case(f"v4 NT 256x256x128 bn{bn}", M=256, N=256, K=128, ty=True, kernel="v4", bn=bn)
        case(f"v4 TN 256x256x128 bn{bn}", M=256, N=256, K=128, tx=True, kernel="v4", bn=bn)
        case(f"v4 TT 256x256x128 bn{bn}", M=256, N=256, K=128, tx=True, ty=True, kernel="v4", bn=bn)
    case("v4 NN 1000x1000x1000", M=1000, N=1000, K=1000, kernel="v4")
    case("v4 TN 1000x1000x1000", M=1000, N=1000, K=1000, tx=True, kernel="v4")
    case("v4 NN 333x200x44 ragged", M=333, N=200, K=44, kernel="v4")
    case("v4 TN 333x200x44 ragged", M=333, N=200, K=44, tx=True, kernel="v4")
    case("v4 NN batched 8x1024x96x768 bB", M=1024, N=96, K=768, nb=8, b_b=True, kernel="v4")
    case("v4 NN batched 96x1024x64x1024 bn64", M=1024, N=64, K=1024, nb=96, kernel="v4", bn=64)
    case("v4 TN batched 96x1024x64x1024 bn64", M=1024, N=64, K=1024, nb=96, tx=True, kernel="v4", bn=64)
    case("v4 NN 4096^3 accuracy U[0,1)", M=4096, N=4096, K=4096, kernel="v4",
         a=np.random.default_rng(5).random((1, 4096, 4096), dtype=np.float32), b=np.random.default_rng(6).random((1, 4096, 4096), dtype=np.float32))
    case("v4 TN 768x3072x8192 accuracy U[0,1)", M=768, N=3072, K=8192, tx=True, kernel="v4",
         a=np.random.default_rng(5).random((1, 8192, 768), dtype=np.float32), b=np.random.default_rng(6).random((1, 8192, 3072), dtype=np.float32))
if which in ("all", "sk"):
    # stream-K: ragged last wave, tiles cut between two CTAs; every operand-major combination, ragged edges, batch,
    # K > 1024 (chunked accumulation inside each segment), then the launch repeated (flags must re-arm themselves)
    case("sk NN 8192x768x768 bn128", M=8192, N=768, K=768, kernel="v4", bn=128, streamk=True)
    case("sk NT 8192x768x768 bn192", M=8192, N=768, K=768, ty=True, kernel="v4", bn=192, streamk=True)
    case("sk TN 8192x768x768 bn128", M=8192, N=768, K=768, tx=True, kernel="v4", bn=128, streamk=True)
    case("sk NN 8192x768x3072 bn128", M=8192, N=768, K=3072, kernel="v4", bn=128, streamk=True)
    case("sk NN 5000x1000x1000 ragged", M=5000, N=1000, K=1000, kernel="v4", bn=128, streamk=True)
    case("sk NN b96 1024x64x1024 bn64", M=1024, N=64, K=1024, nb=96, kernel="v4", bn=64, streamk=True)
    case("sk TN b96 1024x64x1024 bn64", M=1024, N=64, K=1024, nb=96, tx=True, kernel="v4", bn=64, streamk=True)
    case("sk NN 4096^3 U[0,1) bn128", M=4096, N=4096, K=4096, kernel="v4", bn=128, streamk=True,
         a=np.random.default_rng(5).random((1, 4096, 4096), dtype=np.float32), b=np.random.default_rng(6).random((1, 4096, 4096), dtype=np.float32))
    case("sk NN 8192x768x768 x20 launches", M=8192, N=768, K=768, kernel="v4", bn=128, streamk=True, time_it=20)
    shapes = {"4096^3": dict(M=4096, N=4096, K=4096), "8192x768x768": dict(M=8192, N=768, K=768),
              "8192x3072x768": dict(M=8192, N=3072, K=768), "8192x768x3072": dict(M=8192, N=768, K=3072),
              "b96 QK^T 1024x1024x64": dict(M=1024, N=1024, K=64, nb=96), "b96 AV 1024x64x1024": dict(M=1024, N=64, K=1024, nb=96)}
    for name, kw in shapes.items():
        for bn in (64, 128, 192):
            if bn <= kw["N"]:
                for sk in (False, True):
                    case(f"perf sk={int(sk)} bn{bn} {name}", time_it=10, bn=bn, kernel="v4", streamk=sk, **kw)
if which in ("all", "perf4"):
    shapes = {"4096^3": dict(M=4096, N=4096, K=4096), "8192x768x768": dict(M=8192, N=768, K=768),
              "NT 8192x768x768": dict(M=8192, N=768, K=768, ty=True),
              "8192x3072x768": dict(M=8192, N=3072, K=768), "8192x768x3072": dict(M=8192, N=768, K=3072),
              "TN 768x3072x8192": dict(M=768, N=3072, K=8192, tx=True), "TN b4 768x768x2048": dict(M=768, N=768, K=2048, nb=4, tx=True),
              "b96 QK^T 1024x1024x64": dict(M=1024, N=1024, K=64, nb=96), "b96 AV 1024x64x1024": dict(M=1024, N=64, K=1024, nb=96),
              "b96 TN 64x1024x1024": dict(M=64, N=1024, K=1024, nb=96, tx=True)}
    for name, kw in shapes.items():
        for bn in (64, 128, 192, 256):
            if bn <= kw["N"]:
                case(f"perf v4 bn{bn} {name}", time_it=10, bn=bn, kernel="v4", **kw)
if which in ("all", "perf"):
    shapes = {"4096^3": dict(M=4096, N=4096, K=4096), "8192x768x768": dict(M=8192, N=768, K=768),
              "NT 8192x768x768": dict(M=8192, N=768, K=768, ty=True),
              "8192x3072x768": dict(M=8192, N=3072, K=768), "8192x768x3072": dict(M=8192, N=768, K=3072),
              "TN 768x3072x8192": dict(M=768, N=3072, K=8192, tx=True), "TN b4 768x768x2048": dict(M=768, N=768, K=2048, nb=4, tx=True),
              "b96 QK^T 1024x1024x64": dict(M=1024, N=1024, K=64, nb=96), "b96 AV 1024x64x1024": dict(M=1024, N=64, K=1024, nb=96),
              "b96 TN 64x1024x1024": dict(M=64, N=1024, K=1024, nb=96, tx=True)}
    for name, kw in shapes.items():
        for bn in (64, 128, 192, 256):
            if bn <= kw["N"]:
                case(f"perf v2 bn{bn} {name}", time_it=10, bn=bn, kernel="v2", **kw)
        if kw["M"] >= 512 and kw["N"] >= 256:
            for bn in (128, 256):
                case(f"perf v3 bn{bn} {name}", time_it=10, bn=bn, kernel="v3", **kw)
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/tc_probe_{which}.json", "w") as f:
    json.dump(results, f, indent=1)
