"""Plans for the tcgen05 / TMEM / TMA 3xTF32 matmul (codegen/cuda_tcgen05.py).

`try_build` returns a MatmulPlan when the shape qualifies for TMA (16-byte pitches, operands either
fully batched or fully broadcast) and is big enough to be worth a 128-row tile; otherwise None and the
exact-fp32 SIMT kernel runs.  DLGRAD_TC_PASSES=1 selects single-pass TF32 (debug only: ~1e-3 accuracy).
"""
from __future__ import annotations

import os

from dlgrad_b200.codegen import cuda_tcgen05 as tc
from dlgrad_b200.helpers import prod_
from dlgrad_b200.runtime.cuda import compiler, shim

ffi, lib, NULL = shim.ffi, shim.lib, shim.NULL
PASSES = int(os.environ.get("DLGRAD_TC_PASSES", "3"))
ENABLED = os.environ.get("DLGRAD_TC", "1") != "0"
MIN_FLOPS = 1 << 22
SPLIT_WARPS = int(os.environ.get("DLGRAD_TC_SPLIT_WARPS", "4"))
# auto = v4 | v4 = persistent, A operand in TMEM | v3 = CTA pair | v2 = persistent, both operands in smem | v1 = one tile per CTA
KERNEL = os.environ.get("DLGRAD_TC_KERNEL", "auto")


def _tmap(d: dict):
    t = ffi.new("dlg_tmap_desc *")
    t.rank = d["rank"]
    for i, v in enumerate(d["dims"]):
        t.dims[i] = v
    for i, v in enumerate(d["strides"]):
        t.strides_bytes[i] = v
    for i, v in enumerate(d["box"]):
        t.box[i] = v
    t.swizzle = d["swizzle"]
    return t


def qualifies(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, force: bool = False) -> tuple | None:
    if not (ENABLED and aligned):
        return None
    nb = prod_(batch)
    full = (M * K * batch[1], M * K) if batch[1] > 1 else (M * K, 0)

    def mode(bs, mat):
        # fully batched (contiguous over the flattened batch) or fully broadcast
        if nb == 1 or bs == (0, 0):
            return "bcast" if nb > 1 else "full"
        want = (mat * batch[1] if batch[0] > 1 else 0, mat if batch[1] > 1 else 0)
        return "full" if tuple(bs) == want else None
    am, bm = mode(a_bs, M * K), mode(b_bs, K * N)
    if am is None or bm is None:
        return None
    a_mn, b_mn = bool(tx), not bool(ty)
    if (M if a_mn else K) % 4 or (N if b_mn else K) % 4 or N % 4:
        return None                                   # TMA needs 16-byte global pitches
    if not force and (K < 8 or N < 8 or M < 16 or 2.0 * M * N * K * nb < MIN_FLOPS):
        return None
    _ = full
    return a_mn, b_mn, am == "bcast", bm == "bcast", nb


def try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, passes: int | None = None, bn: int | None = None,
              stages: int | None = None, hi_inplace: bool = False, force: bool = False, kernel: str | None = None,
              split_warps: int | None = None, bias: bool = False, streamk: bool | None = None, ksparse: str | None = None,
              a_relu: bool = False, b_relu: bool = False, splitk: int = 0, relu_mask: bool = False, radd: bool = False):
    """`bias=True` builds the variant whose epilogue adds a length-N row (only the TS kernel has one; None otherwise)."""
    q = qualifies(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, force)
    if q is None:
        return None
    a_mn, b_mn, a_bc, b_bc, nb = q
    passes = passes or PASSES
    cluster = 1
    b_rows = None
    flags = None
    which = kernel or KERNEL
    if which == "auto":
        # A-in-TMEM beats both shared-memory-operand kernels on every shape of scripts/tc_probe.py perf/perf4
        # (e.g. 8192x768x768: 210 vs 156 TFLOP/s, 4096^3: 244 vs 235, TN 768x3072x8192: 240 vs 200)
        which = "v4" if passes == 3 else "v2"
    if (bias or ksparse or a_relu or b_relu or splitk or relu_mask or radd) and not (which == "v4" and passes == 3):
        return None
    if which == "v3" and passes == 3 and M > 128 and (bn or 256) % 64 == 0:
        k = tc.matmul_tc_pair(M, N, K, nb, a_mn, b_mn, a_bc, b_bc, bn, stages)
        cluster, b_rows = 2, k.bn // 2
    elif which == "v4" and passes == 3:
        k = tc.matmul_tc_ts(M, N, K, nb, a_mn, b_mn, a_bc, b_bc, bn, stages, None, bias, None, False, ksparse, a_relu, b_relu, splitk, relu_mask,
                            0, radd)
        sk = (tc.wants_streamk(M, N, K, nb, k.bn) if streamk is None else streamk) and not (ksparse or a_relu or b_relu or splitk or relu_mask or radd)
        if sk:
            k = tc.matmul_tc_ts(M, N, K, nb, a_mn, b_mn, a_bc, b_bc, k.bn, stages, None, bias, None, True)
            ntiles = -(-M // 128) * -(-N // k.bn) * nb
            flags = shim.calloc(ntiles * 16)              # uint32[ntiles][4], re-armed by the kernel itself
    elif which == "v1":
        k = tc.matmul_tc(M, N, K, nb, a_mn, b_mn, a_bc, b_bc, passes, bn, stages, hi_inplace)
    else:
        k = tc.matmul_tc_persistent(M, N, K, nb, a_mn, b_mn, a_bc, b_bc, passes, bn, stages, None,
                                    split_warps or SPLIT_WARPS)
    shim.init()
    fn = compiler.get_function(k.name, k.source)
    da, db, dc = tc.tmap_descs(M, N, K, nb, a_mn, b_mn, a_bc, b_bc, k.bn, b_rows)
    ta, tb, tcd = _tmap(da), _tmap(db), _tmap(dc)
    plan = lib.dlg_mm_plan_create(fn, k.grid[0], k.grid[1], k.grid[2], k.threads, k.smem, cluster, k.out_elems * 4,
                                  ta, tb, tcd)
    if plan == NULL:
        raise RuntimeError("dlg_mm_plan_create failed: " + shim.last_error())
    nbytes = k.out_elems * 4

    if splitk > 1:
        # C of the kernel is the scratch of partial sums; the real output and the tile counters ride as aux operands
        counters = shim.calloc(-(-M // 128) * -(-N // k.bn) * 16)
        final_bytes = M * N * 4

        def run(a, b):
            final = shim.alloc(final_bytes)
            scratch = lib.dlg_mm_plan_run_aux(plan, a, b, final, counters)
            if scratch == NULL:
                raise RuntimeError("tensor-core matmul launch failed: " + shim.last_error())
            lib.dlg_free(scratch)               # stream-ordered: whoever gets this block next runs after the kernel
            return final
    elif bias or flags is not None or ksparse or relu_mask or radd:
        fl = NULL if flags is None else flags

        def run(a, b, aux=NULL, aux1=None):
            out = lib.dlg_mm_plan_run_aux(plan, a, b, aux, fl if aux1 is None else aux1)
            if out == NULL:
                raise RuntimeError("tensor-core matmul launch failed: " + shim.last_error())
            return shim.own(out, nbytes)
    else:
        def run(a, b):
            out = lib.dlg_mm_plan_run(plan, a, b)
            if out == NULL:
                raise RuntimeError("tensor-core matmul launch failed: " + shim.last_error())
            return shim.own(out, nbytes)
    from dlgrad_b200.runtime.cuda.matmul import MatmulPlan
    return MatmulPlan(run, 2.0 * nb * M * N * K, (f"tc{passes}" + (f"_sp{ksparse}" if ksparse else "") + ("_radd" if radd else ""), M, N, K, nb))


def build_grouped(mode: str, G: int, M: int, N: int, K: int, a_mn: bool, b_mn: bool, radd: bool = False):
    """Plans for the grouped forms of matmul_tc_ts (separate tensor maps per member, nothing is concatenated):
    mode "n": [A @ B_g for g < G] (K = the shared inner size) -> run(a, [b_g]) returns G output pointers;
    mode "k": sum_g A_g @ B_g (K = inner size of ONE pair)      -> run([a_g], [b_g], r=None) returns one pointer
    (`radd`: + r in the epilogue);
    mode "m": [A_g^T @ B for g < G] as split-K (`radd` carries the number of K slices S; A_g stored [K][M], B stored
    [K][N], K = the full inner size)                             -> run([a_g], b) returns G output pointers.  None when the tensor-core path is disabled or the shape does not qualify."""
    if not ENABLED or PASSES != 3 or KERNEL not in ("auto", "v4"):
        return None
    if (M if a_mn else K) % 4 or (N if b_mn else K) % 4 or N % 4 or K < 8 or N < 8 or M < 128:
        return None
    from dlgrad_b200.runtime.cuda.attention import _tc_plan, _tc_run
    if mode == "m":
        return _build_grouped_splitk(G, M, N, K, radd)
    if mode == "n":
        k = tc.matmul_tc_ts(M, N, K, G, a_mn, b_mn, True, False, group=("n", G))
        da, db, dc = tc.tmap_descs(M, N, K, 1, a_mn, b_mn, False, False, k.bn)
        maps = [da, db, dc] + [db] * (G - 1) + [dc] * (G - 1)
    else:
        if K % tc.BK:
            return None
        k = tc.matmul_tc_ts(M, N, G * K, 1, a_mn, b_mn, False, False, radd=radd, group=("k", G),
                            bn=(int(os.environ.get("DLGRAD_RMASK_BN", "128")) if radd and N >= 128
                                else (int(os.environ["DLGRAD_GROUPK_BN"]) if "DLGRAD_GROUPK_BN" in os.environ else None)))
        da, db, dc = tc.tmap_descs(M, N, K, 1, a_mn, b_mn, False, False, k.bn)
        maps = [da, db, dc] + [db] * (G - 1) + [da] * (G - 1)
    shim.init()
    plan = _tc_plan(k, maps, 3)
    nbytes = M * N * 4
    NULLP = shim.NULL
    what = f"grouped matmul ({mode}, {G})"

    if mode == "n":
        def run(a, bs):
            outs = [shim.alloc(nbytes) for _ in range(G)]
            _tc_run(plan, [a, bs[0], outs[0]] + list(bs[1:]) + outs[1:], [outs[0], NULLP, NULLP], what)
            return outs
    else:
        def run(as_, bs, r=NULLP):
            out = shim.alloc(nbytes)
            _tc_run(plan, [as_[0], bs[0], out] + list(bs[1:]) + list(as_[1:]), [out, r, NULLP], what)
            return out
    from dlgrad_b200.runtime.cuda.matmul import MatmulPlan
    return MatmulPlan(run, 2.0 * G * M * N * K, (f"tc{PASSES}_group_{mode}{G}" + ("_radd" if radd else ""), M, N, K, G))


def _build_grouped_splitk(G: int, M: int, N: int, K: int, S: int):
    """G weight gradients dW_g = A_g^T @ B over the same B (x) as ONE split-K launch of G x S work items per output tile
    (matmul_tc_ts group ("m", G)); S slices of K / S each."""
    from dlgrad_b200.runtime.cuda.attention import _tc_plan, _tc_run
    ks = K // S
    if S < 2 or K % S or ks > 2048 or ks % tc.BK or M % 4 or N % 4:
        return None
    k = tc.matmul_tc_ts(M, N, ks, G * S, True, True, False, False, bn=128 if N >= 128 else None, splitk=S, group=("m", G))
    da, db, _ = tc.tmap_descs(M, N, ks, S, True, True, False, False, k.bn)
    dc = dict(rank=3, dims=(N, M, G * S), strides=(N * 4, M * N * 4), box=(32, 32, 1), swizzle=3)
    shim.init()
    plan = _tc_plan(k, [da, db, dc] + [da] * (G - 1), 3 + (G - 1))
    counters = shim.calloc(G * -(-M // 128) * -(-N // k.bn) * 16)
    nbytes, scratch_bytes = M * N * 4, G * S * M * N * 4

    def run(as_, b):
        finals = [shim.alloc(nbytes) for _ in range(G)]
        scratch = shim.alloc(scratch_bytes)            # freed on return: whoever gets the block next runs after this launch
        _tc_run(plan, [as_[0], b, scratch] + list(as_[1:]), [scratch, finals[0], counters] + finals[1:], f"grouped split-K ({G} x {S})")
        return finals
    from dlgrad_b200.runtime.cuda.matmul import MatmulPlan
    return MatmulPlan(run, 2.0 * G * M * N * K, (f"tc{PASSES}_splitk_group{G}", M, N, K, S))
