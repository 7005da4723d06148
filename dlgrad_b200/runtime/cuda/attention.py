"""Fused attention scores on Device.CUDA:  P = softmax(masked_fill((q @ k^T) * scale, mask, fill), dim=-1).

The reference writes every link of that chain to memory (examples/gpt.py:49-54 -> MATMUL, MUL, MASKED_FILL and the
five softmax kernels).  Here `q @ k.transpose(-2, -1)` of attention-shaped operands is DEFERRED (Buffer pending kind
"mm"); if the next consumers are `* scalar`, `masked_fill(mask2d, v)` and `softmax(dim=-1)`, the softmax dispatch
lands here and ONE tensor-core kernel (codegen/cuda_tcgen05.attn_scores_tc) produces P without S = q @ k^T ever
existing in HBM.  Any other consumer materialises the deferred product with the ordinary GEMM, so values and
semantics are those of the eager chain.  DLGRAD_FUSED_SCORES=0 / DLGRAD_FUSED_ATTN_BWD=0 disable the forward /
backward deferral.
"""
from __future__ import annotations

import math
import os

from dlgrad_b200.buffer import Buffer
from dlgrad_b200.codegen import cuda_tcgen05 as tc
from dlgrad_b200.helpers import prod_
from dlgrad_b200.runtime.cuda import compiler, shim
from dlgrad_b200.runtime.cuda import profile as _profile

import weakref

ENABLED = os.environ.get("DLGRAD_FUSED_SCORES", "1") != "0"
SPARSE = os.environ.get("DLGRAD_ATTN_SPARSE_GEMM", "1") != "0"
# P and dS produced by the fused kernels carry their live-tile table: the GEMMs that read them (P @ v, P^T @ dO,
# dS @ k, dS^T @ q) skip the k-blocks of tiles the mask kills entirely (codegen/cuda_tcgen05.matmul_tc_ts, ksparse)
_tags: dict = {}             # id(Buffer) -> table (Buffer.__eq__ is the elementwise op, so no dict keyed by the object itself)
_fresh: dict = {}            # device address of a just produced P / dS -> (owner of the table, table pointer, Tq, Tk)


def mask_words(Tq: int, Tk: int) -> int:
    return Tq * Tk // 32 + (Tq // 128) * (Tk // 128) + Tq + 1


def claim(buf: Buffer) -> None:
    """Attach the live-tile table of a freshly produced P / dS to the Buffer that now owns the pointer."""
    if _fresh:
        info = _fresh.pop(shim.address(buf._ptr), None)
        _fresh.clear()
        if info is not None:
            _tags[id(buf)] = info
            weakref.finalize(buf, _tags.pop, id(buf), None)


def tag_of(buf: Buffer):
    return _tags.get(id(buf)) if SPARSE else None

# The backward twin (attn_dscores_tc: dS from dO, v, P in one kernel, dP never written): 173 us + 25 us of helper
# kernels against 247 us for GEMM + fused softmax backward at C5; GPT step 34.7 -> 33.9 ms on the same box.
BWD_ENABLED = os.environ.get("DLGRAD_FUSED_ATTN_BWD", "1") != "0"
MIN_SCORES = 1 << 22          # elements of (batch, Tq, Tk) below which the eager chain is just as good
_plans: dict = {}
_digests: dict = {}           # (content tag of the mask, Tq, Tk) -> mask_bits output; a tagged mask (Buffer._const) is a
                              # pure function of shapes and scalars, so its digest is computed once, not per layer and step


def _mask_digest(mask: Buffer, Tq: int, Tk: int):
    from dlgrad_b200.codegen import cuda_kernel as ck
    from dlgrad_b200.runtime.cuda import ops as _ops
    bkey = ("mask_bits", Tq, Tk)
    bplan = _plans.get(bkey)
    if bplan is None:
        bplan = _plans[bkey] = _ops.Plan(ck.mask_bits(Tq, Tk))
    tag = mask._const
    if tag is None:
        return bplan.run(_ops._dev_ptr(mask))
    bits = _digests.get((tag, Tq, Tk))
    if bits is None:
        if len(_digests) >= 64:
            _digests.clear()
        bits = _digests[(tag, Tq, Tk)] = bplan.run(_ops._dev_ptr(mask))
    return bits


def eligible(x: Buffer, y: Buffer) -> bool:
    """`x @ y` with y a deferred transpose of k: (.., Tq, D) @ (.., Tk, D)^T, same batch dims, tile-sized."""
    if not ENABLED or y._pending is None or y._pending[0] != "tr" or x._pending is not None:
        return False
    k = y._pending[1]
    xs, ks = x.metadata.shape, k.metadata.shape
    if len(xs) < 3 or len(xs) != len(ks) or xs[:-2] != ks[:-2] or xs[-1] != ks[-1] or k._pending is not None:
        return False
    if y.metadata.shape != ks[:-2] + (ks[-1], ks[-2]):
        return False
    Tq, D, Tk = xs[-2], xs[-1], ks[-2]
    return (Tq % 128 == 0 and Tk % 128 == 0 and Tk <= 2048 and D in (32, 64) and x.aligned and k.aligned
            and prod_(xs[:-2]) * Tq * Tk >= MIN_SCORES and x.scalar is None and k.scalar is None)


def chain_matches(chain: list, out_shape: tuple) -> tuple | None:
    """[("scale", _, s), ("mfill", _, mask, v)] with a mask that is one (Tq, Tk) plane broadcast over the batch."""
    if len(chain) != 2 or chain[0][0] != "scale" or chain[1][0] != "mfill":
        return None
    mask, val = chain[1][2], chain[1][3]
    ms = tuple(mask.metadata.shape)
    while len(ms) > 2 and ms[0] == 1:
        ms = ms[1:]
    if ms != tuple(out_shape[-2:]) or not mask.aligned or mask.scalar is not None:
        return None
    if not float(chain[0][2]) > 0.0:               # the kernel takes row maxima on the raw scores
        return None
    return float(chain[0][2]), mask, float(val)


def scores(q: Buffer, k: Buffer, scale: float, mask: Buffer, fill: float):
    """Launch the fused kernel; returns the owned device pointer of P (batch.., Tq, Tk)."""
    from dlgrad_b200.runtime.cuda import matmul_tc, ops as _ops
    xs, ks = q.metadata.shape, k.metadata.shape
    nb, Tq, Tk, D = prod_(xs[:-2]), xs[-2], ks[-2], xs[-1]
    key = (nb, Tq, Tk, D, scale, fill)
    plan = _plans.get(key)
    if plan is None:
        kern = tc.attn_scores_tc(Tq, Tk, D, nb, scale, fill)
        fn = compiler.get_function(kern.name, kern.source)
        da, db, dc = tc.tmap_descs(Tq, Tk, D, nb, False, False, False, False, kern.bn)
        p = shim.lib.dlg_mm_plan_create(fn, kern.grid[0], 1, 1, kern.threads, kern.smem, 1, kern.out_elems * 4,
                                        matmul_tc._tmap(da), matmul_tc._tmap(db), matmul_tc._tmap(dc))
        if p == shim.NULL:
            raise RuntimeError("dlg_mm_plan_create failed: " + shim.last_error())
        plan = _plans[key] = (p, kern.out_elems * 4)
    p, nbytes = plan
    bits = _mask_digest(mask, Tq, Tk)                 # the same (Tq, Tk) plane for every head: 1 bit per entry
    pq, pk, pm = _ops._dev_ptr(q), _ops._dev_ptr(k), bits
    timer = _profile.matmul_timer
    t0 = timer.start() if timer is not None else None
    out = shim.lib.dlg_mm_plan_run_aux(p, pq, pk, pm, shim.NULL)
    if timer is not None:
        timer.stop(t0, 2.0 * nb * Tq * Tk * D, ("attn_scores", Tq, Tk, D, nb))
    if out == shim.NULL:
        raise RuntimeError("fused attention-scores launch failed: " + shim.last_error())
    res = shim.own(out, nbytes)
    if math.isinf(fill) and fill < 0:
        # only a -inf fill makes the masked probabilities EXACTLY zero, which is what lets the GEMMs that read P skip
        # the tiles the mask kills; with a finite fill (masked_fill(mask, -5.0)) P is dense and stays untagged
        _fresh[shim.address(res)] = (bits, bits + Tq * Tk // 32, Tq, Tk)
    return res


def dscores(dO: Buffer, v: Buffer, O: Buffer, P: Buffer, mask: Buffer, scale: float):
    """dS = masked_fill(P * (dO @ v^T - rowsum(dO * O)), mask, 0) * scale in one kernel; returns the owned pointer."""
    from dlgrad_b200.codegen import cuda_kernel as ck
    from dlgrad_b200.runtime.cuda import matmul_tc, ops as _ops
    xs, vs = dO.metadata.shape, v.metadata.shape
    nb, Tq, Tk, D = prod_(xs[:-2]), xs[-2], vs[-2], xs[-1]
    key = ("bwd", nb, Tq, Tk, D, scale)
    plan = _plans.get(key)
    if plan is None:
        kern = tc.attn_dscores_tc(Tq, Tk, D, nb, scale)
        fn = compiler.get_function(kern.name, kern.source)
        da, db, dc = tc.tmap_descs(Tq, Tk, D, nb, False, False, False, False, kern.bn)
        p = shim.lib.dlg_mm_plan_create(fn, kern.grid[0], 1, 1, kern.threads, kern.smem, 1, kern.out_elems * 4,
                                        matmul_tc._tmap(da), matmul_tc._tmap(db), matmul_tc._tmap(dc))
        if p == shim.NULL:
            raise RuntimeError("dlg_mm_plan_create failed: " + shim.last_error())
        plan = _plans[key] = (p, kern.out_elems * 4, _ops.Plan(ck.rowdot(nb * Tq, D)), _ops.Plan(ck.mask_bits(Tq, Tk)))
    p, nbytes, dot_plan, bits_plan = plan
    aux = shim.alloc((nb * Tq + mask_words(Tq, Tk)) * 4)     # [rowsum(dO * O) | mask bits, live tiles, live rows]; it travels with dS
    dot_plan.run_into(aux, _ops._dev_ptr(dO), _ops._dev_ptr(O))
    if mask._const is not None:                      # digest of a constant mask: computed once, copied (a few hundred KB)
        shim.check(shim.lib.dlg_d2d(aux + nb * Tq, _mask_digest(mask, Tq, Tk), mask_words(Tq, Tk) * 4), "dlg_d2d")
    else:
        bits_plan.run_into(aux + nb * Tq, _ops._dev_ptr(mask))
    pd, pv, pp = _ops._dev_ptr(dO), _ops._dev_ptr(v), _ops._dev_ptr(P)
    timer = _profile.matmul_timer
    t0 = timer.start() if timer is not None else None
    out = shim.lib.dlg_mm_plan_run_aux(p, pd, pv, pp, aux)
    if timer is not None:
        timer.stop(t0, 2.0 * nb * Tq * Tk * D, ("attn_dscores", Tq, Tk, D, nb))
    if out == shim.NULL:
        raise RuntimeError("fused attention-backward launch failed: " + shim.last_error())
    res = shim.own(out, nbytes)
    _fresh[shim.address(res)] = (aux, aux + nb * Tq + Tq * Tk // 32, Tq, Tk)
    return res


def bwd_chain_matches(chain: list, out_shape: tuple) -> tuple | None:
    """[("mfill", _, mask, 0.0), ("scale", _, s)] — the backward of scale -> masked_fill in reverse order."""
    if len(chain) != 2 or chain[0][0] != "mfill" or chain[1][0] != "scale" or float(chain[0][3]) != 0.0:
        return None
    mask = chain[0][2]
    ms = tuple(mask.metadata.shape)
    while len(ms) > 2 and ms[0] == 1:
        ms = ms[1:]
    if ms != tuple(out_shape[-2:]) or not mask.aligned or mask.scalar is not None:
        return None
    return mask, float(chain[1][2])
