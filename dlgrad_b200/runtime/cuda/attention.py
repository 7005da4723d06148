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


def _stored(b: Buffer) -> bool:
    """materialised, or a deferred head split / merge of a materialised buffer (read in place through strided tensor maps)"""
    pend = b._pending
    if pend is not None and pend[0] == "hsplit":
        src = pend[1]
        if src._pending is not None and src._pending[0] == "lin":
            _ = src.ptr        # a deferred projection under the head split: it runs now, as one launch with its siblings
        return src._pending is None
    return pend is None


def eligible(x: Buffer, y: Buffer) -> bool:
    """`x @ y` with y a deferred transpose of k: (.., Tq, D) @ (.., Tk, D)^T, same batch dims, tile-sized."""
    if not ENABLED or y._pending is None or y._pending[0] != "tr" or not _stored(x):
        return False
    k = y._pending[1]
    xs, ks = x.metadata.shape, k.metadata.shape
    if len(xs) < 3 or len(xs) != len(ks) or xs[:-2] != ks[:-2] or xs[-1] != ks[-1] or not _stored(k):
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


# =====================================================================================================================
# Flash-style attention (codegen/cuda_flash.py): P = softmax(masked_fill((q k^T) * scale, mask, fill)) stays DEFERRED
# (Buffer pending kind "attnP").  `P @ v` then runs ONE kernel that never writes P (flash_forward) and leaves the row
# statistics (m, 1 / l) behind; the backward of that product recomputes P per tile from q, k and the statistics:
# dv = P^T dO (flash_dv) and dS = P (dO v^T - rowsum(dO O)) scale (flash_ds).  dS is still written once and read by the
# two tile-skipping GEMMs dq = dS k, dk = dS^T q.  Whoever reads P itself (att.numpy(), another op) gets it from
# attn_scores_tc as before — same values, the reference's semantics (examples/gpt.py:49-55) untouched.
# DLGRAD_FLASH_ATTN=0 switches the deferral off (the round-1 kernels then run).
# =====================================================================================================================
FLASH = os.environ.get("DLGRAD_FLASH_ATTN", "1") != "0"
FLASH_TAU = float(os.environ.get("DLGRAD_FLASH_TAU", "8"))     # see cuda_flash.flash_attn: 0 forces the O-rescale path
_flash: dict = {}            # id(P Buffer) -> record of the forward that consumed it
_ds_private = [0]            # > 0 while ds_gemm materialises a dS that only the tile-skipping GEMMs will read (flash_ds: lazy_zero)
_tc_plans: dict = {}


class _FlashRecord:
    __slots__ = ("q", "k", "v", "scale", "mask", "fill", "stats", "bits", "geom")


def softmax_deferred(x: Buffer, dim: int):
    """`x.softmax(dim)` where x is the pending chain (q @ k^T) * scale -> masked_fill(mask2d, fill) over attention-shaped
    operands: returns P as a deferred Buffer, or None when the pattern does not apply."""
    from dlgrad_b200.runtime.cuda import ops as _ops
    shape = x.metadata.shape
    if not (FLASH and ENABLED) or x._pending is None or dim % len(shape) != len(shape) - 1:
        return None
    chain, base = _ops._peel_chain(x)
    bp = base._pending
    if bp is None or bp[0] != "mm" or len(bp) != 3:
        return None
    m = chain_matches(chain, shape)
    q, kt = bp[1], bp[2]
    if m is None or kt._pending is None or kt._pending[0] != "tr":
        return None
    return Buffer._deferred(("attnP", q, kt._pending[1], m[0], m[1], m[2]), shape, x.metadata.dtype)


def materialize_p(buf: Buffer):
    """Somebody reads P itself: produce it with the fused scores kernel (round 1), exactly as before."""
    _, q, k, scale, mask, fill = buf._pending
    ptr = scores(q, k, scale, mask, fill)
    buf.ptr = ptr
    claim(buf)


def _geom(q: Buffer, k: Buffer):
    xs, ks = q.metadata.shape, k.metadata.shape
    Tq, D, Tk = xs[-2], xs[-1], ks[-2]
    H = xs[-3] if len(xs) >= 4 else 1
    nb = prod_(xs[:-2])
    return nb, H, nb // H, Tq, Tk, D


def _strides(T: int, D: int, H: int, split: bool = False):
    """element strides (row, head, batch) of a (B, H, T, D) operand: contiguous, or stored as (B, T, H, D) (`split`: a
    deferred head split — the projection's own layout)"""
    return (H * D, D, T * H * D) if split else (D, T * D, H * T * D)


def _resolve(b: Buffer):
    """(buffer that holds the data, stored-as-(B,T,H,D) flag) of an operand that is materialised or a deferred head split"""
    if b._pending is not None and b._pending[0] == "hsplit" and b._pending[1]._pending is None:
        return b._pending[1], True
    return b, False


def _wrap(ptr, shape: tuple, split: bool) -> Buffer:
    """Buffer of logical shape (B, H, T, D) over `ptr`; when the kernel wrote the (B, T, H, D) layout it is handed on as a
    deferred head split of that storage, which the merge (`transpose(1, 2)`) that follows cancels."""
    from dlgrad_b200.device import Device
    if not split:
        return Buffer(ptr, shape, Device.CUDA)
    B, H, T, D = shape
    return Buffer._deferred(("hsplit", Buffer(ptr, (B, T, H, D), Device.CUDA)), shape)


def _tc_plan(kern, maps: list, nptrs: int):
    fn = compiler.get_function(kern.name, kern.source)
    arr = shim.ffi.new("dlg_tmap_desc5[]", len(maps))
    for t, d in zip(arr, maps):
        t.rank = d["rank"]
        for i in range(5):
            t.dims[i] = d["dims"][i] if i < d["rank"] else 1
            t.box[i] = d["box"][i] if i < d["rank"] else 1
        for i in range(4):
            t.strides_bytes[i] = d["strides"][i] if i < d["rank"] - 1 else 0
        t.swizzle = d["swizzle"]
    p = shim.lib.dlg_tc_plan_create(fn, kern.grid[0], kern.threads, kern.smem, len(maps), arr, nptrs)
    if p == shim.NULL:
        raise RuntimeError("dlg_tc_plan_create failed: " + shim.last_error())
    return p


def _tc_run(plan, bases: list, ptrs: list, what: str) -> None:
    b = shim.ffi.new("const void *[]", bases)
    p = shim.ffi.new("void *[]", ptrs)
    if shim.lib.dlg_tc_plan_run(plan, b, p) != 0:
        raise RuntimeError(f"{what} launch failed: " + shim.last_error())


def _timed(tag: tuple, flops: float, fn):
    timer = _profile.matmul_timer
    if timer is None:
        return fn()
    t0 = timer.start()
    out = fn()
    timer.stop(t0, flops, tag)
    return out


def flash_forward(P: Buffer, v: Buffer):
    """O = P @ v for a deferred P, without materialising P.  Returns O as a Buffer (batch.., Tq, D) — in the layout q came
    in: a deferred head split when q is one — or None when v does not fit the pattern (the caller then takes the ordinary
    route, which materialises P)."""
    from dlgrad_b200.codegen import cuda_flash as cf
    from dlgrad_b200.runtime.cuda import ops as _ops
    _, q, k, scale, mask, fill = P._pending
    nb, H, B, Tq, Tk, D = _geom(q, k)
    vs = v.metadata.shape
    if (not _stored(v) or v.scalar is not None or not v.aligned or tuple(vs) != tuple(k.metadata.shape)
            or not _stored(q) or not _stored(k)):
        return None
    (qb, qs), (kb, ks), (vb, vsplit) = _resolve(q), _resolve(k), _resolve(v)
    if len(q.metadata.shape) != 4:
        qs = ks = vsplit = False
    key = ("fwd", nb, H, Tq, Tk, D, scale, fill, FLASH_TAU, qs, ks, vsplit)
    plan = _tc_plans.get(key)
    if plan is None:
        kern = cf.flash_attn("fwd", Tq, Tk, D, H, nb, scale, fill, _strides(Tq, D, H, qs), FLASH_TAU)
        maps = [cf.tmap4(Tq, D, H, B, _strides(Tq, D, H, qs), 128), cf.tmap4(Tk, D, H, B, _strides(Tk, D, H, ks), 64),
                cf.tmap4(Tk, D, H, B, _strides(Tk, D, H, vsplit), 32, mn=True), cf.tmap4(Tq, D, H, B, _strides(Tq, D, H, qs), 32)]
        plan = _tc_plans[key] = _tc_plan(kern, maps, 3)
    bits = _mask_digest(mask, Tq, Tk)
    out = shim.alloc(nb * Tq * D * 4)
    stats = shim.alloc(nb * Tq * 8)
    pq, pk, pv = _ops._dev_ptr(qb), _ops._dev_ptr(kb), _ops._dev_ptr(vb)
    _timed(("attn_flash_fwd", Tq, Tk, D, nb), 4.0 * nb * Tq * Tk * D,
           lambda: _tc_run(plan, [pq, pk, pv, out], [out, stats, bits], "flash attention forward"))
    rec = _FlashRecord()
    rec.q, rec.k, rec.v, rec.scale, rec.mask, rec.fill, rec.stats, rec.bits = q, k, v, scale, mask, fill, stats, bits
    rec.geom = (nb, H, B, Tq, Tk, D)
    _flash[id(P)] = rec
    weakref.finalize(P, _flash.pop, id(P), None)
    return _wrap(out, tuple(q.metadata.shape), qs)


def flash_record(P: Buffer, v: Buffer | None = None):
    rec = _flash.get(id(P))
    if rec is None or (v is not None and rec.v is not v):
        return None
    return rec


def flash_dv(rec, dO: Buffer) -> Buffer:
    """dv = P^T @ dO with P recomputed per tile from q, k and the saved row statistics; dv comes out in v's layout."""
    from dlgrad_b200.codegen import cuda_flash as cf
    from dlgrad_b200.runtime.cuda import ops as _ops
    nb, H, B, Tq, Tk, D = rec.geom
    (qb, qs), (kb, ks), (_, vsplit), (db, dsplit) = _resolve(rec.q), _resolve(rec.k), _resolve(rec.v), _resolve(dO)
    if len(rec.q.metadata.shape) != 4:
        qs = ks = vsplit = dsplit = False
    key = ("dv", nb, H, Tq, Tk, D, rec.scale, rec.fill, qs, ks, vsplit, dsplit)
    plan = _tc_plans.get(key)
    if plan is None:
        kern = cf.flash_attn("dv", Tq, Tk, D, H, nb, rec.scale, rec.fill, _strides(Tk, D, H, vsplit))
        maps = [cf.tmap4(Tk, D, H, B, _strides(Tk, D, H, ks), 128), cf.tmap4(Tq, D, H, B, _strides(Tq, D, H, qs), 64),
                cf.tmap4(Tq, D, H, B, _strides(Tq, D, H, dsplit), 32, mn=True), cf.tmap4(Tk, D, H, B, _strides(Tk, D, H, vsplit), 32)]
        plan = _tc_plans[key] = _tc_plan(kern, maps, 3)
    out = shim.alloc(nb * Tk * D * 4)
    pk, pq, pd = _ops._dev_ptr(kb), _ops._dev_ptr(qb), _ops._dev_ptr(db)
    _timed(("attn_flash_dv", Tq, Tk, D, nb), 4.0 * nb * Tq * Tk * D,
           lambda: _tc_run(plan, [pk, pq, pd, out], [out, rec.stats, rec.bits], "flash attention dV"))
    return _wrap(out, tuple(rec.v.metadata.shape), vsplit)


def flash_ds(rec, dO: Buffer, O: Buffer, dS_owner: Buffer):
    """dS = masked_fill(P * (dO @ v^T - rowsum(dO * O)), mask, 0) * scale with P and dP recomputed per tile; returns the
    owned pointer and registers the live-tile table for the GEMMs that read dS."""
    from dlgrad_b200.codegen import cuda_flash as cf
    from dlgrad_b200.codegen import cuda_kernel as ck
    from dlgrad_b200.runtime.cuda import ops as _ops
    nb, H, B, Tq, Tk, D = rec.geom
    (qb, qs), (kb, ks), (vb, vsplit), (db, dsplit), (ob, osplit) = (_resolve(rec.q), _resolve(rec.k), _resolve(rec.v),
                                                                   _resolve(dO), _resolve(O))
    if len(rec.q.metadata.shape) != 4:
        qs = ks = vsplit = dsplit = osplit = False
    if dsplit != osplit:                       # rowsum(dO * O) walks both operands in storage order: they must agree
        db, dsplit, ob, osplit = dO, False, O, False          # (reading .ptr below materialises the deferred one)
    lazy = bool(_ds_private[0]) and SPARSE and Tk <= 2048
    key = ("ds", nb, H, Tq, Tk, D, rec.scale, rec.fill, qs, ks, vsplit, dsplit, lazy)
    ent = _tc_plans.get(key)
    if ent is None:
        kern = cf.flash_attn("ds", Tq, Tk, D, H, nb, rec.scale, rec.fill, delta_split=dsplit, lazy_zero=lazy)
        maps = [cf.tmap4(Tq, D, H, B, _strides(Tq, D, H, qs), 128), cf.tmap4(Tq, D, H, B, _strides(Tq, D, H, dsplit), 128),
                cf.tmap4(Tk, D, H, B, _strides(Tk, D, H, ks), 64), cf.tmap4(Tk, D, H, B, _strides(Tk, D, H, vsplit), 64),
                dict(rank=3, dims=(Tk, Tq, nb), strides=(Tk * 4, Tq * Tk * 4), box=(32, 32, 1), swizzle=3)]
        ent = _tc_plans[key] = (_tc_plan(kern, maps, 4), _ops.Plan(ck.rowdot(nb * Tq, D)))
    plan, dot_plan = ent
    delta = dot_plan.run(_ops._dev_ptr(db), _ops._dev_ptr(ob))       # rows in storage order: (b, h, t), or (b, t, h) when split
    out = shim.alloc(nb * Tq * Tk * 4)
    pq, pd, pk, pv = _ops._dev_ptr(qb), _ops._dev_ptr(db), _ops._dev_ptr(kb), _ops._dev_ptr(vb)
    _timed(("attn_flash_ds", Tq, Tk, D, nb), 4.0 * nb * Tq * Tk * D,
           lambda: _tc_run(plan, [pq, pd, pk, pv, out], [out, rec.stats, delta, rec.bits], "flash attention dS"))
    _fresh[shim.address(out)] = (rec.bits, rec.bits + Tq * Tk // 32, Tq, Tk)
    return out


def ds_gemm(dS: Buffer, other: Buffer, trans_x: bool):
    """dq = dS @ k (trans_x False) or dk = dS^T @ q (trans_x True) on the tile-skipping GEMM, reading k / q in the layout they
    are stored in and writing the result the same way: when `other` is a deferred head split — (B, T, h, D) storage, the
    projection's own layout — the GEMM addresses it through rank-4 tensor maps and the result is handed on as a deferred
    head split too, which the merge that follows cancels (no permute kernels on either side).  None when dS carries no
    live-tile table or the shapes do not fit; the caller then takes the ordinary matmul route."""
    from dlgrad_b200.codegen import cuda_flash as cf
    from dlgrad_b200.codegen import cuda_tcgen05 as tcg
    from dlgrad_b200.runtime.cuda import ops as _ops
    if not SPARSE or len(dS.metadata.shape) != 4 or tuple(other.metadata.shape[:2]) != tuple(dS.metadata.shape[:2]):
        return None
    ob, split = _resolve(other)
    if not split or not _stored(other) or not other.aligned:
        return None
    if dS._pending is not None:
        # the flash dS kernel runs here and registers the live-tile table.  This dS is the interior gradient of the score
        # matrix: it is consumed by dq = dS k and dk = dS^T q (both on the tile-skipping GEMM) and released, so the
        # tiles the mask kills need not be written (cuda_flash.flash_attn, lazy_zero)
        _ds_private[0] += 1
        try:
            _ops.materialize(dS)
        finally:
            _ds_private[0] -= 1
    info = tag_of(dS)
    if info is None:
        return None
    _, flags, Tq, Tk = info
    Bn, H = dS.metadata.shape[0], dS.metadata.shape[1]
    nb = Bn * H
    T_other, D = other.metadata.shape[-2], other.metadata.shape[-1]
    if dS.metadata.shape[-2:] != (Tq, Tk) or T_other != (Tq if trans_x else Tk) or D % 32 or Tq % 128 or Tk % 128:
        return None
    M, K = (Tk, Tq) if trans_x else (Tq, Tk)
    key = ("dsgemm", nb, H, Tq, Tk, D, trans_x)
    plan = _tc_plans.get(key)
    if plan is None:
        kern = tcg.matmul_tc_ts(M, D, K, nb, trans_x, True, False, False, ksparse="tn" if trans_x else "nn", hsplit=H)
        da, _, _ = tcg.tmap_descs(M, D, K, nb, trans_x, True, False, False, kern.bn)
        so = _strides(T_other, D, H, True)
        sc = _strides(M, D, H, True)
        maps = [da, cf.tmap4(T_other, D, H, Bn, so, 32, mn=True),
                dict(rank=4, dims=(D, M, H, Bn), strides=(sc[0] * 4, sc[1] * 4, sc[2] * 4), box=(32, 32, 1, 1), swizzle=3)]
        plan = _tc_plans[key] = _tc_plan(kern, maps, 3)
    out = shim.alloc(nb * M * D * 4)
    pa, pb = _ops._dev_ptr(dS), _ops._dev_ptr(ob)
    _timed(("tc3_sp" + ("tn" if trans_x else "nn") + "_hs", M, D, K, nb), 2.0 * nb * M * D * K,
           lambda: _tc_run(plan, [pa, pb, out], [out, shim.NULL, flags], "tile-skipping GEMM over the head-split layout"))
    return _wrap(out, tuple(dS.metadata.shape[:2]) + (M, D), True)
