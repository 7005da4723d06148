"""MATMUL on Device.CUDA: picks between the tcgen05/TMEM/TMA 3xTF32 kernel (codegen/cuda_tcgen05.py)
and the exact-fp32 SIMT kernel (codegen/cuda_kernel.matmul_simt) per shape.

Operands arrive rank-aligned (2-D, 3-D or 4-D, dlgrad/buffer.py:234-235); leading dims of size 1
broadcast against the other operand exactly as the reference zeroes the batch stride
(dlgrad/runtime/cpu.py:507-575).  `trans_x` / `trans_y` are an extension of the CUDA backend used by
MatMul.backward so that g @ B^T and A^T @ g need no materialised transpose.
"""
from __future__ import annotations

import os
import weakref

from dlgrad_b200.buffer import Buffer
from dlgrad_b200.codegen import cuda_kernel as ck
from dlgrad_b200.helpers import prod_
from dlgrad_b200.runtime.cuda import ops as _ops
from dlgrad_b200.runtime.cuda import profile as _profile

_plans: dict = {}
FORCE_SIMT = os.environ.get("DLGRAD_MATMUL", "").lower() == "simt"
FUSED_SPLITK = os.environ.get("DLGRAD_FUSED_SPLITK", "1") != "0"


def _geometry(xs: tuple, ys: tuple, tx: bool, ty: bool):
    nd = len(xs)
    bx, by = xs[:-2], ys[:-2]
    M, K = (xs[-1], xs[-2]) if tx else (xs[-2], xs[-1])
    K2, N = (ys[-1], ys[-2]) if ty else (ys[-2], ys[-1])
    if K != K2:
        raise ValueError(f"Matrix inner dims must match: {xs} and {ys}")
    batch = tuple(max(a, b) for a, b in zip(bx, by))
    for a, b in zip(bx, by):
        if a != b and 1 not in (a, b):
            raise ValueError(f"Cannot broadcast batch dims: {xs} and {ys}")
    batch = (1,) * (2 - len(batch)) + batch
    mat_x, mat_y = xs[-1] * xs[-2], ys[-1] * ys[-2]

    def bstrides(b, mat):
        b = (1,) * (2 - len(b)) + tuple(b)
        s1 = mat if b[1] > 1 else 0
        s0 = mat * b[1] if b[0] > 1 else 0
        return (s0 if batch[0] > 1 else 0, s1 if batch[1] > 1 else 0)
    return M, N, K, batch, bstrides(bx, mat_x), bstrides(by, mat_y), nd


def _fold_transpose(b: Buffer, flag: bool):
    pend = b._pending
    if pend is not None and pend[0] == "tr":
        src = pend[1]
        ss = src.metadata.shape
        if b.metadata.shape == ss[:-2] + (ss[-1], ss[-2]):      # not reshaped in place since (unsqueeze/squeeze)
            return src, not flag
    return b, flag


def run_matmul(x: Buffer, y: Buffer, trans_x: bool = False, trans_y: bool = False, bias: Buffer | None = None,
               relu_mask: Buffer | None = None):
    """`bias` (a length-N row) is added by the GEMM epilogue when the shape runs on the TS tensor-core kernel;
    returns (ptr, bias_applied) in that case so the caller adds it separately otherwise.  `relu_mask` (a matrix shaped
    like the product): the epilogue keeps the product only where it is positive; returns (ptr, applied)."""
    # a deferred transpose of an operand (Buffer.transpose) is folded into the kernel's operand-major flag
    x, trans_x = _fold_transpose(x, trans_x)
    y, trans_y = _fold_transpose(y, trans_y)
    if relu_mask is not None:
        out = _relu_masked_matmul(x, y, trans_x, trans_y, relu_mask)
        return (out, True) if out is not None else (run_matmul(x, y, trans_x, trans_y), False)
    if bias is None and ((x._pending is not None and x._pending[0] == "relu") or (y._pending is not None and y._pending[0] == "relu")):
        out = _relu_operand_matmul(x, y, trans_x, trans_y)
        if out is not None:
            return out
    aligned = x.aligned and y.aligned
    if x._pending is not None:
        _ops.materialize(x)            # a deferred dS gets its live-tile table while it is materialised
    sp = _sparse_plan(x, y, trans_x, trans_y, aligned) if bias is None else None
    if sp is not None:
        plan, flags = sp
        args = (_ops._dev_ptr(x), _ops._dev_ptr(y), _ops.NULL, flags)
        timer = _profile.matmul_timer
        if timer is None:
            return plan(*args)
        t0 = timer.start()
        out = plan(*args)
        timer.stop(t0, plan.flops, plan.tag)
        return out
    fused = None
    if bias is not None and bias.aligned and bias.scalar is None:
        bkey = (x.metadata.shape, y.metadata.shape, trans_x, trans_y, aligned, "bias")
        fused = _plans.get(bkey)
        if fused is None and bkey not in _plans:
            fused = _plans[bkey] = _build_bias(x.metadata.shape, y.metadata.shape, trans_x, trans_y, aligned)
    if fused is not None:
        plan, args = fused, (_ops._dev_ptr(x), _ops._dev_ptr(y), _ops._dev_ptr(bias))
    else:
        key = (x.metadata.shape, y.metadata.shape, trans_x, trans_y, aligned)
        plan = _plans.get(key)
        if plan is None:
            plan = _plans[key] = _build(x.metadata.shape, y.metadata.shape, trans_x, trans_y, aligned)
        args = (_ops._dev_ptr(x), _ops._dev_ptr(y))      # materialises deferred operands OUTSIDE the timer bracket
    timer = _profile.matmul_timer
    if timer is None:
        out = plan(*args)
    else:
        t0 = timer.start()
        out = plan(*args)
        timer.stop(t0, plan.flops, plan.tag)
    return out if bias is None else (out, fused is not None)


def _relu_masked_matmul(x: Buffer, y: Buffer, tx: bool, ty: bool, mask: Buffer):
    """x @ y with the relu-backward select on `mask` (the pre-activation) in the epilogue; None when the shape does not
    run on the tensor-core kernel as one launch."""
    if x._pending is not None or y._pending is not None or mask._pending is not None or mask.scalar is not None:
        return None
    aligned = x.aligned and y.aligned and mask.aligned
    key = (x.metadata.shape, y.metadata.shape, tx, ty, aligned, "rmask")
    plan = _plans.get(key)
    if plan is None and key not in _plans:
        M, N, K, batch, a_bs, b_bs, out_shape = _geometry(x.metadata.shape, y.metadata.shape, tx, ty)
        plan = None
        if not FORCE_SIMT and aligned and N % 4 == 0 and prod_(mask.metadata.shape) == prod_(batch) * M * N \
                and _split_k(M, N, K, batch, tx, ty) == 1:
            from dlgrad_b200.runtime.cuda import matmul_tc
            # BN <= 128 keeps two accumulators in TMEM, so the longer epilogue (it also reads the X tile) overlaps the
            # next tile's MMAs; at BN = 192 there is one accumulator and the select costs more than the kernel it replaces
            bn = int(os.environ.get("DLGRAD_RMASK_BN", "128"))
            plan = matmul_tc.try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, relu_mask=True,
                                       bn=bn if N >= bn else None)
        _plans[key] = plan
    if plan is None:
        return None
    args = (_ops._dev_ptr(x), _ops._dev_ptr(y), _ops._dev_ptr(mask))
    timer = _profile.matmul_timer
    if timer is None:
        return plan(*args)
    t0 = timer.start()
    out = plan(*args)
    timer.stop(t0, plan.flops, plan.tag)
    return out


def radd_matmul(x: Buffer, y: Buffer, tx: bool, ty: bool, r: Buffer, a_relu: bool = False):
    """r + x @ y with the addition in the GEMM's epilogue (codegen matmul_tc_ts, `radd`): the residual connection and the
    gradient accumulation that follow a Linear.  `a_relu`: x is relu(stored x) (a deferred relu feeding the Linear).  None
    when the shape does not run on the tensor-core kernel as one launch (the caller then adds separately)."""
    if x._pending is not None and x._pending[0] == "relu" and x._pending[1]._pending is None \
            and x._pending[1].metadata.shape == x.metadata.shape and not a_relu:
        return radd_matmul(x._pending[1], y, tx, ty, r, a_relu=True)      # relu(h) @ W + r: the clamp rides in the A splitter
    if x._pending is not None or y._pending is not None or r._pending is not None or r.scalar is not None:
        return None
    aligned = x.aligned and y.aligned and r.aligned
    key = (x.metadata.shape, y.metadata.shape, tx, ty, aligned, "radd", a_relu)
    plan = _plans.get(key)
    if plan is None and key not in _plans:
        M, N, K, batch, a_bs, b_bs, _ = _geometry(x.metadata.shape, y.metadata.shape, tx, ty)
        plan = None
        if (not FORCE_SIMT and aligned and N % 4 == 0 and prod_(r.metadata.shape) == prod_(batch) * M * N
                and _split_k(M, N, K, batch, tx, ty) == 1):
            from dlgrad_b200.runtime.cuda import matmul_tc
            bn = int(os.environ.get("DLGRAD_RMASK_BN", "128"))      # two accumulators: the longer epilogue overlaps the next tile
            plan = matmul_tc.try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, radd=True, a_relu=a_relu,
                                       bn=bn if N >= bn else None)
        _plans[key] = plan
    if plan is None:
        return None
    args = (_ops._dev_ptr(x), _ops._dev_ptr(y), _ops._dev_ptr(r))
    timer = _profile.matmul_timer
    if timer is None:
        return plan(*args)
    t0 = timer.start()
    out = plan(*args)
    timer.stop(t0, plan.flops, plan.tag)
    return out


def _relu_operand_matmul(x: Buffer, y: Buffer, tx: bool, ty: bool):
    """An operand is a deferred relu (Buffer pending kind "relu"): the tensor-core kernel applies max(., 0) while the
    operand passes through its splitter, so relu(h) is never written.  None when the shape does not run there as one
    launch (the caller then materialises the relu and takes the ordinary route)."""
    a_relu = x._pending is not None and x._pending[0] == "relu"
    b_relu = y._pending is not None and y._pending[0] == "relu"
    xs = x._pending[1] if a_relu else x
    ys = y._pending[1] if b_relu else y
    if xs._pending is not None or ys._pending is not None or xs.metadata.shape != x.metadata.shape or ys.metadata.shape != y.metadata.shape:
        return None
    aligned = xs.aligned and ys.aligned
    key = (xs.metadata.shape, ys.metadata.shape, tx, ty, aligned, "relu", a_relu, b_relu)
    plan = _plans.get(key)
    if plan is None and key not in _plans:
        M, N, K, batch, a_bs, b_bs, _ = _geometry(xs.metadata.shape, ys.metadata.shape, tx, ty)
        plan = None
        if not FORCE_SIMT and _split_k(M, N, K, batch, tx, ty) == 1:
            from dlgrad_b200.runtime.cuda import matmul_tc
            plan = matmul_tc.try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, a_relu=a_relu, b_relu=b_relu)
        _plans[key] = plan
    if plan is None:
        return None
    args = (_ops._dev_ptr(xs), _ops._dev_ptr(ys))
    timer = _profile.matmul_timer
    if timer is None:
        return plan(*args)
    t0 = timer.start()
    out = plan(*args)
    timer.stop(t0, plan.flops, plan.tag)
    return out


def _sparse_plan(x: Buffer, y: Buffer, tx: bool, ty: bool, aligned: bool):
    """(plan, live-tile table) when `x` is a P / dS matrix from the fused attention kernels: the GEMM skips the k-blocks
    of tiles the mask kills entirely."""
    from dlgrad_b200.runtime.cuda import attention
    info = attention.tag_of(x)
    if info is None:
        return None
    _, flags, Tq, Tk = info
    xs = x.metadata.shape
    if xs[-2:] != (Tq, Tk):
        return None
    key = (xs, y.metadata.shape, tx, ty, aligned, "ksparse")
    plan = _plans.get(key)
    if plan is None and key not in _plans:
        M, N, K, batch, a_bs, b_bs, _ = _geometry(xs, y.metadata.shape, tx, ty)
        from dlgrad_b200.runtime.cuda import matmul_tc
        plan = None
        if not FORCE_SIMT and M % 128 == 0 and K % 128 == 0 and K <= 2048:
            plan = matmul_tc.try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, ksparse="tn" if tx else "nn")
        _plans[key] = plan
    return None if plan is None else (plan, flags)


def _build_bias(xs, ys, tx, ty, aligned):
    """The bias-in-epilogue variant, or None when this shape does not run on the TS kernel as one launch."""
    M, N, K, batch, a_bs, b_bs, _ = _geometry(xs, ys, tx, ty)
    if FORCE_SIMT or _split_k(M, N, K, batch, tx, ty) > 1:
        return None
    from dlgrad_b200.runtime.cuda import matmul_tc
    return matmul_tc.try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned, bias=True)


def _build(xs, ys, tx, ty, aligned):
    M, N, K, batch, a_bs, b_bs, _ = _geometry(xs, ys, tx, ty)
    if not FORCE_SIMT:
        from dlgrad_b200.runtime.cuda import matmul_tc
        split = _split_k(M, N, K, batch, tx, ty)
        if split > 1:
            sk = _build_split_k(M, N, K, split, aligned)
            if sk is not None:
                return sk
        tc = matmul_tc.try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned)
        if tc is not None:
            return tc
    k = ck.matmul_simt(M, N, K, batch[0], batch[1], a_bs, b_bs, tx, ty, aligned)
    p = _ops.Plan(k, prod_(batch) * M * N)
    return MatmulPlan(lambda a, b: p.run(a, b), 2.0 * prod_(batch) * M * N * K, ("simt", M, N, K, prod_(batch)))


def _split_k(M: int, N: int, K: int, batch: tuple, tx: bool, ty: bool) -> int:
    """Weight-gradient GEMMs (x^T @ g: small M x N output, K = batch*tokens) leave most of the 148 SMs idle
    with one CTA per output tile.  For the TN form both operands are stored [K][cols], so K-slices ARE a
    batch dimension: run S independent slices as a batched GEMM and add the S partial outputs in a fixed
    order with the column-reduce kernel (deterministic, no atomics)."""
    if not (tx and not ty) or prod_(batch) != 1 or K < 2048:
        return 1
    tiles = -(-M // 128) * -(-N // 128)
    if tiles >= 100:
        return 1
    s = 1
    while s * 2 * tiles <= 160 and K % (s * 2) == 0 and K // (s * 2) >= 512:
        s *= 2
    return s


def _build_split_k(M: int, N: int, K: int, S: int, aligned: bool):
    from dlgrad_b200.runtime.cuda import matmul_tc
    ks = K // S
    if FUSED_SPLITK and ks <= 2048 and N % 4 == 0:
        # one launch: the slice that finishes an output tile last adds the S partial sums itself
        one = matmul_tc.try_build(M, N, ks, (1, S), (0, ks * M), (0, ks * N), True, False, aligned, splitk=S)
        if one is not None:
            return MatmulPlan(one.fn, 2.0 * M * N * K, ("tc3_splitk", M, N, K, S))
    part = matmul_tc.try_build(M, N, ks, (1, S), (0, ks * M), (0, ks * N), True, False, aligned)
    if part is None:
        return None
    reduce_plans = _ops._build_reduce("sum", (S, M * N), 0, True)

    def run(a, b):
        ptr = part(a, b)
        for p in reduce_plans:
            ptr = p.run(ptr)
        return ptr
    return MatmulPlan(run, 2.0 * M * N * K, ("tc3_splitk", M, N, K, S))


class MatmulPlan:
    __slots__ = ("fn", "flops", "tag")

    def __init__(self, fn, flops: float, tag: tuple) -> None:
        self.fn, self.flops, self.tag = fn, flops, tag

    def __call__(self, *operands):
        return self.fn(*operands)


# ---------------------------------------------------------------------------------------------------------
# sibling Linear layers over one input: one grouped launch
# ---------------------------------------------------------------------------------------------------------
GROUP_LINEAR = os.environ.get("DLGRAD_GROUP_LINEAR", "1") != "0"


class LinSlot:
    """Result cell of one deferred Linear output `x2 @ w^T` (Buffer pending kind "lin").  Every Buffer that carries the
    pending tuple — the output itself and deferred reshapes of it — resolves through the slot, so the product runs once.
    Slots of consecutive Linear calls on the SAME stored input with equally shaped weights form a group (the q / k / v
    projections of an attention block, examples/gpt.py:43-45); the first member that is needed computes every member
    that is still open with one grouped tensor-core launch (codegen matmul_tc_ts, group = ("n", G)).
    The group list holds WEAK references: a slot keeps its output storage alive, and a slot <-> list cycle would leave
    25 MB activations to the cyclic garbage collector instead of the reference count."""
    __slots__ = ("ptr", "x2", "w", "group", "__weakref__")

    def __init__(self, x2: Buffer, w: Buffer) -> None:
        self.ptr, self.x2, self.w = None, x2, w
        self.group = [weakref.ref(self)]

    def members(self) -> list:
        return [m for m in (r() for r in self.group) if m is not None]

    def leave(self) -> None:
        self.group[:] = [r for r in self.group if r() is not self and r() is not None]
        self.group = [weakref.ref(self)]


_last_group: list = []


def lin_slot(x2: Buffer, w: Buffer) -> LinSlot:
    global _last_group
    slot = LinSlot(x2, w)
    g = [m for m in (r() for r in _last_group) if m is not None]
    if (GROUP_LINEAR and g and len(g) < 3 and len(g) == len(_last_group) and x2._ptr is not None
            and all(m.ptr is None for m in g) and g[0].x2._ptr is x2._ptr and g[0].x2.metadata.shape == x2.metadata.shape
            and g[0].w.metadata.shape == w.metadata.shape and all(m.w is not w for m in g)):
        _last_group.append(weakref.ref(slot))
        slot.group = _last_group
    _last_group = slot.group
    return slot


def run_lin_slot(slot: LinSlot):
    """Storage of the slot's product; computes it — and its open siblings — if nobody has yet."""
    if slot.ptr is None:
        todo = [m for m in slot.members() if m.ptr is None]
        outs = _grouped_linear(slot.x2, [m.w for m in todo]) if len(todo) >= 2 else None
        if outs is not None:
            for m, o in zip(todo, outs):
                m.ptr = o
        else:
            slot.ptr = run_matmul(slot.x2, slot.w, trans_y=True)
        for m in todo:
            if m.ptr is not None:
                m.x2 = m.w = None                    # the operands are not needed any more
    return slot.ptr


def _grouped_linear(x2: Buffer, ws: list):
    """[x2 @ w^T for w in ws] as ONE launch, or None when the shape does not run on the tensor-core kernel."""
    if FORCE_SIMT or x2._pending is not None or any(w._pending is not None or w.scalar is not None for w in ws):
        return None
    aligned = x2.aligned and all(w.aligned for w in ws)
    G = len(ws)
    key = (x2.metadata.shape, ws[0].metadata.shape, G, aligned, "group_n")
    plan = _plans.get(key)
    if plan is None and key not in _plans:
        M, K = x2.metadata.shape
        N = ws[0].metadata.shape[0]
        plan = None
        if aligned and _split_k(M, N, K, (1, 1), False, True) == 1:
            from dlgrad_b200.runtime.cuda import matmul_tc
            plan = matmul_tc.build_grouped("n", G, M, N, K, False, False)
        _plans[key] = plan
    if plan is None:
        return None
    args = (_ops._dev_ptr(x2), [_ops._dev_ptr(w) for w in ws])
    timer = _profile.matmul_timer
    if timer is None:
        return plan(*args)
    t0 = timer.start()
    out = plan(*args)
    timer.stop(t0, plan.flops, plan.tag)
    return out


def run_lin_dx_sum(gs: tuple, ws: tuple):
    """sum_i gs[i] @ ws[i] — the input gradient shared by sibling Linear layers — as ONE K-concatenated tensor-core launch;
    where the shape does not run there: the first product, then each further one with the add in its epilogue."""
    G = len(gs)
    if G == 1:
        return run_matmul(gs[0], ws[0])
    plan = None
    ok = (not FORCE_SIMT and all(b._pending is None or b._pending[0] != "relu" for b in gs)
          and all(w._pending is None and w.scalar is None for w in ws))
    if ok:
        for g in gs:
            if g._pending is not None:
                _ops.materialize(g)
        aligned = all(b.aligned for b in gs + ws)
        key = (gs[0].metadata.shape, ws[0].metadata.shape, G, aligned, "group_k")
        plan = _plans.get(key)
        if plan is None and key not in _plans:
            M, K = gs[0].metadata.shape
            N = ws[0].metadata.shape[1]
            if aligned and len(gs[0].metadata.shape) == 2 and _split_k(M, N, K, (1, 1), False, False) == 1:
                from dlgrad_b200.runtime.cuda import matmul_tc
                plan = matmul_tc.build_grouped("k", G, M, N, K, False, True)
            _plans[key] = plan
    if plan is not None:
        args = ([_ops._dev_ptr(g) for g in gs], [_ops._dev_ptr(w) for w in ws])
        timer = _profile.matmul_timer
        if timer is None:
            return plan(*args)
        t0 = timer.start()
        out = plan(*args)
        timer.stop(t0, plan.flops, plan.tag)
        return out
    acc = Buffer(run_matmul(gs[0], ws[0]), gs[0].metadata.shape[:-1] + (ws[0].metadata.shape[1],), gs[0].metadata.device)
    for g, w in zip(gs[1:], ws[1:]):
        nxt = radd_matmul(g, w, False, False, acc)
        if nxt is None:
            nxt = _ops.add(acc, Buffer(run_matmul(g, w), acc.metadata.shape, acc.metadata.device))
        acc = Buffer(nxt, acc.metadata.shape, acc.metadata.device)
    return acc.ptr


# ---------------------------------------------------------------------------------------------------------
# weight gradients of sibling Linear layers: one grouped split-K launch
# ---------------------------------------------------------------------------------------------------------
class DwSlot:
    """Result cell of one deferred Linear weight gradient g2^T @ x2 (Buffer pending kind "lin_dw").  Slots of consecutive
    Linear.backward calls over the SAME stored input with equally shaped gradients form a group (dWq, dWk, dWv); a group
    runs as soon as it is complete (three members), when a Linear over another input follows, or when a member is read —
    as ONE grouped split-K launch when two or three members are open.  The slot hands the result to its Buffer at once, so
    the gradient operands (25 MB each at the benchmark's size) are released as early as without the deferral."""
    __slots__ = ("ptr", "g2", "x2", "buf")

    def __init__(self, g2: Buffer, x2: Buffer) -> None:
        self.ptr, self.g2, self.x2, self.buf = None, g2, x2, None


_open_dw: list = []


def lin_dw(g2: Buffer, x2: Buffer, out_f: int, in_f: int, dtype) -> Buffer:
    """dW = g2^T @ x2 of a Linear layer as a Buffer: deferred into the open group when the shape is a split-K weight gradient
    that may get siblings, computed at once otherwise."""
    global _open_dw
    rows = g2.metadata.shape[0]
    groupable = (GROUP_LINEAR and rows >= 2048 and not FORCE_SIMT and g2._pending is None and x2._pending is None and g2._ptr is not None
                 and x2._ptr is not None and g2.aligned and x2.aligned and _split_k(out_f, in_f, rows, (1, 1), True, False) > 1)
    if _open_dw and not (groupable and len(_open_dw) < 3 and _open_dw[0].x2._ptr is x2._ptr
                         and _open_dw[0].g2.metadata.shape == g2.metadata.shape and all(m.ptr is None for m in _open_dw)):
        flush_dw()
    if not groupable:
        return Buffer(run_matmul(g2, x2, trans_x=True), (out_f, in_f), g2.metadata.device, dtype)
    slot = DwSlot(g2, x2)
    out = Buffer._deferred(("lin_dw", g2, x2, slot), (out_f, in_f), dtype)
    slot.buf = weakref.ref(out)
    _open_dw.append(slot)
    if len(_open_dw) == 3:
        flush_dw()
    return out


def flush_dw() -> None:
    """Run the open group of deferred weight gradients."""
    global _open_dw
    todo, _open_dw = [m for m in _open_dw if m.ptr is None], []
    if not todo:
        return
    outs = _grouped_dw([m.g2 for m in todo], todo[0].x2) if len(todo) >= 2 else None
    for i, m in enumerate(todo):
        m.ptr = outs[i] if outs is not None else run_matmul(m.g2, m.x2, trans_x=True)
        m.g2 = m.x2 = None
        b = m.buf() if m.buf is not None else None
        if b is not None and b._pending is not None and b._pending[0] == "lin_dw":
            b.ptr = m.ptr


def run_dw_slot(slot: DwSlot):
    if slot.ptr is None:
        if any(m is slot for m in _open_dw):
            flush_dw()
        else:
            slot.ptr = run_matmul(slot.g2, slot.x2, trans_x=True)
            slot.g2 = slot.x2 = None
    return slot.ptr


def _grouped_dw(gs: list, x2: Buffer):
    G = len(gs)
    rows, M = gs[0].metadata.shape
    N = x2.metadata.shape[1]
    key = (gs[0].metadata.shape, x2.metadata.shape, G, "group_m")
    plan = _plans.get(key)
    if plan is None and key not in _plans:
        from dlgrad_b200.runtime.cuda import matmul_tc
        S = _split_k(M, N, rows, (1, 1), True, False)
        plan = _plans[key] = matmul_tc.build_grouped("m", G, M, N, rows, True, True, radd=S) if S > 1 else None
    if plan is None:
        return None
    args = ([_ops._dev_ptr(g) for g in gs], _ops._dev_ptr(x2))
    timer = _profile.matmul_timer
    if timer is None:
        return plan(*args)
    t0 = timer.start()
    out = plan(*args)
    timer.stop(t0, plan.flops, plan.tag)
    return out
