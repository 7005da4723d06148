"""Device.CUDA runtime: the functions registered with the dispatcher for every (op, Device.CUDA).

Keyword names and return convention are the reference's (dlgrad/runtime/cpu.py): each
function takes Buffers (+ python scalars / tuples) and returns a raw `float *` cdata that owns
freshly allocated, contiguous storage (here: device memory, released to the caching allocator by
`ffi.gc(ptr, dlg_free)`).

Per (op, shape) the first call generates + compiles a shape-specialised kernel and freezes a launch
plan in the shim; later calls are one dict lookup and one FFI call (`dlg_plan_run` = allocate
output + launch).  No `os.path.exists`, no `ffi.new` per call — the reference spends 20-50 µs of
Python per dispatch there.
"""
from __future__ import annotations

import builtins
import os

import numpy as np

from dlgrad_b200.buffer import Buffer, flush_pending
from dlgrad_b200.codegen import cuda_kernel as ck
from dlgrad_b200.device import Device
from dlgrad_b200.dispatch import dispatcher
from dlgrad_b200.dtype import CDataPtr, Scalar
from dlgrad_b200.helpers import (
    BinaryOps,
    BufferOps,
    CustomOps,
    UnaryOps,
    calculate_stride,
    get_broadcast_shape,
    get_broadcast_strides_2,
    prod_,
)
from dlgrad_b200.runtime.cuda import compiler, shim, transfer
from dlgrad_b200.runtime.cuda import profile as _profile

ffi, lib, NULL = shim.ffi, shim.lib, shim.NULL
_own = shim.own
CUDA = Device.CUDA
_plans: dict = {}
_rng_state = {"dev": None, "seed": None, "calls": 0}     # Philox key (device float[4] block) and the next call number
builtins_max, builtins_min = builtins.max, builtins.min


def manual_seed(seed: int) -> None:
    """Seed Tensor.uniform/rand on this backend (the reference reseeds from /dev/random on every
    call, dlgrad/codegen/cpu_kernel.py:774-874, and cannot be made reproducible)."""
    st = _rng_state
    st["seed"], st["calls"] = int(seed) & 0xFFFFFFFFFFFFFFFF, 0
    if st["dev"] is not None:
        _write_rng_key()


def _write_rng_key() -> None:
    st = _rng_state
    words = np.array([st["seed"] & 0xFFFFFFFF, st["seed"] >> 32, 0, 0], dtype=np.uint32)
    transfer.write_bytes(st["dev"], words.tobytes())


def _rng_buffer():
    """Device block [key lo, key hi, call base of the next replay, -].  Set up by an eager call: a block taken inside a
    capture would come from the capture's pool, where the graph's own kernels write."""
    st = _rng_state
    if st["dev"] is None:
        shim.no_capture("the first random draw of the process (graph.CapturedStep's warm-up step does it eagerly)")
        if st["seed"] is None:
            st["seed"] = int.from_bytes(os.urandom(8), "little")
        st["dev"] = shim.alloc(16)
        _write_rng_key()
    return st["dev"]


def _rng_replay_hook(info: dict):
    """Before a replay: reserve the call numbers the graph's draws will use and tell the device where they start."""
    def tick() -> None:
        st = _rng_state
        base = st["calls"] & 0xFFFFFFFF
        st["calls"] += info["n"]
        plan = _plans.get("set_call_base")
        if plan is None:
            plan = _plans["set_call_base"] = Plan(ck.set_call_base())
        plan.run_into(st["dev"], NULL, NULL, NULL, NULL, float(base & 0xFFFFFF), float(base >> 24))
    return tick


class Plan:
    __slots__ = ("c", "nbytes", "kernel")

    def __init__(self, kernel: ck.Kernel, out_elems: int | None = None, cluster: int = 1) -> None:
        shim.init()
        fn = compiler.get_function(kernel.name, kernel.source)
        n = kernel.out_elems if out_elems is None else out_elems
        self.nbytes = n * 4
        self.kernel = kernel
        self.c = lib.dlg_plan_create(fn, *kernel.grid, *kernel.block, kernel.smem, self.nbytes,
                                     int(kernel.zero_out), cluster)
        if self.c == NULL:
            raise RuntimeError("dlg_plan_create failed: " + shim.last_error())

    def run(self, p0=NULL, p1=NULL, p2=NULL, p3=NULL, f0=0.0, f1=0.0, f2=0.0, f3=0.0) -> CDataPtr:
        timer = _profile.kernel_timer
        if timer is not None:
            t0 = timer.start()
            out = lib.dlg_plan_run(self.c, p0, p1, p2, p3, f0, f1, f2, f3)
            timer.stop(t0, float(self.nbytes), (self.kernel.name, self.nbytes))
        else:
            out = lib.dlg_plan_run(self.c, p0, p1, p2, p3, f0, f1, f2, f3)
        if out == NULL:
            raise RuntimeError("kernel launch failed: " + shim.last_error())
        return _own(out, self.nbytes)

    def run_into(self, out, p0=NULL, p1=NULL, p2=NULL, p3=NULL, f0=0.0, f1=0.0, f2=0.0, f3=0.0) -> None:
        timer = _profile.kernel_timer
        t0 = timer.start() if timer is not None else None
        if lib.dlg_plan_run_into(self.c, out, p0, p1, p2, p3, f0, f1, f2, f3) != 0:
            raise RuntimeError("kernel launch failed: " + shim.last_error())
        if timer is not None:
            timer.stop(t0, 0.0, (self.kernel.name, 0))

    def run_into_comm(self, out, p0=NULL, p1=NULL, p2=NULL, p3=NULL, f0=0.0, f1=0.0, f2=0.0, f3=0.0) -> None:
        """run_into on the comm stream (the overlapped data-parallel exchange, dlgrad_b200/distributed.py)."""
        if lib.dlg_comm_plan_run_into(self.c, out, p0, p1, p2, p3, f0, f1, f2, f3) != 0:
            raise RuntimeError("kernel launch on the comm stream failed: " + shim.last_error())


# ------------------------------------------------------------------------------------------------
# deferred pointwise chains (Buffer._pending)
# ------------------------------------------------------------------------------------------------
def _peel_chain(buf: Buffer) -> tuple[list, Buffer]:
    """Pending pointwise ops from `buf` down to its base, innermost first."""
    chain, b = [], buf
    while b._pending is not None and b._pending[0] in ("scale", "mfill"):
        chain.append(b._pending)
        b = b._pending[1]
    chain.reverse()
    return chain, b


def _mask_lead(out_shape: tuple, mask: Buffer):
    """(lead_shape, lead_strides, ok) of a mask broadcast against rows of `out_shape`."""
    ms = get_broadcast_strides_2(out_shape, mask.metadata.shape, calculate_stride(mask.metadata.shape))
    if ms[-1] != 1 and out_shape[-1] != 1:
        return None
    return tuple(out_shape[:-1]), tuple(ms[:-1])


def _row_chain_args(chain: list, out_shape: tuple, aligned: bool):
    """Translate pending ops into softmax-kernel chain descriptors + launch arguments.  Slots: p2/p3 carry
    up to two masks (p0/p1 are the row kernel's own operands), f0/f2 carry scales, f1/f3 fill values."""
    desc, ptrs, fl = [], [], [0.0, 0.0, 0.0, 0.0]
    n_scale = n_mask = 0
    for op in chain:
        if op[0] == "scale":
            if n_scale >= 2:
                return None
            slot = (0, 2)[n_scale]
            fl[slot] = op[2]
            desc.append(("scale", slot))
            n_scale += 1
        else:
            if n_mask >= 2:
                return None
            lead = _mask_lead(out_shape, op[2])
            if lead is None or not op[2].aligned or any(s % 4 for s in lead[1]):
                return None
            fslot, pslot = (1, 3)[n_mask], (2, 3)[n_mask]
            fl[fslot] = op[3]
            desc.append(("mfill", pslot, fslot, lead[0], lead[1]))
            ptrs.append(op[2])
            n_mask += 1
    return tuple(desc), ptrs, fl


def materialize(buf: Buffer) -> None:
    """Run the deferred producer chain of `buf` as ONE kernel and give it real storage."""
    pend = buf._pending
    if pend is None:
        return
    if pend[0] == "full":
        buf.ptr = full(pend[1], pend[2])
        return
    if pend[0] == "arange":
        buf.ptr = arange(pend[1])
        return
    if pend[0] == "tr":
        src = pend[1]
        nd = len(src.metadata.shape)
        buf.ptr = permute(src, tuple(range(nd - 2)) + (nd - 1, nd - 2))
        return
    if pend[0] == "mm":
        buf.ptr = matmul(pend[1], pend[2])
        return
    if pend[0] == "relu":
        buf.ptr = pend[1].relu().ptr
        return
    if pend[0] == "lin":
        from dlgrad_b200.runtime.cuda import matmul as _mm
        buf.ptr = _mm.run_lin_slot(pend[3])
        return
    if pend[0] == "lin_dx":
        buf.ptr = matmul(pend[1], pend[2])
        return
    if pend[0] == "lin_dw":
        from dlgrad_b200.runtime.cuda import matmul as _mm
        buf.ptr = _mm.run_dw_slot(pend[3])
        return
    if pend[0] == "lin_dx_sum":
        from dlgrad_b200.runtime.cuda import matmul as _mm
        buf.ptr = _mm.run_lin_dx_sum(pend[1::2], pend[2::2])
        return
    if pend[0] == "rms_dx":
        buf.ptr = ewise3("rms_dx", "(v0 * v1) * v2", buf.metadata.shape, [pend[1], pend[2], pend[3]])
        return
    if pend[0] == "diff":
        buf.ptr = sub(pend[1], pend[2])
        return
    if pend[0] == "hsplit":
        buf.ptr = permute(pend[1], (0, 2, 1, 3))
        return
    if pend[0] == "attnP":
        from dlgrad_b200.runtime.cuda import attention
        attention.materialize_p(buf)
        return
    chain, base = _peel_chain(buf)
    shape = buf.metadata.shape
    bp = base._pending
    if bp is not None and bp[0] == "softmax_bwd":
        _, y, g, dim, kind = bp
        gp = g._pending
        if gp is not None and gp[0] == "mm" and len(gp) == 4 and kind == "softmax":
            # upstream is the deferred dP = dO @ v^T of an attention block (ops.MatMul.backward): dS in one kernel
            from dlgrad_b200.runtime.cuda import attention
            m = attention.bwd_chain_matches(chain, shape)
            vt = gp[2]
            rec = attention.flash_record(y)
            if (rec is not None and m is not None and vt._pending is not None and vt._pending[0] == "tr" and gp[3] is not None
                    and vt._pending[1] is rec.v and m[0] is rec.mask and m[1] == rec.scale and attention._stored(gp[1])):
                # P was never written (flash forward): recompute it per tile from q, k and the row statistics
                buf.ptr = attention.flash_ds(rec, gp[1], gp[3], buf)
                attention.claim(buf)
                return
            if y._pending is not None:
                materialize(y)                      # the round-1 kernel below reads P
            if m is not None and vt._pending is not None and vt._pending[0] == "tr" and gp[3] is not None:
                buf.ptr = attention.dscores(gp[1], vt._pending[1], gp[3], y, m[0], m[1])
                attention.claim(buf)                # dS carries the live-tile table for the two GEMMs that read it
                return
        rows = _rows_of(shape, dim)
        args = _row_chain_args(chain, shape, y.aligned and g.aligned) if rows is not None else None
        if args is not None:
            desc, masks, fl = args
            ok = y.aligned and g.aligned
            key = (kind + "_bwd_chain", rows, ok, desc)
            plan = _plans.get(key)
            if plan is None:
                plan = _plans[key] = Plan(ck.softmax_bwd_rows(kind, rows[0], rows[1], ok, desc))
            mp = [_dev_ptr(m) for m in masks] + [NULL, NULL]
            buf.ptr = plan.run(_dev_ptr(y), _dev_ptr(g), mp[0], mp[1], fl[0], fl[1], fl[2], fl[3])
            return
        base.ptr = _softmax_bwd_family(kind, y, g, dim)          # could not fold: plain kernel, then the chain
    if base._pending is not None:
        materialize(base)
    if not chain:
        buf.ptr = base._ptr
        return
    # pointwise chain over a materialised base: one generated ewise kernel
    expr, strides, ptrs, fl, nf = "v0", [calculate_stride(shape)], [base], [0.0, 0.0, 0.0, 0.0], 0
    for op in chain:
        if nf >= 4 or len(ptrs) >= 4:
            break
        if op[0] == "scale":
            expr = f"(({expr}) * f{nf})"
            fl[nf] = op[2]
            nf += 1
        else:
            mask = op[2]
            strides.append(get_broadcast_strides_2(shape, mask.metadata.shape, calculate_stride(mask.metadata.shape)))
            expr = f"((v{len(ptrs)} > 0.f) ? f{nf} : ({expr}))"
            ptrs.append(mask)
            fl[nf] = op[3]
            nf += 1
    else:
        ok = all(p.aligned for p in ptrs)
        key = ("chain", expr, shape, tuple(strides), ok)
        plan = _plans.get(key)
        if plan is None:
            plan = _plans[key] = Plan(ck.ewise(expr, shape, tuple(strides), ok, "chain"))
        pp = [_dev_ptr(p) for p in ptrs] + [NULL] * 3
        buf.ptr = plan.run(pp[0], pp[1], pp[2], pp[3], fl[0], fl[1], fl[2], fl[3])
        return
    # chain too long for one kernel: materialise the inner part first
    inner = buf._pending[1]
    materialize(inner)
    materialize(buf)


def _dev_ptr(b: Buffer):
    """Device pointer of a Buffer operand (host buffers are migrated by Buffer._route before we get
    here; this catches direct dispatcher users)."""
    if b.metadata.device is not CUDA:
        b._to_cuda()
    return b.ptr


# ------------------------------------------------------------------------------------------------
# binary elementwise (ADD/SUB/MUL/DIV/EQT/GT/GTE/CMP)        reference: cpu.py:257-311,784-791,835-862,906
# ------------------------------------------------------------------------------------------------
def _binary(name: str, x: Buffer, y: Buffer, out_shape: tuple | None = None) -> CDataPtr:
    xs, ys = x.scalar, y.scalar
    key = ("bin", name, x.metadata.shape, y.metadata.shape, xs is None, ys is None,
           x.aligned and y.aligned, out_shape)
    ent = _plans.get(key)
    if ent is None:
        if xs is not None and ys is not None:
            # scalar (op) scalar: give x real storage and fall through
            return _binary(name, x._materialized(), y, out_shape)
        oshape = out_shape or get_broadcast_shape(x.metadata.shape, y.metadata.shape)
        strides, names = [], []
        for b, sc, slot in ((x, xs, "f0"), (y, ys, "f1")):
            if sc is None:
                names.append(f"v{len(strides)}")
                strides.append(get_broadcast_strides_2(oshape, b.metadata.shape, calculate_stride(b.metadata.shape)))
            else:
                names.append(slot)
        k = ck.ewise(ck.BINARY_EXPR[name].format(a=names[0], b=names[1]), tuple(oshape), tuple(strides),
                     x.aligned and y.aligned, name)
        mode = 0 if (xs is None and ys is None) else (1 if ys is not None else 2)
        ent = _plans[key] = (Plan(k), mode)
    plan, mode = ent
    if mode == 0:
        return plan.run(_dev_ptr(x), _dev_ptr(y))
    if mode == 1:
        return plan.run(_dev_ptr(x), NULL, NULL, NULL, 0.0, ys)
    return plan.run(_dev_ptr(y), NULL, NULL, NULL, xs, 0.0)


@dispatcher.register(BinaryOps.ADD, CUDA)
def add(x: Buffer, y: Buffer) -> CDataPtr:
    return _binary("add", x, y)


@dispatcher.register(BinaryOps.SUB, CUDA)
def sub(x: Buffer, y: Buffer) -> CDataPtr:
    return _binary("sub", x, y)


@dispatcher.register(BinaryOps.MUL, CUDA)
def mul(x: Buffer, y: Buffer) -> CDataPtr:
    return _binary("mul", x, y)


@dispatcher.register(BinaryOps.DIV, CUDA)
def div(x: Buffer, y: Buffer) -> CDataPtr:
    return _binary("div", x, y)


@dispatcher.register(BinaryOps.EQT, CUDA)
def eqt(x: Buffer, y: Buffer) -> CDataPtr:
    return _binary("eq", x, y)


@dispatcher.register(BinaryOps.GT, CUDA)
def gt(x: Buffer, y: Buffer | Scalar) -> CDataPtr:
    return _binary("gt", x, Buffer._as_buffer(y))


@dispatcher.register(BinaryOps.GTE, CUDA)
def gte(x: Buffer, y: Buffer | Scalar) -> CDataPtr:
    return _binary("gte", x, Buffer._as_buffer(y))


@dispatcher.register(BinaryOps.CMP, CUDA)
def cmp(x: Buffer, y: Buffer, out_shape: tuple, mode: str = "<=") -> CDataPtr:
    if mode != "<=":
        raise NotImplementedError(f"cmp mode {mode!r}")
    return _binary("lte", x, y, tuple(out_shape))


# ------------------------------------------------------------------------------------------------
# unary elementwise                                           reference: cpu.py:314-502
# ------------------------------------------------------------------------------------------------
def _unary(tag: str, expr: str, x: Buffer, f0: float = 0.0, f1: float = 0.0) -> CDataPtr:
    key = ("un", tag, expr, x.metadata.numel, x.aligned)
    plan = _plans.get(key)
    if plan is None:
        n = x.metadata.numel
        plan = _plans[key] = Plan(ck.ewise(expr, (n,), ((1,),), x.aligned, tag))
    return plan.run(_dev_ptr(x), NULL, NULL, NULL, f0, f1)


@dispatcher.register(UnaryOps.NEG, CUDA)
def neg(x: Buffer) -> CDataPtr:
    return _unary("neg", ck.UNARY_EXPR["neg"], x)


@dispatcher.register(UnaryOps.EXP, CUDA)
def exp(x: Buffer) -> CDataPtr:
    return _unary("exp", ck.UNARY_EXPR["exp"], x)


@dispatcher.register(UnaryOps.LOG, CUDA)
def log(x: Buffer) -> CDataPtr:
    return _unary("log", ck.UNARY_EXPR["log"], x)


@dispatcher.register(UnaryOps.SQRT, CUDA)
def sqrt(x: Buffer) -> CDataPtr:
    return _unary("sqrt", ck.UNARY_EXPR["sqrt"], x)


@dispatcher.register(UnaryOps.RSQRT, CUDA)
def rsqrt(x: Buffer) -> CDataPtr:
    return _unary("rsqrt", ck.UNARY_EXPR["rsqrt"], x)


@dispatcher.register(UnaryOps.POW, CUDA)
def pow_(x: Buffer, val: Scalar) -> CDataPtr:
    return _unary("pow", ck.pow_expr(int(val)), x)      # int(): the reference truncates, cpu.py:500


@dispatcher.register(UnaryOps.RELU, CUDA)
def relu(x: Buffer) -> CDataPtr:
    return _unary("relu", "(v0 <= 0.f) ? 0.f : v0", x)   # cpu_kernel.relu, :520-533


@dispatcher.register(UnaryOps.CLAMP, CUDA)
def clamp(x: Buffer, min_val: Scalar | None, max_val: Scalar | None) -> CDataPtr:
    # cpu_kernel.clamp (:321-349): `if (val < min) val = min; if (val > max) val = max;`
    e = "v0"
    if min_val is not None:
        e = f"(({e}) < f0 ? f0 : ({e}))"
    if max_val is not None:
        e = f"(({e}) > f1 ? f1 : ({e}))"
    return _unary("clamp", e, x, float(min_val or 0.0), float(max_val or 0.0))


# ------------------------------------------------------------------------------------------------
# select                                                      reference: cpu.py:794-831,865-903
# ------------------------------------------------------------------------------------------------
@dispatcher.register(UnaryOps.WHERE, CUDA)
def where(cond: Buffer, x: Buffer, y: Buffer) -> CDataPtr:
    xs, ys = x.scalar, y.scalar
    key = ("where", cond.metadata.shape, x.metadata.shape, y.metadata.shape, xs is None, ys is None,
           cond.aligned and x.aligned and y.aligned)
    ent = _plans.get(key)
    if ent is None:
        oshape = cond.metadata.shape
        strides = [calculate_stride(oshape)]
        names = []
        for b, sc, slot in ((x, xs, "f0"), (y, ys, "f1")):
            if sc is None:
                names.append(f"v{len(strides)}")
                strides.append(get_broadcast_strides_2(oshape, b.metadata.shape, calculate_stride(b.metadata.shape)))
            else:
                names.append(slot)
        k = ck.ewise(f"(v0 > 0.f) ? {names[0]} : {names[1]}", tuple(oshape), tuple(strides),
                     cond.aligned and x.aligned and y.aligned, "where")
        ent = _plans[key] = Plan(k)
    ptrs = [_dev_ptr(cond)] + [_dev_ptr(b) for b in (x, y) if b.scalar is None]
    while len(ptrs) < 3:
        ptrs.append(NULL)
    return ent.run(ptrs[0], ptrs[1], ptrs[2], NULL, xs or 0.0, ys or 0.0)


@dispatcher.register(UnaryOps.MASKED_FILL, CUDA)
def masked_fill(x: Buffer, mask: Buffer, val: Scalar, out_shape: tuple) -> CDataPtr:
    key = ("mfill", x.metadata.shape, mask.metadata.shape, tuple(out_shape), x.aligned and mask.aligned)
    plan = _plans.get(key)
    if plan is None:
        oshape = tuple(out_shape)
        sx = get_broadcast_strides_2(oshape, x.metadata.shape, calculate_stride(x.metadata.shape))
        sm = get_broadcast_strides_2(oshape, mask.metadata.shape, calculate_stride(mask.metadata.shape))
        plan = _plans[key] = Plan(ck.ewise("(v1 > 0.f) ? f0 : v0", oshape, (sx, sm),
                                           x.aligned and mask.aligned, "masked_fill"))
    return plan.run(_dev_ptr(x), _dev_ptr(mask), NULL, NULL, float(val))


# ------------------------------------------------------------------------------------------------
# creation                                                    reference: cpu.py:194-254
# ------------------------------------------------------------------------------------------------
@dispatcher.register(BufferOps.FULL, CUDA)
def full(shape: tuple, fill_value: Scalar) -> CDataPtr:
    n = prod_(shape)
    key = ("full", n)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.ewise("f0", (n,), (), True, "full"))
    # `int(fill_value)`: the reference's C kernel takes an int (cpu.py:252, cpu_kernel.py:758-770),
    # so Tensor.full(.., 0.5) is all zeros there — and here.
    return plan.run(NULL, NULL, NULL, NULL, float(int(fill_value)))


@dispatcher.register(BufferOps.ARANGE, CUDA)
def arange(shape: tuple) -> CDataPtr:
    n = prod_(shape)
    key = ("arange", n)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.ewise("ei", (n,), (), True, "arange"))
    return plan.run()


@dispatcher.register(BufferOps.UNIFORM, CUDA)
def uniform(shape: tuple, low: float = 0.0, high: float = 1.0) -> CDataPtr:
    """U[low, high) on the device: Philox4x32-10 keyed by manual_seed (codegen uniform_philox).  Every call draws from
    its own counter range, so weight initialisation and dropout masks never repeat; inside a captured graph the call
    number comes from device memory and a pre-replay hook advances it."""
    shim.init()
    n = prod_(shape)
    state = _rng_buffer()
    cap = shim.capture["rng"] if shim.capture["on"] else None
    key = ("uniform_philox", n, cap is not None)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.uniform_philox(n, cap is not None))
    if cap is not None:
        call = cap["n"]                                  # ordinal of this draw inside the captured step
        cap["n"] += 1
        if not cap["hooked"]:
            cap["hooked"] = True
            shim.capture["hooks"].append(_rng_replay_hook(cap))
    else:
        call = _rng_state["calls"]
        _rng_state["calls"] += 1
    return plan.run(state, NULL, NULL, NULL, float(low), float(high), float(call & 0xFFFFFF), float((call >> 24) & 0xFFFFFF))


def sample_last(x: Buffer, temperature: float = 1.0, u: Buffer | None = None) -> CDataPtr:
    """One index per sequence, drawn on the device from softmax(x[..., -1, :] / temperature) — the sampling step of
    examples/gpt.py:216-221,262-266, which the reference does on the host after a D2H of the logits.  x is (V,), (B, V)
    or (B, T, V); returns B floats holding the indices.  `u` (B uniform numbers) may be supplied for testing; by
    default they come from the backend's Philox stream (manual_seed)."""
    shape = x.metadata.shape
    V = shape[-1]
    if len(shape) == 3:
        rows, stride, first = shape[0], shape[1] * V, (shape[1] - 1) * V
    elif len(shape) == 2:
        rows, stride, first = shape[0], V, 0
    elif len(shape) == 1:
        rows, stride, first = 1, V, 0
    else:
        raise ValueError(f"sample_last expects (V,), (B, V) or (B, T, V) logits, got {shape}")
    if not temperature > 0.0:
        raise ValueError("temperature must be positive")
    if u is None:
        u = Buffer(uniform((rows,)), (rows,), CUDA)
    key = ("sample_rows", rows, stride, V)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.sample_rows(rows, stride, V))
    return plan.run(_dev_ptr(x) + first, _dev_ptr(u), NULL, NULL, float(temperature))


# ------------------------------------------------------------------------------------------------
# reductions                                                  reference: cpu.py:605-761,911-959
# ------------------------------------------------------------------------------------------------
def _reduce(op: str, x: Buffer, dim: int) -> CDataPtr:
    shape = x.metadata.shape
    if dim is None:
        dim = -1
    if dim != -1 and dim < 0:
        dim += len(shape)
    key = ("red", op, shape, dim, x.aligned)
    stages = _plans.get(key)
    if stages is None:
        stages = _plans[key] = _build_reduce(op, shape, dim, x.aligned)
    ptr = _dev_ptr(x)
    for plan in stages:
        ptr = plan.run(ptr)
    return ptr


def _build_reduce(op: str, shape: tuple, dim: int, aligned: bool) -> list[Plan]:
    numel = prod_(shape)
    first = "sum" if op == "mean" else op
    if dim == -1 or len(shape) == 0:
        outer, R, inner = 1, numel, 1
    else:
        outer, R, inner = prod_(shape[:dim]), shape[dim], prod_(shape[dim + 1:])
    if inner == 1:
        if outer == 1 and R >= 65536:
            k1 = ck.reduce_all_partial(first, R, aligned)
            G = k1.out_elems
            return [Plan(k1), Plan(ck.reduce_rows(op, 1, G, True, scale_R=R))]
        return [Plan(ck.reduce_rows(op, outer, R, aligned))]
    # strided axis: 128-column strips (each warp reads full 512-byte row segments), 8 warps interleave the
    # rows of a strip.  When outer x strips cannot occupy the chip, R is cut into `splits` row ranges until
    # ~148 x 8 CTAs exist, and a second, tiny pass adds the partials in split order.  Measured on 4096^2
    # (scripts/reduce_probe.py): 32 splits + second pass 13.4 us (5.0 TB/s); a thread-block cluster of 8
    # combining through DSMEM in one launch 22 us; row slabs 25 us.
    V = 4 if (aligned and inner % 4 == 0) else 1
    blocks = -(-(inner // V) // 32) * outer
    splits = 1
    if blocks < 4 * ck.NUM_SMS and R >= 64:
        splits = builtins_max(1, builtins_min(R // 16, -(-8 * ck.NUM_SMS // blocks), 64))
    if splits == 1:
        return [Plan(ck.reduce_cols(op, outer, R, inner, aligned))]
    return [Plan(ck.reduce_cols(first, outer, R, inner, aligned, splits)),
            Plan(ck.reduce_cols(op, 1, splits, outer * inner, aligned and (outer * inner) % 4 == 0, 1, scale_R=R))]


def rms_dw(g: Buffer, x: Buffer, rstd: Buffer) -> CDataPtr:
    """sum over rows of g * (x * rstd) (the RMSNorm weight gradient, dlgrad/nn/__init__.py:28-41 through autograd)
    in the column-reduce kernel itself: the (rows, C) product is never written.  Same strips / splits / order as
    `(g * (x * rstd)).sum(0)`, hence the same bits."""
    C = x.metadata.shape[-1]
    R = x.metadata.numel // C
    ok = g.aligned and x.aligned
    key = ("rms_dw", R, C, ok)
    plans = _plans.get(key)
    if plans is None:
        V = 4 if (ok and C % 4 == 0) else 1
        blocks = -(-(C // V) // 32)
        splits = 1
        if blocks < 4 * ck.NUM_SMS and R >= 64:
            splits = builtins_max(1, builtins_min(R // 16, -(-8 * ck.NUM_SMS // blocks), 64))
        plans = [Plan(ck.reduce_cols("sum", 1, R, C, ok, splits, prod3=True))]
        if splits > 1:
            plans.append(Plan(ck.reduce_cols("sum", 1, splits, C, C % 4 == 0, 1, scale_R=R)))
        _plans[key] = plans
    ptr = plans[0].run(_dev_ptr(g), _dev_ptr(x), _dev_ptr(rstd))
    for p in plans[1:]:
        ptr = p.run(ptr)
    return ptr


@dispatcher.register(UnaryOps.SUM, CUDA)
def sum_(x: Buffer, dim: int) -> CDataPtr:
    return _reduce("sum", x, dim)


@dispatcher.register(UnaryOps.MEAN, CUDA)
def mean(x: Buffer, dim: int) -> CDataPtr:
    return _reduce("mean", x, dim)


@dispatcher.register(UnaryOps.MAX, CUDA)
def max_(x: Buffer, dim: int) -> CDataPtr:
    return _reduce("max", x, dim)


@dispatcher.register(UnaryOps.ARGMAX, CUDA)
def argmax(x: Buffer, dim: int) -> CDataPtr:
    rows, cols = x.metadata.shape
    if dim is None:
        dim = -1
    key = ("argmax", rows, cols, dim)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.argmax(rows, cols, dim))
    return plan.run(_dev_ptr(x))


# ------------------------------------------------------------------------------------------------
# fused row ops
# ------------------------------------------------------------------------------------------------
def _rows_of(shape: tuple, dim: int) -> tuple[int, int] | None:
    if dim < 0:
        dim += len(shape)
    if dim != len(shape) - 1:
        return None
    R = shape[-1]
    if R > 65536:
        return None
    return prod_(shape[:-1]), R


def _softmax_family(kind: str, x: Buffer, dim: int) -> CDataPtr:
    rows = _rows_of(x.metadata.shape, dim)
    if rows is None:
        # not the last axis (or a very long row): the reference's composition
        # (dlgrad/tensor.py:634-669) on the primitive kernels
        mx = x.max(dim=dim, keepdim=True)
        sh = x - mx
        e = sh.exp()
        s = e.sum(dim=dim, keepdim=True)
        return (e / s).ptr if kind == "softmax" else (sh - s.log()).ptr
    if x._pending is not None:
        chain, base = _peel_chain(x)
        bp = base._pending
        if bp is not None and bp[0] == "mm" and kind == "softmax":
            from dlgrad_b200.runtime.cuda import attention
            m = attention.chain_matches(chain, x.metadata.shape)
            q, kt = bp[1], bp[2]
            if m is not None and kt._pending is not None and kt._pending[0] == "tr":
                return attention.scores(q, kt._pending[1], m[0], m[1], m[2])
        if bp is not None:
            materialize(base)                       # a deferred transpose / product under the chain: run it, keep folding
        if chain and base._pending is None:
            args = _row_chain_args(chain, x.metadata.shape, base.aligned)
            if args is not None:
                desc, masks, fl = args
                key = (kind + "_chain", rows, base.aligned, desc)
                plan = _plans.get(key)
                if plan is None:
                    plan = _plans[key] = Plan(ck.softmax_rows(kind, rows[0], rows[1], base.aligned, desc))
                mp = [_dev_ptr(m) for m in masks] + [NULL, NULL]
                return plan.run(_dev_ptr(base), NULL, mp[0], mp[1], fl[0], fl[1], fl[2], fl[3])
    key = (kind, rows, x.aligned)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.softmax_rows(kind, rows[0], rows[1], x.aligned))
    return plan.run(_dev_ptr(x))


def _softmax_bwd_family(kind: str, y: Buffer, grad: Buffer, dim: int) -> CDataPtr:
    rows = _rows_of(y.metadata.shape, dim)
    if rows is None:
        if kind == "softmax":
            return (y * (grad - (grad * y).sum(dim=dim, keepdim=True))).ptr
        return (grad - y.exp() * grad.sum(dim=dim, keepdim=True)).ptr
    key = (kind + "_bwd", rows, y.aligned and grad.aligned)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.softmax_bwd_rows(kind, rows[0], rows[1], y.aligned and grad.aligned))
    return plan.run(_dev_ptr(y), _dev_ptr(grad))


@dispatcher.register(UnaryOps.SOFTMAX, CUDA)
def softmax(x: Buffer, dim: int) -> CDataPtr:
    return _softmax_family("softmax", x, dim)


@dispatcher.register(UnaryOps.LOG_SOFTMAX, CUDA)
def log_softmax(x: Buffer, dim: int) -> CDataPtr:
    return _softmax_family("log_softmax", x, dim)


@dispatcher.register(UnaryOps.SOFTMAX_BACKWARD, CUDA)
def softmax_backward(y: Buffer, grad: Buffer, dim: int) -> CDataPtr:
    return _softmax_bwd_family("softmax", y, grad, dim)


@dispatcher.register(UnaryOps.LOG_SOFTMAX_BACKWARD, CUDA)
def log_softmax_backward(y: Buffer, grad: Buffer, dim: int) -> CDataPtr:
    return _softmax_bwd_family("log_softmax", y, grad, dim)


# ------------------------------------------------------------------------------------------------
# permute                                                     reference: cpu.py:334-368
# ------------------------------------------------------------------------------------------------
@dispatcher.register(UnaryOps.PERMUTE, CUDA)
def permute(x: Buffer, order: tuple) -> CDataPtr:
    shape = x.metadata.shape
    key = ("perm", shape, tuple(order), x.aligned)
    plan = _plans.get(key)
    if plan is None:
        in_str = calculate_stride(shape)
        out_shape = tuple(shape[i] for i in order)
        pstr = tuple(in_str[i] for i in order)
        # collapse first: (B,T,h,d)->(0,2,1,3) keeps d innermost, (1,128,96)->(0,2,1) is 2-D, ...
        cshape, cstr = ck.collapse_dims(out_shape, [calculate_stride(out_shape), pstr])
        cin = cstr[1]
        if cin[-1] == 1 or len(cshape) == 1 or prod_(shape) < 1024:
            k = ck.ewise("v0", out_shape, (pstr,), x.aligned, "permute")
        else:
            tdim = cin.index(1) if 1 in cin else None
            if tdim is None:
                k = ck.ewise("v0", out_shape, (pstr,), x.aligned, "permute")
            else:
                k = ck.transpose_tiled(tuple(cshape), tuple(cin), tdim)
        plan = _plans[key] = Plan(k, prod_(shape))
    return plan.run(_dev_ptr(x))


# ------------------------------------------------------------------------------------------------
# concat (no reference kernel: examples/gpt.py:140-147 round-trips the KV cache through numpy)
# ------------------------------------------------------------------------------------------------
@dispatcher.register(CustomOps.CAT, CUDA)
def cat(x: Buffer, y: Buffer, dim: int) -> CDataPtr:
    xs, ys = x.metadata.shape, y.metadata.shape
    outer = prod_(xs[:dim])
    la, lb = prod_(xs[dim:]), prod_(ys[dim:])
    ok = x.aligned and y.aligned
    key = ("cat2", outer, la, lb, ok)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.cat2(outer, la, lb, ok))
    return plan.run(_dev_ptr(x), _dev_ptr(y))


# ------------------------------------------------------------------------------------------------
# matmul                                                      reference: cpu.py:505-602
# ------------------------------------------------------------------------------------------------
@dispatcher.register(BinaryOps.MATMUL, CUDA)
def matmul(x: Buffer, y: Buffer, trans_x: bool = False, trans_y: bool = False, bias: Buffer | None = None,
           relu_mask: Buffer | None = None):
    """With `bias` or `relu_mask` (CUDA extensions used by ops.Linear) returns (ptr, applied)."""
    from dlgrad_b200.runtime.cuda import matmul as _mm
    return _mm.run_matmul(x, y, trans_x, trans_y, bias, relu_mask)


# ------------------------------------------------------------------------------------------------
# cross-entropy pieces and embedding                          reference: cpu.py:962-1061
# ------------------------------------------------------------------------------------------------
@dispatcher.register(CustomOps.CE_FORWARD, CUDA)
def ce_forward(x: Buffer, y: Buffer) -> CDataPtr:
    n, s0 = x.metadata.shape[0], x.metadata.shape[1]
    key = ("ce_fwd", n, s0)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.ce_forward(n, s0))
    return plan.run(_dev_ptr(x), _dev_ptr(y))


@dispatcher.register(CustomOps.CE_BACKWARD, CUDA)
def ce_backward(x: Buffer, target: Buffer) -> None:
    n, s0 = x.metadata.shape[0], x.metadata.shape[1]
    flush_pending((x,))
    key = ("ce_bwd", n, s0)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.ce_backward(n, s0))
    plan.run_into(NULL, _dev_ptr(x), _dev_ptr(target))


@dispatcher.register(CustomOps.CE_GRAD, CUDA)
def ce_grad(logsm: Buffer, target: Buffer, upstream: Buffer, n: float) -> CDataPtr:
    rows, V = logsm.metadata.shape
    up_ptr = upstream.scalar is None
    key = ("ce_grad", rows, V, up_ptr)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.ce_grad(rows, V, up_ptr))
    if up_ptr:
        return plan.run(_dev_ptr(logsm), _dev_ptr(target), _dev_ptr(upstream), NULL, 0.0, float(n))
    return plan.run(_dev_ptr(logsm), _dev_ptr(target), NULL, NULL, upstream.scalar, float(n))


@dispatcher.register(CustomOps.EMBEDDING, CUDA)
def embedding(x: Buffer, idx: Buffer, out_numel: int, backward: bool = False,
              upstream_grad: Buffer | None = None) -> CDataPtr:
    C = x.metadata.shape[-1]
    n_idx = idx.metadata.numel
    if not backward:
        key = ("emb", n_idx, C, x.aligned)
        plan = _plans.get(key)
        if plan is None:
            plan = _plans[key] = Plan(ck.embedding_fwd(n_idx, C, x.aligned))
        return plan.run(_dev_ptr(x), _dev_ptr(idx))
    rows = x.metadata.shape[0]
    key = ("emb_bwd", rows, n_idx, C)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.embedding_bwd(rows, n_idx, C))
    return plan.run(_dev_ptr(upstream_grad), _dev_ptr(idx))


# ------------------------------------------------------------------------------------------------
# optimizer steps (in place)                                  reference: dlgrad/nn/optim.py:26-70
# ------------------------------------------------------------------------------------------------
@dispatcher.register(CustomOps.ADAM_STEP, CUDA)
def adam_step(param: Buffer, grad: Buffer, m: Buffer, v: Buffer, beta1: float, beta2: float,
              eps: float, bc1: float, bc2: float, lr: float, grad_scale: float = 1.0, scalars=None) -> None:
    ok = param.aligned and grad.aligned and m.aligned and v.aligned
    key = ("adam", param.metadata.numel, beta1, beta2, eps, ok, scalars is not None)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.adam_step(param.metadata.numel, beta1, beta2, eps, ok, scalars is not None))
    plan.run_into(NULL if scalars is None else scalars, _dev_ptr(param), _dev_ptr(grad), _dev_ptr(m), _dev_ptr(v),
                  bc1, bc2, lr, grad_scale)


_adam_tables: dict = {}


@dispatcher.register(CustomOps.ADAM_MULTI, CUDA)
def adam_multi(params: list, grads: list, ms: list, vs: list, beta1: float, beta2: float, eps: float,
               bc1: float, bc2: float, lr: float, grad_scale: float = 1.0, scalars=None) -> bool:
    """One launch for all parameter tensors (codegen adam_multi).  The address table (4 pointers per tensor)
    is staged in pinned host memory and copied asynchronously ahead of the kernel on the compute stream; two
    staging buffers alternate so a table is never rewritten while its copy may still be in flight.  Returns
    False (nothing done) when a tensor does not qualify for the 128-bit path.

    `scalars` (a device float[4], see adam_scalars): the kernel reads bc1, bc2 and lr from it instead of its launch
    arguments — the form recorded into a captured graph, whose staging table is private to that capture."""
    bufs = list(params) + list(grads) + list(ms) + list(vs)
    if not all(b.aligned and b.metadata.numel % 4 == 0 for b in bufs):
        return False
    sizes = tuple(p.metadata.numel for p in params)
    token = id(shim.capture["keep"]) if shim.capture["on"] else 0
    key = ("adam_multi", sizes, beta1, beta2, eps, scalars is not None, token)
    ent = _adam_tables.get(key)
    if ent is None:
        n = len(sizes)
        plan = Plan(ck.adam_multi(sizes, beta1, beta2, eps, scalars is not None))
        host = [lib.dlg_host_alloc(32 * n), lib.dlg_host_alloc(32 * n)]
        if host[0] == NULL or host[1] == NULL:
            raise MemoryError("pinned staging allocation failed: " + shim.last_error())
        views = [ffi.cast("unsigned long long *", h) for h in host]
        ent = _adam_tables[key] = {"plan": plan, "host": host, "views": views, "dev": shim.alloc(32 * n), "flip": 0,
                                   "events": [None, None], "last": None}
    addr = shim.address
    table = []
    for p, g, m, v in zip(params, grads, ms, vs):
        table += [addr(_dev_ptr(p)), addr(_dev_ptr(g)), addr(_dev_ptr(m)), addr(_dev_ptr(v))]
    if table != ent["last"]:
        # the addresses moved (first step, or the allocator handed out other gradient blocks): upload a new table.  The
        # host does not wait for the device between steps, so the staging buffer about to be rewritten may belong to a
        # copy that has not executed yet: each upload is followed by an event, and the event is waited for before its
        # buffer is reused.  In the steady state the table is unchanged and nothing is uploaded at all.
        ent["flip"] ^= 1
        f = ent["flip"]
        if ent["events"][f] is not None:
            shim.check(lib.dlg_event_sync(ent["events"][f]), "dlg_event_sync")
        tab = ent["views"][f]
        for i, a in enumerate(table):
            tab[i] = a
        shim.check(lib.dlg_h2d_async(ent["dev"], ent["host"][f], 32 * len(sizes)), "dlg_h2d_async")
        if not shim.capture["on"]:                 # a captured table is private to its graph and recorded exactly once
            if ent["events"][f] is None:
                ent["events"][f] = lib.dlg_event_create()
            shim.check(lib.dlg_event_record(ent["events"][f]), "dlg_event_record")
        ent["last"] = table
    ent["plan"].run_into(NULL, ent["dev"], NULL if scalars is None else scalars, NULL, NULL, bc1, bc2, lr, grad_scale)
    return True


def adam_scalars(dst, bc1: float, bc2: float, lr: float):
    """Park this step's bias corrections and learning rate in device memory (`dst`, or a fresh float[4])."""
    plan = _plans.get("set_scalars")
    if plan is None:
        plan = _plans["set_scalars"] = Plan(ck.set_scalars())
    if dst is None:
        dst = shim.alloc(16)
    plan.run_into(dst, NULL, NULL, NULL, NULL, bc1, bc2, lr, 0.0)
    return dst


@dispatcher.register(CustomOps.SGD_STEP, CUDA)
def sgd_step(param: Buffer, grad: Buffer, lr: float, grad_scale: float = 1.0) -> None:
    ok = param.aligned and grad.aligned
    key = ("sgd", param.metadata.numel, ok)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.sgd_step(param.metadata.numel, ok))
    plan.run_into(NULL, _dev_ptr(param), _dev_ptr(grad), NULL, NULL, lr, 0.0, 0.0, grad_scale)


def sgd_momentum_step(param: Buffer, grad: Buffer, vel: Buffer, lr: float, momentum: float, grad_scale: float = 1.0) -> None:
    """SGD with momentum as ONE in-place kernel (velocity and parameter both updated where they live)."""
    ok = param.aligned and grad.aligned and vel.aligned
    key = ("sgd_mom", param.metadata.numel, ok)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.sgd_momentum_step(param.metadata.numel, ok))
    plan.run_into(NULL, _dev_ptr(param), _dev_ptr(grad), _dev_ptr(vel), NULL, lr, momentum, 0.0, grad_scale)


@dispatcher.register(CustomOps.RMSNORM, CUDA)
def rmsnorm(x: Buffer, weight: Buffer | None, eps: float):
    """Fused RMSNorm forward -> (y, rstd); nn.RMSNorm composes six kernels (dlgrad/nn/__init__.py:28-41)."""
    shape = x.metadata.shape
    C, rows = shape[-1], prod_(shape[:-1])
    ok = x.aligned and (weight is None or weight.aligned)
    key = ("rmsnorm", rows, C, weight is not None, ok)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = Plan(ck.rmsnorm_rows(rows, C, weight is not None, ok))
    rstd = shim.alloc(rows * 4)
    y = plan.run(_dev_ptr(x), _dev_ptr(weight) if weight is not None else NULL, rstd, NULL, float(eps))
    return y, rstd


def ewise3(tag: str, expr: str, out_shape: tuple, operands: list[Buffer]) -> CDataPtr:
    """out = expr(v0, v1, v2) with every operand broadcast to out_shape (helper for fused backward formulas)."""
    ok = all(b.aligned for b in operands)
    key = ("ew3", tag, expr, out_shape, tuple(b.metadata.shape for b in operands), ok)
    plan = _plans.get(key)
    if plan is None:
        strides = tuple(get_broadcast_strides_2(out_shape, b.metadata.shape, calculate_stride(b.metadata.shape))
                        for b in operands)
        plan = _plans[key] = Plan(ck.ewise(expr, tuple(out_shape), strides, ok, tag))
    pp = [_dev_ptr(b) for b in operands] + [NULL] * 3
    return plan.run(pp[0], pp[1], pp[2], pp[3])


@dispatcher.register(CustomOps.PRINT, CUDA)
def show(x: Buffer) -> None:
    print(x.numpy())


def launch_count() -> int:
    return int(lib.dlg_launch_count())


def fused_add(r: Buffer, d: Buffer):
    """r + d where d is a deferred producer that can absorb the addition: a Linear output ("lin": x @ W^T), a Linear input
    gradient ("lin_dx": g @ W) — the GEMM adds r in its epilogue — or an RMSNorm input gradient ("rms_dx": one elementwise
    kernel).  Returns the owned pointer of the sum, or None when the pattern does not apply."""
    pend = d._pending
    if pend is None or r.scalar is not None or r.metadata.shape != d.metadata.shape or r.metadata.device is not CUDA:
        return None
    if r._pending is not None:
        materialize(r)
    if pend[0] in ("lin", "lin_dx"):
        from dlgrad_b200.runtime.cuda import matmul as _mm
        x2, w = pend[1], pend[2]
        if pend[0] == "lin":
            if pend[3].ptr is not None:
                return None                         # already computed (a sibling's grouped launch): plain add
            pend[3].leave()                         # this product is absorbed here, not by a sibling's grouped launch
        return _mm.radd_matmul(x2, w, False, pend[0] == "lin", r.reshape((x2.metadata.shape[0], -1)))
    if pend[0] == "rms_dx":
        g, w, t = pend[1], pend[2], pend[3]
        if g._pending is not None:
            materialize(g)
        return ewise3("rms_dx_add", "v3 + (v0 * v1) * v2", d.metadata.shape, [g, w, t, r])
    return None
