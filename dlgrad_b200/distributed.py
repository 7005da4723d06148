"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce over NVLink/NVSwitch.

No reference counterpart (dlgrad has no distributed code).  The GPT train step shards by
batch: every rank holds identical weights, runs forward/backward on its own sequences, and the only
exchange is ONE all-reduce of the gradients per step.  Gradients are packed into a single persistent
flat fp32 bucket (one d2d copy each), reduced in place with ncclAllReduce(sum) on the comm stream,
and the parameters' .grad are re-pointed at slices of the bucket; the 1/world averaging is folded
into the fused Adam/SGD kernel via `Optimizer.grad_scale` (exact for power-of-two world sizes).

`distributed.Adam` goes one step further ("fused with its collective"): gradient buckets, parameter
buckets and a few flag words live in peer-mappable memory (dlg_ipc_*), every rank maps every peer's, and ONE
kernel per rank does reduce-scatter (NVLink loads) + Adam on its 1/world shard + all-gather (NVLink stores) —
codegen/cuda_kernel.dp_adam.  Optimizer state is sharded (1/world of m and v per rank), replicas stay
bit-identical by construction, and NCCL is used only once, to exchange the 64-byte memory handles.

The rendezvous (sharing the 128-byte ncclUniqueId) is left to the launcher's side channel; bench.py
uses torch.distributed's gloo store for that and for the timing barrier — plumbing only.
"""
from __future__ import annotations

import os

import numpy as np

from dlgrad_b200.buffer import Buffer, flush_pending
from dlgrad_b200.device import Device
from dlgrad_b200.runtime.cuda import shim
from dlgrad_b200.tensor import Tensor

_state = {"rank": 0, "world": 1, "bucket": None, "bucket_elems": 0}
# seconds a rank waits inside the fused step for a peer before it traps (a dead peer must not hang the node).  The trap
# is a sticky context failure on every waiting rank, so the limit is sized for a DEAD peer, not for ordinary skew: a
# rank that evaluates, writes a checkpoint or JIT-compiles a cold kernel may legitimately be minutes late.
PEER_TIMEOUT_S = float(os.environ.get("DLGRAD_DP_TIMEOUT_S", "600"))


def unique_id() -> bytes:
    buf = shim.ffi.new("char[128]")
    shim.check(shim.lib.dlg_nccl_unique_id(buf), "dlg_nccl_unique_id")
    return bytes(shim.ffi.buffer(buf, 128))


def init(rank: int, world: int, uid: bytes) -> None:
    shim.init()
    shim.check(shim.lib.dlg_nccl_init(rank, world, uid), "dlg_nccl_init")
    _state["rank"], _state["world"] = rank, world


def world_size() -> int:
    return _state["world"]


def rank() -> int:
    return _state["rank"]


def shard_offsets(sizes: list[int], align: int = 64) -> list[int]:
    """Element offsets of each gradient inside the flat bucket (64-element = 256 B aligned so every
    slice keeps 128-bit vector alignment)."""
    offs, acc = [], 0
    for n in sizes:
        offs.append(acc)
        acc += -(-n // align) * align
    offs.append(acc)
    return offs


def all_reduce_gradients(params: list[Tensor]) -> None:
    """Sum .grad of every parameter across ranks (in one NCCL call) and leave the summed gradients in
    the parameters' .grad.  With world == 1 this is a no-op."""
    world = _state["world"]
    live = [p for p in params if p.grad is not None]
    if world == 1 or not live:
        return
    shim.no_capture("all_reduce_gradients() (the collective runs on its own stream; capture forward + backward only)")
    sizes = [p.grad.numel for p in live]
    offs = shard_offsets(sizes)
    total = offs[-1]
    if _state["bucket"] is None or _state["bucket_elems"] != total:
        _state["bucket"] = shim.alloc(total * 4)
        _state["bucket_elems"] = total
        shim.check(shim.lib.dlg_memset32(_state["bucket"], 0, total), "dlg_memset32")
    bucket = _state["bucket"]
    for p, off, n in zip(live, offs, sizes):
        g = p.grad.data
        if g.metadata.device is not Device.CUDA:
            g._to_cuda()
        shim.check(shim.lib.dlg_d2d(bucket + off, g.ptr, n * 4), "dlg_d2d")
    shim.check(shim.lib.dlg_nccl_allreduce_sum(bucket, total), "dlg_nccl_allreduce_sum")
    shim.check(shim.lib.dlg_nccl_wait(), "dlg_nccl_wait")
    for p, off in zip(live, offs):
        view = Buffer(bucket + off, p.grad.shape, Device.CUDA, p.grad.dtype, keep=bucket)
        p.grad = Tensor(view)


# ---------------------------------------------------------------------------------------------------------
# fused reduce-scatter + Adam + all-gather over peer memory
# ---------------------------------------------------------------------------------------------------------
def bucket_layout(sizes: list[int], world: int) -> tuple[list[int], int, int]:
    """(offsets, total, shard) in elements: tensors at 64-element boundaries like shard_offsets, the total padded
    so that it splits into `world` shards of whole float4s."""
    offs = shard_offsets(sizes)
    quantum = 4 * world
    total = -(-offs[-1] // quantum) * quantum
    return offs[:-1], total, total // world


def all_gather_bytes(payload: bytes) -> list[bytes]:
    """Every rank's `payload` (equal lengths), over the NCCL communicator that already exists: each byte travels as
    one float in this rank's row of a zero matrix, and the sum all-reduce of such rows is exact."""
    world, rank_ = _state["world"], _state["rank"]
    n = len(payload)
    host = np.zeros((world, n), dtype=np.float32)
    host[rank_] = np.frombuffer(payload, dtype=np.uint8)
    from dlgrad_b200.runtime.cuda import transfer
    dev = transfer.upload_numpy(host)
    shim.check(shim.lib.dlg_nccl_allreduce_sum(dev, world * n), "dlg_nccl_allreduce_sum")
    shim.check(shim.lib.dlg_nccl_wait(), "dlg_nccl_wait")
    out = transfer.download(dev, world * n).reshape(world, n)
    return [out[r].astype(np.uint8).tobytes() for r in range(world)]


def prebuild(sizes: list[int], worlds: tuple = (2, 4, 8), betas: tuple = (0.9, 0.999), eps: float = 1e-8) -> None:
    """Compile the fused step for the given parameter sizes and world sizes into the kernel cache (build machine)."""
    from dlgrad_b200.codegen import cuda_kernel as ck
    from dlgrad_b200.runtime.cuda import compiler
    for w in worlds:
        offs, _, shard = bucket_layout(sizes, w)
        for k in (ck.dp_adam(w, shard // 4, betas[0], betas[1], eps, PEER_TIMEOUT_S), ck.pack_multi(tuple(sizes), tuple(offs))):
            compiler.get_function(k.name, k.source)


def peer_memory_available() -> bool:
    """Collective probe: can every rank map every other rank's memory (cuIpc handles, peer access over NVLink)?
    All ranks get the same answer, so they can agree to use ncclAllReduce + nn.optim.Adam instead when some
    container or topology forbids it."""
    world, rank_ = _state["world"], _state["rank"]
    if world == 1:
        return True
    ffi, lib = shim.ffi, shim.lib
    shim.init()
    h = ffi.new("unsigned char[64]")
    ptr = lib.dlg_ipc_alloc(1 << 21)
    ok = ptr != shim.NULL and lib.dlg_ipc_export(ptr, h) == 0
    rows = all_gather_bytes(bytes(ffi.buffer(h, 64)) + bytes([int(ok)]))
    if all(r[64] == 1 for r in rows):
        for r in range(world):
            if r == rank_:
                continue
            m = lib.dlg_ipc_open(ffi.new("unsigned char[64]", rows[r][:64]))
            if m == shim.NULL:
                ok = False
                break
            lib.dlg_ipc_close(m)
    else:
        ok = False
    verdict = all(r[0] == 1 for r in all_gather_bytes(bytes([int(ok)])))     # also: nobody frees before all closed
    if ptr != shim.NULL:
        lib.dlg_ipc_free(ptr)
    return verdict


class _PeerBlock:
    """One dlg_ipc_alloc'd block and its peers' mappings; `addr[r]` is rank r's block as seen from this process."""

    def __init__(self, nbytes: int) -> None:
        ffi, lib = shim.ffi, shim.lib
        shim.init()
        self.ptr = lib.dlg_ipc_alloc(nbytes)
        if self.ptr == shim.NULL:
            raise MemoryError(f"peer-memory allocation of {nbytes} bytes failed: {shim.last_error()}")
        self.fptr = ffi.cast("float *", self.ptr)
        world, rank_ = _state["world"], _state["rank"]
        self.mapped = []
        self.addr = [0] * world
        self.addr[rank_] = shim.address(self.ptr)
        if world > 1:
            h = ffi.new("unsigned char[64]")
            shim.check(lib.dlg_ipc_export(self.ptr, h), "dlg_ipc_export")
            handles = all_gather_bytes(bytes(ffi.buffer(h, 64)))
            for r in range(world):
                if r == rank_:
                    continue
                m = lib.dlg_ipc_open(ffi.new("unsigned char[64]", handles[r]))
                if m == shim.NULL:
                    raise RuntimeError(f"mapping rank {r}'s memory failed: {shim.last_error()} "
                                       "(peers must be GPUs of one node, all visible to every process)")
                self.mapped.append(m)
                self.addr[r] = shim.address(m)

    def close(self) -> None:
        for m in self.mapped:
            shim.lib.dlg_ipc_close(m)
        self.mapped = []
        if self.ptr is not None:
            shim.lib.dlg_ipc_free(self.ptr)
            self.ptr = None


class Adam:
    """nn.optim.Adam (dlgrad/nn/optim.py:46-70) for data-parallel replicas: same constructor, `zero_grad()` and
    `step()`; step() averages the gradients over the ranks and applies the update in one kernel per rank.

    Construction re-homes every parameter into one flat peer-mappable bucket (values are kept, the Buffers keep
    their identity) and is collective: every rank must construct it with the same parameter list.  With
    world == 1 it degenerates to a single-GPU fused Adam over the flat bucket (same arithmetic as adam_multi)."""

    def __init__(self, params: list[Tensor], lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8) -> None:
        from dlgrad_b200.codegen import cuda_kernel as ck
        from dlgrad_b200.runtime.cuda import ops as cuda_ops
        self.params = params
        self.lr, (self.beta1, self.beta2), self.eps = lr, betas, eps
        self.t = 0
        world = _state["world"]
        self.sizes = [p.numel for p in params]
        self.offs, self.total, self.shard = bucket_layout(self.sizes, world)
        flush_pending(tuple(p.data for p in params))
        self.grads = _PeerBlock(self.total * 4)
        self.weights = _PeerBlock(self.total * 4)
        self.flags = _PeerBlock(4096)
        lib = shim.lib
        for p, off, n in zip(params, self.offs, self.sizes):
            buf = p.data
            if buf.metadata.device is not Device.CUDA:
                buf._to_cuda()
            shim.check(lib.dlg_d2d(self.weights.fptr + off, buf.ptr, n * 4), "dlg_d2d")
            buf.ptr = self.weights.fptr + off
            buf._keep = self.weights
            buf.aligned = True
        shim.synchronize()                        # the old parameter storage is released only after the copies
        table = np.array(self.grads.addr + self.weights.addr + self.flags.addr + [_state["rank"]], dtype=np.uint64)
        from dlgrad_b200.runtime.cuda import transfer
        self.table = transfer.upload(shim.ffi.from_buffer("char *", table), table.nbytes)
        self.m = shim.calloc(self.shard * 4)
        self.v = shim.calloc(self.shard * 4)
        self.plan = cuda_ops.Plan(ck.dp_adam(world, self.shard // 4, self.beta1, self.beta2, self.eps, PEER_TIMEOUT_S))
        # gradient packing: one launch over a table of source addresses (two pinned staging tables alternate, so
        # one is never rewritten while its copy may still be in flight)
        n = len(params)
        self.pack = cuda_ops.Plan(ck.pack_multi(tuple(self.sizes), tuple(self.offs)))
        self.pack_host = [lib.dlg_host_alloc(8 * n), lib.dlg_host_alloc(8 * n)]
        if shim.NULL in self.pack_host:
            raise MemoryError("pinned staging allocation failed: " + shim.last_error())
        self.pack_view = [shim.ffi.cast("unsigned long long *", h) for h in self.pack_host]
        self.pack_dev = shim.alloc(8 * n)
        self.flip = 0
        self.pack_events = [None, None]
        self.pack_last = None
        shim.synchronize()
        if world > 1:
            all_gather_bytes(b"\0")                # nobody launches before every rank has mapped its peers

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    def step(self) -> None:
        shim.no_capture("distributed.Adam.step() (ranks wait for each other inside the kernel; capture forward + backward only)")
        self.t += 1
        f32 = np.float32
        bc1 = float(f32(1.0 - self.beta1 ** self.t))
        bc2 = float(f32(1.0 - self.beta2 ** self.t))
        lib = shim.lib
        grads = [None if p.grad is None else p.grad.data for p in self.params]
        for g in grads:
            if g is not None and g.metadata.device is not Device.CUDA:
                g._to_cuda()
        if all(g is not None and g.aligned for g in grads):
            table = [shim.address(g.ptr) for g in grads]
            if table != self.pack_last:
                # new gradient addresses: upload the table.  The staging buffer about to be rewritten may still belong
                # to an H2D copy the device has not executed (the host runs ahead), so every upload is followed by an
                # event that is waited for before its buffer is reused; an unchanged table is not uploaded at all.
                self.flip ^= 1
                f = self.flip
                if self.pack_events[f] is not None:
                    shim.check(lib.dlg_event_sync(self.pack_events[f]), "dlg_event_sync")
                view = self.pack_view[f]
                for i, a in enumerate(table):
                    view[i] = a
                shim.check(lib.dlg_h2d_async(self.pack_dev, self.pack_host[f], 8 * len(grads)), "dlg_h2d_async")
                if self.pack_events[f] is None:
                    self.pack_events[f] = lib.dlg_event_create()
                shim.check(lib.dlg_event_record(self.pack_events[f]), "dlg_event_record")
                self.pack_last = table
            self.pack.run_into(self.grads.fptr, self.pack_dev)
        else:
            for g, off, n in zip(grads, self.offs, self.sizes):
                dst = self.grads.fptr + off
                if g is None:
                    shim.check(lib.dlg_memset32(dst, 0, n), "dlg_memset32")
                else:
                    shim.check(lib.dlg_d2d(dst, g.ptr, n * 4), "dlg_d2d")
        flush_pending(tuple(p.data for p in self.params))
        self.plan.run_into(shim.NULL, self.table, self.m, self.v, self.flags.fptr, bc1, bc2, float(f32(self.lr)),
                           1.0 / _state["world"])

    def close(self) -> None:
        shim.synchronize()
        for blk in (self.grads, self.weights, self.flags):
            blk.close()
