"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce over NVLink/NVSwitch.

No reference counterpart (dlgrad has no distributed code).  The GPT train step shards by
batch: every rank holds identical weights, runs forward/backward on its own sequences, and the only
exchange is ONE all-reduce of the gradients per step.  Gradients are packed into a single persistent
flat fp32 bucket (one d2d copy each), reduced in place with ncclAllReduce(sum) on the comm stream,
and the parameters' .grad are re-pointed at slices of the bucket; the 1/world averaging is folded
into the fused Adam/SGD kernel via `Optimizer.grad_scale` (exact for power-of-two world sizes).

The rendezvous (sharing the 128-byte ncclUniqueId) is left to the launcher's side channel; bench.py
uses torch.distributed's gloo store for that and for the timing barrier — plumbing only.
"""
from __future__ import annotations

from dlgrad_b200.buffer import Buffer
from dlgrad_b200.device import Device
from dlgrad_b200.runtime.cuda import shim
from dlgrad_b200.tensor import Tensor

_state = {"rank": 0, "world": 1, "bucket": None, "bucket_elems": 0}


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
