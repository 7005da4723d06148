"""Data-parallel host logic on CPU: world_size-2 `gloo` run (no GPU).

What is checked
  * the flat-bucket layout used by dlgrad_b200.distributed (offsets, alignment, coverage);
  * the DP arithmetic the GPU path relies on: each rank back-propagates its OWN half of the batch on the
    CPU oracle, gradients are summed across ranks (gloo all_reduce standing in for ncclAllReduce) and the
    optimizer runs with grad_scale = 1/world — the updated weights must equal a single-process step on
    the concatenated batch (the reference's CE gradient is a per-rank mean);
  * rank-dependent data seeds and identical weight seeds, as bench.py sets them up.
"""
import os
import socket

import numpy as np
import pytest

from tests import models


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor, nn
    from dlgrad_b200.distributed import shard_offsets
    from oracle import backend
    backend.install()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = models.GPTConfig(vocab_size=17, block_size=16, n_layer=1, n_head=2, n_embd=32, device="cpu")
    gpt = models.GPT(dl, cfg, np.random.default_rng(0))          # identical weights on every rank
    params = nn.utils.get_parameters(gpt)
    opt = nn.optim.Adam(params, lr=1e-3)
    opt.grad_scale = 1.0 / world
    full_x, full_y = models.gpt_batch(cfg, 4 * world, np.random.default_rng(7))
    x, y = full_x[rank * 4:(rank + 1) * 4], full_y[rank * 4:(rank + 1) * 4]   # this rank's shard of the batch
    opt.zero_grad()
    _, loss = gpt(Tensor(x.copy()), Tensor(y.copy()))
    loss.backward()
    live = [p for p in params if p.grad is not None]
    sizes = [p.grad.numel for p in live]
    offs = shard_offsets(sizes)
    flat = np.zeros(offs[-1], dtype=np.float32)
    for p, o, n in zip(live, offs, sizes):
        flat[o:o + n] = p.grad.numpy().reshape(-1)
    t = torch.from_numpy(flat)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                      # the one exchange per step
    for p, o, n in zip(live, offs, sizes):
        p.grad = Tensor(flat[o:o + n].reshape(p.grad.shape).copy())
    opt.step()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=float(loss.numpy()),
             wte=gpt.wte.weight.numpy(), q=gpt.blocks[0].attn.q_proj.weight.numpy(), head=gpt.lm_head.weight.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_offsets_layout():
    from dlgrad_b200.distributed import shard_offsets
    sizes = [96 * 768, 1, 7, 768, 3072 * 768]
    offs = shard_offsets(sizes)
    assert len(offs) == len(sizes) + 1 and offs[0] == 0
    for o, n, nxt in zip(offs, sizes, offs[1:]):
        assert o % 64 == 0 and nxt >= o + n          # 256-byte aligned, non-overlapping
    assert offs[-1] % 64 == 0


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_bucket_layout_and_sharded_update_cover_every_parameter(world):
    """distributed.Adam's layout: the flat bucket splits into `world` shards of whole float4s whose boundaries may
    cut through tensors.  A numpy walk of the kernel's schedule (sum the ranks' buckets in rank order on the owner's
    shard, update, store the shard to every replica) must touch every element exactly once and equal the
    all-reduce + per-tensor update it replaces."""
    from dlgrad_b200.distributed import bucket_layout
    sizes = [96 * 32, 1, 7, 130, 4 * 32 * 32, 33]
    offs, total, shard = bucket_layout(sizes, world)
    assert total == shard * world and shard % 4 == 0 and total >= offs[-1] + sizes[-1]
    assert all(o % 64 == 0 for o in offs)
    rng = np.random.default_rng(world)
    grads = rng.standard_normal((world, total)).astype(np.float32)
    weights = np.tile(rng.standard_normal(total).astype(np.float32), (world, 1))       # identical replicas
    want = weights[0] - np.float32(0.1) * (np.add.reduce(grads, axis=0, dtype=np.float32) * np.float32(1.0 / world))
    touched = np.zeros(total, dtype=np.int32)
    for owner in range(world):
        lo, hi = owner * shard, (owner + 1) * shard
        g = grads[0, lo:hi].copy()
        for r in range(1, world):
            g += grads[r, lo:hi]
        new = weights[owner, lo:hi] - np.float32(0.1) * (g * np.float32(1.0 / world))
        for r in range(world):
            weights[r, lo:hi] = new
        touched[lo:hi] += 1
    assert (touched == 1).all()
    for r in range(world):
        assert np.array_equal(weights[r], weights[0])
    np.testing.assert_allclose(weights[0], want, rtol=1e-6, atol=1e-7)
    for o, n in zip(offs, sizes):                          # every tensor lies inside the bucket, in order
        assert o + n <= total


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_exchange_buckets_partition_the_parameters_in_backward_order(world):
    """distributed.plan_buckets (overlapped exchange): buckets take the parameters in reverse registration order, every
    parameter lies in exactly one bucket, buckets and tensors start on 256-byte boundaries of the flat buffers, every
    bucket splits into `world` shards of whole float4s, and the per-rank moment slices tile the moment arrays."""
    from dlgrad_b200.distributed import plan_buckets
    sizes = [96 * 32, 1024 * 32] + [32, 32 * 32, 32 * 32, 32 * 32, 32 * 32, 32, 128 * 32, 32 * 128] * 3 + [32, 96 * 32]
    buckets = plan_buckets(sizes, world, 3000)
    order = [i for b in buckets for i in b["params"]]
    assert order == list(reversed(range(len(sizes))))
    assert len(buckets) > 3
    flat_end, m_end = 0, 0
    for b in buckets:
        assert b["start"] == flat_end and b["start"] % 64 == 0 and b["mstart"] == m_end
        assert b["total"] == b["shard"] * world and b["shard"] % 4 == 0
        for i, o, nxt in zip(b["params"], b["offs"], b["offs"][1:] + [b["total"]]):
            assert o % 64 == 0 and o + sizes[i] <= nxt
        flat_end += b["total"]
        m_end += b["shard"]
    # all but the last bucket reach the threshold; one huge tensor is a bucket of its own
    assert all(sum(sizes[i] for i in b["params"]) >= 3000 for b in buckets[:-1])


def test_overlapped_exchange_is_enqueued_during_backward():
    """Host-side schedule of distributed.Adam (overlapped) walked with the device stubbed out: the first step learns the
    number of gradient contributions per parameter and exchanges everything in step(); from the second step on every
    bucket but the last ones is enqueued while backward is still running, one bucket late, each exactly once, and
    a second backward before step() fails loudly."""
    import subprocess
    import sys
    code = r"""
import numpy as np
import dlgrad_b200 as dl
from dlgrad_b200 import Tensor, distributed, nn
from dlgrad_b200.runtime.cuda import shim
from tests import models
cfg = models.GPTConfig(vocab_size=96, block_size=32, n_layer=3, n_head=2, n_embd=32, device="cuda")
gpt = models.GPT(dl, cfg, np.random.default_rng(0))
params = nn.utils.get_parameters(gpt)
opt = distributed.Adam(params, lr=1e-3, overlap=True, bucket_mb=0.02)
assert len(opt.buckets) >= 4, len(opt.buckets)
x, y = models.gpt_batch(cfg, 2, np.random.default_rng(1))
log = []
orig = opt._exchange
def traced(k):
    log.append((k, in_backward[0]))
    orig(k)
opt._exchange = traced
in_backward = [False]
for step in range(3):
    del log[:]
    opt.zero_grad()
    _, loss = gpt(Tensor(x, device="cuda"), Tensor(y, device="cuda"))
    in_backward[0] = True
    loss.backward()
    in_backward[0] = False
    opt.step()
    assert sorted(k for k, _ in log) == list(range(len(opt.buckets))), log
    early = [k for k, b in log if b]
    if step == 0:
        assert not early
    else:
        assert early == list(range(len(early))) and len(early) >= len(opt.buckets) - 2, log
assert opt.t == 3 and all(e == 1 for e in opt._expect)
opt.zero_grad()
_, loss = gpt(Tensor(x, device="cuda"), Tensor(y, device="cuda"))
loss.backward()
try:
    opt.zero_grad()
    raise SystemExit("zero_grad() after a started exchange must raise")
except RuntimeError:
    pass
print("SCHEDULE-OK")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DLGRAD_CUDA_COMPILE_ONLY="1", PYTHONPATH=root)
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0 and "SCHEDULE-OK" in res.stdout, res.stdout[-1500:] + res.stderr[-3000:]


@pytest.mark.timeout(300)
def test_two_rank_dp_step_equals_full_batch_step(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("wte", "q", "head"):
        assert np.array_equal(r0[k], r1[k]), f"ranks diverged on {k}"
    # single process, full batch of 8
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor, nn
    from oracle import backend
    backend.install()
    cfg = models.GPTConfig(vocab_size=17, block_size=16, n_layer=1, n_head=2, n_embd=32, device="cpu")
    gpt = models.GPT(dl, cfg, np.random.default_rng(0))
    opt = nn.optim.Adam(nn.utils.get_parameters(gpt), lr=1e-3)
    full_x, full_y = models.gpt_batch(cfg, 8, np.random.default_rng(7))
    opt.zero_grad()
    _, loss = gpt(Tensor(full_x.copy()), Tensor(full_y.copy()))
    loss.backward()
    opt.step()
    # the sum-reduced loss adds across ranks; updated weights agree to fp32 rounding of the split sums
    assert abs((float(r0["loss"]) + float(r1["loss"])) - float(loss.numpy())) <= 1e-5 * abs(float(loss.numpy()))
    for k, ref in (("wte", gpt.wte.weight.numpy()), ("q", gpt.blocks[0].attn.q_proj.weight.numpy()),
                   ("head", gpt.lm_head.weight.numpy())):
        np.testing.assert_allclose(r0[k], ref, rtol=1e-5, atol=1e-6, err_msg=k)
