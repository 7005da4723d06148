"""Multi-GPU check of distributed.Adam (one kernel per rank: reduce-scatter + Adam + all-gather over NVLink peer
memory) against the path it replaces (ncclAllReduce of the flat gradient bucket, then the local Adam kernel).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/dp_fused_check.py

Every rank builds TWO copies of the char-GPT (configs[3]) with identical seeded weights, feeds both the same
rank-specific tokens, and trains one copy with each exchange for STEPS steps.  Checked:
  * all ranks hold bit-identical parameters after the fused steps (CRC exchange over gloo);
  * fused vs a float32 numpy restatement of the kernel's schedule (every rank's gradients gathered over gloo, added
    in rank order, IEEE float32 Adam in the kernel's operation order): BIT-IDENTICAL at any world size;
  * fused vs NCCL: bit-identical at world <= 2 (a + b is order-free).  Beyond that NCCL adds in a different order, and
    Adam divides by sqrt(v): where eight gradients cancel to ~1e-8 the last-bit difference of the sums is a percent of
    the result, so single parameters differ by up to ~lr per step — bounded here by 2 * lr * steps, with the mean
    difference far below it and the losses tracking each other.
torch.distributed (gloo) is the launcher's side channel only: rendezvous and the CRC exchange.
"""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

STEPS = 4
LR, BETAS, EPS = 1e-3, (0.9, 0.999), 1e-8
OVERLAP = os.environ.get("DLGRAD_DP_OVERLAP", "0") == "1"      # which schedule of distributed.Adam is checked
BUCKET_MB = 0.25          # several exchange buckets for a 0.43 M parameter model: steps 2.. overlap with backward


def adam_float32(p, g_sum, m, v, t, world):
    """codegen/cuda_kernel.dp_adam's per-element arithmetic, one IEEE float32 operation per line."""
    f = np.float32
    gg = g_sum * f(1.0 / world)
    m[:] = m * f(BETAS[0]) + gg * f(1.0 - BETAS[0])
    v[:] = v * f(BETAS[1]) + (gg * gg) * f(1.0 - BETAS[1])
    mh = m / f(1.0 - BETAS[0] ** t)
    vh = v / f(1.0 - BETAS[1] ** t)
    p[:] = p - (mh / (np.sqrt(vh) + f(EPS))) * f(LR)


def main() -> int:
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("DLGRAD_CUDA_DEVICE", str(local))
    import torch.distributed as dist

    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor, distributed, nn
    from dlgrad_b200.runtime.cuda import shim
    from tests import models
    shim.init(local)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = [distributed.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    distributed.init(rank, world, uid[0])

    cfg = models.GPTConfig(vocab_size=96, block_size=128, n_layer=2, n_head=4, n_embd=128, device="cuda")
    nets = [models.GPT(dl, cfg, np.random.default_rng(0)) for _ in range(2)]
    plist = [nn.utils.get_parameters(n) for n in nets]
    if os.environ.get("DLGRAD_CUDA_COMPILE_ONLY") == "1":      # build machine: compile the 2/4/8-rank kernels
        distributed.prebuild([p.numel for p in plist[0]], bucket_mb=BUCKET_MB)
    flat = lambda tensors: np.concatenate([t.numpy().reshape(-1) for t in tensors])     # noqa: E731
    ref_p = flat(plist[0]).copy()
    ref_m, ref_v = np.zeros_like(ref_p), np.zeros_like(ref_p)
    fused = distributed.Adam(plist[0], lr=LR, betas=BETAS, eps=EPS, overlap=OVERLAP, bucket_mb=BUCKET_MB)
    plain = nn.optim.Adam(plist[1], lr=LR, betas=BETAS, eps=EPS)
    plain.grad_scale = 1.0 / world
    rng = np.random.default_rng(100 + rank)
    ok = True
    for step in range(STEPS):
        x, y = models.gpt_batch(cfg, 4, rng)
        losses = []
        for net, opt, params in ((nets[0], fused, plist[0]), (nets[1], plain, plist[1])):
            opt.zero_grad()
            _, loss = net(Tensor(x.copy(), device="cuda"), Tensor(y.copy(), device="cuda"))
            loss.backward()
            if opt is plain:
                distributed.all_reduce_gradients(params)
            else:
                every = [None] * world
                dist.all_gather_object(every, flat([p.grad for p in params]))
                g_sum = every[0].copy()
                for r in range(1, world):
                    g_sum += every[r]                       # rank order, like the kernel
                adam_float32(ref_p, g_sum, ref_m, ref_v, step + 1, world)
            opt.step()
            losses.append(float(loss.numpy()))
        flat_f, flat_p = flat(plist[0]), flat(plist[1])
        exact = np.array_equal(flat_f.view(np.uint32), ref_p.view(np.uint32))
        crc = zlib.crc32(flat_f.tobytes())
        crcs = [None] * world
        dist.all_gather_object(crcs, crc)
        same_bits = np.array_equal(flat_f.view(np.uint32), flat_p.view(np.uint32))
        diff = np.abs(flat_f - flat_p)
        replicas = len(set(crcs)) == 1
        near = same_bits if world <= 2 else (diff.max() <= 2 * LR * (step + 1) and diff.mean() <= 1e-2 * LR)
        good = replicas and exact and near and np.isfinite(losses).all() and abs(losses[0] - losses[1]) <= 1e-4 * abs(losses[1])
        ok = ok and good
        if rank == 0:
            print(f"step {step}: loss fused {losses[0]:.4f} nccl {losses[1]:.4f} | replicas identical: {replicas} | "
                  f"vs float32 restatement: {'bit-identical' if exact else 'DIFFERENT'} | vs nccl+adam: "
                  + ("bit-identical" if same_bits else f"max |diff| {diff.max():.2e} mean {diff.mean():.2e} (lr {LR:g})")
                  + f" | {'ok' if good else 'FAIL'}", flush=True)
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if rank == 0:
        print(f"world {world}: {flat_f.size} parameters, optimizer state per rank {fused.shard} of {fused.total} elements "
              f"-> {'PASS' if all(oks) else 'FAIL'}", flush=True)
    shim.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    if os.environ.get("DLGRAD_CUDA_COMPILE_ONLY") == "1":
        return 0                                                  # nothing ran: the walk only filled the kernel cache
    return 0 if all(oks) else 1


if __name__ == "__main__":
    sys.exit(main())
