#!/usr/bin/env python
"""bench.py — the driver's measurement contract for dlgrad_b200.

    python bench.py --gpus N --steps K --warmup W            # this package on N B200s (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path, same metric

Metric (BASELINE.json): GPT train steps/s on the scaled dlgrad GPT — n_layer 12, n_embd 768,
n_head 12, block_size 1024 (configs[4]), V = 96 (the reference's char vocabulary), per-GPU batch 8,
Adam lr 1e-4, synthetic tokens, random-init weights.  One "step" = forward + backward + optimizer over
ONE per-GPU batch; with N ranks (data parallel, weak scaling) every iteration completes N of them, so
`value` = N * K / max-over-ranks device time.  The op microbenches of configs[1] (4096^2 elementwise /
reduce in GB/s, 4096^3 matmul in TFLOP/s) ride along in "ops" at N = 1.

Timed regions (each: barrier + device sync on both sides, CUDA events on the launching stream):
  value  K steps with the token batches already resident in HBM, loss read back once at the end;
  e2e    K steps through the public API with HOST inputs: Tensor(np, device="cuda") H2D every step and
         loss.numpy() D2H every step.
Inputs and activations are far larger than L2 (one step touches > 20 GB), so no explicit L2 flush.
"""
from __future__ import annotations

import argparse
import atexit
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import models  # noqa: E402  (framework-agnostic GPT / MLP definitions)

C5 = dict(vocab_size=96, block_size=1024, n_layer=12, n_head=12, n_embd=768)
C4 = dict(vocab_size=96, block_size=128, n_layer=6, n_head=4, n_embd=128)


def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except (OSError, KeyError, ValueError):
        # fallback stated in /opt/skills/guides/B200_PROFILING.md
        return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant matmul shape, from the committed
    `ncu --set full` capture (profiles/r02_ncu_mm_traffic.json); null when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_mm_traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions (A and B, the same load) run.

    ONE nvidia-smi process per job (rank 0 samples all the job's GPUs), started BEFORE the warm-up steps: its start-up
    (NVML initialisation, attaching to every GPU of the box) stalls work submission on a GPU for ~30 ms — measured as one
    58 ms step in a 20-step region when every rank started its own sampler at the top of the timed region, and with a
    cross-rank barrier in every step one stalled rank stalls all of them.  Rows are kept only while `active`."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, indices) -> None:
        self.rows, self.proc, self.indices, self.active, self.seen = [], None, list(indices), False, 0

    def __enter__(self):
        if not self.indices:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--id=" + ",".join(str(i) for i in self.indices), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            atexit.register(self.__exit__)
        except OSError:
            self.proc = None
        return self

    def wait_started(self, timeout_s: float = 5.0) -> None:
        """Block until nvidia-smi has printed its first row (its start-up is over) or `timeout_s` has passed."""
        t_end = time.perf_counter() + timeout_s
        while self.proc is not None and self.seen == 0 and self.proc.poll() is None and time.perf_counter() < t_end:
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.seen += 1
            if self.active:
                self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc and self.proc.poll() is None:
            self.proc.terminate()

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU restatement of the reference's kernels, same autograd
# layer) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_block_sample(cfg_kw: dict, repeats: int = 1) -> tuple[float, str]:
    """Seconds for ONE transformer block, forward + backward, on ONE sequence of the benchmark config,
    on the CPU oracle (single thread, as the reference)."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor
    from oracle import backend
    backend.install()
    cfg = models.GPTConfig(device="cpu", **cfg_kw)
    rng = np.random.default_rng(0)
    blk = models._Block(dl, cfg, rng)
    x = (rng.random((1, cfg.block_size, cfg.n_embd), dtype=np.float32) - 0.5).astype(np.float32)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        xt = Tensor(x.copy(), requires_grad=True)
        y = blk(xt)
        y.sum().backward()
        best = min(best, time.perf_counter() - t0)
    return best, (f"1 of {cfg.n_layer} transformer blocks x 1 sequence (T={cfg.block_size}, C={cfg.n_embd}), "
                  f"forward+backward on {cpu_kind_text()}, single thread")


def cpu_kind() -> str:
    """'reference' when every kernel of the CPU sample was the reference's own generated C (oracle/_ref,
    compiled from /root/reference/dlgrad/codegen/cpu_kernel.py with the reference's flags), 'port' when the
    restatement in oracle/kernels.c served any of them."""
    from oracle import backend
    o = backend.kernel_origin
    return "reference" if (o["reference"] > 0 and o["port"] == 0) else "port"


def cpu_kind_text() -> str:
    return ("the reference's own generated C kernels (oracle/_ref) under this package's autograd layer"
            if cpu_kind() == "reference" else "the CPU oracle (oracle/kernels.c restatement)")


def cpu_equiv_steps_per_s(block_seconds: float, cfg_kw: dict, batch: int) -> float:
    """Extrapolate the block sample to full steps: L blocks x B sequences, plus the embedding / head /
    loss / optimizer share taken from the FLOP model."""
    cfg = models.GPTConfig(**cfg_kw)
    full = models.gpt_train_flops(cfg, 1)
    T, C, L = cfg.block_size, cfg.n_embd, cfg.n_layer
    blocks = 3.0 * L * (24.0 * T * C * C + 4.0 * T * T * C)
    step_seconds = block_seconds * L * batch * (full / blocks)
    return 1.0 / step_seconds


_cpu_model: dict = {}


def cpu_full_step_sample(cfg_kw: dict) -> tuple[float, str]:
    """Seconds for ONE complete train step (forward + backward + Adam, all layers, embedding, head, loss) of the
    benchmark config at batch 1 on the CPU oracle — 1/B of the workload's step, measured, not extrapolated."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor, nn
    from oracle import backend
    backend.install()
    key = tuple(sorted(cfg_kw.items()))
    if key not in _cpu_model:
        cfg = models.GPTConfig(device="cpu", **cfg_kw)
        gpt = models.GPT(dl, cfg, np.random.default_rng(0))
        params = nn.utils.get_parameters(gpt)
        _cpu_model[key] = (cfg, gpt, nn.optim.Adam(params, lr=1e-4), np.random.default_rng(1000))
    cfg, gpt, opt, rng = _cpu_model[key]
    xb, yb = models.gpt_batch(cfg, 1, rng)
    t0 = time.perf_counter()
    opt.zero_grad()
    _, loss = gpt(Tensor(xb), Tensor(yb))
    loss.backward()
    opt.step()
    _ = float(loss.numpy())
    dt = time.perf_counter() - t0
    return dt, (f"one full train step (all {cfg.n_layer} layers, embeddings, head, loss, Adam) at batch 1 "
                f"(T={cfg.block_size}, C={cfg.n_embd}) on {cpu_kind_text()}, single thread")


def run_reference(args) -> None:
    """The reference's CPU path on the benchmark config.  One timed "step" of this arm is a BOUNDED SAMPLE of the
    workload's step: the complete train step at batch 1 (the workload's batch is `--batch` independent sequences, the
    CPU path is single-threaded and linear in the batch), so `ms_per_step` is the measured time of one sample and
    `value` = 1 / (batch x sample seconds).  At most `--steps` samples, and no more than fit in REF_BUDGET_S seconds;
    the warm-up is one transformer-block sample (it loads every kernel), which also cross-checks the full step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_kw = C5 if args.config == "c5" else C4
    budget = float(os.environ.get("DLGRAD_REF_BUDGET_S", "150"))
    t_block, _ = cpu_block_sample(cfg_kw)                       # warm-up: every kernel of the step loaded and run once
    times, sample, t_start = [], "", time.perf_counter()
    while len(times) < max(1, args.steps):
        t, sample = cpu_full_step_sample(cfg_kw)
        times.append(t)
        if time.perf_counter() - t_start + t > budget:           # the next sample would overrun the budget
            break
    mean = float(np.mean(times))
    value = 1.0 / (mean * args.batch)     # the CPU path does not scale with --gpus: one host, one thread
    extrap = cpu_equiv_steps_per_s(t_block, cfg_kw, args.batch)
    line = {
        "impl": "reference", "metric": "gpt_train_steps_per_s", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": 1, "ms_per_step": 1000.0 * mean,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg_kw, args.batch, args.gpus),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": 1, "kind": cpu_kind(),
                         "sample": sample + f"; {len(times)} samples, mean {mean:.2f} s each; value = 1 / ({args.batch} x "
                                            f"sample): the workload's step is {args.batch} such sequences and the reference "
                                            "is single-threaded",
                         "cross_check": {"block_sample_s": t_block, "extrapolated_steps_per_s": extrap,
                                         "ratio_measured_over_extrapolated": value / extrap}},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cpu_count": os.cpu_count()},
    }
    print(json.dumps(line))


def workload_config(cfg_kw: dict, batch: int, gpus: int) -> dict:
    return {"workload": f"scaled dlgrad GPT train step (BASELINE.json configs[4]): n_layer={cfg_kw['n_layer']} "
                        f"n_embd={cfg_kw['n_embd']} n_head={cfg_kw['n_head']} block_size={cfg_kw['block_size']} "
                        f"vocab={cfg_kw['vocab_size']}, Adam lr 1e-4",
            "per_gpu_batch": batch, "global_batch": batch * gpus, "seq_len": cfg_kw["block_size"],
            "parallelism": f"dp{gpus}", "l2": "working set per step >> 126 MB L2; no explicit flush",
            "arithmetic": "fp32 everywhere; matmul = 3xTF32 on tcgen05 tensor cores (hi/lo split, fp32 accumulate)"}


# ---------------------------------------------------------------------------------------------------
# this package
# ---------------------------------------------------------------------------------------------------
def build_gpt(cfg_kw: dict, device: str, fused_dp: bool = False):
    import dlgrad_b200 as dl
    from dlgrad_b200 import distributed, nn
    cfg = models.GPTConfig(device=device, **cfg_kw)
    gpt = models.GPT(dl, cfg, np.random.default_rng(0))       # identical weights on every rank
    params = nn.utils.get_parameters(gpt)
    # data parallel: reduce-scatter + Adam + all-gather in one kernel over NVLink peer memory
    opt = distributed.Adam(params, lr=1e-4) if fused_dp else nn.optim.Adam(params, lr=1e-4)
    return cfg, gpt, params, opt


def train_step(gpt, opt, params, x, y, world):
    from dlgrad_b200 import distributed
    opt.zero_grad()
    _, loss = gpt(x, y)
    loss.backward()
    if world > 1 and not isinstance(opt, distributed.Adam):
        distributed.all_reduce_gradients(params)     # DLGRAD_DP_FUSED=0: ncclAllReduce, then the local Adam kernel
    opt.step()
    return loss


def op_microbench(peaks: dict, only: tuple | None = None) -> dict:
    """configs[1]: broadcast add/mul and sum/max over axis on 4096x4096 fp32; matmul 4096^3 fwd.

    Timing: every op is launched ROUNDS x SETS times back to back (captured in one CUDA graph, replayed
    between ONE pair of CUDA events), walking round-robin over SETS = 6 independent input sets (6 x 64 MiB per
    operand = 384 MiB > 3 x L2), so no launch finds its input in the 126 MB L2 ("inputs larger than L2") and
    neither per-launch event latency nor Python dispatch is billed to a 20 us kernel.  2 warm-up rounds.  A device-to-device copy of the same 64 MiB arrays is timed
    the same way as the size-matched ceiling (MEASURED_PEAKS' copy figure is for 2 GiB buffers)."""
    from dlgrad_b200 import Tensor
    from dlgrad_b200.runtime.cuda import profile, shim
    rng = np.random.default_rng(0)
    n, SETS, ROUNDS = 4096, 6, 4
    mk = lambda *shape: [Tensor(rng.random(shape, dtype=np.float32), device="cuda") for _ in range(SETS)]   # noqa: E731
    a, b, row, col = mk(n, n), mk(n, n), mk(1, n), mk(n, 1)
    full, half, one = 3 * n * n * 4, 2 * n * n * 4, n * n * 4
    specs = [("add_same", lambda i: a[i] + b[i], full, "GB/s"), ("mul_same", lambda i: a[i] * b[i], full, "GB/s"),
             ("add_row", lambda i: a[i] + row[i], half, "GB/s"), ("mul_col", lambda i: a[i] * col[i], half, "GB/s"),
             ("sum_axis0", lambda i: a[i].sum(dim=0), one, "GB/s"), ("sum_axis1", lambda i: a[i].sum(dim=1), one, "GB/s"),
             ("max_axis0", lambda i: a[i].max(dim=0), one, "GB/s"), ("max_axis1", lambda i: a[i].max(dim=1), one, "GB/s"),
             ("sum_all", lambda i: a[i].sum(), one, "GB/s"),
             ("exp", lambda i: a[i].exp(), half, "GB/s"), ("transpose", lambda i: a[i].permute((1, 0)), half, "GB/s"),
             ("softmax_rows", lambda i: a[i].softmax(dim=1), half, "GB/s"),
             ("matmul_4096", lambda i: a[i] @ b[i], 2.0 * n ** 3, "TFLOP/s"),
             # the two backward GEMMs of configs[1]'s "matmul 4096^3 fwd+bwd", as MatMul.backward issues them: the
             # stored operands are read transposed inside the kernel (no materialised transposes)
             ("matmul_4096_bwd_dA", lambda i: mm_t(a[i], b[i], trans_y=True), 2.0 * n ** 3, "TFLOP/s"),
             ("matmul_4096_bwd_dB", lambda i: mm_t(a[i], b[i], trans_x=True), 2.0 * n ** 3, "TFLOP/s")]
    out = {}
    scratch = [shim.alloc(one) for _ in range(SETS)]

    def mm_t(x, y, **kw):
        from dlgrad_b200.buffer import Buffer
        from dlgrad_b200.device import Device
        from dlgrad_b200.dispatch import dispatcher
        from dlgrad_b200.helpers import BinaryOps
        return Buffer(dispatcher.dispatch(op=BinaryOps.MATMUL, device=Device.CUDA, x=x.data, y=y.data, **kw), (n, n), Device.CUDA)

    def timed(fn, rounds):
        """Seconds per launch.  The ROUNDS x SETS launches are captured into ONE CUDA graph and the graph is
        replayed between a pair of events, so host dispatch (~10 us per op in Python) is not billed to
        kernels that run for 10-30 us; the replay with the best time of three is reported."""
        for r in range(2):
            for i in range(SETS):
                fn(i)
        shim.synchronize()
        shim.check(shim.lib.dlg_graph_begin(), "dlg_graph_begin")
        keep = []
        for r in range(rounds):
            for i in range(SETS):
                keep.append(fn(i))          # outputs stay referenced: no block is recycled inside the graph
        graph = shim.lib.dlg_graph_end()
        if graph == shim.NULL:
            raise RuntimeError("graph capture failed: " + shim.last_error())
        best = float("inf")
        for _ in range(4):
            e0 = profile.Event().record()
            shim.check(shim.lib.dlg_graph_launch(graph), "dlg_graph_launch")
            e1 = profile.Event().record()
            e1.sync()
            best = min(best, e1.ms_since(e0))
        shim.lib.dlg_graph_destroy(graph)
        del keep
        return best * 1e-3 / (rounds * SETS)

    t = timed(lambda i: shim.check(shim.lib.dlg_d2d(scratch[i], a[i].data.ptr, one), "d2d"), ROUNDS)
    out["d2d_copy_same_size"] = {"achieved": half / t / 1e9, "unit": "GB/s", "ms": t * 1e3,
                                 "frac_of_hbm_peak": half / t / 1e9 / peaks["hbm_gbs"], "algorithmic_MB": half / 1e6}
    for name, fn, work, unit in specs:
        if only and name not in only:
            continue
        t = timed(fn, ROUNDS if unit == "GB/s" else 2)
        if unit == "GB/s":
            ach = work / t / 1e9
            out[name] = {"achieved": ach, "unit": unit, "ms": t * 1e3, "frac_of_hbm_peak": ach / peaks["hbm_gbs"],
                         "algorithmic_MB": work / 1e6}
        else:
            ach = work / t / 1e12
            out[name] = {"achieved": ach, "unit": unit, "ms": t * 1e3,
                         "frac_of_3xtf32_peak": ach / (peaks["bf16_burst"] / 6.0)}
    if all(k in out for k in ("matmul_4096", "matmul_4096_bwd_dA", "matmul_4096_bwd_dB")):
        ms = sum(out[k]["ms"] for k in ("matmul_4096", "matmul_4096_bwd_dA", "matmul_4096_bwd_dB"))
        ach = 3 * 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        out["matmul_4096_fwd_bwd"] = {"achieved": ach, "unit": "TFLOP/s", "ms": ms, "gflop": 3 * 2.0 * n ** 3 / 1e9,
                                      "frac_of_3xtf32_peak": ach / (peaks["bf16_burst"] / 6.0)}
    shim.synchronize()
    return out


def run_b200(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("DLGRAD_CUDA_DEVICE", str(local))
    from dlgrad_b200 import Tensor, distributed
    from dlgrad_b200.runtime.cuda import ops as cuda_ops
    from dlgrad_b200.runtime.cuda import profile, shim, transfer
    shim.init(local)
    dist = None
    if world > 1:
        import torch.distributed as dist      # rendezvous + timing barrier only (gloo: no GPU context of its own)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        uid = [distributed.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        distributed.init(rank, world, uid[0])

    def barrier():
        shim.synchronize()
        if dist is not None:
            dist.barrier()

    cfg_kw = C5 if args.config == "c5" else C4
    # DLGRAD_DP_FUSED=force: the data-parallel optimizer on ONE GPU too (its exchange degenerates to local loads and stores) —
    # a cheap way to measure what the comm-stream kernels cost the backward they overlap with
    fused_dp = (world > 1 and os.environ.get("DLGRAD_DP_FUSED", "1") != "0") or os.environ.get("DLGRAD_DP_FUSED") == "force"
    if fused_dp and world > 1 and not distributed.peer_memory_available():
        fused_dp = False
        if rank == 0:
            print("[bench] peer memory cannot be mapped on this box: gradient exchange falls back to ncclAllReduce "
                  f"({shim.last_error()})", file=sys.stderr)
    # rank 0 samples the clocks of all the job's GPUs (local rank = GPU index under the single-node launcher); started here so
    # that nvidia-smi's start-up falls into the model build and the warm-up, not into a timed region
    sampling = rank == 0 and not (args.prebuild or args.profile)
    clocks = ClockSampler(range(world) if sampling else ()).__enter__()
    cfg, gpt, params, opt = build_gpt(cfg_kw, "cuda", fused_dp)
    opt.grad_scale = 1.0 / world
    data_rng = np.random.default_rng(1000 + rank)            # rank-specific tokens
    batches = [models.gpt_batch(cfg, args.batch, data_rng) for _ in range(4)]
    dev_batches = [(Tensor(x, device="cuda"), Tensor(y, device="cuda")) for x, y in batches]

    if args.prebuild:       # compile-only census of every kernel the bench launches
        train_step(gpt, opt, params, *dev_batches[0], world)
        distributed.prebuild([p.numel for p in params])           # the fused data-parallel step at 2, 4 and 8 ranks
        op_microbench(load_peaks())
        return

    for i in range(args.warmup):
        loss = train_step(gpt, opt, params, *dev_batches[i % 4], world)
    _ = loss.numpy() if args.warmup else None

    if args.profile:
        # every launch bracketed by events (serialises nothing, but adds event overhead): shares, not a bench value
        profile.kernel_timer = profile.KernelTimer()
        profile.matmul_timer = profile.KernelTimer()
        nsteps = 3
        e0 = profile.Event().record()
        for i in range(nsteps):
            loss = train_step(gpt, opt, params, *dev_batches[i % 4], world)
        e1 = profile.Event().record()
        e1.sync()
        kt, mt = profile.kernel_timer.collect(), profile.matmul_timer.collect()
        profile.kernel_timer = profile.matmul_timer = None
        tot = e1.ms_since(e0) / nsteps
        print(f"step {tot:.2f} ms; generated-kernel launches {kt['launches'] / nsteps:.0f}/step {kt['ms'] / nsteps:.2f} ms; "
              f"matmul launches {mt['launches'] / nsteps:.0f}/step {mt['ms'] / nsteps:.2f} ms")
        for tag, v in sorted(kt["by_tag"].items(), key=lambda kv: -kv[1][1]):
            print(f"  {tag[0]:24s} out_MB={tag[1] / 1e6:8.2f} n/step={v[0] / nsteps:7.1f}  ms/step={v[1] / nsteps:8.3f}  avg_us={1e3 * v[1] / v[0]:8.1f}")
        for tag, v in sorted(mt["by_tag"].items(), key=lambda kv: -kv[1][1]):
            print(f"  mm {tag}  n/step={v[0] / nsteps:6.1f}  ms/step={v[1] / nsteps:7.3f}  TF/s={v[2] / (v[1] * 1e-3) / 1e12:7.1f}")
        return

    peaks = load_peaks()
    # Python's cyclic collector: a full (generation-2) pass walks every live object of the process — 10 ms here, but ~80 ms
    # once torch.distributed is imported for the N > 1 rendezvous (175 k objects) — and the host is never more than one
    # launch queue (~2 steps) ahead of the device, so one such pause inside a timed region starves the GPU for 20-35 ms
    # (seen as ONE 47-62 ms step per region at N = 2 and none at N = 1; with a cross-rank barrier in every step one paused
    # rank pauses all).  Standard remedy for training loops: collect now, then freeze the survivors out of later passes.
    if os.environ.get("DLGRAD_BENCH_GC_FREEZE", "1") != "0":
        gc.collect()
        gc.freeze()
    # ---- region A: device-resident inputs ----------------------------------------------------------
    barrier()
    launches0 = cuda_ops.launch_count()
    # roofline evidence is taken LIVE inside this region: every matmul-class launch of every 10th timed step is bracketed by
    # CUDA events on the launching stream (runtime/cuda/profile.py).  Two event records per launch cost ~2 us of stream
    # time each — 1.2 ms per step when every step carries them (round 1 did) — so nine steps in ten run bare.
    timer = profile.KernelTimer()
    timed_steps = 0
    clocks.wait_started()
    clocks.active = True                              # keeps sampling through region B; closed after it
    e0 = e0_a = profile.Event().record()
    step_times = os.environ.get("DLGRAD_BENCH_STEP_TIMES", "0") == "1"      # diagnosis: per-step device times of both regions
    marks_a, marks_b = [], []
    runahead = int(os.environ.get("DLGRAD_BENCH_RUNAHEAD", "0"))
    step_events = []
    # (round 2, measured: a step that carries the events takes +2.5 ms — 29.2 against 26.6 ms — so every 10th step does)
    every = int(os.environ.get("DLGRAD_BENCH_TIMER_EVERY", "10"))      # 0: no per-launch events at all (A/B runs only)
    for i in range(args.steps):
        sampled = every > 0 and i % every == 0
        profile.matmul_timer = timer if sampled else None
        timed_steps += int(sampled)
        loss = train_step(gpt, opt, params, *dev_batches[i % 4], world)
        if step_times:
            marks_a.append(profile.Event().record())
        if runahead > 0:
            # bounded run-ahead: the host enqueues at most `runahead` steps beyond the one the device is executing
            step_events.append(profile.Event().record())
            if len(step_events) > runahead:
                step_events.pop(0).sync()
    profile.matmul_timer = None
    e1 = profile.Event().record()
    e1.sync()
    barrier()
    mm = timer.collect()
    timed_steps = max(timed_steps, 1)
    ms_a = e1.ms_since(e0)
    launches = cuda_ops.launch_count() - launches0
    last_loss = float(loss.numpy())
    # ---- region B: end to end from host buffers ------------------------------------------------------
    transfer.reset_counters()
    barrier()
    e0 = profile.Event().record()
    for i in range(args.steps):
        x, y = batches[i % 4]
        loss = train_step(gpt, opt, params, Tensor(x, device="cuda"), Tensor(y, device="cuda"), world)
        _ = loss.numpy()
        if step_times:
            marks_b.append(profile.Event().record())
    e1 = profile.Event().record()
    e1.sync()
    barrier()
    ms_b = e1.ms_since(e0)
    if step_times:
        for name, start, marks in (("A", e0_a, marks_a), ("B", e0, marks_b)):
            prev, row = start, []
            for m in marks:
                row.append(round(m.ms_since(prev), 2))
                prev = m
            print(f"[bench] rank {rank} region {name} per-step ms: {row}", file=sys.stderr, flush=True)
    io = transfer.counters()
    if rank == 0 and not clocks.rows:                 # a very short run: keep the load up until one sample exists
        t_end = time.perf_counter() + 1.5
        fill = Tensor(np.ones((2048, 2048), dtype=np.float32), device="cuda")
        while not clocks.rows and time.perf_counter() < t_end:      # rank-local load only: no collective in here
            _ = float((fill @ fill).sum().numpy())
    clocks.active = False
    clocks.__exit__(None, None, None)

    replicas_identical = None
    if dist is not None:      # max over ranks; launches summed over ranks
        import torch
        import zlib
        # correctness of the data-parallel step travels with the speed: every rank hashes its parameters (all bytes,
        # after the timed steps) and the hashes must agree — replicas that drifted apart would make `value` meaningless
        crc = 0
        for p in params:
            crc = zlib.crc32(p.numpy().tobytes(), crc)
        hashes = [None] * world
        dist.all_gather_object(hashes, crc)
        replicas_identical = len(set(hashes)) == 1
        t = torch.tensor([ms_a, ms_b], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_a, ms_b = float(t[0]), float(t[1])
        n = torch.tensor([launches], dtype=torch.int64)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        launches = int(n[0])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * args.steps / (ms_a * 1e-3)
    e2e = world * args.steps / (ms_b * 1e-3)
    # dense GEMMs vs attention-class kernels (fused attention + the GEMMs that skip masked tiles): the latter are
    # credited their NOMINAL (dense) flops by the per-launch timer; under the causal mask only (nq + 1) / (2 nq) of the
    # 128 x 128 tile pairs are executed, so both figures are reported
    nq = cfg.block_size // 128
    live = (nq + 1) / (2.0 * nq) if nq else 1.0
    split = {"dense": [0.0, 0.0, 0], "attention": [0.0, 0.0, 0]}
    for t, v in mm["by_tag"].items():
        cls = "attention" if (str(t[0]).startswith("attn") or "_spnn" in str(t[0]) or "_sptn" in str(t[0])) else "dense"
        split[cls][0] += v[2]
        split[cls][1] += v[1]
        split[cls][2] += v[0]

    def _tf(work, ms):
        return work / (ms * 1e-3) / 1e12 if ms > 0 else None
    flops_step = models.gpt_train_flops(cfg, args.batch)
    mm_tflops = (mm["work"] / (mm["ms"] * 1e-3) / 1e12) if mm["ms"] > 0 else 0.0
    peak_3xtf32 = peaks["bf16_sustained"] / 6.0      # TF32 dense = bf16/2, three MMAs per product
    line = {
        "metric": "gpt_train_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_a / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(cfg_kw, args.batch, world), **({"gradient_exchange": (
            (f"overlapped with backward: {len(opt.buckets)} parameter buckets, each exchanged on the comm stream as soon as its "
             "gradients are final by ONE kernel per rank (reduce-scatter over NVLink peer loads + Adam on the 1/world shard + "
             "all-gather over peer stores); cross-rank arrivals are stream memory operations" if getattr(opt, "overlap", False) else
             "one kernel per rank after backward: reduce-scatter (NVLink peer loads) + Adam on the 1/world shard + all-gather (peer stores)")
            if fused_dp else "ncclAllReduce of the flat gradient bucket, then the local Adam kernel")} if world > 1 else {})),
        "tokens_per_s": value * args.batch * cfg.block_size,
        "model_tflops": value * flops_step / 1e12,
        "last_loss": last_loss,
        "gpu_launches": int(launches),
        "hbm": {"peak_gb": int(shim.lib.dlg_mem_peak()) / 1e9, "reserved_gb": int(shim.lib.dlg_mem_reserved()) / 1e9,
                "note": "caching allocator, per rank; B200 has 180 GB"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": io["h2d_bytes"] // args.steps,
                "d2h_bytes_per_step": io["d2h_bytes"] // args.steps},
        "roofline": {"bound": "tensor", "kernel": "matmul (all launches of the timed steps, CUDA events per launch)",
                     "achieved": mm_tflops, "peak": peak_3xtf32, "unit": "TFLOP/s",
                     "frac": mm_tflops / peak_3xtf32 if peak_3xtf32 else None, "traffic": load_traffic(),
                     "peak_note": f"3xTF32 = sustained bf16 {peaks['bf16_sustained']} TF/s / 6 ({peaks['source']}: "
                                  "MEASURED_PEAKS.json has no TF32 entry; TF32 dense = bf16/2, 3 MMAs per product)",
                     "matmul_share_of_step": (mm["ms"] / timed_steps) / (ms_a / args.steps) if ms_a else None,
                     "matmul_launches": mm["launches"], "steps_with_per_launch_events": timed_steps,
                     "tf32_in_box_note": "cuBLAS TF32 measured on a B200 of this pool: 723 burst / 599 sustained TFLOP/s "
                                         "(profiles/r01_tf32_peak_cublas.txt), i.e. 241 / 200 for 3xTF32; the stricter "
                                         "bf16/6 figure is the denominator used",
                     "split": {
                         "dense_gemms": {"launches": split["dense"][2], "ms_per_step": split["dense"][1] / timed_steps,
                                         "tflops": _tf(split["dense"][0], split["dense"][1]),
                                         "frac": (_tf(split["dense"][0], split["dense"][1]) or 0.0) / peak_3xtf32},
                         "attention_class": {"launches": split["attention"][2], "ms_per_step": split["attention"][1] / timed_steps,
                                             "nominal_tflops": _tf(split["attention"][0], split["attention"][1]),
                                             "executed_tflops": _tf(split["attention"][0] * live, split["attention"][1]),
                                             "executed_fraction_of_nominal": live,
                                             "frac_executed": (_tf(split["attention"][0] * live, split["attention"][1]) or 0.0) / peak_3xtf32,
                                             "note": "fused attention kernels and tile-skipping GEMMs; the causal mask kills "
                                                     f"{nq * (nq - 1) // 2} of {nq * nq} 128x128 tile pairs per head"}},
                     "by_shape": [{"kernel": t[0], "M": t[1], "N": t[2], "K": t[3], "batch": t[4], "launches": v[0],
                                   "ms_per_step": v[1] / timed_steps, "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] else None}
                                  for t, v in sorted(mm["by_tag"].items(), key=lambda kv: -kv[1][1])[:12]]},
    }
    if world > 1:
        line["replicas_identical"] = replicas_identical
    if world == 1 and not args.no_ops:
        line["ops"] = op_microbench(peaks)
    if world == 1 and not args.no_small:
        # BASELINE configs[0], [2], [3] — the reference's own CPU-sized workloads — eager and as captured CUDA graphs
        # (host-dispatch bound; scripts/small_configs.py has the full per-config report incl. the CPU arm)
        from scripts import small_configs as sc
        small = {}
        for name, what, make, n_gpu, _ in sc.CONFIGS:
            g, c = sc.gpu_arm(make, max(10, n_gpu // 2)), sc.graph_arm(make, max(10, n_gpu // 2))
            small[name] = {"config": what, "unit": "steps/s", "eager": g["value"], "eager_e2e": g["e2e"],
                           "launches_per_step": g["launches_per_step"], "us_per_launch": g["us_per_launch"],
                           "captured": c["value"], "captured_e2e": c["e2e"]}
        line["small_configs"] = small
    if world == 1 and not args.no_cpu:
        # last: installing the CPU arm registers a host runtime, which changes where default-device tensors live
        t, sample = cpu_block_sample(cfg_kw, repeats=3)
        v = cpu_equiv_steps_per_s(t, cfg_kw, args.batch)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": 1, "kind": cpu_kind(),
                                "sample": sample + f" (best of 3: {t:.2f} s), scaled by L x B x FLOP share — `--impl reference` "
                                f"measures whole batch-1 steps and reports how well this extrapolation holds; host has "
                                f"{os.cpu_count()} cores, the reference is single-threaded"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c5", choices=["c5", "c4"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-ops", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the configs[0]/[2]/[3] extras (MLP, GAN, char-GPT)")
    ap.add_argument("--prebuild", action="store_true", help="compile-only census (used by __graft_entry__.build)")
    ap.add_argument("--ops-only", action="store_true", help="run only the configs[1] op microbenches")
    ap.add_argument("--profile", action="store_true", help="per-kernel-family CUDA-event breakdown of 3 steps (not a bench value)")
    args = ap.parse_args()
    if args.ops_only:
        from dlgrad_b200.runtime.cuda import shim
        shim.init(int(os.environ.get("LOCAL_RANK", "0")))
        res = op_microbench(load_peaks())
        for k, v in res.items():
            print(k, json.dumps({a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
