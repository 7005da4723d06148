"""BASELINE.json configs[0], [2], [3] — the reference's own CPU-sized workloads — on Device.CUDA and on the CPU arm.

    python scripts/small_configs.py > gpurun_out/small_configs.jsonl        (one B200)

  C1  examples/mnist_mlp.py   MLP 784-64-10, batch 128, SGD lr 1e-3, cross-entropy
  C3  examples/mnist_gan.py   G 128-256-512-1024-784 / D 784-1024-512-256-1, batch 100, Adam lr 5e-4 betas (0.5, 0.999):
                              one discriminator step + one generator step
  C4  examples/gpt.py         char-GPT L6 h4 C128 T128, batch 16, Adam lr 1e-4

These steps are launch- and host-bound on a B200: a step is a few hundred kernels of microseconds each,
issued by a single Python thread.  Reported per config: steps/s with the batch resident in HBM (`value`), steps/s with
host batches and the loss read back every step (`e2e`), kernel launches per step, the same with forward + backward
replayed as a captured CUDA graph (`captured`, dlgrad_b200/graph.py), and the same step on the CPU arm —
this package's autograd layer over the reference's own generated C kernels (oracle/_ref; the oracle restatement if
that library is absent), single thread like the reference.  Synthetic MNIST-shaped data (no network for the real files).
The GPU runs come first: installing the CPU arm registers a host runtime, which changes where default-device tensors live.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from tests import models  # noqa: E402


def make_mlp(dl, device):
    from dlgrad_b200 import Tensor, nn
    rng = np.random.default_rng(0)
    mlp = models.MLP(dl, device, rng)
    opt = nn.optim.SGD(nn.utils.get_parameters(mlp), lr=1e-3)
    x = rng.random((128, 784), dtype=np.float32)
    y = rng.integers(0, 10, size=(128, 1)).astype(np.float32)

    def fwd_bwd(xt, yt):
        opt.zero_grad()
        loss = mlp(xt).cross_entropy_loss(yt)
        loss.backward()
        return loss
    return [(fwd_bwd, opt)], (x, y), lambda a: Tensor(a, device=device)


def make_gan(dl, device):
    from dlgrad_b200 import Tensor, nn
    rng = np.random.default_rng(1)
    gan = models.GANPair(dl, device, rng)
    opt_g = nn.optim.Adam(nn.utils.get_parameters(gan.g_layers), lr=0.0005, betas=(0.5, 0.999))
    opt_d = nn.optim.Adam(nn.utils.get_parameters(gan.d_layers), lr=0.0005, betas=(0.5, 0.999))
    real = (2.0 * rng.random((100, 784), dtype=np.float32) - 1.0).astype(np.float32)
    noise = (2.0 * rng.random((100, 128), dtype=np.float32) - 1.0).astype(np.float32)

    def d_phase(real_t, noise_t):
        fake = gan.generator(noise_t).detach()
        opt_d.zero_grad()
        d_loss = (gan.discriminator(real_t).bcewithlogitsloss(Tensor.full((100, 1), 1.0, device=device))
                  + gan.discriminator(fake).bcewithlogitsloss(Tensor.full((100, 1), 0.0, device=device))) / 2.0
        d_loss.backward()
        return d_loss

    def g_phase(real_t, noise_t):
        opt_g.zero_grad()
        g_loss = gan.discriminator(gan.generator(noise_t)).bcewithlogitsloss(Tensor.full((100, 1), 1.0, device=device))
        g_loss.backward()
        return g_loss
    return [(d_phase, opt_d), (g_phase, opt_g)], (real, noise), lambda a: Tensor(a, device=device)


def make_gpt(dl, device):
    from dlgrad_b200 import Tensor, nn
    cfg = models.GPTConfig(device=device, **bench.C4)
    gpt = models.GPT(dl, cfg, np.random.default_rng(0))
    params = nn.utils.get_parameters(gpt)
    opt = nn.optim.Adam(params, lr=1e-4)
    x, y = models.gpt_batch(cfg, 16, np.random.default_rng(7))

    def fwd_bwd(xt, yt):
        opt.zero_grad()
        _, loss = gpt(xt, yt)
        loss.backward()
        return loss
    return [(fwd_bwd, opt)], (x, y), lambda a: Tensor(a, device=device)


CONFIGS = (("c1_mlp", "configs[0] examples/mnist_mlp.py MLP 784-64-10, batch 128, SGD", make_mlp, 200, 20),
           ("c3_gan", "configs[2] examples/mnist_gan.py GAN, batch 100, Adam, D step + G step", make_gan, 100, 5),
           ("c4_gpt", "configs[3] examples/gpt.py char-GPT L6 h4 C128 T128, batch 16, Adam", make_gpt, 50, 2))


def eager_step(phases):
    def step(*inputs):
        for fwd_bwd, opt in phases:
            loss = fwd_bwd(*inputs)
            opt.step()
        return loss
    return step


def graph_arm(make, steps: int) -> dict:
    """Every phase (forward + backward) captured once and replayed as one CUDA graph launch; optimizer steps eager."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import graph
    from dlgrad_b200.runtime.cuda import profile, shim
    phases, host, to_dev = make(dl, "cuda")
    dev = tuple(to_dev(a) for a in host)
    def whole(fwd_bwd, opt):             # the optimizer step is part of the graph (Adam reads 1 - beta^t from device memory)
        def run(*inputs):
            loss = fwd_bwd(*inputs)
            opt.step()
            return loss
        return run
    caps = [graph.CapturedStep(whole(fwd_bwd, opt), *dev, params=opt.params) for fwd_bwd, opt in phases]

    def step(*arrays):
        for i, cap in enumerate(caps):
            loss = cap(*arrays) if i == 0 else cap()          # later phases reuse the input buffers just filled
        return loss
    for _ in range(3):
        loss = step()
    _ = loss.numpy()
    n0 = int(shim.lib.dlg_launch_count())
    e0 = profile.Event().record()
    for _ in range(steps):
        loss = step()
    e1 = profile.Event().record()
    e1.sync()
    ms = e1.ms_since(e0) / steps
    eager_launches = (int(shim.lib.dlg_launch_count()) - n0) / steps
    e0 = profile.Event().record()
    for _ in range(steps):
        last = float(step(*host).numpy())
    e1 = profile.Event().record()
    e1.sync()
    ms_e2e = e1.ms_since(e0) / steps
    for cap in caps:
        cap.close()
    return {"value": 1000.0 / ms, "ms_per_step": ms, "e2e": 1000.0 / ms_e2e, "graph_launches_per_step": len(caps),
            "eager_launches_per_step": eager_launches, "last_loss": last}


def gpu_arm(make, steps: int) -> dict:
    import dlgrad_b200 as dl
    from dlgrad_b200.runtime.cuda import profile, shim
    phases, host, to_dev = make(dl, "cuda")
    step = eager_step(phases)
    dev = tuple(to_dev(a) for a in host)
    for _ in range(5):
        loss = step(*dev)
    _ = loss.numpy()
    n0 = int(shim.lib.dlg_launch_count())
    e0 = profile.Event().record()
    for _ in range(steps):
        loss = step(*dev)
    e1 = profile.Event().record()
    e1.sync()
    launches = (int(shim.lib.dlg_launch_count()) - n0) / steps
    ms = e1.ms_since(e0) / steps
    e0 = profile.Event().record()
    for _ in range(steps):
        last = float(step(*(to_dev(a.copy()) for a in host)).numpy())
    e1 = profile.Event().record()
    e1.sync()
    ms_e2e = e1.ms_since(e0) / steps
    return {"value": 1000.0 / ms, "ms_per_step": ms, "e2e": 1000.0 / ms_e2e, "launches_per_step": launches,
            "us_per_launch": 1000.0 * ms / launches, "last_loss": last}


def cpu_arm(make, steps: int) -> dict:
    import dlgrad_b200 as dl
    phases, host, to_dev = make(dl, "cpu")
    step = eager_step(phases)
    step(*(to_dev(a.copy()) for a in host))
    t0 = time.perf_counter()
    for _ in range(steps):
        last = float(step(*(to_dev(a.copy()) for a in host)).numpy())
    dt = (time.perf_counter() - t0) / steps
    return {"value": 1.0 / dt, "ms_per_step": 1000.0 * dt, "steps": steps, "last_loss": last}


def main() -> None:
    from dlgrad_b200.runtime.cuda import shim
    shim.init()
    gpu = {name: gpu_arm(make, n_gpu) for name, _, make, n_gpu, _ in CONFIGS}
    captured = {name: graph_arm(make, n_gpu) for name, _, make, n_gpu, _ in CONFIGS}
    from oracle import backend
    backend.install()
    for name, what, make, n_gpu, n_cpu in CONFIGS:
        for k in backend.kernel_origin:
            backend.kernel_origin[k] = 0                    # `kind` is per config
        cpu = cpu_arm(make, n_cpu)
        g = gpu[name]
        print(json.dumps({
            "config": what, "metric": "train_steps_per_s", "unit": "steps/s", "data": "synthetic", "dtype": "f32",
            "value": g["value"], "ms_per_step": g["ms_per_step"], "steps": n_gpu, "warmup": 5,
            "e2e": {"value": g["e2e"], "unit": "steps/s"},
            "gpu_launches_per_step": g["launches_per_step"], "us_per_launch": g["us_per_launch"],
            "bound": "host dispatch / launch latency (one Python thread issues every kernel)",
            "cpu_baseline": {"value": cpu["value"], "unit": "steps/s", "cores": 1, "kind": bench.cpu_kind(),
                             "sample": f"{cpu['steps']} full steps, {cpu['ms_per_step']:.1f} ms each, on {bench.cpu_kind_text()}; "
                                       f"host has {os.cpu_count()} cores, the reference is single-threaded"},
            "captured": {"value": captured[name]["value"], "ms_per_step": captured[name]["ms_per_step"],
                         "e2e": captured[name]["e2e"], "unit": "steps/s",
                         "graph_launches_per_step": captured[name]["graph_launches_per_step"],
                         "eager_launches_per_step": captured[name]["eager_launches_per_step"],
                         "what": "forward + backward of every phase replayed as one CUDA graph (dlgrad_b200/graph.py), "
                                 "optimizer step included (Adam reads its bias corrections from device memory, "
                                 "refreshed by one tiny eager launch per replay)"},
            "speedup_vs_cpu": max(g["e2e"], captured[name]["e2e"]) / cpu["value"],
        }), flush=True)


if __name__ == "__main__":
    main()
