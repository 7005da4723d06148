"""Generator 128-256-512-1024-784 / discriminator 784-1024-512-256-1 on synthetic 28x28 data, Adam (5e-4, betas 0.5 / 0.999),
BCE-with-logits — the workload of the reference's mnist_gan example: one discriminator step and one generator step per iteration."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200 import Tensor, graph, nn  # noqa: E402


def stack(widths, device):
    return [nn.Linear(a, b, device=device) for a, b in zip(widths[:-1], widths[1:])]


def generator(layers, z):
    for i, lin in enumerate(layers):
        z = lin(z)
        z = z.leaky_relu(0.2) if i < len(layers) - 1 else z.tanh()
    return z


def discriminator(layers, x):
    for i, lin in enumerate(layers):
        x = lin(x)
        if i < len(layers) - 1:
            x = x.leaky_relu(0.2)
    return x


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--captured", action="store_true")
    args = ap.parse_args()
    B = args.batch
    rng = np.random.default_rng(0)
    G, D = stack((128, 256, 512, 1024, 784), "cuda"), stack((784, 1024, 512, 256, 1), "cuda")
    opt_g = nn.optim.Adam(nn.utils.get_parameters(G), lr=5e-4, betas=(0.5, 0.999))
    opt_d = nn.optim.Adam(nn.utils.get_parameters(D), lr=5e-4, betas=(0.5, 0.999))
    ones = lambda: Tensor.full((B, 1), 1.0, device="cuda")       # noqa: E731
    zeros = lambda: Tensor.full((B, 1), 0.0, device="cuda")      # noqa: E731

    def d_step(real, noise):
        fake = generator(G, noise).detach()
        opt_d.zero_grad()
        loss = (discriminator(D, real).bcewithlogitsloss(ones()) + discriminator(D, fake).bcewithlogitsloss(zeros())) / 2.0
        loss.backward()
        opt_d.step()
        return loss

    def g_step(real, noise):
        opt_g.zero_grad()
        loss = discriminator(D, generator(G, noise)).bcewithlogitsloss(ones())
        loss.backward()
        opt_g.step()
        return loss

    def data():
        real = np.tanh(rng.standard_normal((B, 784))).astype(np.float32)
        noise = (2.0 * rng.random((B, 128), dtype=np.float32) - 1.0).astype(np.float32)
        return real, noise

    r0, n0 = data()
    rt, nt = Tensor(r0, device="cuda"), Tensor(n0, device="cuda")
    caps = None
    if args.captured:       # two phases, two graphs sharing the input buffers; .grad bindings are restored per replay
        caps = (graph.CapturedStep(d_step, rt, nt, params=opt_d.params), graph.CapturedStep(g_step, rt, nt, params=opt_g.params))
    t0 = time.perf_counter()
    for i in range(args.steps):
        real, noise = data()
        if caps:
            dl_, gl_ = caps[0](real, noise), caps[1]()
        else:
            real_t, noise_t = Tensor(real, device="cuda"), Tensor(noise, device="cuda")
            dl_, gl_ = d_step(real_t, noise_t), g_step(real_t, noise_t)
        if i % 25 == 0 or i == args.steps - 1:
            print(f"step {i:4d}  d_loss {float(dl_.numpy()):.4f}  g_loss {float(gl_.numpy()):.4f}")
    print(f"{args.steps / (time.perf_counter() - t0):.0f} iterations/s ({'captured graphs' if caps else 'eager'})")


if __name__ == "__main__":
    main()
