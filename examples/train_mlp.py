"""MLP 784-64-10 on MNIST-shaped synthetic data, SGD, cross-entropy — the workload of the reference's mnist_mlp example."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200 import Tensor, graph, nn  # noqa: E402


class Net:
    def __init__(self, device):
        self.fc1 = nn.Linear(784, 64, bias=True, device=device)
        self.fc2 = nn.Linear(64, 10, bias=True, device=device)

    def __call__(self, x):
        return self.fc2(self.fc1(x).relu())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--captured", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    # ten fixed class prototypes + noise: something an MLP can actually learn
    protos = rng.random((10, 784), dtype=np.float32)
    net = Net("cuda")
    opt = nn.optim.SGD(nn.utils.get_parameters(net), lr=1e-2)

    def batch():
        y = rng.integers(0, 10, size=(args.batch, 1))
        x = protos[y[:, 0]] + 0.3 * rng.standard_normal((args.batch, 784)).astype(np.float32)
        return x.astype(np.float32), y.astype(np.float32)

    def step(x, y):
        opt.zero_grad()
        loss = net(x).cross_entropy_loss(y)
        loss.backward()
        opt.step()
        return loss

    x0, y0 = batch()
    xt, yt = Tensor(x0, device="cuda"), Tensor(y0, device="cuda")
    run = graph.CapturedStep(step, xt, yt, params=opt.params) if args.captured else None
    t0 = time.perf_counter()
    for i in range(args.steps):
        x, y = batch()
        loss = run(x, y) if run else step(Tensor(x, device="cuda"), Tensor(y, device="cuda"))
        if i % 50 == 0 or i == args.steps - 1:
            print(f"step {i:4d}  loss/sample {float(loss.numpy()) / args.batch:.4f}")
    dt = time.perf_counter() - t0
    x, y = batch()
    acc = float((net(Tensor(x, device="cuda")).argmax(dim=1).numpy() == y).mean())
    print(f"{args.steps / dt:.0f} steps/s ({'captured graph' if run else 'eager'}), accuracy on a fresh batch {acc:.2f}")


if __name__ == "__main__":
    main()
