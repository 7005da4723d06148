"""Character-level GPT (RMSNorm, causal self-attention, relu MLP, no biases) on synthetic tokens, Adam 1e-4 — the workload
of the reference's gpt example.  Defaults are the reference's CPU-sized configuration; `--layers 12 --embd 768 --heads 12
--block 1024 --batch 8` is the scaled one that bench.py measures."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200 import Tensor, graph, nn  # noqa: E402


class Block:
    def __init__(self, C, heads, device):
        self.heads, self.hd = heads, C // heads
        self.ln1, self.ln2 = nn.RMSNorm(C), nn.RMSNorm(C)
        self.q, self.k, self.v, self.o = (nn.Linear(C, C, bias=False, device=device) for _ in range(4))
        self.fc = nn.Linear(C, 4 * C, bias=False, device=device)
        self.proj = nn.Linear(4 * C, C, bias=False, device=device)

    def attend(self, x):
        B, T, C = x.shape
        split = lambda t: t.reshape((B, T, self.heads, self.hd)).transpose(1, 2)      # noqa: E731
        q, k, v = split(self.q(x)), split(self.k(x)), split(self.v(x))
        att = (q @ k.transpose(2, 3)) * (self.hd ** -0.5)
        causal = Tensor.tril(Tensor.ones((T, T)), k=0.0)
        att = att.masked_fill(causal == Tensor(0.0), float("-inf")).softmax(dim=3)
        return self.o((att @ v).transpose(1, 2).reshape((B, T, C)))

    def __call__(self, x):
        x = x + self.attend(self.ln1(x))
        return x + self.proj(self.fc(self.ln2(x)).relu())


class GPT:
    def __init__(self, vocab, block, layers, heads, C, device):
        self.vocab = vocab
        self.wte, self.wpe = nn.Embedding(vocab, C), nn.Embedding(block, C)
        self.blocks = [Block(C, heads, device) for _ in range(layers)]
        self.ln_f = nn.RMSNorm(C)
        self.head = nn.Linear(C, vocab, bias=False, device=device)

    def __call__(self, idx, targets):
        B, T = idx.shape
        x = self.wte(idx) + self.wpe(Tensor(np.arange(T, dtype=np.float32), device=idx.device))
        for blk in self.blocks:
            x = blk(x)
        logits = self.head(self.ln_f(x))
        return logits.reshape((B * T, self.vocab)).cross_entropy_loss(targets.reshape((B * T, 1)))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--embd", type=int, default=128)
    ap.add_argument("--block", type=int, default=128)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--vocab", type=int, default=96)
    ap.add_argument("--captured", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    gpt = GPT(args.vocab, args.block, args.layers, args.heads, args.embd, "cuda")
    params = nn.utils.get_parameters(gpt)
    for p in params:                       # the embedding tables start at U[0, 1) in the reference; centre them
        if p.shape[0] in (args.vocab, args.block) and len(p.shape) == 2 and p.shape[1] == args.embd:
            p.copy_from(((rng.random(p.shape, dtype=np.float32) - 0.5) * 0.1).astype(np.float32).tobytes())
    opt = nn.optim.Adam(params, lr=1e-4)
    # a learnable stream: the next token is a fixed permutation of the current one, with 10 % noise
    perm = rng.permutation(args.vocab)

    def batch():
        toks = np.empty((args.batch, args.block + 1), dtype=np.int64)
        toks[:, 0] = rng.integers(0, args.vocab, size=args.batch)
        for t in range(args.block):
            nxt = perm[toks[:, t]]
            noise = rng.random(args.batch) < 0.1
            toks[:, t + 1] = np.where(noise, rng.integers(0, args.vocab, size=args.batch), nxt)
        return toks[:, :-1].astype(np.float32), toks[:, 1:].astype(np.float32)

    def step(x, y):
        opt.zero_grad()
        loss = gpt(x, y)
        loss.backward()
        opt.step()
        return loss

    x0, y0 = batch()
    run = graph.CapturedStep(step, Tensor(x0, device="cuda"), Tensor(y0, device="cuda"), params=params) if args.captured else None
    per_tok = 1.0 / (args.batch * args.block)          # the reference's cross-entropy is a SUM over the rows
    t0 = time.perf_counter()
    for i in range(args.steps):
        x, y = batch()
        loss = run(x, y) if run else step(Tensor(x, device="cuda"), Tensor(y, device="cuda"))
        if i % 10 == 0 or i == args.steps - 1:
            print(f"step {i:4d}  loss/token {float(loss.numpy()) * per_tok:.4f}")
    print(f"{args.steps / (time.perf_counter() - t0):.1f} steps/s ({'captured graph' if run else 'eager'})")


if __name__ == "__main__":
    main()
