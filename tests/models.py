"""The three workloads of BASELINE.json (MLP, GAN, char-GPT) written against the dlgrad API and
parameterised by the module that provides it, so that the SAME model code runs on
  * the real reference (`import dlgrad`, only in the build container, tests/golden/make_golden.py),
  * dlgrad_b200 on the CPU oracle (device="cpu" with oracle.backend installed),
  * dlgrad_b200 on Device.CUDA.
Architectures follow examples/mnist_mlp.py:12-22, examples/mnist_gan.py:30-58 and
examples/gpt.py:33-119 of the reference.  Weights are injected from a seeded numpy generator
because the reference's RNG cannot be seeded.
"""
from __future__ import annotations

import math

import numpy as np


def _inject(dl, tensor_shape, rng, low, high, device):
    arr = (low + (high - low) * rng.random(tensor_shape, dtype=np.float32)).astype(np.float32)
    return dl.Tensor(arr, requires_grad=True, device=device)


def seed_linear(dl, layer, rng, device):
    out_f, in_f = layer.weight.shape
    k = 1.0 / math.sqrt(in_f)
    layer.weight = _inject(dl, (out_f, in_f), rng, -k, k, device)
    if layer.bias is not None:
        layer.bias = _inject(dl, (1, out_f), rng, -k, k, device)


class MLP:
    """examples/mnist_mlp.py: Linear(784,64) -> relu -> Linear(64,10)."""

    def __init__(self, dl, device, rng, in_dim=784, hidden=64, ncls=10):
        nn = dl.nn
        self.l1 = nn.Linear(in_dim, hidden, bias=True, device=device)
        self.l2 = nn.Linear(hidden, ncls, bias=True, device=device)
        seed_linear(dl, self.l1, rng, device)
        seed_linear(dl, self.l2, rng, device)
        self.layers = [self.l1, dl.Tensor.relu, self.l2]

    def __call__(self, x):
        return x.sequential(self.layers)


class GANPair:
    """examples/mnist_gan.py generator / discriminator stacks (widths scalable for tests)."""

    def __init__(self, dl, device, rng, widths_g=(128, 256, 512, 1024, 784), widths_d=(784, 1024, 512, 256, 1)):
        nn = dl.nn
        self.g_layers, self.d_layers = [], []
        for a, b in zip(widths_g[:-1], widths_g[1:]):
            lin = nn.Linear(a, b, device=device)
            seed_linear(dl, lin, rng, device)
            self.g_layers.append(lin)
        for a, b in zip(widths_d[:-1], widths_d[1:]):
            lin = nn.Linear(a, b, device=device)
            seed_linear(dl, lin, rng, device)
            self.d_layers.append(lin)

    def generator(self, x):
        for i, lin in enumerate(self.g_layers):
            x = lin(x)
            x = x.leaky_relu(0.2) if i < len(self.g_layers) - 1 else x.tanh()
        return x

    def discriminator(self, x):
        for i, lin in enumerate(self.d_layers):
            x = lin(x)
            if i < len(self.d_layers) - 1:
                x = x.leaky_relu(0.2)
        return x


class GPTConfig:
    def __init__(self, vocab_size=96, block_size=128, n_layer=6, n_head=4, n_embd=128, device="cpu"):
        self.vocab_size, self.block_size = vocab_size, block_size
        self.n_layer, self.n_head, self.n_embd = n_layer, n_head, n_embd
        self.device = device


class _Attention:
    def __init__(self, dl, cfg, rng):
        nn = dl.nn
        self.n_head = cfg.n_head
        self.head_dim = cfg.n_embd // cfg.n_head
        self.scale = self.head_dim ** -0.5
        self.q_proj = nn.Linear(cfg.n_embd, cfg.n_embd, bias=False, device=cfg.device)
        self.k_proj = nn.Linear(cfg.n_embd, cfg.n_embd, bias=False, device=cfg.device)
        self.v_proj = nn.Linear(cfg.n_embd, cfg.n_embd, bias=False, device=cfg.device)
        self.c_proj = nn.Linear(cfg.n_embd, cfg.n_embd, bias=False, device=cfg.device)
        for lin in (self.q_proj, self.k_proj, self.v_proj, self.c_proj):
            seed_linear(dl, lin, rng, cfg.device)

    def __call__(self, x):
        Tensor = type(x)     # the framework's Tensor (no module reference is kept: get_parameters walks __dict__)
        B, T, C = x.shape
        q = self.q_proj(x).reshape((B, T, self.n_head, self.head_dim)).transpose(1, 2)
        k = self.k_proj(x).reshape((B, T, self.n_head, self.head_dim)).transpose(1, 2)
        v = self.v_proj(x).reshape((B, T, self.n_head, self.head_dim)).transpose(1, 2)
        att = (q @ k.transpose(2, 3)) * self.scale
        mask = Tensor.tril(Tensor.ones((T, T)), k=0.0)
        att = att.masked_fill(mask == Tensor(0.0), float("-inf"))
        att = att.softmax(dim=3)
        y = att @ v
        y = y.transpose(1, 2).reshape((B, T, C))
        return self.c_proj(y)


class _Block:
    def __init__(self, dl, cfg, rng):
        nn = dl.nn
        self.ln1 = nn.RMSNorm(cfg.n_embd)
        self.attn = _Attention(dl, cfg, rng)
        self.ln2 = nn.RMSNorm(cfg.n_embd)
        self.c_fc = nn.Linear(cfg.n_embd, 4 * cfg.n_embd, bias=False, device=cfg.device)
        self.c_proj = nn.Linear(4 * cfg.n_embd, cfg.n_embd, bias=False, device=cfg.device)
        seed_linear(dl, self.c_fc, rng, cfg.device)
        seed_linear(dl, self.c_proj, rng, cfg.device)

    def __call__(self, x):
        x = x + self.attn(self.ln1(x))
        return x + self.c_proj(self.c_fc(self.ln2(x)).relu())


class GPT:
    """examples/gpt.py:94-119 (training branch): wte + wpe, n_layer blocks, RMSNorm, lm_head,
    cross-entropy over the flattened (B*T, V) logits."""

    def __init__(self, dl, cfg: GPTConfig, rng):
        nn = dl.nn
        self.cfg = cfg
        self.wte = nn.Embedding(cfg.vocab_size, cfg.n_embd)
        self.wpe = nn.Embedding(cfg.block_size, cfg.n_embd)
        # the reference's Embedding takes no device: its table is a host tensor (U[0,1))
        self.wte.weight = _inject(dl, (cfg.vocab_size, cfg.n_embd), rng, 0.0, 1.0, "cpu")
        self.wpe.weight = _inject(dl, (cfg.block_size, cfg.n_embd), rng, 0.0, 1.0, "cpu")
        self.blocks = [_Block(dl, cfg, rng) for _ in range(cfg.n_layer)]
        self.ln_f = nn.RMSNorm(cfg.n_embd)
        self.lm_head = nn.Linear(cfg.n_embd, cfg.vocab_size, bias=False)
        seed_linear(dl, self.lm_head, rng, "cpu")

    def __call__(self, idx, targets):
        Tensor = type(idx)
        B, T = idx.shape
        tok = self.wte(idx)
        pos = self.wpe(Tensor(np.arange(T).astype(np.float32), device=idx.device))
        x = tok + pos
        for blk in self.blocks:
            x = blk(x)
        x = self.ln_f(x)
        logits = self.lm_head(x)
        loss = logits.reshape((B * T, self.cfg.vocab_size)).cross_entropy_loss(targets.reshape((B * T, 1)))
        return logits, loss


def gpt_batch(cfg: GPTConfig, batch: int, rng, T: int | None = None):
    T = T or cfg.block_size
    toks = rng.integers(0, cfg.vocab_size, size=(batch, T + 1))
    return toks[:, :-1].astype(np.float32).copy(), toks[:, 1:].astype(np.float32).copy()


def gpt_train_flops(cfg: GPTConfig, batch: int, T: int | None = None) -> float:
    """3*[L*(24*B*T*C^2 + 4*B*T^2*C) + 2*B*T*C*V]."""
    T = T or cfg.block_size
    C, L, V = cfg.n_embd, cfg.n_layer, cfg.vocab_size
    return 3.0 * (L * (24.0 * batch * T * C * C + 4.0 * batch * T * T * C) + 2.0 * batch * T * C * V)


def generate_on_device(gpt: GPT, idx, max_new_tokens: int, temperature: float = 0.7):
    """The KV-cache decode loop of examples/gpt.py:131-153,173-229 with nothing but token indices leaving the device:
    keys / values are appended with Tensor.cat, the next token is drawn with Tensor.sample_last and fed straight back
    into the embedding.  `idx` is a (1, T0) prompt; returns the (1, T0 + max_new_tokens) token tensor."""
    Tensor = type(idx)
    cfg = gpt.cfg
    dev = idx.device
    prompt = idx.numpy().reshape(-1) if idx.shape[1] > 1 else None
    caches = [None] * len(gpt.blocks)
    toks = idx
    n_prompt = idx.shape[1]
    cur = idx if n_prompt == 1 else Tensor(np.array([[prompt[0]]], dtype=np.float32), device=dev)
    for pos in range(n_prompt + max_new_tokens - 1):
        x = gpt.wte(cur) + gpt.wpe(Tensor(np.array([float(pos)], dtype=np.float32), device=dev))
        for i, blk in enumerate(gpt.blocks):
            a = blk.attn
            h = blk.ln1(x)
            q = a.q_proj(h).reshape((1, 1, a.n_head, a.head_dim)).transpose(1, 2)
            k = a.k_proj(h).reshape((1, 1, a.n_head, a.head_dim)).transpose(1, 2)
            v = a.v_proj(h).reshape((1, 1, a.n_head, a.head_dim)).transpose(1, 2)
            if caches[i] is not None:
                k, v = Tensor.cat([caches[i][0], k], dim=2), Tensor.cat([caches[i][1], v], dim=2)
            caches[i] = (k, v)
            att = ((q @ k.transpose(2, 3)) * a.scale).softmax(dim=3)
            y = (att @ v).transpose(1, 2).reshape((1, 1, cfg.n_embd))
            x = x + a.c_proj(y)
            x = x + blk.c_proj(blk.c_fc(blk.ln2(x)).relu())
        if pos + 1 < n_prompt:
            cur = Tensor(np.array([[prompt[pos + 1]]], dtype=np.float32), device=dev)
            continue
        cur = gpt.lm_head(gpt.ln_f(x)).sample_last(temperature)      # (1, 1) index tensor, stays on the device
        toks = Tensor.cat([toks, cur], dim=1)
    return toks
