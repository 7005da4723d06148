"""Generates tests/golden/*.npz by running the REAL reference (python package at /root/reference,
CPU path: generated C compiled with gcc -O3 -march=native -ffast-math) on seeded inputs.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own (its tests compare with torch on unseeded inputs,
SURVEY §4), so these files are the pin: oracle/ must reproduce them (tests/test_oracle_golden.py)
and the CUDA path is compared with both.  Every case stores its inputs next to its outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


def _import_reference():
    os.makedirs(os.path.expanduser("~/.cache/dlgrad"), exist_ok=True)   # the reference never creates it
    sys.path.insert(0, REFERENCE)
    import dlgrad  # noqa: F401
    from dlgrad import Tensor, nn  # noqa: F401
    return dlgrad


def run_ops(dl, out_path: str) -> None:
    sys.path.insert(0, REPO)
    from tests.cases import op_cases
    Tensor = dl.Tensor
    rng = np.random.default_rng(1234)
    store = {}
    for name, (fn, inputs, diff) in op_cases(rng).items():
        ts = [Tensor(a.copy(), requires_grad=diff) for a in inputs]
        y = fn(*ts)
        store[f"{name}/n_in"] = np.array(len(inputs))
        for i, a in enumerate(inputs):
            store[f"{name}/in{i}"] = a
        store[f"{name}/out"] = np.array(y.numpy(), dtype=np.float32).reshape(y.shape)
        if diff:
            w = np.random.default_rng(99).random(y.shape, dtype=np.float32).astype(np.float32) if y.shape else None
            loss = (y * Tensor(w)).sum() if w is not None else y
            if w is not None:
                store[f"{name}/w"] = w
            try:
                loss.backward()
            except Exception as exc:      # e.g. Sqrt.backward multiplies a Buffer by an int and raises
                print(f"  reference cannot differentiate {name}: {type(exc).__name__}: {exc}")
                continue
            for i, t in enumerate(ts):
                if t.grad is not None:
                    store[f"{name}/grad{i}"] = np.array(t.grad.numpy(), dtype=np.float32).reshape(t.grad.shape)
    # integer / index work: bit-exact cases
    a = np.random.default_rng(5).integers(0, 4, size=(37, 11)).astype(np.float32)     # many ties
    for d in (0, 1):
        store[f"argmax_d{d}/in0"] = a
        store[f"argmax_d{d}/out"] = np.array(Tensor(a.copy()).argmax(dim=d).numpy(), dtype=np.float32)
    x = (np.random.default_rng(6).random((5, 6), dtype=np.float32) - 0.5).astype(np.float32)
    store["gt/in0"] = x
    store["gt/out"] = np.array((Tensor(x.copy()) > 0.1).numpy())
    store["gte/out"] = np.array((Tensor(x.copy()) >= 0.1).numpy())
    store["eq/out"] = np.array((Tensor(np.round(x * 3)) == Tensor(np.zeros_like(x))).numpy())
    cond = (np.random.default_rng(7).random((5, 6)) > 0.5).astype(np.float32)
    store["where/cond"] = cond
    store["where/out"] = np.array(Tensor.where(Tensor(cond.copy()), Tensor(x.copy()), 3.0).numpy())
    big = (np.random.default_rng(8).random((2, 3, 6, 6), dtype=np.float32)).astype(np.float32)
    tm = Tensor.tril(Tensor.ones((6, 6)), k=0.0)
    store["tril/out"] = np.array(tm.numpy())
    mf = Tensor(big.copy()).masked_fill(tm == Tensor(0.0), float("-inf"))
    store["masked_fill/in0"] = big
    store["masked_fill/out"] = np.array(mf.numpy())
    store["full_0p5/out"] = np.array(Tensor.full((3, 4), 0.5).numpy())          # int truncation quirk
    store["full_2p9/out"] = np.array(Tensor.full((3, 4), 2.9).numpy())
    # embedding: forward gather + ordered scatter-add gradient
    W = np.random.default_rng(9).random((11, 8), dtype=np.float32)
    idx = np.random.default_rng(10).integers(0, 11, size=(6, 9)).astype(np.float32)
    wt = Tensor(W.copy(), requires_grad=True)
    e = Tensor.embedding(wt, Tensor(idx.copy()))
    g = np.random.default_rng(11).random(e.shape, dtype=np.float32)
    (e * Tensor(g)).sum().backward()
    store.update({"embedding/W": W, "embedding/idx": idx, "embedding/g": g,
                  "embedding/out": np.array(e.numpy()), "embedding/grad": np.array(wt.grad.numpy())})
    # cross entropy (sum-reduced quirk) value + grad
    logits = (np.random.default_rng(12).random((32, 10), dtype=np.float32) * 6 - 3).astype(np.float32)
    tgt = np.random.default_rng(13).integers(0, 10, size=(32, 1)).astype(np.float32)
    lt = Tensor(logits.copy(), requires_grad=True)
    loss = lt.cross_entropy_loss(Tensor(tgt.copy()))
    loss.backward()
    store.update({"ce/logits": logits, "ce/target": tgt, "ce/loss": np.array(loss.numpy()).reshape(()),
                  "ce/grad": np.array(lt.grad.numpy())})
    # BCE with logits
    p = (np.random.default_rng(14).random((16, 1), dtype=np.float32) * 4 - 2).astype(np.float32)
    tb = (np.random.default_rng(15).random((16, 1)) > 0.5).astype(np.float32)
    pt = Tensor(p.copy(), requires_grad=True)
    bl = pt.bcewithlogitsloss(Tensor(tb.copy()))
    bl.backward()
    store.update({"bce/p": p, "bce/t": tb, "bce/loss": np.array(bl.numpy()).reshape(()), "bce/grad": np.array(pt.grad.numpy())})
    # RMSNorm forward + backward (Mean.backward-is-zero quirk)
    xr = (np.random.default_rng(16).random((4, 9, 32), dtype=np.float32) - 0.5).astype(np.float32)
    norm = dl.nn.RMSNorm(32)
    xt = Tensor(xr.copy(), requires_grad=True)
    yn = norm(xt)
    gw = np.random.default_rng(17).random(yn.shape, dtype=np.float32)
    (yn * Tensor(gw)).sum().backward()
    store.update({"rmsnorm/x": xr, "rmsnorm/g": gw, "rmsnorm/out": np.array(yn.numpy()),
                  "rmsnorm/grad_x": np.array(xt.grad.numpy()), "rmsnorm/grad_w": np.array(norm.weight.grad.numpy())})
    np.savez_compressed(out_path, **store)
    print(f"wrote {out_path}: {len(store)} arrays")


def run_models(dl, out_path: str) -> None:
    sys.path.insert(0, REPO)
    from tests import models
    Tensor, nn = dl.Tensor, dl.nn
    store = {}
    # ---- MLP (config 1): 784-64-10, batch 128, SGD 1e-3, 3 steps --------------------------------
    rng = np.random.default_rng(0)
    mlp = models.MLP(dl, "cpu", rng)
    x = rng.random((128, 784), dtype=np.float32)
    y = rng.integers(0, 10, size=(128, 1)).astype(np.float32)
    opt = nn.optim.SGD(nn.utils.get_parameters(mlp), lr=1e-3)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = mlp(Tensor(x.copy())).cross_entropy_loss(Tensor(y.copy()))
        losses.append(float(loss.numpy()))
        loss.backward()
        if _ == 0:
            store["mlp/grad_w1_step0"] = np.array(mlp.l1.weight.grad.numpy())
        opt.step()
    pred = mlp(Tensor(x.copy())).argmax(dim=1)
    store.update({"mlp/x": x, "mlp/y": y, "mlp/losses": np.array(losses, dtype=np.float64),
                  "mlp/w1": np.array(mlp.l1.weight.numpy()), "mlp/b2": np.array(mlp.l2.bias.numpy()),
                  "mlp/pred": np.array(pred.numpy())})
    # ---- GAN (config 3, narrow): one D step + one G step, Adam(5e-4, (0.5, 0.999)) ---------------
    rng = np.random.default_rng(1)
    gan = models.GANPair(dl, "cpu", rng, widths_g=(16, 32, 48), widths_d=(48, 32, 1))
    real = (rng.random((10, 48), dtype=np.float32) * 2 - 1).astype(np.float32)
    noise1 = (rng.random((10, 16), dtype=np.float32) * 2 - 1).astype(np.float32)
    noise2 = (rng.random((10, 16), dtype=np.float32) * 2 - 1).astype(np.float32)
    opt_g = nn.optim.Adam(nn.utils.get_parameters(gan.g_layers), lr=0.0005, betas=(0.5, 0.999))
    opt_d = nn.optim.Adam(nn.utils.get_parameters(gan.d_layers), lr=0.0005, betas=(0.5, 0.999))
    fake = gan.generator(Tensor(noise1.copy())).detach()
    opt_d.zero_grad()
    d_loss = (gan.discriminator(Tensor(real.copy())).bcewithlogitsloss(Tensor.full((10, 1), 1.0))
              + gan.discriminator(fake).bcewithlogitsloss(Tensor.full((10, 1), 0.0))) / 2.0
    d_loss.backward()
    opt_d.step()
    opt_g.zero_grad()
    g_loss = gan.discriminator(gan.generator(Tensor(noise2.copy()))).bcewithlogitsloss(Tensor.full((10, 1), 1.0))
    g_loss.backward()
    opt_g.step()
    store.update({"gan/real": real, "gan/noise1": noise1, "gan/noise2": noise2,
                  "gan/d_loss": np.array(float(d_loss.numpy())), "gan/g_loss": np.array(float(g_loss.numpy())),
                  "gan/g_w0": np.array(gan.g_layers[0].weight.numpy()), "gan/d_w0": np.array(gan.d_layers[0].weight.numpy())})
    # ---- char-GPT (config 4 architecture, reduced: L2 h2 C32 T16 B4 V17), Adam 1e-4, 3 steps ------
    rng = np.random.default_rng(2)
    cfg = models.GPTConfig(vocab_size=17, block_size=16, n_layer=2, n_head=2, n_embd=32, device="cpu")
    gpt = models.GPT(dl, cfg, rng)
    xb, yb = models.gpt_batch(cfg, 4, rng)
    opt = nn.optim.Adam(nn.utils.get_parameters(gpt), lr=1e-4)
    losses = []
    for step in range(3):
        opt.zero_grad()
        _, loss = gpt(Tensor(xb.copy()), Tensor(yb.copy()))
        losses.append(float(loss.numpy()))
        loss.backward()
        if step == 0:
            store["gpt/grad_wte_step0"] = np.array(gpt.wte.weight.grad.numpy())
            store["gpt/grad_q0_step0"] = np.array(gpt.blocks[0].attn.q_proj.weight.grad.numpy())
            store["gpt/grad_lnf_step0"] = np.array(gpt.ln_f.weight.grad.numpy())
        opt.step()
    store.update({"gpt/x": xb, "gpt/y": yb, "gpt/losses": np.array(losses, dtype=np.float64),
                  "gpt/wte": np.array(gpt.wte.weight.numpy()), "gpt/q0": np.array(gpt.blocks[0].attn.q_proj.weight.numpy()),
                  "gpt/head": np.array(gpt.lm_head.weight.numpy())})
    np.savez_compressed(out_path, **store)
    print(f"wrote {out_path}: {len(store)} arrays; mlp losses {store['mlp/losses']}, gpt losses {store['gpt/losses']}")


if __name__ == "__main__":
    if not os.path.isdir(REFERENCE):
        sys.exit("the reference is not mounted here; golden files can only be regenerated in the build container")
    ref = _import_reference()
    run_ops(ref, os.path.join(HERE, "ops.npz"))
    run_models(ref, os.path.join(HERE, "models.npz"))
