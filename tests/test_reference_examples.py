"""The north star's "dlgrad scripts (MLP, GAN, GPT) run unchanged with device='cuda'": the reference's OWN example
sources (/root/reference/examples/{mnist_mlp,mnist_gan,gpt}.py) are loaded with `dlgrad` aliased to this package and
their classes / functions / training-loop bodies are executed

  * on the CPU oracle (device="cpu", `-m "not gpu"`): a real numerical run of the unchanged script code on this
    package's Tensor / nn / optim layer, here in the build container;
  * on Device.CUDA (`-m gpu`): the same, on a B200 — and walked in compile-only mode by __graft_entry__.build().

The scripts cannot travel (reference sources are never copied into this repository), so every test here skips where
/root/reference is absent (the GPU box).  What is substituted, and only that: the `device = "cpu"` line, loop counts,
plotting / image-grid modules that are not installed (matplotlib, torchvision), and synthetic data for the MNIST
files and the TinyStories text (no network).
"""
import importlib.util
import os
import re
import struct
import sys
import types

import numpy as np
import pytest

from tests.conftest import COMPILE_ONLY

EXAMPLES = "/root/reference/examples"
needs_reference = pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason="/root/reference is not mounted here")


def _alias_dlgrad():
    """`import dlgrad` -> dlgrad_b200 (and its submodules), plus stubs for plotting modules the image lacks."""
    import dlgrad_b200
    import dlgrad_b200.nn
    import dlgrad_b200.nn.datasets
    import dlgrad_b200.nn.optim
    import dlgrad_b200.nn.utils
    saved = {k: sys.modules.get(k) for k in ("dlgrad", "dlgrad.nn", "dlgrad.nn.utils", "dlgrad.nn.datasets", "dlgrad.nn.optim",
                                               "matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.utils")}
    sys.modules["dlgrad"] = dlgrad_b200
    sys.modules["dlgrad.nn"] = dlgrad_b200.nn
    sys.modules["dlgrad.nn.utils"] = dlgrad_b200.nn.utils
    sys.modules["dlgrad.nn.datasets"] = dlgrad_b200.nn.datasets
    sys.modules["dlgrad.nn.optim"] = dlgrad_b200.nn.optim
    for name in ("matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.utils"):
        if saved[name] is None:
            m = types.ModuleType(name)
            m.__dict__.update(plot=lambda *a, **k: None, title=lambda *a, **k: None, xlabel=lambda *a, **k: None,
                              ylabel=lambda *a, **k: None, savefig=lambda *a, **k: None,
                              make_grid=lambda t, **k: t, save_image=lambda *a, **k: None)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["torchvision"].utils = sys.modules["torchvision.utils"]
    return saved


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _load_script(name: str):
    """Import an example file as a module (its `if __name__ == "__main__":` part does not run)."""
    spec = importlib.util.spec_from_file_location(f"_ref_example_{name}", os.path.join(EXAMPLES, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _device_fixture(device, request):
    if device == "cpu":
        request.getfixturevalue("oracle")


def _run_gpt(device: str) -> dict:
    saved = _alias_dlgrad()
    try:
        mod = _load_script("gpt")
        from dlgrad_b200 import Tensor, nn
        cfg = mod.config
        cfg.vocab_size, cfg.block_size, cfg.n_layer, cfg.n_head, cfg.n_embd = 17, 16, 2, 2, 32
        cfg.batch_size, cfg.device = 4, device
        model = mod.GPT(cfg)                                   # examples/gpt.py:94-119
        opt = nn.optim.Adam(params=nn.utils.get_parameters(model), lr=cfg.learning_rate)
        data = np.random.default_rng(0).integers(0, cfg.vocab_size, size=4000).astype(np.int32)
        rng = np.random.default_rng(1)
        losses = []
        for _ in range(2):                                     # the loop body of examples/gpt.py:366-387
            ix = rng.integers(0, len(data) - cfg.block_size, (cfg.batch_size,))
            x_batch = np.stack([data[j: j + cfg.block_size] for j in ix]).astype(np.float32)
            y_batch = np.stack([data[j + 1: j + cfg.block_size + 1] for j in ix]).astype(np.float32)
            x_t, y_t = Tensor(x_batch, device=cfg.device), Tensor(y_batch, device=cfg.device)
            opt.zero_grad()
            logits, loss = model(x_t, y_t)
            losses.append(float(loss.numpy()))
            loss.backward()
            opt.step()
        # inference: the script's own KV-cache generate loop (examples/gpt.py:173-229), 3 tokens
        np.random.seed(0)
        out = model.generate(Tensor(np.array([[1.0]], dtype=np.float32), device=cfg.device), max_new_tokens=3)
        return {"losses": losses, "generated": out.numpy().reshape(-1), "logits_shape": logits.shape}
    finally:
        _restore(saved)


def _run_gan(device: str) -> dict:
    saved = _alias_dlgrad()
    try:
        mod = _load_script("mnist_gan")
        from dlgrad_b200 import Tensor, nn
        mod.device = device
        gen, disc = mod.LinearGen(), mod.LinearDisc()           # examples/mnist_gan.py:30-58
        optim_g = nn.optim.Adam(nn.utils.get_parameters(gen), lr=0.0005, betas=(0.5, 0.999))
        optim_d = nn.optim.Adam(nn.utils.get_parameters(disc), lr=0.0005, betas=(0.5, 0.999))
        images = np.random.default_rng(0).integers(0, 256, size=(500, 28, 28)).astype(np.float32)
        np.random.seed(0)
        data_real = mod.make_batch(images)                      # :60-63
        # :121-126 — the script passes INT bounds to Tensor.uniform (the reference itself cannot compile that)
        noise = Tensor.uniform((mod.BS, 128), low=-1, high=1, device=device)
        data_fake = gen(noise).detach()
        loss_d = mod.train_discriminator(disc, optim_d, data_real, data_fake)      # :69-83
        noise = Tensor.uniform((mod.BS, 128), low=-1, high=1, device=device)
        loss_g = mod.train_generator(disc, optim_g, gen(noise))                     # :85-94
        fake = gen(noise).detach().numpy()
        return {"loss_d": float(loss_d.numpy()), "loss_g": float(loss_g.numpy()), "fake_shape": fake.shape,
                "fake_range": (float(np.nanmin(fake)), float(np.nanmax(fake)))}
    finally:
        _restore(saved)


def _write_idx(path, arr, images):
    with open(path, "wb") as f:
        if images:
            f.write(struct.pack(">IIII", 2051, arr.shape[0], 28, 28))
        else:
            f.write(struct.pack(">II", 2049, arr.shape[0]))
        f.write(arr.astype(np.uint8).tobytes())


def _run_mlp(device: str, tmp_path) -> dict:
    """examples/mnist_mlp.py trains at import time: the whole file is executed, with synthetic MNIST-sized idx files."""
    rng = np.random.default_rng(0)
    _write_idx(tmp_path / "train-images-idx3-ubyte", rng.integers(0, 256, size=(60000, 784)), True)
    _write_idx(tmp_path / "train-labels-idx1-ubyte", rng.integers(0, 10, size=(60000,)), False)
    _write_idx(tmp_path / "t10k-images-idx3-ubyte", rng.integers(0, 256, size=(10000, 784)), True)
    _write_idx(tmp_path / "t10k-labels-idx1-ubyte", rng.integers(0, 10, size=(10000,)), False)
    src = open(os.path.join(EXAMPLES, "mnist_mlp.py")).read()
    src, n1 = re.subn(r'^device = "cpu".*$', f'device = "{device}"', src, count=1, flags=re.M)
    src, n2 = re.subn(r"^EPOCHS = 5$", "EPOCHS = 1", src, count=1, flags=re.M)
    src, n3 = re.subn(r"^STEPS = .*$", "STEPS = 3", src, count=1, flags=re.M)
    assert (n1, n2, n3) == (1, 1, 1), "examples/mnist_mlp.py no longer has the lines this harness substitutes"
    saved = _alias_dlgrad()
    old = os.environ.get("DLGRAD_DATA")
    os.environ["DLGRAD_DATA"] = str(tmp_path)
    try:
        ns = {"__name__": "_ref_example_mnist_mlp"}
        exec(compile(src, os.path.join(EXAMPLES, "mnist_mlp.py"), "exec"), ns)
        return {"loss": float(ns["loss"].numpy()), "test_acc": float(np.asarray(ns["test_acc"].numpy()).reshape(-1)[0]),
                "pred_shape": ns["y_pred"].shape}
    finally:
        if old is None:
            os.environ.pop("DLGRAD_DATA", None)
        else:
            os.environ["DLGRAD_DATA"] = old
        _restore(saved)


def _check_gpt(res):
    assert res["logits_shape"] == (4, 16, 17)
    if COMPILE_ONLY:
        return
    assert all(np.isfinite(res["losses"])) and res["losses"][0] > 0
    # sum-reduced cross-entropy: ~ B*T*ln(V) at init
    assert 0.5 * 64 * np.log(17) < res["losses"][0] < 40 * 64 * np.log(17)
    assert res["generated"].shape == (4,) and ((res["generated"] >= 0) & (res["generated"] < 17)).all()


def _check_gan(res):
    assert res["fake_shape"] == (100, 784)
    if COMPILE_ONLY:
        return
    assert np.isfinite(res["loss_d"]) and np.isfinite(res["loss_g"]) and res["loss_d"] > 0 and res["loss_g"] > 0
    assert -1.0 <= res["fake_range"][0] <= res["fake_range"][1] <= 1.0          # tanh output


def _check_mlp(res):
    assert res["pred_shape"] == (10000, 1)
    if COMPILE_ONLY:
        return
    assert np.isfinite(res["loss"]) and 0.0 <= res["test_acc"] <= 100.0


@needs_reference
def test_reference_gpt_script_classes_run_on_the_oracle(oracle):
    _check_gpt(_run_gpt("cpu"))


@needs_reference
def test_reference_gan_script_classes_run_on_the_oracle(oracle):
    _check_gan(_run_gan("cpu"))


@needs_reference
def test_reference_mlp_script_runs_on_the_oracle(oracle, tmp_path):
    _check_mlp(_run_mlp("cpu", tmp_path))


@needs_reference
@pytest.mark.gpu
def test_reference_gpt_script_classes_run_on_cuda():
    _check_gpt(_run_gpt("cuda"))


@needs_reference
@pytest.mark.gpu
def test_reference_gan_script_classes_run_on_cuda():
    _check_gan(_run_gan("cuda"))


@needs_reference
@pytest.mark.gpu
def test_reference_mlp_script_runs_on_cuda(tmp_path):
    _check_mlp(_run_mlp("cuda", tmp_path))
