"""Round-2 parity gates (`-m gpu`): the headline config's own shapes against the oracle and float64.

VERDICT r1 "weak 1" listed the holes these close:
  (a) one C5 transformer block (T=1024, C=768, h=12), B=1, forward + backward vs the oracle (oracle/_ref when built):
      y, dx and all six weight gradients — the block bench.py's CPU arm times (`cpu_block_sample`);
  (b) every GEMM shape of the bench line's `by_shape` list, through the variants that run inside the step
      (plain NT / NN / TN, split-K, bias, relu operand A and B, relu-mask epilogue) vs float64;
  (c) the full-width GAN of examples/mnist_gan.py:30-58 (G 128-256-512-1024-784, D 784-1024-512-256-1, batch 100),
      one discriminator + one generator Adam step vs the oracle — including its N=1 / K=1 GEMMs;
  (d) the C4 char-GPT loss curve over 30 Adam steps, CUDA vs oracle, tolerance stated in the test;
  (e) an attention block with a FINITE mask fill (P is dense then; no tile may be skipped).
"""
import numpy as np
import pytest

from tests import models, runner
from tests.conftest import COMPILE_ONLY, check

pytestmark = pytest.mark.gpu


def _weights_of_block(blk):
    a = blk.attn
    return {"q_proj": a.q_proj.weight, "k_proj": a.k_proj.weight, "v_proj": a.v_proj.weight, "c_proj(attn)": a.c_proj.weight,
            "c_fc": blk.c_fc.weight, "c_proj(mlp)": blk.c_proj.weight}


def test_c5_transformer_block_forward_backward_vs_oracle(oracle):
    """(a) BASELINE configs[4] dimensions, one sequence: T = 1024, C = 768, 12 heads.  Every fusion of the headline
    step engages at this size (tensor-core GEMMs at the bench shapes' N / K, fused attention, deferred relu, fused
    RMSNorm).  Tolerances: activations and input gradient 1e-4 of the largest magnitude (matmul bound of the north
    star), weight gradients 1e-4 as well (K = 1024 tokens summed)."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor
    cfg_kw = dict(vocab_size=96, block_size=1024, n_layer=12, n_head=12, n_embd=768)
    x = (np.random.default_rng(0).random((1, 1024, 768), dtype=np.float32) - 0.5).astype(np.float32)
    coef = np.linspace(0.5, 1.5, x.size, dtype=np.float32).reshape(x.shape)
    res, pre = {}, {}
    for dev in ("cuda", "cpu"):
        cfg = models.GPTConfig(device=dev, **cfg_kw)
        blk = models._Block(dl, cfg, np.random.default_rng(7))
        xt = Tensor(x.copy(), requires_grad=True, device=dev)
        y = blk(xt)
        (y * Tensor(coef.copy(), device=dev)).sum().backward()
        out = {"y": y.numpy(), "dx": xt.grad.numpy()}
        for name, w in _weights_of_block(blk).items():
            out["d" + name] = w.grad.numpy()
        out["dln1"], out["dln2"] = blk.ln1.weight.grad.numpy(), blk.ln2.weight.grad.numpy()
        res[dev] = out
        # the relu's pre-activation, recomputed with the block's own layers (no gradient): see below
        x0 = Tensor(x.copy(), device=dev)
        x1 = x0 + blk.attn(blk.ln1(x0))
        pre[dev] = blk.c_fc(blk.ln2(x1)).numpy().reshape(1024, 3072)
    if COMPILE_ONLY:
        return
    # relu is discontinuous: a pre-activation within rounding of zero lands on different sides in the oracle's
    # sequential-fp32 GEMM and in 3xTF32 (expected here: 3.1 M units x P(|h| < 1e-6) ~ a handful).  One flipped unit
    # (token r, hidden unit c) moves row r of dx and row c of dc_fc by O(1e-3) of their scale — a property of the
    # function, not an error of either side.  Flips must be few and must sit within 1e-5 of zero; the rows they touch
    # are compared at 1e-2, everything else at the stated 1e-4.
    flips = np.argwhere((pre["cuda"] > 0) != (pre["cpu"] > 0))
    assert len(flips) <= 64, f"{len(flips)} relu units have different signs on the two sides"
    if len(flips):
        assert np.abs(pre["cpu"][flips[:, 0], flips[:, 1]]).max() < 1e-5 * np.abs(pre["cpu"]).max()
    tok, unit = np.unique(flips[:, 0]) if len(flips) else [], np.unique(flips[:, 1]) if len(flips) else []
    for k in res["cpu"]:
        a, b = res["cuda"][k], res["cpu"][k]
        if k == "dx" and len(tok):
            keep = np.ones(1024, dtype=bool)
            keep[tok] = False
            scale = float(np.abs(b).max())
            check(a[0][~keep], b[0][~keep], rtol=1e-2, atol=1e-2 * scale, what="C5 block dx, tokens with a flipped relu unit")
            # the flipped tokens' gradient also reaches earlier tokens through the attention backward (dk, dv): second
            # order, measured <= 1.6e-4 of the largest magnitude on 18 of 783 k entries -> 3e-4 for the other tokens
            runner.rel_check(check, a[0][keep], b[0][keep], 3e-4, "C5 block dx")
        elif k == "dc_fc" and len(unit):
            keep = np.ones(3072, dtype=bool)
            keep[unit] = False
            scale = float(np.abs(b).max())
            check(a[~keep], b[~keep], rtol=1e-2, atol=1e-2 * scale, what="C5 block dc_fc, hidden units with a flipped relu")
            runner.rel_check(check, a[keep], b[keep], runner.TOL_MATMUL, "C5 block dc_fc")
        elif k in ("dln1", "dln2") and len(flips):
            # sums over all tokens of a gradient that contains the flipped rows: each flip moves an entry by ~5e-4 of the scale
            runner.rel_check(check, a, b, 2e-3, f"C5 block {k}")
        elif k in ("dq_proj", "dk_proj"):
            # gradients that pass through the softmax backward: the oracle differentiates the reference's UNFUSED fp32 chain
            # (max / sub / exp / sum / div over 1024-long rows), which is ill-conditioned on small entries — its own result
            # moves by ~1e-5 with the summation order (tests/test_oracle_golden.py) and sits ~5e-4 of the largest magnitude
            # from float64 at this size, while the flash kernels stay within 1e-4 of float64
            # (test_flash_attention_forward_and_backward[8-12-1024-64]).  Held to 1e-3, like the GPT-level gradients.
            runner.rel_check(check, a, b, 1e-3, f"C5 block {k}")
        else:
            runner.rel_check(check, a, b, runner.TOL_MATMUL, f"C5 block {k}")


def _f64(a):
    return np.asarray(a, dtype=np.float64)


@pytest.mark.parametrize("rows,in_f,out_f,bias", [(8192, 768, 768, False), (8192, 768, 768, True), (8192, 768, 96, False)])
def test_bench_linear_shapes_vs_float64(rows, in_f, out_f, bias):
    """(b) the projection GEMMs of the C5 step through ops.Linear: forward x @ W^T (8192 x out x in, NT), dX = g @ W
    (NN), dW = g^T @ x (TN; 768 x 768 x 8192 and 96 x 768 x 8192 run as in-kernel split-K), bias in the epilogue.
    All against float64: 1e-4 of the largest magnitude AND, per entry, 1e-4 x (|A| |B|)_ij (runner.gemm_check)."""
    from dlgrad_b200 import Tensor
    rng = np.random.default_rng(rows + in_f + out_f + int(bias))
    x = (rng.random((rows, in_f), dtype=np.float32) - 0.5).astype(np.float32)
    w = ((rng.random((out_f, in_f), dtype=np.float32) - 0.5) * 0.1).astype(np.float32)
    b = (rng.random((1, out_f), dtype=np.float32) - 0.5).astype(np.float32)
    g = (rng.random((rows, out_f), dtype=np.float32) - 0.5).astype(np.float32)
    X = Tensor(x.copy(), device="cuda", requires_grad=True)
    W = Tensor(w.copy(), device="cuda", requires_grad=True)
    Bt = Tensor(b.copy(), device="cuda", requires_grad=True) if bias else None
    y = X.linear(W, Bt)
    (y * Tensor(g.copy(), device="cuda")).sum().backward()
    runner.gemm_check(check, y.numpy(), x, w.T, runner.TOL_MATMUL, f"linear {rows}x{out_f}x{in_f} forward", bias=b if bias else None)
    runner.gemm_check(check, X.grad.numpy(), g, w, runner.TOL_MATMUL, f"linear {rows}x{in_f}x{out_f} dX")
    runner.gemm_check(check, W.grad.numpy(), g.T, x, runner.TOL_MATMUL, f"linear {out_f}x{in_f}x{rows} dW")
    if bias:
        runner.rel_check(check, Bt.grad.numpy(), _f64(g).sum(0, keepdims=True), 1e-5, "linear db")


@pytest.mark.parametrize("B", [8, 2])
def test_sibling_projections_run_as_grouped_launches(B):
    """The q / k / v projections of an attention block (examples/gpt.py:43-45: three Linear layers on the same normed x,
    each followed by the head split) at the C5 sizes.  Forward: ONE grouped launch [x @ Wq^T, x @ Wk^T, x @ Wv^T]
    (matmul_tc_ts group ("n", 3): 7.8 tile waves instead of 3 x 2.6) — every output BIT-IDENTICAL to the single launch it
    replaces and within the componentwise 1e-4 bound of float64.  Backward: the three input gradients meet in the
    autograd driver's `grad + g` and run as ONE K-concatenated launch dq @ Wq + dk @ Wk + dv @ Wv (group ("k", 3)),
    checked against float64 with the bound of the concatenated dot product; the three weight gradients dW_g = dY_g^T @ x
    run as ONE grouped split-K launch (group ("m", 3): 3 x 4 K-slices per output tile, each member's partial sums added in
    slice order by the slice that finishes last) — bit-identical to the single split-K launches."""
    from dlgrad_b200 import Tensor
    from dlgrad_b200.runtime.cuda import profile
    T, C, H = 1024, 768, 12
    rows = B * T
    rng = np.random.default_rng(40 + B)
    x = (rng.random((B, T, C), dtype=np.float32) - 0.5).astype(np.float32)
    w = [((rng.random((C, C), dtype=np.float32) - 0.5) * 0.1).astype(np.float32) for _ in range(3)]
    g = [(rng.random((B, H, T, C // H), dtype=np.float32) - 0.5).astype(np.float32) for _ in range(3)]
    X = Tensor(x.copy(), device="cuda", requires_grad=True)
    W = [Tensor(a.copy(), device="cuda", requires_grad=True) for a in w]
    if not COMPILE_ONLY:
        profile.matmul_timer = profile.KernelTimer()
    heads = [X.linear(Wi, None).reshape((B, T, H, C // H)).transpose(1, 2) for Wi in W]      # (B, h, T, d), like the script
    loss = None
    for hd, gi in zip(heads, g):
        term = (hd * Tensor(gi.copy(), device="cuda")).sum()
        loss = term if loss is None else loss + term
    loss.backward()
    outs = [hd.numpy() for hd in heads]
    dx = X.grad.numpy()                                   # the deferred sum of the three input gradients runs here
    if not COMPILE_ONLY:
        tags = [t[0] for t in profile.matmul_timer.collect()["by_tag"]]
        profile.matmul_timer = None
        assert "tc3_group_n3" in tags and "tc3_group_k3" in tags and "tc3_splitk_group3" in tags, tags
    profile.matmul_timer = None
    x2 = x.reshape(rows, C)
    g2 = [gi.transpose(0, 2, 1, 3).reshape(rows, C) for gi in g]          # gradient of each projection output, (rows, C)
    for i in range(3):
        got = outs[i].transpose(0, 2, 1, 3).reshape(rows, C)
        runner.gemm_check(check, got, x2, w[i].T, runner.TOL_MATMUL, f"grouped projection {i} forward (B={B})")
        single = (Tensor(x.copy(), device="cuda").linear(Tensor(w[i].copy(), device="cuda"), None)).numpy().reshape(rows, C)
        check(got, single, exact=True, what=f"grouped projection {i} is bit-identical to the single launch (B={B})")
        runner.gemm_check(check, W[i].grad.numpy(), g2[i].T, x2, runner.TOL_MATMUL, f"grouped projection {i} dW (B={B})")
    runner.gemm_check(check, dx.reshape(rows, C), np.concatenate(g2, axis=1), np.concatenate(w, axis=0),
                      runner.TOL_MATMUL, f"K-concatenated input gradient dq Wq + dk Wk + dv Wv (B={B})")


def test_grouped_launches_are_optional_and_change_no_values(monkeypatch):
    """DLGRAD_GROUP_LINEAR=0 (three launches + two add epilogues) and the grouped launches give the same forward bits and
    input gradients within 1e-5 of the largest magnitude (one accumulator instead of three rounded products; measured 4.5e-6)."""
    from dlgrad_b200 import Tensor
    from dlgrad_b200.runtime.cuda import matmul as mm
    B, T, C, H = 2, 1024, 768, 12
    rng = np.random.default_rng(77)
    x = (rng.random((B, T, C), dtype=np.float32) - 0.5).astype(np.float32)
    w = [((rng.random((C, C), dtype=np.float32) - 0.5) * 0.1).astype(np.float32) for _ in range(3)]
    res = {}
    for grouped in (True, False):
        monkeypatch.setattr(mm, "GROUP_LINEAR", grouped)
        X = Tensor(x.copy(), device="cuda", requires_grad=True)
        W = [Tensor(a.copy(), device="cuda", requires_grad=True) for a in w]
        q, k, v = [X.linear(Wi, None).reshape((B, T, H, C // H)).transpose(1, 2) for Wi in W]
        ((q * k).sum() + (v * v).sum()).backward()
        res[grouped] = (q.numpy(), v.numpy(), X.grad.numpy(), W[1].grad.numpy())
    check(res[True][0], res[False][0], exact=True, what="q: grouped vs single launches")
    check(res[True][1], res[False][1], exact=True, what="v: grouped vs single launches")
    check(res[True][3], res[False][3], exact=True, what="dWk: grouped vs single launches")
    runner.rel_check(check, res[True][2], res[False][2], 1e-5, "dx: one K-concatenated launch vs three launches with add epilogues")


def test_bench_mlp_block_shapes_vs_float64():
    """(b) c_fc -> relu -> c_proj at the C5 sizes (8192 tokens, 768 -> 3072 -> 768): forward GEMMs 8192 x 3072 x 768 and
    8192 x 768 x 3072 (A operand = deferred relu), backward dH = g @ W2 with the relu-mask epilogue, dW2 = g^T @ relu(h)
    (B operand = deferred relu, 768 x 3072 x 8192 TN), dW1 = dH^T @ x (3072 x 768 x 8192 TN), dX = dH @ W1."""
    from dlgrad_b200 import Tensor
    rng = np.random.default_rng(3072)
    x = (rng.random((8, 1024, 768), dtype=np.float32) - 0.5).astype(np.float32)
    w1 = ((rng.random((3072, 768), dtype=np.float32) - 0.5) * 0.1).astype(np.float32)
    w2 = ((rng.random((768, 3072), dtype=np.float32) - 0.5) * 0.1).astype(np.float32)
    g = (rng.random((8, 1024, 768), dtype=np.float32) - 0.5).astype(np.float32)
    X = Tensor(x.copy(), device="cuda", requires_grad=True)
    W1 = Tensor(w1.copy(), device="cuda", requires_grad=True)
    W2 = Tensor(w2.copy(), device="cuda", requires_grad=True)
    H = X.linear(W1, None)
    y = H.relu().linear(W2, None)
    (y * Tensor(g.copy(), device="cuda")).sum().backward()
    x2, g2 = _f64(x).reshape(8192, 768), _f64(g).reshape(8192, 768)
    # relu is discontinuous: a pre-activation within rounding of zero may fall on either side in float64 and in 3xTF32,
    # and one flipped unit moves a whole row of dW1 / dX by O(1).  The float64 references below therefore take the
    # relu pattern of the DEVICE's own pre-activation (which is itself checked against float64 first).
    h_dev = H.numpy().reshape(8192, 3072)
    runner.gemm_check(check, h_dev, x2, _f64(w1).T, runner.TOL_MATMUL, "mlp pre-activation")
    on = h_dev > 0
    r = np.where(on, _f64(h_dev), 0.0)
    dh = (g2 @ _f64(w2)) * on
    runner.gemm_check(check, y.numpy().reshape(8192, 768), r, _f64(w2).T, runner.TOL_MATMUL, "mlp forward (relu operand A)")
    runner.gemm_check(check, W2.grad.numpy(), g2.T, r, runner.TOL_MATMUL, "mlp dW2 (relu operand B)")
    runner.gemm_check(check, W1.grad.numpy(), dh.T, x2, runner.TOL_MATMUL, "mlp dW1 (through the relu-mask epilogue)")
    runner.gemm_check(check, X.grad.numpy().reshape(8192, 768), dh, _f64(w1), runner.TOL_MATMUL, "mlp dX")


def test_full_width_gan_step_vs_oracle(oracle):
    """(c) examples/mnist_gan.py:30-58 at full width, batch 100: one discriminator step and one generator step with
    Adam(5e-4, betas (0.5, 0.999)) on Device.CUDA vs the oracle.  Includes the N = 1 output GEMM of the discriminator
    and its K = 1 backward.  Tolerances: losses 1e-5; first-step gradients 1e-4 of their largest magnitude; updated
    weights: Adam moves every element by ~lr whatever its gradient, so elements whose gradient is rounding noise may
    differ by 2 * lr (stated absolute bound)."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor, nn
    data = np.random.default_rng(11)
    real = (data.random((100, 784), dtype=np.float32) * 2 - 1).astype(np.float32)
    noise1 = (data.random((100, 128), dtype=np.float32) * 2 - 1).astype(np.float32)
    noise2 = (data.random((100, 128), dtype=np.float32) * 2 - 1).astype(np.float32)
    res = {}
    for dev in ("cuda", "cpu"):
        gan = models.GANPair(dl, dev, np.random.default_rng(1))
        opt_g = nn.optim.Adam(nn.utils.get_parameters(gan.g_layers), lr=0.0005, betas=(0.5, 0.999))
        opt_d = nn.optim.Adam(nn.utils.get_parameters(gan.d_layers), lr=0.0005, betas=(0.5, 0.999))
        T = lambda a: Tensor(np.array(a), device=dev)     # noqa: E731
        fake = gan.generator(T(noise1)).detach()
        opt_d.zero_grad()
        d_loss = (gan.discriminator(T(real)).bcewithlogitsloss(Tensor.full((100, 1), 1.0, device=dev))
                  + gan.discriminator(fake).bcewithlogitsloss(Tensor.full((100, 1), 0.0, device=dev))) / 2.0
        d_loss.backward()
        gd = [gan.d_layers[0].weight.grad.numpy(), gan.d_layers[3].weight.grad.numpy(), gan.d_layers[3].bias.grad.numpy()]
        opt_d.step()
        opt_g.zero_grad()
        g_loss = gan.discriminator(gan.generator(T(noise2))).bcewithlogitsloss(Tensor.full((100, 1), 1.0, device=dev))
        g_loss.backward()
        gg = [gan.g_layers[0].weight.grad.numpy(), gan.g_layers[3].weight.grad.numpy()]
        opt_g.step()
        res[dev] = dict(d_loss=np.array(float(d_loss.numpy())), g_loss=np.array(float(g_loss.numpy())), fake=fake.numpy(),
                        gd0=gd[0], gd3=gd[1], gdb3=gd[2], gg0=gg[0], gg3=gg[1],
                        d_w3=gan.d_layers[3].weight.numpy(), g_w0=gan.g_layers[0].weight.numpy())
    a, b = res["cuda"], res["cpu"]
    for k in ("d_loss", "g_loss"):
        runner.rel_check(check, a[k], b[k], 1e-5, f"GAN {k}")
    runner.rel_check(check, a["fake"], b["fake"], runner.TOL_MATMUL, "GAN generator output (tanh)")
    for k in ("gd0", "gd3", "gdb3", "gg0", "gg3"):
        runner.rel_check(check, a[k], b[k], runner.TOL_MATMUL, f"GAN first-step gradient {k}")
    for k in ("d_w3", "g_w0"):
        check(a[k], b[k], rtol=1e-5, atol=2 * 0.0005, what=f"GAN updated weight {k}")


@pytest.mark.timeout(900)
def test_c4_char_gpt_loss_curve_30_steps_vs_oracle(oracle):
    """(d) BASELINE configs[3] (n_layer 6, n_head 4, n_embd 128, block 128, batch 16, V 96), 30 Adam steps (lr 1e-4) on
    the same 4 rotating batches, Device.CUDA vs the oracle.  Stated tolerance on the loss curve: 2e-4 relative at every
    step — Adam's update is ~ +-lr per element whatever the size of its gradient, so elements whose gradient is
    rounding noise drift apart by up to lr per step; the LOSS, which depends on all of them, is the quantity the
    north star bounds ("within a stated tolerance on loss curves").  The first step, which depends on no update, is
    held to 1e-5."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor, nn
    steps = 2 if COMPILE_ONLY else 30          # walking the test to pre-build its kernels needs no curve
    curves = {}
    for dev in ("cuda", "cpu"):
        cfg = models.GPTConfig(vocab_size=96, block_size=128, n_layer=6, n_head=4, n_embd=128, device=dev)
        gpt = models.GPT(dl, cfg, np.random.default_rng(4))
        brng = np.random.default_rng(5)
        batches = [models.gpt_batch(cfg, 16, brng) for _ in range(4)]
        opt = nn.optim.Adam(nn.utils.get_parameters(gpt), lr=1e-4)
        losses = []
        for step in range(steps):
            xb, yb = batches[step % 4]
            opt.zero_grad()
            _, loss = gpt(Tensor(xb.copy(), device=dev), Tensor(yb.copy(), device=dev))
            losses.append(float(loss.numpy()))
            loss.backward()
            opt.step()
        curves[dev] = np.array(losses)
    if COMPILE_ONLY:
        return
    a, b = curves["cuda"], curves["cpu"]
    rel = np.abs(a - b) / np.abs(b)
    print(f"\nC4 loss curve: first {b[0]:.3f} last {b[-1]:.3f}; max relative deviation {rel.max():.3g} at step {int(rel.argmax())}")
    assert b[-1] < b[0], "the oracle's loss must go down over 30 steps"
    assert rel[0] <= 1e-5, f"step 0 loss differs by {rel[0]:.3g}"
    assert rel.max() <= 2e-4, f"loss curve deviates by {rel.max():.3g} (step {int(rel.argmax())}); stated tolerance 2e-4"


@pytest.mark.parametrize("fill", [-5.0, 0.0])
def test_attention_with_finite_mask_fill_is_dense(oracle, fill):
    """(e) masked_fill(mask, finite) leaves NON-zero probabilities in the masked tiles: the GEMMs that read P must not
    skip them (ADVICE r1 — only a -inf fill makes the dead tiles exactly zero).  y, dq, dk, dv vs the oracle at a size
    where every attention fusion engages (B*H*T*T = 4 Mi scores)."""
    from dlgrad_b200 import Tensor
    B, H, T, D = 1, 16, 512, 64
    rng = np.random.default_rng(99)
    q, k, v = ((rng.random((B, H, T, D), dtype=np.float32) - 0.5).astype(np.float32) for _ in range(3))
    mask = np.triu(np.ones((T, T), dtype=np.float32), 1)
    w = np.linspace(0.5, 1.5, B * H * T * D, dtype=np.float32).reshape(B, H, T, D)
    scale = float(1.0 / np.sqrt(D))

    def run(dev):
        Q, K, V = (Tensor(a.copy(), device=dev, requires_grad=True) for a in (q, k, v))
        att = ((Q @ K.transpose(2, 3)) * scale).masked_fill(Tensor(mask.copy(), device=dev), fill).softmax(dim=3)
        y = att @ V
        (y * Tensor(w.copy(), device=dev)).sum().backward()
        return [y.numpy(), Q.grad.numpy(), K.grad.numpy(), V.grad.numpy()]
    got, ref = run("cuda"), run("cpu")
    for a, b, what in zip(got, ref, ("y", "dq", "dk", "dv")):
        runner.rel_check(check, a, b, 1e-4, f"attention with fill {fill}: {what} vs oracle")


@pytest.mark.parametrize("shape,temperature", [((4, 7, 96), 0.7), ((3, 50257), 1.0), ((96,), 1.3), ((64, 1, 96), 0.7)])
def test_device_side_sampling_matches_inverse_cdf(shape, temperature):
    """§8f2: Tensor.sample_last draws the next token on the device (examples/gpt.py:216-221, 262-266 do it on the host
    after a D2H of the logits).  With the uniform numbers supplied, the chosen index must be the inverse-CDF index of
    softmax(logits[..., -1, :] / temperature) computed in float64 — except where u falls within 1e-5 of a CDF step,
    where float32 summation may legitimately pick the neighbour."""
    from dlgrad_b200 import Tensor
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) * 3).astype(np.float32)
    rows = 1 if len(shape) == 1 else shape[0]
    u = rng.random(rows, dtype=np.float32)
    u[0] = np.float32(1.0) - np.float32(2.0 ** -24)          # the largest value the generator can produce
    got = Tensor(x.copy(), device="cuda").sample_last(temperature, u=Tensor(u.copy(), device="cuda")).numpy().reshape(rows)
    if COMPILE_ONLY:
        return
    last = x.reshape(rows, -1, shape[-1])[:, -1, :].astype(np.float32)
    z = (last / np.float32(temperature)).astype(np.float64)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    cdf = np.cumsum(e, axis=1) / e.sum(axis=1, keepdims=True)
    for r in range(rows):
        want = min(int(np.searchsorted(cdf[r], float(u[r]), side="right")), shape[-1] - 1)
        g = int(got[r])
        assert 0 <= g < shape[-1] and float(got[r]) == g
        if g != want:
            edge = cdf[r][min(g, want)]
            assert abs(edge - float(u[r])) < 1e-5 and abs(g - want) == 1, f"row {r}: got {g}, inverse CDF gives {want} (u={u[r]}, step at {edge})"


def test_device_side_sampling_distribution_and_generate_loop():
    """Sampling from the backend's own Philox stream reproduces the distribution (chi-square against the float64
    probabilities over 32768 draws), and a KV-cache generate loop that never copies the distribution to the host — sampled
    index fed straight back into the embedding, keys / values appended with Tensor.cat — runs end to end."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor
    from dlgrad_b200.runtime.cuda import ops as cuda_ops
    from dlgrad_b200.runtime.cuda import transfer
    cuda_ops.manual_seed(1234)
    V, N = 96, 32768
    row = (np.random.default_rng(0).standard_normal(V) * 2).astype(np.float32)
    logits = Tensor(np.tile(row, (N, 1)), device="cuda")
    idx = logits.sample_last(0.7).numpy().reshape(N).astype(np.int64)
    if not COMPILE_ONLY:
        z = row.astype(np.float64) / 0.7
        p = np.exp(z - z.max())
        p /= p.sum()
        counts = np.bincount(idx, minlength=V).astype(np.float64)
        keep = p * N >= 5
        chi2 = float((((counts - p * N) ** 2) / (p * N))[keep].sum())
        assert chi2 < 2.0 * keep.sum() + 40, f"chi-square {chi2:.1f} over {int(keep.sum())} bins"
    # generate loop on a tiny GPT: device-side KV cache + device-side sampling; only the printed tokens cross PCIe
    cfg = models.GPTConfig(vocab_size=V, block_size=32, n_layer=2, n_head=2, n_embd=32, device="cuda")
    gpt = models.GPT(dl, cfg, np.random.default_rng(3))
    transfer.reset_counters()
    toks = models.generate_on_device(gpt, Tensor(np.array([[5.0]], dtype=np.float32), device="cuda"), 12, temperature=0.7)
    io = transfer.counters()
    out = toks.numpy().reshape(-1)
    if not COMPILE_ONLY:
        assert out.shape == (13,) and out[0] == 5 and ((out >= 0) & (out < V)).all() and (out == np.round(out)).all()
        assert io["d2h_bytes"] == 0, f"the generate loop copied {io['d2h_bytes']} bytes to the host"


def _attention_inputs(B, H, T, D, mask_kind, seed):
    rng = np.random.default_rng(seed)
    q, k, v = ((rng.random((B, H, T, D), dtype=np.float32) - 0.5).astype(np.float32) * 2 for _ in range(3))
    if mask_kind == "causal":
        mask = np.triu(np.ones((T, T), dtype=np.float32), 1)
    elif mask_kind == "random":
        mask = np.minimum((rng.random((T, T)) > 0.6).astype(np.float32), 1.0 - np.eye(T, dtype=np.float32))      # no dead rows
    elif mask_kind == "growing":         # scores that keep growing along the keys: the running maximum moves at every step
        mask = np.triu(np.ones((T, T), dtype=np.float32), 1)
        k = (k * (1.0 + np.arange(T, dtype=np.float32)[None, None, :, None] / 32.0)).astype(np.float32)
        q = np.abs(q).astype(np.float32)
        k = np.abs(k).astype(np.float32)
    else:                                # "dead_rows": some query rows have no unmasked key at all
        mask = np.maximum(rng.random((T, T)) > 0.5, (np.arange(T)[:, None] % 37 == 0)).astype(np.float32)
    w = np.linspace(0.5, 1.5, B * H * T * D, dtype=np.float32).reshape(B, H, T, D)
    return q, k, v, mask, w


def _attention_block(dev, q, k, v, mask, w, fill, scale, backward=True):
    from dlgrad_b200 import Tensor
    Q, K, V = (Tensor(a.copy(), device=dev, requires_grad=True) for a in (q, k, v))
    att = ((Q @ K.transpose(2, 3)) * scale).masked_fill(Tensor(mask.copy(), device=dev), fill).softmax(dim=3)
    y = att @ V
    out = [y.numpy()]
    if backward:
        (y * Tensor(w.copy(), device=dev)).sum().backward()
        out += [Q.grad.numpy(), K.grad.numpy(), V.grad.numpy()]
    return out, att


def _attention_f64(q, k, v, mask, w, fill, scale):
    """float64 forward and backward of y = softmax(masked_fill(q k^T scale, mask, fill)) v, loss = sum(y * w)."""
    q, k, v, w = (a.astype(np.float64) for a in (q, k, v, w))
    s = np.einsum("bhqd,bhkd->bhqk", q, k) * scale
    s = np.where(mask > 0, fill, s)
    with np.errstate(invalid="ignore"):
        e = np.exp(s - s.max(axis=3, keepdims=True))
        p = e / e.sum(axis=3, keepdims=True)
    y = p @ v
    dv = np.einsum("bhqk,bhqd->bhkd", p, w)
    dp = np.einsum("bhqd,bhkd->bhqk", w, v)
    ds = p * (dp - (dp * p).sum(axis=3, keepdims=True))
    ds = np.where(mask > 0, 0.0, ds) * scale
    return [y, ds @ k, np.einsum("bhqk,bhqd->bhkd", ds, q), dv]


@pytest.mark.parametrize("B,H,T,D,mask_kind,fill,tau", [
    (2, 16, 512, 64, "causal", float("-inf"), 8.0),
    (2, 16, 512, 64, "growing", float("-inf"), 0.0),        # forces the rescale of O in tensor memory at every step
    (2, 16, 512, 64, "growing", float("-inf"), 8.0),        # the lazy maximum under steadily growing scores
    (1, 16, 1024, 32, "random", float("-inf"), 8.0),        # head dim 32, a mask without structure
    (1, 16, 512, 64, "causal", -5.0, 8.0),                   # finite fill: P is dense, nothing may be skipped
    (8, 12, 1024, 64, "causal", float("-inf"), 8.0),        # the benchmark's attention
])
def test_flash_attention_forward_and_backward(oracle, B, H, T, D, mask_kind, fill, tau):
    """The flash-style kernels (codegen/cuda_flash.py: P never written forward or backward; dv and dS recomputed from
    q, k and the saved row statistics) against float64 — y, dq, dk, dv — and, at the sizes the oracle finishes in
    seconds, against the oracle, which runs the reference's unfused graph (examples/gpt.py:49-55).  Tolerances: y and
    dv 1e-4 of the largest magnitude (two chained matmuls), dq / dk 1e-4 as well (the round-1 kernels needed 5e-5
    against each other; float64 is the stricter judge)."""
    from dlgrad_b200.runtime.cuda import attention
    q, k, v, mask, w = _attention_inputs(B, H, T, D, mask_kind, T + D + B)
    scale = float(1.0 / np.sqrt(D))
    old = (attention.FLASH, attention.FLASH_TAU, attention.MIN_SCORES)
    attention.FLASH, attention.FLASH_TAU, attention.MIN_SCORES = True, tau, 1 << 20
    try:
        got, att = _attention_block("cuda", q, k, v, mask, w, fill, scale)
        if not COMPILE_ONLY:
            assert att.data._pending is not None and att.data._pending[0] == "attnP", "P must never have been materialised"
            kinds = {key[0] for key in attention._tc_plans}
            assert {"fwd", "dv", "ds"} <= kinds, f"expected the three flash kernels, got {kinds}"
        ref = _attention_f64(q, k, v, mask, w, fill, scale)
        for a, b, what in zip(got, ref, ("y", "dq", "dk", "dv")):
            runner.rel_check(check, a, b, 1e-4, f"flash attention {what} vs float64 ({mask_kind}, tau {tau})")
        if B * H * T * T <= (1 << 23):
            orc, _ = _attention_block("cpu", q, k, v, mask, w, fill, scale)
            for a, b, what in zip(got, orc, ("y", "dq", "dk", "dv")):
                runner.rel_check(check, a, b, 1e-4, f"flash attention {what} vs oracle ({mask_kind}, tau {tau})")
        # the probabilities themselves, read after the fact, are what the fused scores kernel / the eager chain give
        p = att.numpy()
        if not COMPILE_ONLY:
            s = np.einsum("bhqd,bhkd->bhqk", q.astype(np.float64), k.astype(np.float64)) * scale
            s = np.where(mask > 0, fill, s)
            e = np.exp(s - s.max(axis=3, keepdims=True))
            # ("growing" drives the logits to ~50: an fp32 logit then carries ~5e-6 of absolute error, which IS the relative
            # error of its probability — the per-element 1e-5 applies to logits of ordinary size)
            big = mask_kind == "growing"
            runner.rel_check(check, p, e / e.sum(axis=3, keepdims=True), 1e-4 if big else runner.TOL_ELEMENTWISE,
                             "P read after the flash forward", elementwise=not big)
        # same bits when run again (fixed summation order everywhere)
        again, _ = _attention_block("cuda", q, k, v, mask, w, fill, scale)
        for a, b, what in zip(got, again, ("y", "dq", "dk", "dv")):
            check(a, b, exact=True, what=f"flash attention {what}: second run")
    finally:
        attention.FLASH, attention.FLASH_TAU, attention.MIN_SCORES = old


def test_flash_attention_dead_rows_and_round1_kernels(oracle):
    """(1) Query rows without any unmasked key: y is NaN exactly where the unfused chain has NaN, finite rows agree.
    (2) With DLGRAD_FLASH_ATTN off the round-1 kernels (attn_scores_tc / attn_dscores_tc + tile-skipping GEMMs), which
    remain the materialisers of P, still give the same block."""
    from dlgrad_b200.runtime.cuda import attention
    B, H, T, D = 4, 8, 512, 32
    q, k, v, mask, w = _attention_inputs(B, H, T, D, "dead_rows", 5)
    scale = float(1.0 / np.sqrt(D))
    old = (attention.FLASH, attention.MIN_SCORES)
    attention.MIN_SCORES = 1 << 20
    try:
        attention.FLASH = True
        flash, _ = _attention_block("cuda", q, k, v, mask, w, float("-inf"), scale, backward=False)
        ref = _attention_f64(q, k, v, mask, w, float("-inf"), scale)[0]
        if not COMPILE_ONLY:
            assert np.isnan(ref).any() and (np.isnan(flash[0]) == np.isnan(ref)).all()
            ok = ~np.isnan(ref)
            runner.rel_check(check, flash[0][ok], ref[ok], 1e-4, "flash forward with dead rows vs float64")
        q, k, v, mask, w = _attention_inputs(2, 16, 512, 64, "causal", 6)
        attention.FLASH = False
        r1, att = _attention_block("cuda", q, k, v, mask, w, float("-inf"), 0.125)
        attention.FLASH = True
        fl, _ = _attention_block("cuda", q, k, v, mask, w, float("-inf"), 0.125)
        if not COMPILE_ONLY:
            assert att.data._pending is None, "with the flash path off P is materialised by the fused scores kernel"
        for a, b, what in zip(fl, r1, ("y", "dq", "dk", "dv")):
            runner.rel_check(check, a, b, 1e-4, f"{what}: flash vs round-1 fused kernels")
    finally:
        attention.FLASH, attention.MIN_SCORES = old


def test_attention_in_the_projections_layout_needs_no_permute_kernels(oracle):
    """The attention block exactly as examples/gpt.py:44-64 writes it — q / k / v = Linear(x).reshape(B, T, h, d)
    .transpose(1, 2), y.transpose(1, 2).reshape(B, T, C) — on CUDA: the head split / merge transposes are deferred (Buffer
    pending kind "hsplit"), the flash kernels and the two tile-skipping GEMMs read q, k, v, dO and write O, dq, dk, dv in
    the (B, T, h, d) layout through rank-4 tensor maps, so NO permute kernel runs forward or backward (the reference runs
    8 per layer).  Output and every gradient vs the oracle, which executes the reference's graph."""
    from dlgrad_b200 import Tensor
    from dlgrad_b200.runtime.cuda import attention, profile
    B, T, H, D = 1, 512, 16, 64
    C = H * D
    rng = np.random.default_rng(21)
    x = (rng.random((B, T, C), dtype=np.float32) - 0.5).astype(np.float32)
    ws = [((rng.random((C, C), dtype=np.float32) - 0.5) * (2.0 / np.sqrt(C))).astype(np.float32) for _ in range(3)]
    coef = np.linspace(0.5, 1.5, B * T * C, dtype=np.float32).reshape(B, T, C)
    old = attention.MIN_SCORES
    attention.MIN_SCORES = 1 << 20
    res = {}
    try:
        for dev in ("cuda", "cpu"):
            X = Tensor(x.copy(), device=dev, requires_grad=True)
            W = [Tensor(w.copy(), device=dev, requires_grad=True) for w in ws]
            if dev == "cuda" and not COMPILE_ONLY:
                profile.kernel_timer = profile.KernelTimer()
            q, k, v = (X.linear(w, None).reshape((B, T, H, D)).transpose(1, 2) for w in W)
            att = (q @ k.transpose(2, 3)) * (D ** -0.5)
            mask = Tensor.tril(Tensor.ones((T, T)), k=0.0)
            att = att.masked_fill(mask == Tensor(0.0), float("-inf")).softmax(dim=3)
            y = (att @ v).transpose(1, 2).reshape((B, T, C))
            (y * Tensor(coef.copy(), device=dev)).sum().backward()
            res[dev] = [y.numpy(), X.grad.numpy()] + [w.grad.numpy() for w in W]
            if dev == "cuda" and not COMPILE_ONLY:
                tags = profile.kernel_timer.collect()["by_tag"]
                profile.kernel_timer = None
                names = sorted({t[0] for t in tags})
                assert not any("permute" in n or "transpose" in n for n in names), f"permute kernels ran: {names}"
                assert att.data._pending is not None, "P must not have been materialised"
    finally:
        attention.MIN_SCORES = old
        profile.kernel_timer = None
    for a, b, what in zip(res["cuda"], res["cpu"], ("y", "dx", "dWq", "dWk", "dWv")):
        runner.rel_check(check, a, b, 1e-4 if what in ("y", "dWv") else 1e-3, f"attention in the (B, T, h, d) layout: {what}")


def test_adds_ride_on_the_gemm_that_produces_their_operand(oracle):
    """`x + proj(h)` (the residual connection, examples/gpt.py:88-89) and `grad + g @ W` (the autograd driver's gradient
    accumulation, dlgrad/tensor.py:805) run as ONE GEMM with the addition in its epilogue (matmul_tc_ts `radd`); the
    RMSNorm input gradient is added to the residual stream's gradient by the kernel that computes it.  A two-branch
    residual block against the oracle; a Linear output that was consumed by a fused add and is read afterwards — even
    after the weight has been updated in place — still has its forward value."""
    import dlgrad_b200 as dl
    from dlgrad_b200 import Tensor, nn
    from dlgrad_b200.runtime.cuda import profile
    rows, C = 2048, 512
    rng = np.random.default_rng(8)
    x = (rng.random((2, rows // 2, C), dtype=np.float32) - 0.5).astype(np.float32)
    w = [((rng.random((C, C), dtype=np.float32) - 0.5) * 0.1).astype(np.float32) for _ in range(3)]
    coef = np.linspace(0.5, 1.5, x.size, dtype=np.float32).reshape(x.shape)
    res = {}
    for dev in ("cuda", "cpu"):
        X = Tensor(x.copy(), device=dev, requires_grad=True)
        W = [Tensor(a.copy(), device=dev, requires_grad=True) for a in w]
        norm = nn.RMSNorm(C)
        if dev == "cuda" and not COMPILE_ONLY:
            profile.kernel_timer = profile.KernelTimer()
        h = norm(X)
        a, b = h.linear(W[0], None), h.linear(W[1], None)          # two consumers of h: their input gradients are accumulated
        lin = (a * b).linear(W[2], None)
        y = X + lin                                                 # residual
        (y * Tensor(coef.copy(), device=dev)).sum().backward()
        if dev == "cuda" and not COMPILE_ONLY:
            names = [t[0] for t in profile.kernel_timer.collect()["by_tag"]]
            profile.kernel_timer = None
            assert "add" not in names, f"a separate add kernel ran: {sorted(set(names))}"
            assert lin.data._pending is not None, "the consumed Linear output must still be deferred"
        opt = nn.optim.SGD([W[2]], lr=0.5)
        opt.step()                                                  # overwrites W[2] in place
        res[dev] = [y.numpy(), X.grad.numpy(), W[0].grad.numpy(), W[2].grad.numpy(), norm.weight.grad.numpy(), lin.numpy()]
    profile.kernel_timer = None
    for got, exp, what in zip(res["cuda"], res["cpu"], ("y", "dx", "dW0", "dW2", "dnorm", "linear output read after the update")):
        runner.rel_check(check, got, exp, 1e-4, f"fused adds: {what}")
