"""Bring-up aid for the flash attention kernels: where (which rows / columns / tiles) does the forward differ from float64?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200 import Tensor
from dlgrad_b200.runtime.cuda import attention

B, H, T, D = 1, 2, 512, int(os.environ.get("D", "64"))
attention.MIN_SCORES = 1 << 10
rng = np.random.default_rng(0)
q, k, v = ((rng.random((B, H, T, D), dtype=np.float32) - 0.5).astype(np.float32) * 2 for _ in range(3))
mask = np.triu(np.ones((T, T), dtype=np.float32), 1)
scale = float(1 / np.sqrt(D))
Q, K, V = (Tensor(a.copy(), device="cuda", requires_grad=True) for a in (q, k, v))
att = ((Q @ K.transpose(2, 3)) * scale).masked_fill(Tensor(mask, device="cuda"), float("-inf")).softmax(dim=3)
y = att @ V
got = y.numpy()
s = np.einsum("bhqd,bhkd->bhqk", q.astype(np.float64), k.astype(np.float64)) * scale
s = np.where(mask > 0, -np.inf, s)
e = np.exp(s - s.max(axis=3, keepdims=True)); p = e / e.sum(axis=3, keepdims=True)
ref = p @ v.astype(np.float64)
err = np.abs(got - ref)
print("max err", err.max(), "nan", np.isnan(got).sum())
for h in range(H):
    for qt in range(T // 128):
        blk = err[0, h, qt * 128:(qt + 1) * 128]
        print(f"head {h} qtile {qt}: cols0-31 {blk[:, :32].max():.3e} cols32-63 {blk[:, 32:].max():.3e}  rows by quarter",
              [f"{blk[i*32:(i+1)*32].max():.2e}" for i in range(4)])
# is it a P problem? compare against P truncated to the first / second 32 keys of each 64 block
for name, sel in (("only keys%64<32", (np.arange(T) % 64) < 32), ("only keys%64>=32", (np.arange(T) % 64) >= 32)):
    alt = (p * sel[None, None, None, :]) @ v.astype(np.float64)
    print(name, "max |got - alt|", np.abs(got - alt).max())
(y * Tensor(np.ones_like(q), device="cuda")).sum().backward()
dv_ref = np.einsum("bhqk,bhqd->bhkd", p, np.ones_like(q, dtype=np.float64))
print("dv err", np.abs(V.grad.numpy() - dv_ref).max())
dp = np.einsum("bhqd,bhkd->bhqk", np.ones_like(q, dtype=np.float64), v.astype(np.float64))
ds = p * (dp - (dp * p).sum(axis=3, keepdims=True)); ds = np.where(mask > 0, 0, ds) * scale
print("dq err", np.abs(Q.grad.numpy() - ds @ k.astype(np.float64)).max(), "dk err", np.abs(K.grad.numpy() - np.einsum("bhqk,bhqd->bhkd", ds, q.astype(np.float64))).max())
# the saved row statistic against float64: lse2 = log2(sum_k 2^(s_k c)) with masked entries left out
rec = attention._flash.get(id(att.data))
if rec is not None:
    from dlgrad_b200.runtime.cuda import transfer
    lse = transfer.download(rec.stats, B * H * T).reshape(B, H, T)
    s2 = s * np.log2(np.e)
    mx = s2.max(axis=3)
    ref_lse = mx + np.log2(np.exp2(s2 - mx[..., None]).sum(axis=3))
    d = np.abs(lse - ref_lse)
    print("lse max err", d.max(), "per q tile", [f"{d[0, 0, i*128:(i+1)*128].max():.2e}" for i in range(T // 128)])
    print("lse row0..3", lse[0, 0, :4], ref_lse[0, 0, :4])
    # y reconstructed from the kernel's own lse: tells P/V/O handling apart from the statistics
    p_k = np.exp2(s2 - lse[..., None].astype(np.float64))
    print("max |got - P(lse_kernel) v|", np.abs(got - p_k @ v.astype(np.float64)).max())
    print("row 0 of y:", got[0, 0, 0, :6], "v[0]:", v[0, 0, 0, :6])
    print("row 1 of y:", got[0, 0, 1, :6], "expected", ref[0, 0, 1, :6])
    print("y cols 32..37 row 0:", got[0, 0, 0, 32:38], v[0, 0, 0, 32:38])
