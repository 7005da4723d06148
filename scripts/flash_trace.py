"""Pipeline trace of a flash attention kernel (DLGRAD_FLASH_TRACE=fwd|dv|ds): CTA 0 stamps clock64 at every hand-off of its
first 48 steps; this prints, per step, when each role passed each point (cycles since the first stamp).
  0 S-issuer: stage split   1 S-issuer: S buffer free   2 S-issuer: MMAs issued
  3 O-issuer: P ready       4 O-issuer: MMAs issued
  5 row warp 2: S full      6 row: exponentials done    7 row: previous PV done     8 row: P written, arrive"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = os.environ.setdefault("DLGRAD_FLASH_TRACE", "fwd")
from dlgrad_b200 import Tensor
from dlgrad_b200.runtime.cuda import attention, shim, transfer

B, H, T, D = 8, 12, 1024, 64
rng = np.random.default_rng(0)
q, k, v = ((rng.random((B, H, T, D), dtype=np.float32) - 0.5).astype(np.float32) for _ in range(3))
mask = Tensor(np.triu(np.ones((T, T), dtype=np.float32), 1), device="cuda")
Q, K, V = (Tensor(a, device="cuda", requires_grad=True) for a in (q, k, v))
att = ((Q @ K.transpose(2, 3)) * (D ** -0.5)).masked_fill(mask, float("-inf")).softmax(dim=3)
y = att @ V
if mode != "fwd":
    (y * Tensor(np.ones((B, H, T, D), dtype=np.float32), device="cuda")).sum().backward()
shim.synchronize()
rec = attention._flash[id(att.data)]
raw = transfer.download(rec.stats, 2 * B * H * T)[B * H * T:]
tr = raw.view(np.int64)[:24 * 48].reshape(24, 48)
t0 = tr[tr > 0].min()
names = ["P:free", "P:issued", "X:full", "X:split", "S:split", "S:free", "S:issued", "O:Pready", "O:issued", "R:Sfull", "R:exp", "R:pvdone", "R:Pwritten"]
kinds = [16, 17, 18, 19, 0, 1, 2, 3, 4, 5, 6, 7, 8]
print("step " + " ".join(f"{n:>9s}" for n in names))
for g in range(40):
    print(f"{g:4d} " + " ".join(f"{(tr[kind][g] - t0) if tr[kind][g] > 0 else -1:9d}" for kind in kinds))
print("per item: row: split begin | A landed | A free | split done || S-issuer: at item | A split seen || row: item end done")
for n in range(6):
    print(f"{n:4d} " + " ".join(f"{(tr[kind][n] - t0) if tr[kind][n] > 0 else -1:10d}" for kind in (9, 10, 11, 12, 13, 14, 15)))
print("item end (row warp 2): l exchanged | last PV done | O read | O stored | done")
for n in range(5):
    print(f"{n:4d} " + " ".join(f"{(tr[kind][n] - t0) if tr[kind][n] > 0 else -1:10d}" for kind in (20, 21, 22, 23, 15)))
d = np.diff(tr[8][2:18])
print("row-warp step period (steps 2..17):", d.tolist())
