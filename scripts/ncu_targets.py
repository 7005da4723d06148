"""A handful of launches of the dominant kernels, for `ncu --set full` (run once plain, then under ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlgrad_b200 import Tensor, nn
from dlgrad_b200.runtime.cuda import shim

rng = np.random.default_rng(0)
T = lambda *s: Tensor(rng.random(s, dtype=np.float32), device="cuda")   # noqa: E731
x = T(8, 1024, 768)
lin = nn.Linear(768, 768, bias=False, device="cuda")
a, b = T(4096, 4096), T(4096, 4096)
att = T(8, 12, 1024, 1024)
mask = Tensor(np.triu(np.ones((1024, 1024), dtype=np.float32), 1), device="cuda")
for _ in range(2):
    y = lin(x)                                            # persistent tcgen05 matmul, A in TMEM, (8192, 768, 768) NT
    c = a @ b                                             # same kernel family, 4096^3
    s = a + b                                             # ewise
    r0 = a.sum(dim=0)                                     # strip column reduce (two passes)
    r1 = a.sum(dim=1)                                     # row reduce
    p = (att * 0.125).masked_fill(mask, float("-inf")).softmax(dim=3)   # fused scale+mask+softmax
    t = a.permute((1, 0))                                 # tiled transpose (a.T alone is deferred into its consumer)
    _ = float(r1.sum().numpy())
shim.synchronize()
print("ok")
