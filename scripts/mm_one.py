"""One matmul shape through one tcgen05 kernel variant, a few launches (ncu target).

    python scripts/mm_one.py M N K bn kernel [tx] [ty] [nb]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dlgrad_b200.runtime.cuda import matmul_tc, shim, transfer  # noqa: E402

M, N, K, bn = (int(x) for x in sys.argv[1:5])
kernel = sys.argv[5] if len(sys.argv) > 5 else "v4"
tx = len(sys.argv) > 6 and sys.argv[6] == "1"
ty = len(sys.argv) > 7 and sys.argv[7] == "1"
nb = int(sys.argv[8]) if len(sys.argv) > 8 else 1
shim.init()
rng = np.random.default_rng(0)
a = rng.random((nb, K, M) if tx else (nb, M, K), dtype=np.float32)
b = rng.random((nb, N, K) if ty else (nb, K, N), dtype=np.float32)
bs = lambda n: (0, 0) if nb == 1 else (0, n)  # noqa: E731
plan = matmul_tc.try_build(M, N, K, (1, nb), bs(M * K), bs(K * N), tx, ty, True, bn=bn or None, force=True, kernel=kernel)
da, db = transfer.upload_numpy(a), transfer.upload_numpy(b)
for _ in range(3):
    out = plan(da, db)
shim.synchronize()
print("ok", M, N, K, bn, kernel)
