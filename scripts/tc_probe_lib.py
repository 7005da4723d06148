import numpy as np
from dlgrad_b200.runtime.cuda import matmul_tc, profile, shim, transfer


def time_shape(M, N, K, nb=1, tx=False, ty=False, bn=None, iters=10):
    shim.init()
    rng = np.random.default_rng(0)
    a = rng.random((nb, K, M) if tx else (nb, M, K), dtype=np.float32)
    b = rng.random((nb, N, K) if ty else (nb, K, N), dtype=np.float32)
    a_bs = (0, 0) if nb == 1 else (0, M * K)
    b_bs = (0, 0) if nb == 1 else (0, K * N)
    plan = matmul_tc.try_build(M, N, K, (1, nb), a_bs, b_bs, tx, ty, True, bn=bn, force=True)
    da, db = transfer.upload_numpy(a), transfer.upload_numpy(b)
    for _ in range(3):
        plan(da, db)
    e0 = profile.Event().record()
    for _ in range(iters):
        plan(da, db)
    e1 = profile.Event().record()
    e1.sync()
    ms = e1.ms_since(e0) / iters
    return ms, 2.0 * M * N * K * nb / (ms * 1e-3) / 1e12
