"""Does the C5 GPT actually learn on the device?  150 Adam steps over four fixed synthetic batches (the bench's
workload); the sum-reduced loss (reference quirk: cross entropy is a SUM over the 8192 tokens) must fall, and the
caching allocator must be in steady state (no growth after the first steps).

    python scripts/train_curve.py [steps] [lr]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from dlgrad_b200 import Tensor  # noqa: E402
from dlgrad_b200.runtime.cuda import shim  # noqa: E402
from tests import models  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
shim.init()
cfg, gpt, params, opt = bench.build_gpt(bench.C5, "cuda")
if len(sys.argv) > 2:
    opt.lr = float(sys.argv[2])
rng = np.random.default_rng(1000)
batches = [models.gpt_batch(cfg, 8, rng) for _ in range(4)]
dev = [(Tensor(x, device="cuda"), Tensor(y, device="cuda")) for x, y in batches]
ntok = 8 * cfg.block_size
print(f"uniform-guess loss = {ntok * np.log(cfg.vocab_size):.1f} (sum over {ntok} tokens)")
reserved = []
for i in range(steps):
    loss = bench.train_step(gpt, opt, params, *dev[i % 4], 1)
    if i % 10 == 0 or i == steps - 1:
        v = float(loss.numpy())
        reserved.append(int(shim.lib.dlg_mem_reserved()))
        print(f"step {i:4d}  loss {v:12.2f}  per-token {v / ntok:7.4f}  reserved {reserved[-1] / 1e9:6.2f} GB  live {int(shim.lib.dlg_mem_live()) / 1e9:6.2f} GB", flush=True)
if not shim.COMPILE_ONLY:
    assert reserved[-1] == reserved[2], "allocator grew after warm-up"
    print("allocator steady:", reserved[-1] == reserved[2])
