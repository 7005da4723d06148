"""Host-side logic of the grouped sibling-projection launches (dlgrad_b200/runtime/cuda/matmul.py: LinSlot, DwSlot,
run_lin_dx_sum) walked with the device stubbed out (DLGRAD_CUDA_COMPILE_ONLY=1: fake addresses, launches counted, every
kernel still generated and compiled for sm_100a).  What the GPU tests check numerically
(tests/test_parity_r2.py::test_sibling_projections_run_as_grouped_launches) is checked here structurally: which launches an
attention-shaped block issues, that nothing is computed twice, and that no reference cycle keeps activations alive."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import collections, gc
import numpy as np
from dlgrad_b200 import Tensor
from dlgrad_b200.buffer import Buffer
from dlgrad_b200.runtime.cuda import matmul as mm, attention as att

tags = collections.Counter()
orig_call = mm.MatmulPlan.__call__
def call(self, *a):
    tags[self.tag[0]] += 1
    return orig_call(self, *a)
mm.MatmulPlan.__call__ = call

B, T, C, H = 2, 1024, 768, 12
rng = np.random.default_rng(0)
x = (rng.random((B, T, C), dtype=np.float32) - 0.5).astype(np.float32)
w = [((rng.random((C, C), dtype=np.float32) - 0.5) * 0.1).astype(np.float32) for _ in range(4)]

def block(grouped):
    mm.GROUP_LINEAR = grouped
    tags.clear()
    X = Tensor(x.copy(), device="cuda", requires_grad=True)
    W = [Tensor(a.copy(), device="cuda", requires_grad=True) for a in w]
    q, k, v = [X.linear(Wi, None).reshape((B, T, H, C // H)).transpose(1, 2) for Wi in W[:3]]
    # all three projections are still deferred: nothing has been launched for them
    assert all(t.data._pending is not None for t in (q, k, v))
    assert sum(tags.values()) == 0, tags
    y = (q * k + v).transpose(1, 2).reshape((B, T, C)).linear(W[3], None)       # a fourth Linear over ANOTHER input
    (y * y).sum().backward()
    _ = X.grad.data.ptr                                                          # the deferred sum of input gradients runs here
    for Wi in W:
        assert Wi.grad is not None and Wi.grad.data._pending is None or Wi.grad.data.ptr is not None
    return dict(tags)

g = block(True)
assert g.get("tc3_group_n3") == 1, g                  # q, k, v forward: one launch
assert g.get("tc3_group_k3") == 1, g                  # dq Wq + dk Wk + dv Wv: one launch
assert g.get("tc3_splitk_group3") == 1, g             # dWq, dWk, dWv: one launch
assert g.get("tc3_splitk") == 1, g                    # the fourth layer's weight gradient stays a single split-K launch
assert g.get("tc3", 0) == 2 and "tc3_radd" not in g, g   # fourth layer forward + its input gradient
s = block(False)
assert "tc3_group_n3" not in s and "tc3_group_k3" not in s and "tc3_splitk_group3" not in s, s
assert s.get("tc3_splitk") == 4 and s.get("tc3", 0) + s.get("tc3_radd", 0) == 8, s
assert sum(s.values()) - sum(g.values()) == 6, (s, g)   # 3 -> 1 three times

# a slot resolves once: reading the reshaped view and the Linear output itself gives the same storage, one launch
mm.GROUP_LINEAR = True
tags.clear()
X = Tensor(x.copy(), device="cuda")
a = X.linear(Tensor(w[0].copy(), device="cuda"), None)
b = a.reshape((B, T, H, C // H))
assert a.data._pending is not None and b.data._pending is not None
pb, pa = b.data.ptr, a.data.ptr
assert pa is pb and sum(tags.values()) == 1, tags

# no cyclic garbage: with the collector off, a block leaves no Buffer behind
gc.collect(); gc.disable()
live = lambda: sum(1 for o in gc.get_objects() if isinstance(o, Buffer))
base = live()
block(True)
mm._last_group = []; mm.flush_dw()
assert live() - base <= 0, live() - base
print("GROUPING-OK")
"""


@pytest.mark.timeout(600)
def test_sibling_projection_grouping_on_the_host():
    env = dict(os.environ, DLGRAD_CUDA_COMPILE_ONLY="1", PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-c", CODE], cwd=ROOT, env=env, capture_output=True, text=True, timeout=580)
    assert res.returncode == 0 and "GROUPING-OK" in res.stdout, res.stdout[-1500:] + res.stderr[-3000:]
