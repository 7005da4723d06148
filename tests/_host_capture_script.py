"""Host-side invariants of constant folding, captured steps and the device RNG, run with the stand-in library
(DLGRAD_CUDA_COMPILE_ONLY=1: fake addresses, launches ignored) by tests/test_host_logic.py in a subprocess.
Prints one line per check; exits non-zero on the first failure."""
import os
import sys

os.environ["DLGRAD_CUDA_COMPILE_ONLY"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import dlgrad_b200 as dl  # noqa: E402
from dlgrad_b200 import Tensor, buffer, graph, nn  # noqa: E402
from dlgrad_b200.runtime.cuda import ops as cuda_ops  # noqa: E402
from dlgrad_b200.runtime.cuda import shim  # noqa: E402


def ok(cond, what):
    print(("ok   " if cond else "FAIL ") + what, flush=True)
    if not cond:
        sys.exit(1)


# ---- constant folding -------------------------------------------------------------------------------------------
launches = lambda: shim.lib.launches      # noqa: E731
ones = Tensor.ones((256, 256), device="cuda")
ok(ones.data._pending is not None and ones.data._pending[0] == "full", "FULL is lazy")
n0 = launches()
m1 = Tensor.tril(Tensor.ones((256, 256), device="cuda"), k=0.0)
first = launches() - n0
n0 = launches()
m2 = Tensor.tril(Tensor.ones((256, 256), device="cuda"), k=0.0)
z2 = m2 == Tensor(0.0)
z1 = m1 == Tensor(0.0)
ok(first >= 3 and launches() - n0 == 1, f"second tril(ones) launches nothing, the first `== 0` one kernel ({first} then {launches() - n0})")
ok(shim.address(m1.data.ptr) == shim.address(m2.data.ptr) and m1.data is not m2.data, "same tag -> same storage, separate Buffers")
ok(m1.data._cow and m2.data._cow and z1.data._const == z2.data._const, "shared storage is copy-on-write, derived tags agree")
before = shim.address(m2.data.ptr)
buffer.flush_pending((m2.data,))
ok(shim.address(m2.data.ptr) != before and m2.data._const is None and not m2.data._cow, "an in-place writer gets a private copy and loses the tag")
ok(shim.address(m1.data.ptr) == before, "the other holder keeps the shared storage")
a, b = Tensor.ones((8, 8), device="cuda"), Tensor.ones((8, 8), device="cuda")
ok(shim.address(a.data.ptr) != shim.address(b.data.ptr), "FULL results themselves are never shared")
ok(Tensor.tril(Tensor.ones((256, 256), device="cuda"), k=1.0).data._const != m1.data._const, "another diagonal is another tag")

c = Tensor.tril(Tensor.ones((256, 256), device="cuda"), k=0.0)
shared = shim.address(c.data.ptr)
c.copy_from(np.zeros((256, 256), dtype=np.float32).tobytes())
ok(c.data._const is None and shim.address(c.data.ptr) != shared, "copy_from writes a private, untagged copy (the cached constant is untouched)")
o = Tensor.ones((4, 4), device="cuda")
o.copy_from(np.full((4, 4), 7.0, dtype=np.float32).tobytes())
ok(o.data._const is None and (o * 2.0).data._const is None, "a loaded tensor is no longer a constant: nothing derived from it is folded")
u = Tensor.ones((3, 1), device="cuda")
u.data.unsqueeze(0)
ok(u.data._const[0] == "reshape" and u.data._const[2] == (1, 3, 1), "an in-place relabel carries the tag to the new shape")

# ---- captured steps: guards, hooks, state restore ------------------------------------------------------------------
w = Tensor(np.ones((16, 16), dtype=np.float32), device="cuda", requires_grad=True)
opt = nn.optim.Adam([w], lr=1e-2)
x0 = Tensor(np.ones((16, 16), dtype=np.float32), device="cuda")


def step(x):
    opt.zero_grad()
    loss = (w * x).sum()
    loss.backward()
    opt.step()
    return loss


xin = Tensor.ones((16, 16), device="cuda")
capin = graph.CapturedStep(lambda t: t * 2.0, xin, warmup=1)
ok(xin.data._const is None and not xin.data._cow, "a constant passed as a graph input becomes an ordinary buffer")
capin.close()
cap = graph.CapturedStep(step, x0, warmup=2, params=[w])
ok(opt.t == 2 and len(cap.hooks) == 1, "warm-up steps are real steps, the captured call is not; Adam registered its hook")
g_captured = w.grad
for i in range(3):
    w.grad = None
    cap(np.full((16, 16), i, dtype=np.float32))
ok(opt.t == 5 and w.grad is g_captured, "every replay advances t and re-binds .grad to the captured gradient")
ok(not shim.capture["on"] and shim.capture["hooks"] is None and shim.capture["rng"] is None, "capture state is cleared")
cap.close()


def reads_back(x):
    return Tensor(float((w * x).sum().numpy()))


try:
    graph.CapturedStep(lambda x: reads_back(x) if shim.capture["on"] else x, x0, warmup=1)
    ok(False, "numpy() inside a capture must raise")
except RuntimeError as e:
    ok("captured" in str(e) and not shim.capture["on"], "numpy() inside a capture raises and leaves eager mode usable")
try:
    graph.CapturedStep(lambda x: x, Tensor(np.ones((2, 2), dtype=np.float32)))
    ok(False, "host inputs must be rejected")
except ValueError:
    ok(True, "inputs must live on Device.CUDA")

# ---- device RNG: every draw has its own call number; a replay reserves its range ----------------------------------------
dl.manual_seed(5)
Tensor.uniform((4, 4), device="cuda")
Tensor.uniform((4, 4), device="cuda")
ok(cuda_ops._rng_state["calls"] == 2 and cuda_ops._rng_state["seed"] == 5, "two eager draws, two call numbers")
xt = Tensor(np.ones((64, 64), dtype=np.float32), device="cuda")
cap = graph.CapturedStep(lambda t: t.dropout(0.5).dropout(0.5), xt, warmup=1)
base = cuda_ops._rng_state["calls"]
ok(base == 4 and len(cap.hooks) == 1, "the warm-up drew eagerly, the capture did not advance the host counter; one RNG hook")
cap()
cap()
ok(cuda_ops._rng_state["calls"] == base + 4, "each replay reserves as many call numbers as the step draws")
Tensor.uniform((4, 4), device="cuda")
ok(cuda_ops._rng_state["calls"] == base + 5, "eager draws continue after the reserved ranges")
cap.close()
dl.manual_seed(5)
ok(cuda_ops._rng_state["calls"] == 0, "manual_seed restarts the sequence")
print("all host checks passed")
