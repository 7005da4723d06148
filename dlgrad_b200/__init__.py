"""dlgrad_b200 — a Blackwell-native Device.CUDA backend behind dlgrad's Tensor / ops / nn / optim API.

    from dlgrad_b200 import Tensor, nn          # as `from dlgrad import Tensor, nn`
    x = Tensor(np_array, device="cuda")

Compute happens on Device.CUDA only; there is no CPU fallback (see dispatch.py).
"""
from dlgrad_b200.device import Device  # noqa: F401
from dlgrad_b200.tensor import Tensor  # noqa: F401
from dlgrad_b200 import nn  # noqa: F401, E402
from dlgrad_b200.runtime.cuda.ops import manual_seed  # noqa: F401, E402


def synchronize() -> None:
    from dlgrad_b200.runtime.cuda import shim
    shim.synchronize()
