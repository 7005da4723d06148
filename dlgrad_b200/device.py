"""Device enum.  Mirrors dlgrad/device.py:4-13 and adds CUDA, the only device this package computes on."""
from enum import Enum, auto


class Device(Enum):
    __hash__ = object.__hash__        # identity hash in C: these are dict keys on every dispatch (Enum's is a Python call)
    CPU = auto()     # host staging memory (numpy views, checkpoints); no kernels of its own here
    METAL = auto()   # kept so that dlgrad code naming it still imports; never dispatched to
    CUDA = auto()    # B200 / sm_100a

    @staticmethod
    def from_str(d: str) -> "Device":
        # The reference prints and returns None for unknown names (device.py:9-13), which lets a
        # typo in device= travel until the first dispatch; raise at the source instead.
        key = str(d).strip().lower().split(":")[0]
        alias = {"cpu": Device.CPU, "metal": Device.METAL, "cuda": Device.CUDA, "gpu": Device.CUDA}
        if key not in alias:
            raise ValueError(f"Invalid device: {d!r} (expected 'cuda' or 'cpu')")
        return alias[key]


def as_device(d) -> Device:
    if isinstance(d, Device):
        return d
    if d is None:
        return Device.CPU
    return Device.from_str(d)
