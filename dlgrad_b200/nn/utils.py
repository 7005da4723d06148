"""Parameter discovery and safetensors checkpoints (dlgrad/nn/utils.py:10-132).  The file format is
the reference's — 8-byte little-endian header length, JSON header {name: {dtype:"F32", shape,
data_offsets}}, raw float32 blobs — so checkpoints interchange between dlgrad's CPU models and CUDA
models of this package.  Tensor.tobytes / Tensor.copy_from do the D2H / H2D."""
from __future__ import annotations

import json
import os
import struct

from dlgrad_b200.tensor import Tensor


def _walk(obj, path: str, visit) -> None:
    if isinstance(obj, Tensor):
        visit(path, obj)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            _walk(v, f"{path}.{k}" if path else str(k), visit)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            _walk(v, f"{path}.{i}" if path else str(i), visit)
    elif hasattr(obj, "__dict__"):
        for k, v in vars(obj).items():
            if not k.startswith("_"):
                _walk(v, f"{path}.{k}" if path else str(k), visit)


def get_state_dict(obj, prefix: str = "") -> dict[str, Tensor]:
    """{'path.to.tensor': Tensor} for every Tensor reachable through attributes, lists and dicts."""
    out: dict[str, Tensor] = {}
    _walk(obj, prefix, lambda name, t: out.__setitem__(name, t))
    return out


def get_parameters(obj) -> list[Tensor]:
    """Every reachable Tensor, in attribute order (the reference returns non-trainable ones too)."""
    found: list[Tensor] = []

    def rec(o) -> None:
        if isinstance(o, Tensor):
            found.append(o)
        elif isinstance(o, dict):
            for v in o.values():
                rec(v)
        elif isinstance(o, (list, tuple)):
            for v in o:
                rec(v)
        elif hasattr(o, "__dict__"):
            rec(vars(o))
    rec(obj)
    return found


def save_model(model, path: str) -> None:
    state = get_state_dict(model)
    header, blobs, offset = {}, [], 0
    for name, t in state.items():
        raw = t.tobytes()
        header[name] = {"dtype": "F32", "shape": list(t.shape), "data_offsets": [offset, offset + len(raw)]}
        blobs.append(raw)
        offset += len(raw)
    head = json.dumps(header).encode("utf-8")
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(head)))
        f.write(head)
        for raw in blobs:
            f.write(raw)
    print(f"Saved {len(state)} tensors to {path}")


def load_model(model, path: str) -> None:
    if not os.path.exists(path):
        raise FileNotFoundError(f"Checkpoint {path} not found")
    state = get_state_dict(model)
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
        base = 8 + hlen
        loaded = 0
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            t = state.get(name)
            if t is None:
                print(f"Warning: {name} found in file but not in model. Skipping.")
                continue
            if tuple(t.shape) != tuple(meta["shape"]):
                print(f"Shape mismatch for {name}: Model {t.shape} vs File {meta['shape']}")
                continue
            lo, hi = meta["data_offsets"]
            f.seek(base + lo)
            t.copy_from(f.read(hi - lo))
            loaded += 1
    print(f"Successfully loaded {loaded} tensors.")
