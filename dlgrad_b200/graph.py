"""Replay a training step as ONE CUDA graph launch.

No reference counterpart: dlgrad issues every kernel from Python, and so does this package in eager mode — about
10-15 us of interpreter time per kernel, against 2-5 us of device time for the kernels of the reference's own
workloads (MLP, GAN, char-GPT: SURVEY §8d "launch/host-bound").  A steady-state step launches the same kernels on
the same shapes every time, so it can be recorded once and replayed:

    step = graph.CapturedStep(lambda x, y: train_step(model, opt, x, y), x0, y0, params=opt.params)
    for x, y in batches:
        loss = step(x, y)             # copies x, y into the captured input buffers, launches the graph

How it stays correct
  * memory: every allocation made while capturing comes from a private pool of the caching allocator (dlg_pool_*):
    blocks freed by Python — during the capture or later — return to the pool only, so no eager allocation can alias
    memory the graph writes on replay.  The Tensors created by the captured call (loss, .grad of the parameters)
    keep their addresses; a replay refreshes their contents in place.
  * inputs are the Tensors passed at capture; a replay first copies the new host arrays into them.
  * host data created inside the step (examples/gpt.py builds its position indices from numpy every call) is frozen:
    staged once in pinned memory owned by the graph and re-copied by every replay.  Anything that must differ between
    steps has to be a graph input.
  * host round trips (numpy(), synchronize) raise inside a capture instead of silently recording something else.
  * random numbers (dropout) come from a counter-based device generator; inside a capture the counter base lives in
    device memory and a hook advances it before every replay, so every replay draws fresh masks.
  * optimizer steps may be part of the captured call.  SGD's launch arguments never change (with momentum the velocity
    is updated in place by the same kernel).  Adam's bias corrections
    1 - beta^t do: inside a capture Adam records its multi-tensor kernel in the form that reads them from device
    memory, and a hook advances `t` and rewrites those scalars (one tiny eager launch) before every replay.
  * a capture records, it does not execute: the warm-up calls before it are real steps, the captured call is not.
  * deferred buffers (Buffer._pending) still alive when the captured call returns are materialised inside the graph.
  * Python-side effects are not replayed.  The one that matters is which Tensor each parameter's `.grad` names: pass
    `params=` and every replay re-binds `.grad` to the gradient Tensors of the captured call (two captured phases that
    both touch the same parameters — a GAN's discriminator and generator steps — would otherwise read each other's).
"""
from __future__ import annotations

import numpy as np

from dlgrad_b200.buffer import flush_pending
from dlgrad_b200.device import Device
from dlgrad_b200.runtime.cuda import shim
from dlgrad_b200.tensor import Tensor


class CapturedStep:
    def __init__(self, fn, *inputs: Tensor, warmup: int = 2, params: list | None = None) -> None:
        for t in inputs:
            if not isinstance(t, Tensor) or t.device is not Device.CUDA:
                raise ValueError("CapturedStep inputs must be Tensors on Device.CUDA (their storage becomes the graph's input)")
        self.fn, self.inputs = fn, inputs
        flush_pending(tuple(t.data for t in inputs))     # replays overwrite the inputs in place: private, untagged storage
        lib = shim.lib
        for _ in range(max(1, warmup)):            # compiles every kernel, fills the plan / constant caches
            out = fn(*inputs)
        del out
        flush_pending()
        shim.synchronize()
        self.pool = lib.dlg_pool_begin()
        if self.pool <= 0:
            raise RuntimeError("dlg_pool_begin failed: " + shim.last_error())
        self.keep: list = []
        self.hooks: list = []                      # run before every replay (Adam: advance t, refresh its device scalars)
        self.refs: list = []                       # device storage the graph reads by address but does not own (constants)
        shim.capture["on"], shim.capture["keep"], shim.capture["hooks"] = True, self.keep, self.hooks
        shim.capture["refs"] = self.refs
        shim.capture["rng"] = {"n": 0, "hooked": False}   # random draws recorded in this step (runtime/cuda/ops.uniform)
        self.exec = shim.NULL
        try:
            shim.check(lib.dlg_graph_begin(), "dlg_graph_begin")
            try:
                self.out = fn(*inputs)
                for t in self._tensors(self.out):
                    _ = t.data.ptr
                flush_pending()
            finally:
                self.exec = lib.dlg_graph_end()
        finally:
            shim.capture["on"], shim.capture["keep"], shim.capture["hooks"], shim.capture["rng"] = False, None, None, None
            shim.capture["refs"] = None
            lib.dlg_pool_end()
        if self.exec == shim.NULL:
            raise RuntimeError("capturing the step failed: " + shim.last_error())
        self.grads = [(p, p.grad) for p in (params or [])]
        self.replays = 0

    @staticmethod
    def _tensors(out) -> list:
        if isinstance(out, Tensor):
            return [out]
        if isinstance(out, (tuple, list)):
            return [t for t in out if isinstance(t, Tensor)]
        return []

    def __call__(self, *arrays):
        """Replay with new input values: each argument is a float32 numpy array of the captured input's shape, or None
        to keep what the input buffer holds.  Returns the Tensors of the captured call (same objects every time)."""
        if len(arrays) not in (0, len(self.inputs)):
            raise ValueError(f"expected {len(self.inputs)} inputs, got {len(arrays)}")
        for t, a in zip(self.inputs, arrays):
            if a is None:
                continue
            a = np.ascontiguousarray(a, dtype=np.float32)
            if a.size != t.numel:
                raise ValueError(f"input of {a.size} elements for a captured buffer of shape {t.shape}")
            shim.check(shim.lib.dlg_h2d(t.data.ptr, shim.ffi.from_buffer("float *", a), a.nbytes), "dlg_h2d")
        for hook in self.hooks:
            hook()
        shim.check(shim.lib.dlg_graph_launch(self.exec), "dlg_graph_launch")
        for p, g in self.grads:
            p.grad = g
        self.replays += 1
        return self.out

    def __enter__(self) -> "CapturedStep":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:          # interpreter shutdown: the library may already be gone
            pass

    def close(self) -> None:
        """Destroy the graph and release its memory pool; the Tensors of the captured call stay valid until dropped."""
        if getattr(self, "exec", shim.NULL) != shim.NULL:
            shim.synchronize()
            shim.lib.dlg_graph_destroy(self.exec)
            self.exec = shim.NULL
            self.out, self.grads = None, []
            shim.lib.dlg_pool_destroy(self.pool)
            for pin in self.keep:
                shim.lib.dlg_host_free(pin)
            self.keep = []
            self.refs = []
