"""Optimizers (dlgrad/nn/optim.py:4-70).  Same constructor arguments and update rules; on CUDA each
parameter is updated by ONE fused in-place kernel (codegen/cuda_kernel.adam_step / sgd_step) instead
of the reference's chain of ~14 (Adam) or 2 (SGD) elementwise dispatches per parameter, each of
which allocates a fresh tensor.  The per-element float32 operation sequence is the reference's.
"""
from __future__ import annotations

import numpy as np

from dlgrad_b200.buffer import Buffer, flush_pending
from dlgrad_b200.device import Device
from dlgrad_b200.dispatch import dispatcher
from dlgrad_b200.helpers import CustomOps
from dlgrad_b200.tensor import Tensor

_f32 = np.float32


class Optimizer:
    def __init__(self, params: list[Tensor], lr: float = 1e-1) -> None:
        self.params = params
        self.lr = lr
        self.grad_scale = 1.0     # set to 1/world by data-parallel training (dlgrad_b200/distributed.py)

    def _device_for(self, *bufs: Buffer) -> Device:
        return Buffer._route(*bufs)

    def _before_inplace(self, extra: list | None = None) -> None:
        """The updates below write parameters (and optimizer state) in place: deferred pointwise producers
        that still read one of those buffers must run first (dlgrad_b200/buffer.py)."""
        flush_pending(tuple(p.data for p in self.params) + tuple(extra or ()))

    def step(self) -> None:
        self._before_inplace()
        for p in self.params:
            if p.grad is None:
                continue
            assert p.shape == p.grad.shape, (
                f"The shape of the parameter {p.shape} and its grad {p.grad.shape} do not match")
            dev = self._device_for(p.data, p.grad.data)
            dispatcher.dispatch(op=CustomOps.SGD_STEP, device=dev, param=p.data, grad=p.grad.data,
                                lr=float(_f32(self.lr)), grad_scale=self.grad_scale)

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None


class SGD(Optimizer):
    def __init__(self, params: list[Tensor], lr: float = 1e-3, momentum: float = 0.0) -> None:
        super().__init__(params, lr)
        self.momentum = momentum
        if momentum != 0:
            # the reference's momentum branch raises TypeError (float * Tensor without __rmul__,
            # dlgrad/nn/optim.py:40); velocity = momentum*velocity + grad is what it spells
            self.velocities = {id(p): Tensor.zeros_like(p) for p in params}

    def step(self) -> None:
        if self.momentum == 0:
            return super().step()
        self._before_inplace([v.data for v in self.velocities.values()])
        for p in self.params:
            if p.grad is None:
                continue
            vel = self.velocities[id(p)]
            dev = self._device_for(p.data, p.grad.data, vel.data)
            if dev is Device.CUDA:
                # velocity and parameter are updated IN PLACE by one kernel: the launch arguments are the same every
                # step, so a captured step (dlgrad_b200/graph.py) replays it correctly; rebinding self.velocities to
                # a fresh Tensor every step (the host form below) is a Python side effect a replay would not see
                from dlgrad_b200.runtime.cuda import ops as cuda_ops
                cuda_ops.sgd_momentum_step(p.data, p.grad.data, vel.data, float(_f32(self.lr)),
                                           float(_f32(self.momentum)), self.grad_scale)
                continue
            v = (vel * self.momentum + p.grad).detach()
            self.velocities[id(p)] = v
            dispatcher.dispatch(op=CustomOps.SGD_STEP, device=dev, param=p.data, grad=v.data,
                                lr=float(_f32(self.lr)), grad_scale=self.grad_scale)


class Adam(Optimizer):
    def __init__(self, params: list[Tensor], lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8) -> None:
        self.params = params
        self.lr = lr
        self.beta1, self.beta2 = betas
        self.eps = eps
        self.t = 0
        self.grad_scale = 1.0
        self.m = {id(p): Tensor.zeros_like(p) for p in params}
        self.v = {id(p): Tensor.zeros_like(p) for p in params}
        self._scal = None         # device float[4] holding 1 - beta1^t, 1 - beta2^t, lr for steps inside a captured graph

    def step(self) -> None:
        from dlgrad_b200.runtime.cuda import shim
        if shim.capture["on"]:
            return self._captured_step()
        self.t += 1
        # python-double scalars rounded to float32 once, exactly where the reference wraps them in a
        # 0-d Tensor (dlgrad/nn/optim.py:64-69)
        bc1 = float(_f32(1.0 - self.beta1 ** self.t))
        bc2 = float(_f32(1.0 - self.beta2 ** self.t))
        lr = float(_f32(self.lr))
        self._before_inplace([t.data for t in self.m.values()] + [t.data for t in self.v.values()])
        b1, b2, eps = self.beta1, self.beta2, self.eps     # rounded to float32 inside the kernel generator
        live = [p for p in self.params if p.grad is not None]
        if live:
            dev = self._device_for(*[p.data for p in live], *[p.grad.data for p in live])
            if self._scal is None and dev is Device.CUDA:
                # allocated by an EAGER step on purpose: a block taken during a capture comes from the capture's pool,
                # where the graph's own kernels may write (graph.CapturedStep always runs a warm-up step first)
                from dlgrad_b200.runtime.cuda import shim as _shim
                self._scal = _shim.alloc(16)
            if dispatcher.has(CustomOps.ADAM_MULTI, dev):
                done = dispatcher.dispatch(
                    op=CustomOps.ADAM_MULTI, device=dev, params=[p.data for p in live],
                    grads=[p.grad.data for p in live], ms=[self.m[id(p)].data for p in live],
                    vs=[self.v[id(p)].data for p in live], beta1=b1, beta2=b2, eps=eps, bc1=bc1, bc2=bc2, lr=lr,
                    grad_scale=self.grad_scale)
                if done:
                    return
        for p in self.params:
            if p.grad is None:
                continue
            m, v = self.m[id(p)], self.v[id(p)]
            dev = self._device_for(p.data, p.grad.data, m.data, v.data)
            dispatcher.dispatch(op=CustomOps.ADAM_STEP, device=dev, param=p.data, grad=p.grad.data,
                                m=m.data, v=v.data, beta1=b1, beta2=b2, eps=eps, bc1=bc1, bc2=bc2, lr=lr,
                                grad_scale=self.grad_scale)

    # ------------------------------------------------------------------ inside a captured graph (dlgrad_b200/graph.py)
    def _scalars(self) -> tuple[float, float, float]:
        return (float(_f32(1.0 - self.beta1 ** self.t)), float(_f32(1.0 - self.beta2 ** self.t)), float(_f32(self.lr)))

    def _tick(self) -> None:
        """Before every replay: the step count advances on the host, the new scalars go to device memory."""
        from dlgrad_b200.runtime.cuda import ops as cuda_ops
        self.t += 1
        cuda_ops.adam_scalars(self._scal, *self._scalars())

    def _captured_step(self) -> None:
        """Record the multi-tensor kernel in the form that reads 1 - beta^t and lr from device memory; the capture
        itself executes nothing, so `t` does not advance here — `_tick` runs before each replay."""
        from dlgrad_b200.runtime.cuda import shim
        self._before_inplace([t.data for t in self.m.values()] + [t.data for t in self.v.values()])
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        if self._scal is None:
            raise RuntimeError("Adam.step() inside a captured step needs one eager step first (graph.CapturedStep's "
                               "warm-up provides it)")
        kw = dict(beta1=self.beta1, beta2=self.beta2, eps=self.eps, bc1=0.0, bc2=0.0, lr=0.0, grad_scale=self.grad_scale,
                  scalars=self._scal)
        done = dispatcher.dispatch(
            op=CustomOps.ADAM_MULTI, device=Device.CUDA, params=[p.data for p in live], grads=[p.grad.data for p in live],
            ms=[self.m[id(p)].data for p in live], vs=[self.v[id(p)].data for p in live], **kw)
        if not done:                               # some tensor is not float4-shaped: one launch per tensor, like eager
            for p in live:
                dispatcher.dispatch(op=CustomOps.ADAM_STEP, device=Device.CUDA, param=p.data, grad=p.grad.data,
                                    m=self.m[id(p)].data, v=self.v[id(p)].data, **kw)
        shim.capture["hooks"].append(self._tick)
