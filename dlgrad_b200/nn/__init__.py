"""Layers (dlgrad/nn/__init__.py:8-41): Linear, Embedding, RMSNorm — pure Tensor-level code.

Embedding and RMSNorm take no device in the reference and therefore always create host tensors
(dlgrad/nn/__init__.py:17-31).  Their parameters are host buffers here too until the first op
touches them, at which point Buffer migrates the storage to HBM in place, once.
"""
import math

from dlgrad_b200.device import Device
from dlgrad_b200.nn import optim, utils  # noqa: F401
from dlgrad_b200.tensor import Tensor


class Linear:
    """y = x @ W^T + b with W (out, in), b (1, out), both U(-1/sqrt(in), 1/sqrt(in))."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device: Device = Device.CPU) -> None:
        k = 1.0 / math.sqrt(in_features)
        self.weight = Tensor.uniform((out_features, in_features), low=-k, high=k, requires_grad=True, device=device)
        self.bias = (Tensor.uniform((1, out_features), low=-k, high=k, requires_grad=True, device=device)
                     if bias else None)

    def __call__(self, x: Tensor) -> Tensor:
        return x.linear(self.weight, self.bias)


class Embedding:
    def __init__(self, num_embeddings: int, embedding_dim: int) -> None:
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.weight = Tensor.uniform((num_embeddings, embedding_dim), requires_grad=True)

    def __call__(self, idx: Tensor) -> Tensor:
        return Tensor.embedding(self.weight, idx)


class RMSNorm:
    def __init__(self, dim: int, eps: float = 1e-6, elementwise_affine: bool = True) -> None:
        self.eps = eps
        self.weight = Tensor.ones((1, dim), requires_grad=True) if elementwise_affine else None

    def _norm(self, x: Tensor) -> Tensor:
        ms = (x ** 2).mean(x.ndim - 1, keepdim=True)
        return x * (ms + self.eps).rsqrt()

    def __call__(self, x: Tensor) -> Tensor:
        if not x._on_host_runtime():
            # Device.CUDA: one fused node (codegen/cuda_kernel.rmsnorm_rows); same arithmetic and the same
            # gradient the composition below produces through the reference's autograd
            from dlgrad_b200 import ops
            if self.weight is None:
                return ops.RMSNorm.execute(x, eps=self.eps)
            return ops.RMSNorm.execute(x, self.weight, eps=self.eps)
        y = self._norm(x)
        return y if self.weight is None else y * self.weight
