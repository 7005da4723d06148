"""Tensor: dlgrad's user API (creation functions, operator overloads, autograd driver) on top of
Buffer.  Method names, defaults and argument meaning follow dlgrad/tensor.py so that the MLP / GAN /
GPT scripts run unchanged with device="cuda"; the file itself is written against this package's
device model:

  * `OP.execute` stamps the result with the device the op actually ran on.  The reference copies the
    first parent's label (dlgrad/tensor.py:80-84), which is only harmless on unified memory.
  * a numpy array given to `Tensor(...)` is wrapped zero-copy as a host buffer and moved to HBM the
    first time an op touches it (or immediately when device="cuda").
  * `numpy()`, `tobytes()`, `copy_from()` and slicing go through D2H / H2D / offset views.
"""
from __future__ import annotations

from functools import reduce

from dlgrad_b200.buffer import Buffer
from dlgrad_b200.device import Device, as_device
from dlgrad_b200.dtype import DType, Scalar
from dlgrad_b200.runtime.cuda import ops as _cuda_runtime  # registers every (op, Device.CUDA)  # noqa: F401

_CUDA = Device.CUDA
# set by dlgrad_b200/distributed.py (overlapped data-parallel exchange): called with every tensor whose `.grad` the
# backward sweep has just assigned or accumulated into, so a bucket of parameters can be exchanged as soon as it is final
_grad_hook = None


class OP:
    """Base of the differentiable ops in dlgrad_b200/ops.py; one instance per executed op is the
    autograd node (`parents`, `req_grad`), as in dlgrad/tensor.py:15-33."""

    def __init__(self, *data: "Tensor") -> None:
        self.parents: tuple = data
        self.req_grad: list[bool] = [t.requires_grad for t in data]
        self.requires_grad: bool = any(self.req_grad)

    def forward(self, *args, **kwargs):
        raise NotImplementedError(f"forward not implemented for {type(self)}")

    def backward(self, *args, **kwargs):
        raise RuntimeError(f"backward not implemented for {type(self)}")

    @classmethod
    def execute(cls, *data: "Tensor", **kwargs) -> "Tensor":
        if isinstance(data[0], tuple):
            data = data[0]
        ctx = cls(*data)
        kwargs.pop("device", None)
        dtype = kwargs.pop("dtype", None)
        out = Tensor.__new__(Tensor)
        out.data = ctx.forward(*[t.data for t in data], **kwargs)
        if dtype is not None:
            out.data.metadata.dtype = dtype
        out.requires_grad = ctx.requires_grad
        out._ctx = ctx if ctx.requires_grad else None
        out.grad = None
        return out


import dlgrad_b200.ops as ops  # noqa: E402  (ops imports OP from this module)


def _scalar_like(v, ref: "Tensor") -> "Tensor":
    return Tensor(float(v), device=ref.device)


class Tensor:
    def __init__(self, data: "Buffer | np.ndarray | Scalar", requires_grad: bool = False,  # noqa: F821
                 device: Device | str = Device.CPU) -> None:
        self.requires_grad = requires_grad
        self._ctx: OP | None = None
        self.grad: Tensor | None = None
        dev = as_device(device)
        if isinstance(data, Buffer):
            self.data = data
        elif isinstance(data, (int, float)):
            self.data = Buffer.from_scalar(float(data))
        elif type(data).__module__ == "numpy" and hasattr(data, "dtype"):
            if str(data.dtype) != "float32":
                raise ValueError("dlgrad only supports float32 dtype")
            if data.ndim > 4:
                raise ValueError("dlgrad only supports upto 4D tensors")
            if not data.flags["C_CONTIGUOUS"]:
                import numpy as np
                data = np.ascontiguousarray(data)
            self.data = Buffer.from_numpy(data)
        else:
            raise ValueError("The data must be of type Buffer, np.ndarray or float")
        if dev is _CUDA and self.data.scalar is None:
            self.data._to_cuda()

    # ------------------------------------------------------------------ host access
    def numpy(self) -> "np.ndarray":  # noqa: F821
        return self.data.numpy()

    def tobytes(self) -> bytes:
        """Raw little-endian float32 contents (D2H first when on CUDA); dlgrad/tensor.py:125-129."""
        return self.data.numpy().tobytes()

    def copy_from(self, source_bytes: bytes) -> None:
        """Overwrite the storage with raw bytes (H2D when on CUDA); dlgrad/tensor.py:131-138."""
        if len(source_bytes) != self.nbytes:
            raise ValueError(f"Buffer size mismatch! Expected {self.nbytes} bytes, got {len(source_bytes)}")
        from dlgrad_b200.buffer import flush_pending
        flush_pending()
        flush_pending((self.data,))       # in-place write: private storage, and the content tag (if any) no longer holds
        if self.data.metadata.device is _CUDA:
            from dlgrad_b200.runtime.cuda import transfer
            transfer.write_bytes(self.data.ptr, bytes(source_bytes))
        else:
            from dlgrad_b200.helpers import ffi
            ffi.memmove(self.data.ptr, source_bytes, len(source_bytes))

    def show(self) -> None:
        self.data.show()

    def to(self, device: Device | str) -> "Tensor":
        if as_device(device) is _CUDA:
            self.data._to_cuda()
        return self

    # ------------------------------------------------------------------ creation
    @staticmethod
    def uniform(shape: tuple, device: str | Device | None = Device.CPU, dtype: str | DType | None = DType.FLOAT32,
                low: float = 0.0, high: float = 1.0, requires_grad: bool = False) -> "Tensor":
        return Tensor(Buffer.uniform(shape, device=device, dtype=dtype, low=low, high=high), requires_grad)

    @staticmethod
    def rand(shape: tuple, device: str | Device | None = Device.CPU, dtype: str | DType | None = DType.FLOAT32,
             requires_grad: bool = False) -> "Tensor":
        return Tensor.uniform(shape, device, dtype, requires_grad=requires_grad)

    @staticmethod
    def full(shape: tuple, fill_value: Scalar, device: Device = Device.CPU, dtype: DType = DType.FLOAT32,
             requires_grad: bool = False) -> "Tensor":
        return Tensor(Buffer.full(shape, fill_value=fill_value, device=device, dtype=dtype), requires_grad)

    @staticmethod
    def ones(shape: tuple, device: Device = Device.CPU, dtype: DType = DType.FLOAT32,
             requires_grad: bool = False) -> "Tensor":
        return Tensor.full(shape, 1.0, device, dtype, requires_grad)

    @staticmethod
    def ones_like(input: "Tensor", device: Device = Device.CPU, dtype: DType = DType.FLOAT32,  # noqa: A002
                  requires_grad: bool = False) -> "Tensor":
        return Tensor.full(input.shape, 1.0, device, dtype, requires_grad)

    @staticmethod
    def zeros_like(input: "Tensor", device: Device = Device.CPU, dtype: DType = DType.FLOAT32,  # noqa: A002
                   requires_grad: bool = False) -> "Tensor":
        return Tensor.full(input.shape, 0.0, device, dtype, requires_grad)

    @staticmethod
    def arange(shape: tuple, device: Device = Device.CPU, dtype: DType = DType.FLOAT32,
               requires_grad: bool = False) -> "Tensor":
        return Tensor(Buffer.arange(shape, device=device, dtype=dtype), requires_grad)

    # ------------------------------------------------------------------ arithmetic
    @staticmethod
    def _pair(x, y) -> tuple["Tensor", "Tensor"]:
        if not isinstance(y, Tensor):
            y = _scalar_like(y, x)
        elif not isinstance(x, Tensor):
            x = _scalar_like(x, y)
        return x, y

    @staticmethod
    def add(x: "Tensor", y: "Tensor | Scalar") -> "Tensor":
        return ops.Add.execute(*Tensor._pair(x, y))

    @staticmethod
    def sub(x: "Tensor", y: "Tensor | Scalar") -> "Tensor":
        return ops.Sub.execute(*Tensor._pair(x, y))

    @staticmethod
    def mul(x: "Tensor", y: "Tensor | Scalar") -> "Tensor":
        return ops.Mul.execute(*Tensor._pair(x, y))

    @staticmethod
    def div(x: "Tensor", y: "Tensor | Scalar") -> "Tensor":
        return ops.Div.execute(*Tensor._pair(x, y))

    @staticmethod
    def matmul(x: "Tensor", y: "Tensor") -> "Tensor":
        return ops.MatMul.execute(x, y)

    def linear(self, weight: "Tensor", bias: "Tensor | None") -> "Tensor":
        """`self @ weight.T + bias` (dlgrad/tensor.py:520-535)."""
        if not self._on_host_runtime() and weight.ndim == 2 and self.ndim >= 2:
            Buffer._route(self.data, weight.data)          # host-created layers migrate to HBM here, once
            return ops.Linear.execute(self, weight, bias) if bias is not None else ops.Linear.execute(self, weight)
        out = self @ weight.T
        return out + bias if bias is not None else out

    # ------------------------------------------------------------------ shape
    def reshape(self, new_shape: tuple) -> "Tensor":
        return ops.Reshape.execute(self, shape=new_shape)

    def permute(self, order: tuple) -> "Tensor":
        return ops.Permute.execute(self, order=order)

    def transpose(self, dim0: int, dim1: int) -> "Tensor":
        nd = self.ndim
        d0, d1 = dim0 % nd, dim1 % nd
        if {d0, d1} == {nd - 2, nd - 1} and not self._on_host_runtime():
            return ops.Transpose.execute(self, dim0=d0, dim1=d1)      # deferred on CUDA, folded into matmul
        order = list(range(nd))
        order[d0], order[d1] = order[d1], order[d0]
        return self.permute(tuple(order))

    @staticmethod
    def cat(tensors: "list[Tensor] | tuple", dim: int = 0) -> "Tensor":
        """Concatenate along `dim` on the tensors' device (extension: the reference's scripts use
        np.concatenate on host copies, examples/gpt.py:140-147)."""
        tensors = list(tensors)
        if not tensors:
            raise ValueError("Tensor.cat needs at least one tensor")
        out = tensors[0]
        d = dim % out.ndim
        for t in tensors[1:]:
            out = ops.Cat.execute(out, t, dim=d)
        return out

    def squeeze(self, dim: list[int] | int) -> "Tensor":
        return ops.Squeeze.execute(self, dim=dim)

    def unsqueeze(self, dim: list[int] | int) -> "Tensor":
        return ops.Unsqueeze.execute(self, dim=dim)

    @property
    def T(self) -> "Tensor":  # noqa: N802
        return self.transpose(1, 0)

    # ------------------------------------------------------------------ reductions
    def sum(self, dim: int = -1, keepdim: bool = False) -> "Tensor":
        return ops.Sum.execute(self, dim=dim, keepdim=keepdim)

    def mean(self, dim: int = -1, keepdim: bool = False) -> "Tensor":
        return ops.Mean.execute(self, dim=dim, keepdim=keepdim)

    def max(self, dim: int = -1, keepdim: bool = False) -> "Tensor":
        return ops.Max.execute(self, dim=dim, keepdim=keepdim)

    def argmax(self, dim: int = -1) -> "Tensor":
        return Tensor(self.data.argmax(dim))

    # ------------------------------------------------------------------ activations / math
    def relu(self) -> "Tensor":
        return ops.Relu.execute(self)

    def leaky_relu(self, neg_slope: Scalar = 0.01) -> "Tensor":
        return ops.LeakyRelu.execute(self, neg_slope=neg_slope)

    def tanh(self) -> "Tensor":
        return ops.Tanh.execute(self)

    def sigmoid(self) -> "Tensor":
        return ops.Sigmoid.execute(self)

    def exp(self) -> "Tensor":
        return ops.Exp.execute(self)

    def log(self) -> "Tensor":
        return ops.Log.execute(self)

    def sqrt(self) -> "Tensor":
        return ops.Sqrt.execute(self)

    def rsqrt(self) -> "Tensor":
        return ops.RSqrt.execute(self)

    def clamp(self, min: int | None = None, max: int | None = None) -> "Tensor":  # noqa: A002
        return ops.Clamp.execute(self, min=min, max=max)

    def _on_host_runtime(self) -> bool:
        """True when this tensor would be computed by a registered host runtime (the test oracle)."""
        return Buffer._route(self.data) is not _CUDA

    def softmax(self, dim: int = -1) -> "Tensor":
        """Softmax along `dim` (negative dims count from the end).  On CUDA this is one fused node;
        on a host runtime it is the reference's max/sub/exp/sum/div graph (dlgrad/tensor.py:634-649)
        so that the oracle differentiates through exactly the reference's chain."""
        dim = dim % self.ndim
        if self._on_host_runtime():
            e = (self - self.max(dim=dim, keepdim=True)).exp()
            return e / e.sum(dim=dim, keepdim=True)
        return ops.Softmax.execute(self, dim=dim)

    def log_softmax(self, dim: int = 1) -> "Tensor":
        dim = dim % self.ndim
        if self._on_host_runtime():      # dlgrad/tensor.py:652-669
            m = self - self.max(dim=dim, keepdim=True)
            return m - m.exp().sum(dim=dim, keepdim=True).log()
        return ops.LogSoftmax.execute(self, dim=dim)

    def dropout(self, p: Scalar) -> "Tensor":
        return ops.Dropout.execute(self, p=p)

    def sequential(self, layers: list) -> "Tensor":
        return reduce(lambda acc, layer: layer(acc), layers, self)

    # ------------------------------------------------------------------ losses / indexing ops
    def cross_entropy_loss(self, target: "Tensor") -> "Tensor":
        if not isinstance(target, Tensor):
            target = Tensor(target)
        return ops.CrossEntropy.execute(self, target)

    def bcewithlogitsloss(self, target: "Tensor") -> "Tensor":
        if not isinstance(target, Tensor):
            target = Tensor(target)
        return ops.BCEWithLogitsLoss.execute(self, target)

    @staticmethod
    def embedding(weight: "Tensor", idx: "Tensor") -> "Tensor":
        return ops.Embedding.execute(weight, idx)

    @staticmethod
    def tril(inp: "Tensor", k: Scalar = 0.0) -> "Tensor":
        if not isinstance(k, Tensor):
            k = Tensor(float(k))
        return ops.Tril.execute(inp, k)

    def masked_fill(self, mask: "Tensor", value: Scalar) -> "Tensor":
        return ops.MaskedFill.execute(self, mask, value=value)

    @staticmethod
    def where(cond: "Tensor", inp: "Tensor | Scalar", other: "Tensor | Scalar") -> "Tensor":
        if not isinstance(inp, Tensor):
            inp = Tensor(inp)
        if not isinstance(other, Tensor):
            other = Tensor(other)
        return ops.Where.execute(cond, inp, other)

    def sample_last(self, temperature: float = 1.0, u: "Tensor | None" = None) -> "Tensor":
        """Draw the next token of every sequence from softmax(self[..., -1, :] / temperature) ON THE DEVICE and return
        the indices as a (B, 1) float tensor (like argmax).  Extension for the inference path: the reference's
        generate loops (examples/gpt.py:216-221, 262-266) copy the logits to the host and call np.random.choice for
        every token; with this op only the chosen index crosses PCIe — or nothing, when the result is fed straight back
        into the embedding.  On a host runtime (the test oracle) the same inverse-CDF rule runs in numpy."""
        import numpy as np
        rows = 1 if self.ndim == 1 else self.shape[0]
        if self._on_host_runtime():
            x = self.numpy().reshape(rows, -1, self.shape[-1])[:, -1, :].astype(np.float32)
            un = (np.random.default_rng().random(rows) if u is None else u.numpy().reshape(rows)).astype(np.float32)
            z = (x / np.float32(temperature)).astype(np.float32)
            e = np.exp(z - z.max(axis=1, keepdims=True)).astype(np.float64)
            cdf = np.cumsum(e, axis=1)
            pick = [min(int(np.searchsorted(cdf[r], float(un[r]) * cdf[r, -1], side="right")), x.shape[1] - 1) for r in range(rows)]
            return Tensor(np.array(pick, dtype=np.float32).reshape(rows, 1))
        self.data._to_cuda()
        if u is not None:
            u.data._to_cuda()
        ptr = _cuda_runtime.sample_last(self.data, float(temperature), None if u is None else u.data)
        return Tensor(Buffer(ptr, (rows, 1), _CUDA, self.dtype))

    def detach(self) -> "Tensor":
        return Tensor(self.data, requires_grad=False)

    # ------------------------------------------------------------------ autograd
    def backward(self) -> None:
        """Reverse-mode sweep over the recorded graph (dlgrad/tensor.py:777-809): topological order
        by depth-first search from the loss, seed 1.0, accumulate into `.grad` of every tensor that
        requires it."""
        assert all(d == 1 for d in self.shape), "backward must be called on a scalar Tensor"
        order: list[Tensor] = []
        seen: set[int] = set()
        stack: list[tuple[Tensor, bool]] = [(self, False)]
        while stack:                      # iterative DFS: no recursion limit on deep graphs
            node, expanded = stack.pop()
            if expanded:
                order.append(node)
                continue
            if id(node) in seen or node._ctx is None:
                continue
            seen.add(id(node))
            stack.append((node, True))
            for parent in reversed(node._ctx.parents):
                stack.append((parent, False))
        self.grad = Tensor(1.0)
        for node in reversed(order):
            if node.grad is None:
                raise RuntimeError(f"Tensor {node} has no grad")
            grads = node._ctx.backward(node.grad.data)
            for parent, g in zip(node._ctx.parents, grads):
                if g is None or not parent.requires_grad:
                    continue
                g_t = Tensor(g)
                parent.grad = g_t if parent.grad is None else parent.grad + g_t
                if _grad_hook is not None:
                    _grad_hook(parent)
            if node is not self:
                # an interior node's gradient has been consumed: release it (a (B,h,T,T) attention gradient
                # is 400 MB at the benchmark config; the reference keeps every one alive until the graph dies)
                node.grad = None
        if order:
            self._ctx = None

    def __getitem__(self, idx: slice) -> "Tensor":
        """Leading-axis slices only (dlgrad/tensor.py:811-823); the view shares storage and keeps
        its base alive."""
        if not (isinstance(idx, slice) and idx.start is not None and idx.stop is not None and idx.step is None):
            raise ValueError("dlgrad only supports slicing or start, stop are None")
        offset = idx.start * self.stride[0]
        shape = (idx.stop - idx.start, *self.shape[1:])
        base = self.data
        view = Buffer(base.ptr + offset, shape, base.device, base.dtype,
                      aligned=base.aligned and offset % 4 == 0, keep=(base.ptr, base._keep))
        view._cow = base._cow            # a slice of storage shared with the constant cache is copy-on-write as well
        return Tensor(view, requires_grad=self.requires_grad)

    # ------------------------------------------------------------------ operators
    def __add__(self, other):
        return Tensor.add(self, other)

    def __radd__(self, other):
        return Tensor.add(self, other)

    def __mul__(self, other):
        return Tensor.mul(self, other)

    def __rmul__(self, other):
        return Tensor.mul(self, other)

    def __sub__(self, other):
        return Tensor.sub(self, other)

    def __rsub__(self, other):
        return Tensor.sub(_scalar_like(other, self), self)

    def __truediv__(self, other):
        return Tensor.div(self, other)

    def __pow__(self, val: Scalar) -> "Tensor":
        return ops.Pow.execute(self, y=float(val))

    def __neg__(self) -> "Tensor":
        return Tensor(-self.data, requires_grad=self.requires_grad)

    def __matmul__(self, other: "Tensor") -> "Tensor":
        return Tensor.matmul(self, other)

    def __gt__(self, other: Scalar) -> "Tensor":
        return Tensor(self.data > (other.data if isinstance(other, Tensor) else other))

    def __ge__(self, other: Scalar) -> "Tensor":
        return Tensor(self.data >= (other.data if isinstance(other, Tensor) else other))

    def __eq__(self, other: "Tensor") -> "Tensor":  # type: ignore[override]
        return Tensor(self.data == (other.data if isinstance(other, Tensor) else Buffer.from_scalar(float(other))))

    def __hash__(self) -> int:
        return id(self)

    def __repr__(self) -> str:
        return (f"Tensor<dtype: {self.dtype} device: {self.device}, shape: {self.shape}, ndim: {self.ndim}, "
                f"size: {self.numel * 4 / 1e6}mb, requires_grad: {self.requires_grad}>")

    # ------------------------------------------------------------------ metadata
    @property
    def numel(self) -> int:
        return self.data.metadata.numel

    @property
    def shape(self) -> tuple:
        return self.data.metadata.shape

    @property
    def stride(self) -> tuple:
        return self.data.metadata.stride

    @property
    def ndim(self) -> int:
        return self.data.metadata.ndim

    @property
    def dtype(self) -> DType:
        return self.data.metadata.dtype

    @property
    def device(self) -> Device:
        return self.data.metadata.device

    @property
    def nbytes(self) -> int:
        return self.data.metadata.nbytes
