"""Differentiable ops: forward/backward in Buffer algebra (class names follow dlgrad/ops.py so code
that names `ops.Add`, `ops.MatMul`, ... keeps working).

The gradient formulas are the reference's; where the CUDA backend has a fused kernel the op calls
it (softmax / log-softmax, cross-entropy gradient, transposed matmuls in MatMul.backward) instead of
composing primitives.  Reference behaviours that look like bugs but define the oracle's numbers are
kept and marked QUIRK:
  * Mean.backward multiplies by Buffer.full(1/N), and `full` truncates to int -> zeros;
  * CrossEntropy.forward returns the SUM of the per-row NLL while backward divides by N.
"""
from __future__ import annotations

import os
import weakref

from dlgrad_b200.buffer import Buffer
from dlgrad_b200.device import Device
from dlgrad_b200.dispatch import dispatcher
from dlgrad_b200.dtype import Scalar
from dlgrad_b200.helpers import CustomOps, UnaryOps, broadcast_shapes, check_broadcast
from dlgrad_b200.tensor import OP


def _expand_dims(g: Buffer, dim: int, keepdim: bool) -> Buffer:
    """View of a reduced gradient with the reduced axis restored as size 1 (the reference edits the
    incoming Buffer's metadata in place, dlgrad/ops.py:69-71; a reshaped view leaves it alone)."""
    if keepdim:
        return g
    shp = list(g.shape)
    shp.insert(0 if dim == -1 else dim, 1)
    return g.reshape(tuple(shp))


def unbroadcast(grad: Buffer, original_shape: tuple) -> Buffer:
    """Sum a gradient back to the shape of a broadcast operand (dlgrad/ops.py:287-295 and
    OP.reduce_grad_for_broadcasting, dlgrad/tensor.py:35-54)."""
    while grad.ndim > len(original_shape):
        grad = grad.sum(0, keepdim=False)
    for i, (gd, od) in enumerate(zip(grad.shape, original_shape)):
        if od == 1 and gd > 1:
            grad = grad.sum(i, keepdim=True)
    return grad


# ------------------------------------------------------------------ shape ops
class Reshape(OP):
    def forward(self, x: Buffer, shape: tuple) -> Buffer:
        self.input_shape = x.shape
        return x.reshape(shape)

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return (upstream_grad.reshape(self.input_shape),)


class Permute(OP):
    def forward(self, x: Buffer, order: tuple) -> Buffer:
        self.order = tuple(order)
        return x.permute(self.order)

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        inverse = [0] * len(self.order)
        for dst, src in enumerate(self.order):
            inverse[src] = dst
        return (upstream_grad.permute(tuple(inverse)),)


class Transpose(OP):
    def forward(self, x: Buffer, dim0: int, dim1: int) -> Buffer:
        self.dims = (dim0, dim1)
        return x.transpose(dim0, dim1)

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return (upstream_grad.transpose(self.dims[1], self.dims[0]),)


class Cat(OP):
    """Concatenate two tensors along `dim`.  An addition of the CUDA backend for the inference path: the
    reference appends to its KV cache with np.concatenate on the host (examples/gpt.py:140-147).  On CUDA it is
    one kernel; on a host runtime it is the same numpy concatenation.  Backward hands each operand its slice of
    the upstream gradient."""
    def forward(self, x: Buffer, y: Buffer, dim: int) -> Buffer:
        self.dim, self.xs, self.ys = dim, x.shape, y.shape
        if x.shape[:dim] != y.shape[:dim] or x.shape[dim + 1:] != y.shape[dim + 1:]:
            raise ValueError(f"Cannot concatenate {x.shape} and {y.shape} along dim {dim}")
        out_shape = x.shape[:dim] + (x.shape[dim] + y.shape[dim],) + x.shape[dim + 1:]
        dev = Buffer._route(x, y)
        if dev is Device.CUDA:
            from dlgrad_b200.helpers import CustomOps
            return Buffer(dispatcher.dispatch(op=CustomOps.CAT, device=dev, x=x, y=y, dim=dim), out_shape, dev, x.dtype)
        import numpy as np
        return Buffer.from_numpy(np.ascontiguousarray(np.concatenate((x.numpy(), y.numpy()), axis=dim)))

    def backward(self, upstream_grad: Buffer) -> tuple:
        import numpy as np
        g = upstream_grad.numpy()                       # inference-path op: the gradient split stays simple
        n0 = self.xs[self.dim]
        idx_a = tuple(slice(None) if d != self.dim else slice(0, n0) for d in range(len(self.xs)))
        idx_b = tuple(slice(None) if d != self.dim else slice(n0, None) for d in range(len(self.xs)))
        dev = upstream_grad.device

        def up(a):
            b = Buffer.from_numpy(np.ascontiguousarray(a))
            if dev is Device.CUDA:
                b._to_cuda()
            return b
        return (up(g[idx_a]) if self.req_grad[0] else None, up(g[idx_b]) if self.req_grad[1] else None)


class Squeeze(OP):
    def forward(self, x: Buffer, dim: list[int] | int) -> Buffer:
        self.dim = dim
        x.squeeze(dim)
        return x

    def backward(self, grad_output: Buffer) -> tuple[Buffer]:
        grad_output.unsqueeze(self.dim)
        return (grad_output,)


class Unsqueeze(OP):
    def forward(self, x: Buffer, dim: list[int] | int) -> Buffer:
        self.dim = dim
        x.unsqueeze(dim)
        return x

    def backward(self, grad_output: Buffer) -> tuple[Buffer]:
        grad_output.squeeze(self.dim)
        return (grad_output,)


# ------------------------------------------------------------------ reductions
class Sum(OP):
    def forward(self, x: Buffer, dim: int = -1, keepdim: bool = False) -> Buffer:
        self.inp_shape, self.device, self.dtype = x.shape, x.device, x.dtype
        self.dim, self.keepdim = dim, keepdim
        return x.sum(dim=dim, keepdim=keepdim)

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        ones = Buffer.full(shape=self.inp_shape, fill_value=1.0, device=self.device, dtype=self.dtype)
        return (ones * _expand_dims(upstream_grad, self.dim, self.keepdim),)


class Mean(OP):
    def forward(self, x: Buffer, dim: int = -1, keepdim: bool = False) -> Buffer:
        self.inp_shape, self.device, self.dtype = x.shape, x.device, x.dtype
        self.dim, self.keepdim = dim, keepdim
        self.N = x.numel if dim == -1 else x.shape[dim]
        return x.mean(dim=dim, keepdim=keepdim)

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        # QUIRK: full() truncates 1/N to int(0) (dlgrad/ops.py:86-91 + dlgrad/runtime/cpu.py:252),
        # so the reference's mean gradient is all zeros.  Kept: it defines RMSNorm's input gradient.
        scale = Buffer.full(self.inp_shape, fill_value=1.0 / self.N, device=self.device, dtype=self.dtype)
        return (scale * _expand_dims(upstream_grad, self.dim, self.keepdim),)


class Max(OP):
    def forward(self, x: Buffer, dim: int = -1, keepdim: bool = False) -> Buffer:
        self.x, self.dim, self.keepdim = x, dim, keepdim
        self.out = x.max(dim=dim, keepdim=keepdim)
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        is_max = self.x == _expand_dims(self.out, self.dim, self.keepdim)
        return (is_max * _expand_dims(upstream_grad, self.dim, self.keepdim),)


# ------------------------------------------------------------------ unary
class Pow(OP):
    def forward(self, x: Buffer, y: float) -> Buffer:
        self.x, self.y = x, y
        return x ** y

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return (upstream_grad * ((self.x ** (self.y - 1.0)) * self.y),)


class Exp(OP):
    def forward(self, x: Buffer) -> Buffer:
        self.out = x.exp()
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return (upstream_grad * self.out,)


class Log(OP):
    def forward(self, x: Buffer) -> Buffer:
        self.x = x
        return x.log()

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return (upstream_grad / self.x,)


class Sqrt(OP):
    def forward(self, x: Buffer) -> Buffer:
        self.out = x.sqrt()
        return self.out

    def backward(self, grad_output: Buffer) -> tuple[Buffer]:
        return (grad_output / (self.out * 2.0),)


class RSqrt(OP):
    def forward(self, x: Buffer) -> Buffer:
        self.out = x.rsqrt()
        return self.out

    def backward(self, grad_output: Buffer) -> tuple[Buffer]:
        return (grad_output * ((self.out ** 3) * -0.5),)


# gradient Buffers that a Linear backward produced with the relu-backward select already applied in the GEMM epilogue:
# id(gradient Buffer) -> id(pre-activation Buffer)   (Buffer.__eq__ is the elementwise op, so no dict keyed by the object)
_premasked: dict = {}
RELU_BWD_EPILOGUE = os.environ.get("DLGRAD_RELU_BWD_EPILOGUE", "1") != "0"
LAZY_LINEAR = os.environ.get("DLGRAD_LAZY_LINEAR", "1") != "0"      # deferred Linear outputs / input gradients / rms_dx (fused adds)


class Relu(OP):
    """On CUDA a large relu is DEFERRED (Buffer pending kind "relu"): the GEMM that consumes it applies max(., 0) in
    its operand splitter and relu(x) is never written; any other consumer materialises it with the reference's
    `where(x > 0, x, 0)` kernel (dlgrad/ops.py:151-158).  The backward selects on the pre-activation
    (x > 0 <=> relu(x) > 0), so it does not need the output either."""
    def forward(self, x: Buffer) -> Buffer:
        from dlgrad_b200.buffer import LAZY_MIN_NUMEL
        self.x = x
        if Buffer._route(x) is Device.CUDA and x.scalar is None and x.numel >= LAZY_MIN_NUMEL:
            self.out = Buffer._deferred(("relu", x), x.shape, x.dtype)
        else:
            self.out = x.relu()
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        if _premasked.get(id(upstream_grad)) == id(self.x):
            return (upstream_grad,)      # the GEMM that produced this gradient already selected on x > 0 (Linear.backward)
        if Buffer._route(self.x, upstream_grad) is Device.CUDA and self.x.shape == upstream_grad.shape:
            # one select instead of the reference's where(1, 0) followed by a multiply (same values)
            return (self.x.where(upstream_grad, 0.0),)
        return (self.out.where(inp=1.0, other=0.0) * upstream_grad,)


class LeakyRelu(OP):
    def forward(self, x: Buffer, neg_slope: Scalar = 0.01) -> Buffer:
        self.neg_slope = neg_slope
        self.out = x.leaky_relu(neg_slope=neg_slope)
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return (self.out.where(inp=1.0, other=self.neg_slope) * upstream_grad,)


class Sigmoid(OP):
    def forward(self, x: Buffer) -> Buffer:
        self.out = x.sigmoid()
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return ((self.out * -(self.out - 1.0)) * upstream_grad,)


class Tanh(OP):
    def forward(self, x: Buffer) -> Buffer:
        self.out = x.tanh()
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        one = Buffer.full(self.out.shape, 1.0, self.out.device, self.out.dtype)
        return ((one - self.out ** 2) * upstream_grad,)


class Clamp(OP):
    def forward(self, x: Buffer, min: int | None, max: int | None) -> Buffer:  # noqa: A002
        self.min, self.max = min, max
        return x.clamp(min, max)

    def backward(self, grad_output: Buffer) -> tuple[Buffer]:
        # the reference clamps the gradient itself (dlgrad/ops.py:201-202); kept
        return (grad_output.clamp(self.min, self.max),)


class Softmax(OP):
    """Tensor.softmax as one node (fused kernel on CUDA).  The reference builds it from
    max/sub/exp/sum/div nodes (dlgrad/tensor.py:634-649); value and gradient are the same function."""
    def forward(self, x: Buffer, dim: int) -> Buffer:
        self.dim = dim
        self.out = x.softmax(dim)
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        dev = Buffer._route(self.out, upstream_grad)
        if dev is Device.CUDA and self.dim == self.out.ndim - 1 and self.out.numel >= (1 << 16):
            # deferred: the masked_fill / scale backward that usually follow are folded into this kernel
            return (Buffer._deferred(("softmax_bwd", self.out, upstream_grad, self.dim, "softmax"), self.out.shape,
                                     self.out.dtype),)
        g = dispatcher.dispatch(op=UnaryOps.SOFTMAX_BACKWARD, device=dev, y=self.out, grad=upstream_grad, dim=self.dim)
        return (Buffer(g, self.out.shape, dev, self.out.dtype),)


class LogSoftmax(OP):
    def forward(self, x: Buffer, dim: int) -> Buffer:
        self.dim = dim
        self.out = x.log_softmax(dim)
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        dev = Buffer._route(self.out, upstream_grad)
        g = dispatcher.dispatch(op=UnaryOps.LOG_SOFTMAX_BACKWARD, device=dev, y=self.out, grad=upstream_grad,
                                dim=self.dim)
        return (Buffer(g, self.out.shape, dev, self.out.dtype),)


class RMSNorm(OP):
    """nn.RMSNorm as one node on CUDA (dlgrad/nn/__init__.py:28-41 composes pow / mean / add / rsqrt /
    mul / mul).  The backward is what the reference's autograd yields for that composition INCLUDING its
    zero Mean.backward (QUIRK, dlgrad/ops.py:86-91): the rsqrt branch receives no gradient, so
        dx = (g * w) * rstd                      dw = sum_rows g * (x * rstd)."""
    def forward(self, x: Buffer, weight: Buffer | None = None, eps: float = 1e-6) -> Buffer:
        from dlgrad_b200.helpers import CustomOps as _C
        dev = Buffer._route(x) if weight is None else Buffer._route(x, weight)
        self.x, self.w = x, weight
        y, rstd = dispatcher.dispatch(op=_C.RMSNORM, device=dev, x=x, weight=weight, eps=eps)
        self.rstd = Buffer(rstd, x.shape[:-1] + (1,), dev, x.dtype)
        return Buffer(y, x.shape, dev, x.dtype)

    def backward(self, upstream_grad: Buffer) -> tuple:
        from dlgrad_b200.runtime.cuda import ops as _rt
        g, x, w, t = upstream_grad, self.x, self.w, self.rstd
        gx = gw = None
        if self.req_grad[0]:
            if w is not None and LAZY_LINEAR and x.numel >= (1 << 20) and g.shape == x.shape:
                # deferred ("rms_dx"): usually added to the residual stream's gradient next — one kernel for both
                gx = Buffer._deferred(("rms_dx", g, w, t), x.shape, x.dtype)
            elif w is not None:
                gx = Buffer(_rt.ewise3("rms_dx", "(v0 * v1) * v2", x.shape, [g, w, t]), x.shape, Device.CUDA, x.dtype)
            else:
                gx = Buffer(_rt.ewise3("rms_dx", "v0 * v1", x.shape, [g, t]), x.shape, Device.CUDA, x.dtype)
        if w is not None and len(self.req_grad) > 1 and self.req_grad[1]:
            gw = Buffer(_rt.rms_dw(g, x, t), w.shape, Device.CUDA, x.dtype)       # sum_rows g * (x * rstd), product never written
        return (gx, gw)


# ------------------------------------------------------------------ binary
class Add(OP):
    def forward(self, x: Buffer, y: Buffer) -> Buffer:
        self.xs, self.ys = x.shape, y.shape
        check_broadcast(x.shape, y.shape)
        return x + y

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer | None, Buffer | None]:
        gx = unbroadcast(upstream_grad, self.xs) if self.req_grad[0] else None
        gy = unbroadcast(upstream_grad, self.ys) if self.req_grad[1] else None
        return gx, gy


class Sub(OP):
    def forward(self, x: Buffer, y: Buffer) -> Buffer:
        self.xs, self.ys = x.shape, y.shape
        check_broadcast(x.shape, y.shape)
        return x - y

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer | None, Buffer | None]:
        gx = unbroadcast(upstream_grad, self.xs) if self.req_grad[0] else None
        gy = unbroadcast(-upstream_grad, self.ys) if self.req_grad[1] else None
        return gx, gy


class Mul(OP):
    def forward(self, x: Buffer, y: Buffer) -> Buffer:
        self.x, self.y = x, y
        check_broadcast(x.shape, y.shape)
        return x * y

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer | None, Buffer | None]:
        gx = unbroadcast(upstream_grad * self.y, self.x.shape) if self.req_grad[0] else None
        gy = unbroadcast(upstream_grad * self.x, self.y.shape) if self.req_grad[1] else None
        return gx, gy


class Div(OP):
    def forward(self, x: Buffer, y: Buffer) -> Buffer:
        self.x, self.y = x, y
        check_broadcast(x.shape, y.shape)
        return x / y

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer | None, Buffer | None]:
        gx = unbroadcast(upstream_grad / self.y, self.x.shape) if self.req_grad[0] else None
        gy = (unbroadcast((-upstream_grad * self.x) / self.y ** 2, self.y.shape)
              if self.req_grad[1] else None)
        return gx, gy


class MatMul(OP):
    def forward(self, x: Buffer, y: Buffer) -> Buffer:
        self.p1, self.p2, self.out_shape = broadcast_shapes(x.shape, y.shape)
        self.x, self.y = x, y
        self.out = x @ y                 # kept for the fused attention backward: rowsum(dP * P) = rowsum(dO * O)
        return self.out

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer | None, Buffer | None]:
        g = upstream_grad if upstream_grad.shape == self.out_shape else upstream_grad.reshape(self.out_shape)
        xr = self.x if self.x.shape == self.p1 else self.x.reshape(self.p1)
        yr = self.y if self.y.shape == self.p2 else self.y.reshape(self.p2)
        dev = Buffer._route(g, xr, yr)
        if dev is Device.CUDA:
            # CUDA: dX = g @ Y^T and dY = X^T @ g read the stored operands transposed inside the
            # kernel; the reference materialises both transposes first (dlgrad/ops.py:319-327).
            # When an operand is itself a deferred transpose s^T (q @ k.transpose(2, 3)), the gradient is
            # produced for the STORED matrix s and handed on as a deferred transpose, which the Transpose
            # node's backward cancels: ds = (dX)^T = Y @ g^T resp. (dY)^T = g^T @ X.
            from dlgrad_b200.helpers import BinaryOps

            def mm(a, b, shape, **kw):
                return Buffer(dispatcher.dispatch(op=BinaryOps.MATMUL, device=dev, x=a, y=b, **kw), shape, dev, g.dtype)

            def lazy_t(b):
                return Buffer._deferred(("tr", b), b.shape[:-2] + (b.shape[-1], b.shape[-2]), b.dtype)
            batch = self.out_shape[:-2]
            gx = gy = None
            if self.req_grad[0]:
                if xr._pending is not None and xr._pending[0] == "tr":
                    gx = lazy_t(mm(yr, g, batch + (self.p1[-1], self.p1[-2]), trans_y=True))
                else:
                    gx = None
                    from dlgrad_b200.runtime.cuda import attention as _att
                    if yr._pending is not None and yr._pending[0] == "tr":
                        # dq = dS @ k with k (and dq) in the projections' (B, T, h, D) layout: no permute on either side
                        gx = _att.ds_gemm(g, yr._pending[1], False)
                    if gx is None and _att._stored(yr) and _att._stored(g):
                        # dP = dO @ v^T of an attention block: deferred, so that the softmax / masked_fill / scale
                        # backward that follow can absorb it into one kernel (runtime/cuda/attention.dscores)
                        from dlgrad_b200.runtime.cuda import attention
                        yt = lazy_t(yr)
                        if attention.BWD_ENABLED and attention.eligible(g, yt):
                            gx = Buffer._deferred(("mm", g, yt, self.out), batch + (self.p1[-2], self.p1[-1]), g.dtype)
                    if gx is None:
                        gx = mm(g, yr, batch + (self.p1[-2], self.p1[-1]), trans_y=True)
                gx = unbroadcast(gx, self.x.shape)
            if self.req_grad[1]:
                rec = None
                if xr._pending is not None or yr._pending is None:
                    from dlgrad_b200.runtime.cuda import attention
                    rec = attention.flash_record(xr, yr)
                if rec is not None and attention._stored(g) and g.aligned:
                    # x is the attention matrix P that the flash forward consumed without writing it: dv = P^T @ dO with
                    # P recomputed per tile from q, k and the saved row statistics
                    gy = attention.flash_dv(rec, g)
                elif yr._pending is not None and yr._pending[0] == "tr":
                    from dlgrad_b200.runtime.cuda import attention as _att2
                    dk = _att2.ds_gemm(g, xr, True)              # dk = dS^T @ q, layouts as for dq above
                    gy = lazy_t(dk if dk is not None else mm(g, xr, batch + (self.p2[-1], self.p2[-2]), trans_x=True))
                else:
                    gy = mm(xr, g, batch + (self.p2[-2], self.p2[-1]), trans_x=True)
                gy = unbroadcast(gy, self.y.shape)
            return gx, gy
        nd = xr.ndim
        gx = unbroadcast(g @ yr.transpose(nd - 2, nd - 1), self.x.shape)
        gy = unbroadcast(xr.transpose(nd - 2, nd - 1) @ g, self.y.shape)
        return gx, gy


class Linear(OP):
    """`x @ weight.T (+ bias)` as one node on CUDA (Tensor.linear, dlgrad/tensor.py:520-535, builds it from
    a materialised transpose, a batched matmul whose second operand is broadcast, and an add).  Here the
    stored (out, in) weight is read K-major by the tensor-core kernel, so no transpose is written; in the
    backward dX = g @ W, and dW = g^T @ x runs as ONE GEMM over the flattened (batch*tokens) axis instead
    of per-batch products followed by a sum over the batch and a transpose back."""
    def forward(self, x: Buffer, weight: Buffer, bias: Buffer | None = None) -> Buffer:
        from dlgrad_b200.helpers import BinaryOps
        self.x, self.w, self.b = x, weight, bias
        # the input is a deferred relu (ops.Relu): its backward is a select on the pre-activation, which the dX GEMM of
        # this Linear can apply in its epilogue
        self.relu_src = x._pending[1] if (x._pending is not None and x._pending[0] == "relu") else None
        out_f, in_f = weight.shape
        rows = x.numel // in_f
        x2 = x.reshape((rows, in_f))
        if bias is not None and bias.numel == out_f:
            # the bias row rides on the GEMM epilogue when the shape runs on the tensor-core kernel
            y, applied = dispatcher.dispatch(op=BinaryOps.MATMUL, device=Device.CUDA, x=x2, y=weight, trans_y=True, bias=bias)
            out = Buffer(y, x.shape[:-1] + (out_f,), Device.CUDA, x.dtype)
            return out if applied else out + bias
        if bias is None and LAZY_LINEAR and rows * out_f >= (1 << 20) and weight._pending is None:
            # deferred (Buffer pending kind "lin"): when the consumer is an addition — the residual connection
            # `x + proj(h)`, examples/gpt.py:88-89 — the GEMM runs with the add in its epilogue (matmul_tc_ts `radd`) and
            # the three-pass ADD kernel disappears; every other consumer materialises it through `.ptr`
            # ... and when sibling layers read the same input — the q / k / v projections, examples/gpt.py:43-45 — the
            # first one that is needed runs all of them as ONE grouped launch (runtime/cuda/matmul.LinSlot)
            from dlgrad_b200.runtime.cuda import matmul as _mm
            return Buffer._deferred(("lin", x2, weight, _mm.lin_slot(x2, weight)), x.shape[:-1] + (out_f,), x.dtype)
        y = dispatcher.dispatch(op=BinaryOps.MATMUL, device=Device.CUDA, x=x2, y=weight, trans_y=True)
        out = Buffer(y, x.shape[:-1] + (out_f,), Device.CUDA, x.dtype)
        return out if bias is None else out + bias

    def backward(self, upstream_grad: Buffer) -> tuple:
        from dlgrad_b200.helpers import BinaryOps
        out_f, in_f = self.w.shape
        rows = self.x.numel // in_f
        g2 = upstream_grad.reshape((rows, out_f))
        gx = gw = gb = None
        if self.req_grad[0] and self.relu_src is not None and RELU_BWD_EPILOGUE and self.relu_src._pending is None:
            ptr, applied = dispatcher.dispatch(op=BinaryOps.MATMUL, device=Device.CUDA, x=g2, y=self.w,
                                               relu_mask=self.relu_src.reshape((rows, in_f)))
            gx = Buffer(ptr, self.x.shape, Device.CUDA, self.x.dtype)
            if applied:
                _premasked[id(gx)] = id(self.relu_src)
                weakref.finalize(gx, _premasked.pop, id(gx), None)
        elif self.req_grad[0]:
            if LAZY_LINEAR and rows * in_f >= (1 << 20) and self.w._pending is None:
                # deferred ("lin_dx"): the gradient accumulation `grad + g @ W` of the autograd driver (dlgrad/tensor.py:805)
                # then rides on this GEMM's epilogue
                gx = Buffer._deferred(("lin_dx", g2, self.w), self.x.shape, self.x.dtype)
            else:
                gx = Buffer(dispatcher.dispatch(op=BinaryOps.MATMUL, device=Device.CUDA, x=g2, y=self.w),
                            self.x.shape, Device.CUDA, self.x.dtype)
        if self.req_grad[1]:
            # dW = g^T @ x; the weight gradients of sibling layers (dWq, dWk, dWv: same x, split-K shapes) run as one grouped
            # launch (runtime/cuda/matmul.lin_dw)
            from dlgrad_b200.runtime.cuda import matmul as _mm
            gw = _mm.lin_dw(g2, self.x.reshape((rows, in_f)), out_f, in_f, self.w.dtype)
        if self.b is not None and self.req_grad[2]:
            gb = g2.sum(0, keepdim=True).reshape(self.b.shape)
        return (gx, gw, gb) if self.b is not None else (gx, gw)


# ------------------------------------------------------------------ losses
class CrossEntropy(OP):
    def forward(self, logits: Buffer, target: Buffer, dim: int = 1) -> Buffer:
        assert logits.shape[0] == target.shape[0]
        self.target = target
        self.log_softmax = logits.log_softmax(dim)          # kept for backward, (N, V)
        picked = Buffer.ce_forward(self.log_softmax, self.target)
        self.N = float(logits.shape[0])
        # QUIRK: sum, not mean (dlgrad/ops.py:357) — the printed loss is N x the per-row loss
        return -picked.sum()

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        ls = self.log_softmax
        dev = Buffer._route(ls, self.target, upstream_grad)
        g = dispatcher.dispatch(op=CustomOps.CE_GRAD, device=dev, logsm=ls, target=self.target,
                                upstream=upstream_grad, n=self.N)
        return (Buffer(g, ls.shape, dev, ls.dtype),)


class BCEWithLogitsLoss(OP):
    def forward(self, probs: Buffer, target: Buffer) -> Buffer:
        assert probs.shape == target.shape, f"shape mismatch {probs.shape} vs {target.shape}"
        self.probs, self.target = probs, target
        max_val = (-probs).clamp(min=0)
        loss = ((1.0 - target) * probs) + max_val + ((-max_val).exp() + (-(probs + max_val)).exp()).log()
        return loss.mean()

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        n = float(self.probs.numel)
        return (((self.probs.sigmoid() - self.target) * upstream_grad) / n,)


class Embedding(OP):
    def forward(self, weight: Buffer, idx: Buffer) -> Buffer:
        self.weight, self.idx = weight, idx
        return weight.embedding(idx)

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        return (self.weight.embedding(self.idx, backward=True, upstream_grad=upstream_grad),)


# ------------------------------------------------------------------ select
class Where(OP):
    def forward(self, cond: Buffer, inp: Buffer | Scalar, other: Buffer | Scalar) -> Buffer:
        self.cond = cond
        return cond.where(inp, other)

    def backward(self, upstream_grad: Buffer) -> tuple:
        return (None, self.cond.where(upstream_grad, 0.0), self.cond.where(0.0, upstream_grad))


class MaskedFill(OP):
    def forward(self, data: Buffer, mask: Buffer, value: Scalar) -> Buffer:
        self.mask = mask
        return Buffer.masked_fill(x=data, mask=mask, val=value)

    def backward(self, upstream_grad: Buffer) -> tuple:
        return (Buffer.masked_fill(x=upstream_grad, mask=self.mask, val=0.0), None, None)


class Tril(OP):
    def forward(self, x: Buffer, k: Buffer | int = 0) -> Buffer:
        # 2-D only, like the reference (dlgrad/ops.py:110-126): mask = cols <= rows + k
        h, w = x.shape[0], x.shape[1]
        rows = Buffer.arange((h, 1), x.device)
        cols = Buffer.arange((1, w), x.device)
        self.mask = cols <= (rows + k)
        return self.mask.where(x, 0.0)

    def backward(self, upstream_grad: Buffer) -> tuple:
        return (self.mask.where(upstream_grad, 0.0), None)


class Dropout(OP):
    def forward(self, x: Buffer, p: Scalar) -> Buffer:
        self.p = p
        if p == 1.0:
            return Buffer.full(x.shape, 0.0, x.device, x.dtype)
        if p == 0.0:
            return x
        self.mask = Buffer.uniform(x.shape, x.device, x.dtype, low=0.0, high=1.0) >= p
        self.scale = 1.0 / (1.0 - p)
        return self.mask.where(x, 0.0) * self.scale

    def backward(self, upstream_grad: Buffer) -> tuple[Buffer]:
        if self.p == 1.0:
            return (Buffer.full(upstream_grad.shape, 0.0, upstream_grad.device, upstream_grad.dtype),)
        if self.p == 0.0:
            return (upstream_grad,)
        return (self.mask.where(upstream_grad, 0.0) * self.scale,)
