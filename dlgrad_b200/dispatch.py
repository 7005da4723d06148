"""(op, device) -> runtime function table.  Same contract as dlgrad/dispatch.py:7-57: runtimes
register plain functions under an (OpEnum, Device) key and Buffer code calls
dispatcher.dispatch(op=..., device=..., **kwargs); the function returns a raw `float *` cdata.

This package registers Device.CUDA entries only (dlgrad_b200/runtime/cuda_ops.py).  There is no
CPU runtime in the product: dispatching to a key that is absent raises, it never falls back.
(The test-suite's oracle registers Device.CPU entries from oracle/backend.py; nothing under
dlgrad_b200/ imports it.)
"""
from __future__ import annotations

from collections.abc import Callable
from enum import Enum


class Dispatcher:
    def __init__(self) -> None:
        self._dispatch_table: dict[tuple[Enum, Enum], Callable] = {}

    def register(self, op: Enum, device: Enum) -> Callable:
        def decorator(func: Callable) -> Callable:
            self._dispatch_table[(op, device)] = func
            return func   # the reference returns None here (dispatch.py:32-34); keeping the name usable is harmless
        return decorator

    def has(self, op: Enum, device: Enum) -> bool:
        return (op, device) in self._dispatch_table

    def has_device(self, device: Enum) -> bool:
        return any(k[1] is device for k in self._dispatch_table)

    def dispatch(self, op: Enum, device: Enum, *args, **kwargs):
        try:
            fn = self._dispatch_table[(op, device)]
        except KeyError:
            raise NotImplementedError(
                f"no runtime registered for {op} on {device}; dlgrad_b200 computes on Device.CUDA only"
            ) from None
        return fn(*args, **kwargs)


dispatcher = Dispatcher()
