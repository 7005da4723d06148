"""Tensor-core (tcgen05 + TMEM + TMA, 3xTF32) matmul plans.  Filled in by codegen/cuda_tcgen05.py;
until a shape is supported `try_build` returns None and the SIMT kernel runs."""
from __future__ import annotations


def try_build(M, N, K, batch, a_bs, b_bs, tx, ty, aligned):
    return None
