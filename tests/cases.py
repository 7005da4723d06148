"""Op-level parity cases shared by the golden generator (run on the real reference) and the tests
(run on the oracle and on Device.CUDA).  Shapes follow the reference's own tests — all one-sided
broadcast patterns of (2,3,4,5) (test/test_ops.py:29-75), per-axis reductions (:109-153), 2-/3-/4-D
matmul with batch broadcast (:77-91,319-345), every-axis softmax — plus the GPT shapes of SURVEY
appendix A.2 at reduced size."""
from __future__ import annotations

import numpy as np


def op_cases(rng) -> dict:
    """name -> (callable on Tensors, list of numpy inputs, differentiate?)"""
    r = lambda *s: rng.random(s, dtype=np.float32)                      # noqa: E731
    rs = lambda *s: (rng.random(s, dtype=np.float32) * 4 - 2).astype(np.float32)   # noqa: E731
    cases = {}
    base = (2, 3, 4, 5)
    others = {"same": base, "b0": (1, 3, 4, 5), "b1": (2, 1, 4, 5), "b2": (2, 3, 1, 5), "b3": (2, 3, 4, 1),
              "r1": (5,), "r2": (4, 5), "r3": (3, 4, 5), "one": (1,)}
    for oname, fn in (("add", lambda a, b: a + b), ("sub", lambda a, b: a - b), ("mul", lambda a, b: a * b),
                      ("div", lambda a, b: a / b)):
        for k, shp in others.items():
            cases[f"{oname}_{k}"] = (fn, [rs(*base), r(*shp) + 0.5], True)
        cases[f"{oname}_scalar"] = (lambda a, _f=fn: _f(a, 1.7), [rs(*base)], True)
    cases["sub_rev"] = (lambda a, b: a - b, [r(4, 5), rs(2, 3, 4, 5)], True)
    cases["div_rev"] = (lambda a, b: a / b, [r(4, 5), r(2, 3, 4, 5) + 0.5], True)
    cases["add_rowvec"] = (lambda a, b: a + b, [rs(64, 48), rs(1, 48)], True)
    cases["mul_colvec"] = (lambda a, b: a * b, [rs(64, 48), rs(64, 1)], True)
    for k, fn in (("neg", lambda a: -a), ("exp", lambda a: a.exp()), ("relu", lambda a: a.relu()),
                  ("leaky_relu", lambda a: a.leaky_relu(0.2)), ("sigmoid", lambda a: a.sigmoid()),
                  ("tanh", lambda a: a.tanh()), ("pow2", lambda a: a ** 2), ("pow3", lambda a: a ** 3),
                  ("pow_half", lambda a: a ** 0.5), ("clamp", lambda a: a.clamp(-0.5, 0.7)),
                  ("clamp_min", lambda a: a.clamp(min=0))):
        cases[k] = (fn, [rs(6, 37)], k not in ("neg", "clamp", "clamp_min"))
    for k, fn in (("log", lambda a: a.log()), ("sqrt", lambda a: a.sqrt()), ("rsqrt", lambda a: a.rsqrt()),
                  ("pow_m1", lambda a: a ** -1)):
        cases[k] = (fn, [r(6, 37) + 0.1], True)
    for shp in ((2, 3, 4, 5), (64, 48), (131, 70)):
        tag = "x".join(map(str, shp))
        for d in list(range(len(shp))) + [-1]:
            for keep in (False, True):
                if d == -1 and keep:
                    continue          # reference labels a 1-element result with the input shape (quirk 12)
                for oname in ("sum", "mean", "max"):
                    cases[f"{oname}_{tag}_d{d}_k{int(keep)}"] = (
                        lambda a, _o=oname, _d=d, _k=keep: getattr(a, _o)(dim=_d, keepdim=_k), [rs(*shp)],
                        len(shp) == 4 or (shp == (64, 48) and not keep))
    for order in ((1, 0),):
        cases["permute_2d"] = (lambda a, _o=order: a.permute(_o), [rs(33, 70)], True)
    for order in ((0, 2, 1), (2, 0, 1), (1, 0, 2)):
        cases[f"permute_3d_{''.join(map(str, order))}"] = (lambda a, _o=order: a.permute(_o), [rs(5, 33, 36)], True)
    for order in ((0, 2, 1, 3), (0, 1, 3, 2), (3, 2, 1, 0), (1, 0, 2, 3)):
        cases[f"permute_4d_{''.join(map(str, order))}"] = (lambda a, _o=order: a.permute(_o), [rs(2, 5, 12, 36)], True)
    mm = {"2d": ((33, 20), (20, 17)), "2d_k1": ((16, 1), (1, 8)), "2d_n1": ((100, 256), (256, 1)),
          "3d": ((4, 33, 20), (4, 20, 17)), "3d_bx": ((1, 33, 20), (4, 20, 17)), "3d_by": ((4, 33, 20), (1, 20, 17)),
          "3d_2d": ((4, 33, 20), (20, 17)), "4d": ((2, 3, 16, 8), (2, 3, 8, 16)),
          "4d_b": ((2, 3, 16, 8), (1, 1, 8, 16)), "gpt_qk": ((2, 2, 16, 8), (2, 2, 8, 16)),
          "tc_128": ((128, 64), (64, 128)), "tc_batched": ((2, 128, 32), (1, 32, 64))}
    for k, (s1, s2) in mm.items():
        cases[f"matmul_{k}"] = (lambda a, b: a @ b, [rs(*s1), rs(*s2)], True)
    for d in range(3):
        cases[f"softmax_d{d}"] = (lambda a, _d=d: a.softmax(dim=_d), [rs(4, 7, 12)], True)
        cases[f"log_softmax_d{d}"] = (lambda a, _d=d: a.log_softmax(dim=_d), [rs(4, 7, 12)], True)
    cases["softmax_rows128"] = (lambda a: a.softmax(dim=3), [rs(2, 2, 16, 128)], True)
    return cases


