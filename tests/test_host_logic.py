"""Host-side logic that needs no GPU: shape rules (dlgrad/helpers.py), the dispatcher contract
(dlgrad/dispatch.py:19-34), Tensor argument checking (dlgrad/tensor.py:100-123) and the matmul planner's
geometry / tile / split-K decisions."""
import numpy as np
import pytest

from dlgrad_b200 import Tensor
from dlgrad_b200.codegen import cuda_tcgen05 as tc
from dlgrad_b200.device import Device
from dlgrad_b200.dispatch import Dispatcher
from dlgrad_b200.helpers import (BinaryOps, broadcast_shapes, cal_sum_max_out_shape, calculate_stride, check_broadcast,
                                 get_broadcast_shape, get_broadcast_strides_2)


def test_strides_and_broadcast_shapes():
    assert calculate_stride((2, 3, 4)) == (12, 4, 1)
    assert calculate_stride(()) == ()
    assert get_broadcast_shape((2, 1, 4), (3, 1)) == (2, 3, 4)
    assert get_broadcast_strides_2((2, 3, 4), (3, 1), (1, 1)) == (0, 1, 0)
    with pytest.raises(ValueError):
        get_broadcast_shape((2, 3), (4, 3))
    with pytest.raises(AssertionError, match="Cannot broadcast"):
        check_broadcast((2, 3), (4, 3))
    assert check_broadcast((2, 3), (1, 3))


def test_matmul_rank_alignment_and_errors():
    p1, p2, out = broadcast_shapes((4, 5), (2, 3, 5, 6))
    assert (p1, p2, out) == ((1, 1, 4, 5), (2, 3, 5, 6), (2, 3, 4, 6))
    with pytest.raises(ValueError, match="inner dims"):
        broadcast_shapes((4, 5), (6, 7))
    with pytest.raises(ValueError, match="batch dims"):
        broadcast_shapes((2, 4, 5), (3, 5, 6))


def test_reduce_out_shape_quirk():
    # dim == -1 means "all"; with keepdim the reference returns the INPUT shape there
    assert cal_sum_max_out_shape(3, -1, (2, 3, 4)) == ()
    assert cal_sum_max_out_shape(3, -1, (2, 3, 4), keepdim=True) == (2, 3, 4)
    assert cal_sum_max_out_shape(3, 1, (2, 3, 4)) == (2, 4)
    assert cal_sum_max_out_shape(3, 1, (2, 3, 4), keepdim=True) == (2, 1, 4)


def test_dispatcher_contract():
    d = Dispatcher()
    d.register(BinaryOps.ADD, Device.CPU)(lambda x, y: ("first", x + y))     # (the reference's decorator returns None;
    # this one hands the function back so that runtime ops can call each other by name)
    assert d.dispatch(op=BinaryOps.ADD, device=Device.CPU, x=1, y=2) == ("first", 3)
    d.register(BinaryOps.ADD, Device.CPU)(lambda x, y: ("second", x + y))
    assert d.dispatch(op=BinaryOps.ADD, device=Device.CPU, x=1, y=2)[0] == "second"      # last registration wins
    assert d.has(BinaryOps.ADD, Device.CPU) and not d.has(BinaryOps.MUL, Device.CPU)
    with pytest.raises(Exception):
        d.dispatch(op=BinaryOps.MUL, device=Device.CPU, x=1, y=2)


def test_tensor_argument_checks():
    with pytest.raises(ValueError, match="float32"):
        Tensor(np.zeros((2, 2), dtype=np.float64))
    with pytest.raises(ValueError, match="4D"):
        Tensor(np.zeros((1, 1, 1, 1, 2), dtype=np.float32))
    with pytest.raises(ValueError):
        Tensor("nope")
    t = Tensor(np.arange(6, dtype=np.float32).reshape(2, 3))
    assert t.shape == (2, 3) and t.ndim == 2 and t.numel == 6 and t.nbytes == 24
    with pytest.raises(ValueError, match="Buffer size mismatch"):
        t.copy_from(b"\0" * 8)
    with pytest.raises(ValueError, match="Cannot reshape"):
        t.data.reshape((4, 2))
    assert t.data.reshape((3, -1)).shape == (3, 2)
    np.testing.assert_array_equal(Tensor(np.ascontiguousarray(np.arange(6, dtype=np.float32).reshape(2, 3).T)).numpy(),
                                  np.arange(6, dtype=np.float32).reshape(2, 3).T)


def test_host_scalar_and_views_need_no_runtime():
    s = Tensor(2.5)
    assert s.shape == () and float(s.numpy()) == 2.5
    t = Tensor(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    b = t.data
    b.unsqueeze(0)
    assert b.shape == (1, 2, 3, 4)
    b.squeeze(0)
    assert b.shape == (2, 3, 4)
    with pytest.raises(AssertionError):
        b.unsqueeze([1, 1])


@pytest.mark.parametrize("M,N,K,nb", [(4096, 4096, 4096, 1), (8192, 768, 768, 1), (8192, 96, 768, 1), (1024, 64, 1024, 96),
                                      (200, 96, 72, 1), (64, 1024, 1024, 96), (8192, 768, 96, 1), (2048, 128, 128, 1)])
def test_tile_width_choice_is_legal(M, N, K, nb):
    kblocks = -(-K // 32)
    bn = tc.pick_bn_ts(M, N, nb, kblocks)
    assert bn in (32, 64, 96, 128, 192, 256) and bn % 16 == 0
    assert bn <= max(32, -(-N // 32) * 32)
    k = tc.matmul_tc_ts(M, N, K, nb, False, True, False, False)
    assert k.bn == bn and 2 <= k.stages <= 6 and k.smem <= 227 * 1024 and k.grid[0] <= 148
    # TMEM budget: accumulator buffers + 64 columns per ring stage never exceed the 512 columns of an SM
    acc = 2 * bn if 2 * bn + 64 * min(k.stages, 3) <= 512 else bn
    assert acc + 64 * k.stages <= 512
    assert k.out_elems == nb * M * N


def test_matmul_geometry_and_split_k():
    from dlgrad_b200.runtime.cuda import matmul as mm
    M, N, K, batch, a_bs, b_bs, nd = mm._geometry((8, 12, 1024, 64), (8, 12, 64, 1024), False, False)
    assert (M, N, K, batch) == (1024, 1024, 64, (8, 12)) and a_bs == (12 * 1024 * 64, 1024 * 64)
    M, N, K, batch, a_bs, b_bs, nd = mm._geometry((8192, 768), (8192, 3072), True, False)       # x^T @ g
    assert (M, N, K) == (768, 3072, 8192) and batch == (1, 1)
    M, N, K, batch, a_bs, b_bs, nd = mm._geometry((4, 5, 6), (1, 6, 7), False, False)           # broadcast batch
    assert batch == (1, 4) and b_bs == (0, 0)
    with pytest.raises(ValueError):
        mm._geometry((4, 5), (6, 7), False, False)
    assert mm._split_k(768, 768, 8192, (1, 1), True, False) == 4          # Linear dW: 36 tiles -> 4 K-slices
    assert mm._split_k(768, 3072, 8192, (1, 1), True, False) == 1         # 144 tiles already fill the GPU
    assert mm._split_k(768, 768, 8192, (1, 1), False, False) == 1         # only the TN form has K as a batch axis
    assert mm._split_k(768, 768, 512, (1, 1), True, False) == 1


def test_stream_k_is_opt_in(monkeypatch):
    monkeypatch.delenv("DLGRAD_TC_STREAMK", raising=False)
    assert tc.wants_streamk(8192, 768, 768, 1, 128) is False
    monkeypatch.setenv("DLGRAD_TC_STREAMK", "1")
    assert tc.wants_streamk(8192, 768, 768, 1, 128) is True               # 2.59 waves of tiles
    assert tc.wants_streamk(8192, 3072, 768, 1, 192) is False             # 6.92 waves: nothing to recover
    assert tc.wants_streamk(1024, 64, 1024, 96, 64) is False              # batched: measured slower
    k = tc.matmul_tc_ts(8192, 768, 768, 1, False, True, False, False, bn=128, streamk=True)
    assert "flags + t * 4u + q" in k.source and k.name.endswith("_sk")


def test_every_kernel_gets_the_dependent_launch_prologue():
    from dlgrad_b200.codegen import cuda_kernel as ck
    from dlgrad_b200.runtime.cuda import compiler
    for k in (ck.ewise("v0 + v1", (64, 64), ((64, 1), (64, 1)), True, "add"), tc.matmul_tc_ts(256, 256, 128, 1, False, True, False, False)):
        src = compiler.with_pdl_prologue(k.source)
        body = src[src.index('extern "C" __global__'):]
        first = body[body.index("{") + 1:]
        assert first.lstrip().startswith("//") and "griddepcontrol.launch_dependents" in first[:400]
        assert first.index("griddepcontrol.wait") < first.index("blockIdx") if "blockIdx" in first else True


def _cuda_buf(shape):
    from dlgrad_b200.buffer import Buffer
    return Buffer(None, shape, Device.CUDA)


def test_fused_attention_eligibility_and_chain_matching():
    """runtime/cuda/attention.py decides on the host which q @ k^T products are deferred and which pending chains the
    fused kernels may absorb; everything else must fall back to the eager kernels."""
    from dlgrad_b200.buffer import Buffer
    from dlgrad_b200.runtime.cuda import attention
    q, k = _cuda_buf((8, 12, 1024, 64)), _cuda_buf((8, 12, 1024, 64))
    kt = Buffer._deferred(("tr", k), (8, 12, 64, 1024))
    assert attention.eligible(q, kt)
    assert not attention.eligible(q, _cuda_buf((8, 12, 64, 1024)))                    # k^T materialised: plain GEMM
    assert not attention.eligible(_cuda_buf((8, 12, 1000, 64)), Buffer._deferred(("tr", _cuda_buf((8, 12, 1000, 64))), (8, 12, 64, 1000)))
    assert not attention.eligible(_cuda_buf((8, 12, 1024, 128)), Buffer._deferred(("tr", _cuda_buf((8, 12, 1024, 128))), (8, 12, 128, 1024)))
    assert not attention.eligible(_cuda_buf((2, 2, 128, 64)), Buffer._deferred(("tr", _cuda_buf((2, 2, 128, 64))), (2, 2, 64, 128)))  # too small
    assert not attention.eligible(_cuda_buf((4, 12, 1024, 64)), kt)                   # batch dims differ
    mask = _cuda_buf((1024, 1024))
    s = _cuda_buf((8, 12, 1024, 1024))
    fwd = [("scale", s, 0.125), ("mfill", s, mask, float("-inf"))]
    assert attention.chain_matches(fwd, (8, 12, 1024, 1024)) == (0.125, mask, float("-inf"))
    assert attention.chain_matches(fwd[::-1], (8, 12, 1024, 1024)) is None            # masked_fill before the scale
    assert attention.chain_matches([("scale", s, -1.0), fwd[1]], (8, 12, 1024, 1024)) is None      # row max taken on raw scores
    assert attention.chain_matches([fwd[0], ("mfill", s, _cuda_buf((12, 1024, 1024)), 0.0)], (8, 12, 1024, 1024)) is None  # per-head mask
    assert attention.chain_matches([fwd[0], ("mfill", s, _cuda_buf((1, 1, 1024, 1024)), -1e9)], (8, 12, 1024, 1024)) is not None
    bwd = [("mfill", s, mask, 0.0), ("scale", s, 0.125)]
    assert attention.bwd_chain_matches(bwd, (8, 12, 1024, 1024)) == (mask, 0.125)
    assert attention.bwd_chain_matches([("mfill", s, mask, 1.0), bwd[1]], (8, 12, 1024, 1024)) is None
    assert attention.mask_words(1024, 1024) == 1024 * 32 + 64 + 1024 + 1
    # a Buffer's live-tile table is found by identity (Buffer.__eq__ is the elementwise op) and dies with the Buffer
    p = _cuda_buf((8, 12, 1024, 1024))
    attention._tags[id(p)] = ("owner", "flags", 1024, 1024)
    assert attention.tag_of(p) == ("owner", "flags", 1024, 1024) and attention.tag_of(_cuda_buf((8, 12, 1024, 1024))) is None
    attention._tags.pop(id(p))


def test_fused_attention_kernels_generate_for_every_supported_head_dim():
    for D in (32, 64):
        kf = tc.attn_scores_tc(512, 1024, D, 6, 0.25, float("-inf"))
        kb = tc.attn_dscores_tc(512, 1024, D, 6, 0.25)
        for k in (kf, kb):
            assert k.threads == 448 and k.smem <= 227 * 1024 and k.grid[0] <= 148 and k.out_elems == 6 * 512 * 1024
        assert "live_tiles" in kf.source and "cp.async.cg.shared.global" in kb.source
    assert "return 255u;" in tc.attn_scores_tc(1024, 1024, 64, 4, 0.125, -1e9).source     # finite fill: no tile is skipped
    with pytest.raises(AssertionError):
        tc.attn_scores_tc(1024, 1024, 128, 4, 0.125, float("-inf"))                         # TMEM cannot hold two Q buffers
    ks = tc.matmul_tc_ts(1024, 64, 1024, 96, True, True, False, False, bn=64, ksparse="tn")
    assert ks.name.endswith("_sptn") and "last_kb" in ks.source


def test_philox_reference_matches_random123_known_answers():
    """tests/philox_ref.py — the numpy checker of the device generator — against Random123's kat_vectors for
    philox4x32 with 10 rounds (counter, key -> output)."""
    from tests import philox_ref
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for counter, key, want in kats:
        assert tuple(int(x) for x in philox_ref.philox4x32_10(counter, key)) == want
    u = philox_ref.uniform(1001, seed=7, call=3, low=-1.0, high=1.0)
    assert u.shape == (1001,) and u.dtype == np.float32 and (u >= -1.0).all() and (u < 1.0).all()
    assert not np.array_equal(u, philox_ref.uniform(1001, seed=7, call=4, low=-1.0, high=1.0))


@pytest.mark.parametrize("batch,ntile,grid", [(96, 8, 148), (8, 8, 64), (3, 5, 15), (200, 2, 148), (1, 8, 8)])
def test_balanced_work_order_covers_every_item_once_and_evens_out_a_causal_mask(batch, ntile, grid):
    """The persistent attention kernels and the tile-skipping GEMMs walk (head, tile) items in an order that spreads the
    cost of a causal mask (tile i has i + 1 live key tiles) over the CTAs; round-robin leaves 36 vs 13 units at C5."""
    items = batch * ntile
    per_cta = [tc.balanced_items(c, grid, items, batch, ntile) for c in range(grid)]
    assert sorted(w for lst in per_cta for w in lst) == list(range(items))
    for cost in (lambda w: w % ntile + 1, lambda w: ntile - w % ntile):
        loads = [sum(cost(w) for w in lst) for lst in per_cta]
        mean = sum(loads) / grid
        assert max(loads) <= mean + ntile, (max(loads), mean)
    if (batch, ntile, grid) == (96, 8, 148):
        rr = [sum((w % 8) + 1 for w in range(c, items, grid)) for c in range(grid)]
        assert max(rr) == 36 and max(sum(w % 8 + 1 for w in lst) for lst in per_cta) <= 24


def test_constant_folding_capture_and_rng_host_invariants():
    """tests/_host_capture_script.py with the stand-in library (no GPU): lazy FULL / ARANGE and shared copy-on-write
    storage of constant masks, the guards and pre-replay hooks of graph.CapturedStep (Adam's step count, .grad
    re-binding), and the call-number bookkeeping of the device RNG."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_host_capture_script.py")], cwd=root,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "all host checks passed" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_one_mma_issuing_warp_by_default(monkeypatch):
    """Two MMA-issuing warps per CTA are faster on narrow tiles but make the accumulation order, hence the bits of a
    run, vary (profiles/r01_tc_pipeline_findings.md): the default must stay at one issuer, two only on request."""
    monkeypatch.delenv("DLGRAD_TC_ISSUERS", raising=False)
    narrow = tc.matmul_tc_ts(1024, 64, 1024, 96, False, False, False, False)
    assert narrow.bn <= 96 and narrow.threads == 448 and "warp == 14" not in narrow.source
    two = tc.matmul_tc_ts(1024, 64, 1024, 96, False, False, False, False, issuers=2)
    assert two.threads == 480 and "warp == 14" in two.source


def test_sample_last_on_the_host_runtime_is_the_inverse_cdf(oracle):
    """Tensor.sample_last on a host runtime (the oracle): inverse CDF of softmax(x[..., -1, :] / temperature) with the
    supplied uniform numbers — the rule the CUDA kernel implements (tests/test_parity_r2.py checks it on the device)."""
    import numpy as np
    from dlgrad_b200 import Tensor
    x = np.zeros((2, 3, 4), dtype=np.float32)
    x[0, -1] = [0.0, np.log(2.0), np.log(3.0), np.log(4.0)]          # probabilities 0.1 0.2 0.3 0.4
    x[1, -1] = [5.0, -50.0, -50.0, -50.0]
    for u, want in ((0.05, 0), (0.15, 1), (0.45, 2), (0.75, 3), (0.999999, 3)):
        got = Tensor(x.copy()).sample_last(1.0, u=Tensor(np.array([u, 0.5], dtype=np.float32))).numpy().reshape(-1)
        assert got.tolist() == [float(want), 0.0], (u, got)
