"""Flash-style attention in 3xTF32 on tcgen05 / TMEM / TMA: the probabilities P = softmax(mask(q k^T * scale)) of
examples/gpt.py:49-55 never reach HBM, neither forward nor backward.

The reference executes `att = (q @ k^T) * scale; att = masked_fill(att, mask, fill); att = softmax(att); y = att @ v`
as MATMUL, MUL, MASKED_FILL, five softmax kernels and MATMUL over (B, h, T, T) tensors (dlgrad/tensor.py:634-649).
Round 1 kept S = q k^T and dP on chip but still wrote P and dS (403 MB each per layer at C5) and re-read them in four
GEMMs: about 2.3 GB of HBM traffic per layer against 0.1 GB of algorithmic bytes (q, k, v, O, dO in; O, dq, dk, dv out).
One generator, three kernels that share a skeleton ("A-side" tile of 128 rows resident in tensor memory, the other
sequence streamed in 64-row blocks through a TMA ring, rows of the score tile owned by threads):

    fwd   A = Q tile.  per key block:  S = Q K^T -> online softmax in registers -> P (hi | lo) written to TMEM by the
          row threads -> O += P V as a TS-form MMA.  Output O (normalised) and the row statistics (m, 1 / l).
    dv    A = K tile.  per query block: S^T = K Q^T -> P^T = 2^(S^T c - m_q) / l_q from the saved statistics ->
          TMEM -> dV += P^T dO.
    ds    A = Q tile and dO tile.  per key block: S = Q K^T and dP = dO V^T -> dS = P (dP - delta) scale, masked -> HBM
          (delta = rowsum(dO * O)).  dS is the one (B, h, T, T) tensor that is still materialised: dq = dS k and
          dk = dS^T q run on the tile-skipping GEMM (cuda_tcgen05.matmul_tc_ts, ksparse).  Folding those two into this
          kernel needs 576 TMEM columns in 3xTF32 (Q, dO, S, dP, dS hi | lo, dQ) — 64 more than an SM has.

Every product is 3xTF32 (a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulate), so the results keep fp32 accuracy
(north star: matmul <= 1e-4, softmax <= 1e-5 relative).  kind::tf32 truncates its inputs, so a staged fp32 tile IS its
own `hi`; only `lo` tiles are computed (splitter warps for the streamed operands, the row threads for A and P).

Warp roles (512 threads, one persistent CTA per SM):
    warp 0       TMA producer: A tile(s) of the next item, L2 prefetch of coming tiles, and per step the first streamed tile (ring 1)
    warp 1, 14   MMA issuers (one elected lane each): warp 1 the score tiles, warp 14 the second product of the step; each
                 owns its accumulators, so every accumulation order is one thread's program order (deterministic)
    warps 2-9    row warps: thread = row of the 128-row score tile = TMEM lane; two warps per lane quarter, each owns 32
                 of the 64 score columns of a step (the row work — TMEM reads, exponentials, hi | lo splits, TMEM writes —
                 paces these kernels, not the tensor pipe: with four row warps the forward ran at 50 % pipe utilisation)
    warps 10-13  splitter: lo = x - trunc_tf32(x) into the twin buffer of a stage (same swizzled layout); 10-11 ring 1, 12-13 ring 2
    warp 15      second TMA producer (ring 2)
"""
from __future__ import annotations

import math
from functools import cache

import numpy as np

from dlgrad_b200.codegen.cuda_tcgen05 import _HELPERS, _HELPERS_TS, NUM_SMS, TcKernel, _f32lit, _idesc
from dlgrad_b200.runtime.cuda.compiler import KNAME

_EXTRA = r"""
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tma_prefetch_4d(const TMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 :: "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
#define TMEM_LD32(addr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]), "=r"((v)[5]), "=r"((v)[6]), "=r"((v)[7]), \
      "=r"((v)[8]), "=r"((v)[9]), "=r"((v)[10]), "=r"((v)[11]), "=r"((v)[12]), "=r"((v)[13]), "=r"((v)[14]), "=r"((v)[15]), \
      "=r"((v)[16]), "=r"((v)[17]), "=r"((v)[18]), "=r"((v)[19]), "=r"((v)[20]), "=r"((v)[21]), "=r"((v)[22]), "=r"((v)[23]), \
      "=r"((v)[24]), "=r"((v)[25]), "=r"((v)[26]), "=r"((v)[27]), "=r"((v)[28]), "=r"((v)[29]), "=r"((v)[30]), "=r"((v)[31]) \
    : "r"(addr))
#define TMEM_ST32(addr, v) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" \
    :: "r"(addr), "r"((v)[0]), "r"((v)[1]), "r"((v)[2]), "r"((v)[3]), "r"((v)[4]), "r"((v)[5]), "r"((v)[6]), "r"((v)[7]), \
      "r"((v)[8]), "r"((v)[9]), "r"((v)[10]), "r"((v)[11]), "r"((v)[12]), "r"((v)[13]), "r"((v)[14]), "r"((v)[15]), \
      "r"((v)[16]), "r"((v)[17]), "r"((v)[18]), "r"((v)[19]), "r"((v)[20]), "r"((v)[21]), "r"((v)[22]), "r"((v)[23]), \
      "r"((v)[24]), "r"((v)[25]), "r"((v)[26]), "r"((v)[27]), "r"((v)[28]), "r"((v)[29]), "r"((v)[30]), "r"((v)[31]) : "memory")
#define TMEM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define TMEM_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define TC_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TC_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
// hi | lo split of 32 fp32 values (hi = the 19 bits kind::tf32 keeps, lo = the rest, exact) straight into tensor memory
__device__ __forceinline__ void split_to_tmem(unsigned t_hi, unsigned t_lo, const float* x) {
    unsigned hi[32], lo[32];
    #pragma unroll
    for (unsigned e = 0; e < 32u; ++e) {
        hi[e] = __float_as_uint(x[e]) & 0xffffe000u;
        lo[e] = __float_as_uint(x[e] - __uint_as_float(hi[e]));
    }
    TMEM_ST32(t_hi, hi);
    TMEM_ST32(t_lo, lo);
}
"""

NEG_INF = "__int_as_float(0xff800000)"
QNAN = "__int_as_float(0x7fc00000)"


def _sub(src: str, **kw) -> str:
    for k, v in kw.items():
        src = src.replace(f"${k}$", str(v))
    assert "$" not in src, src[src.index("$") - 40:src.index("$") + 40]
    return src


@cache
def flash_attn(mode: str, Tq: int, Tk: int, D: int, H: int, NB: int, scale: float, fill: float,
               out_strides: tuple = (0, 0, 0), tau: float = 8.0, delta_split: bool = False, lazy_zero: bool = False) -> TcKernel:
    """mode "fwd" | "dv" | "ds" (module docstring).  Tq, Tk multiples of 128, D in (32, 64), NB = batch * H score planes,
    H heads (a plane z is head z % H of sequence z / H: every operand is addressed through a rank-4 tensor map
    (D, T, H, B), so both (B, h, T, D) and the projection's own (B, T, h, D) layout work without a permute kernel).
    `out_strides` = element strides (row, head, batch) of the direct-store output (O for fwd, dV for dv).
    `delta_split` (ds): delta = rowsum(dO * O) is indexed (batch, row, head) — dO and O stored as (B, T, h, D).
    `lazy_zero` (ds): tiles of dS that the mask kills entirely are left UNWRITTEN when every query row has a live entry —
    the tile-skipping GEMMs that consume dS never read them in that case (same test, same table), and they are 44 % of
    the 403 MB the kernel would otherwise write per layer at C5.  Only for a dS that nobody else can read.
    `tau`: the running maximum of the online softmax is only replaced when it grows by more than `tau` (base-2 units):
    any m gives the same P / l in exact arithmetic, and 2^tau bounds the un-normalised P, so O in tensor memory is
    rescaled almost never instead of almost always (tau = 0 forces the rescale path; the tests use it).

    Kernel parameters
      fwd: tmQ, tmK, tmV, tmO, O, lse (one float per row: log2 of the row's sum of 2^(s c), i.e. P = 2^(s c - lse); NaN for a
           row without a live entry), bits
      dv : tmK, tmQ, tmDO, tmDV, dV, lse, bits
      ds : tmQ, tmDO, tmK, tmV, tmDS, dS, lse, delta, bits            (bits: cuda_kernel.mask_bits of the (Tq, Tk) mask)"""
    assert mode in ("fwd", "dv", "ds") and Tq % 128 == 0 and Tk % 128 == 0 and D in (32, 64) and scale > 0
    import os
    # diagnostic (scripts/flash_trace.py): CTA 0 timestamps the first 48 steps of every role into the unused second half of
    # the lse allocation (NB * TA * 8 bytes are allocated, NB * TA * 4 used)
    trace = os.environ.get("DLGRAD_FLASH_TRACE", "") == mode
    lse_rows = NB * Tq

    def TR(kind: int, idx: str, cond: str = "lane == 0u") -> str:
        return (f"if (blockIdx.x == 0u && ({idx}) < 48u && ({cond})) reinterpret_cast<long long*>(const_cast<float*>(lse) + {lse_rows}u)"
                f"[{kind * 48}u + ({idx})] = clock64();") if trace else ""
    assert max(Tq, Tk) // 128 <= 32
    kbq = D // 32
    nq, nk = Tq // 128, Tk // 128
    nA, nS = (nk, nq) if mode == "dv" else (nq, nk)          # tiles on the resident side / on the streamed side
    TA, TS = (Tk, Tq) if mode == "dv" else (Tq, Tk)
    items = NB * nA
    grid = min(items, NUM_SMS)
    skip = math.isinf(fill) and fill < 0
    tile_b = kbq * 8192                                      # one streamed tile: 64 rows x D fp32
    raw_b = 2 * tile_b
    per_stage = 2 * raw_b                                    # per step: ring 1 stage (B1 raw | B1 lo) + ring 2 stage (B2 raw | B2 lo)
    rstage = 2 * tile_b                                      # one stage of one ring
    a_tiles = 2 if mode == "ds" else 1
    a_bytes = a_tiles * kbq * 16384
    epi_bytes = 8 * 4096 if mode == "ds" else 0              # one 32x32 staging box per row warp for the dS store
    tail_bytes = 256 + 1024                                  # barriers | end-of-item exchange of the row sums (fwd)
    budget = 227 * 1024 - 1024 - tail_bytes - a_bytes - epi_bytes
    stages = max(2, min(3, budget // per_stage))
    smem = a_bytes + stages * per_stage + epi_bytes + 1024 + tail_bytes
    ring2_off = stages * rstage
    if mode == "ds":
        S0, DP0, A0, A1 = 0, 128, 256, 256 + 2 * D
        PHI = PLO = OACC = 0
        assert A1 + 2 * D <= 512
    else:
        S0, PHI, PLO, OACC, A0 = 0, 128, 192, 256, 320
        DP0 = A1 = 0
        assert A0 + 2 * D <= 512
    log2e = np.float32(math.log2(math.e))
    sc2 = _f32lit(float(np.float32(scale) * log2e))
    fl2 = _f32lit(fill if math.isinf(fill) else float(np.float32(fill) * log2e))
    idesc_s = _idesc(128, 64, False, False)                  # S / dP: N = 64 streamed rows, B K-major
    idesc_o = _idesc(128, D, False, True)                    # O / dV: N = D, B MN-major
    sT, sH, sB = out_strides
    # which streamed 128-row tiles are live for resident tile `at`: one word per resident tile, computed ONCE per CTA into
    # shared memory (each role looked it up with global loads at every item boundary before: ~3000 clk of L2 latency under
    # load, at the one point of the pipeline where nothing else can hide it)
    if skip:
        if mode == "dv":
            live_fill = f"""for (unsigned j = 0; j < {nq}u; ++j) lm |= (__ldg(mask + {Tq * Tk // 32}u + j * {nk}u + threadIdx.x) != 0u ? 1u : 0u) << j;"""
        else:
            live_fill = f"""for (unsigned j = 0; j < {nk}u; ++j) lm |= (__ldg(mask + {Tq * Tk // 32}u + threadIdx.x * {nk}u + j) != 0u ? 1u : 0u) << j;"""
    else:
        live_fill = f"lm = {(1 << nS) - 1 if nS < 32 else 0xffffffff}u;"
    live = "return live_tab[at];"
    # heavy tiles first: under a causal mask query tile i has i + 1 live key tiles, key tile i has nq - i live query tiles
    tile_of = f"p / {NB}u" if mode == "dv" else f"{nA - 1}u - p / {NB}u"

    if mode == "ds":
        params = ("const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmA2, const __grid_constant__ TMap tmB1, "
                  "const __grid_constant__ TMap tmB2, const __grid_constant__ TMap tmOut, float* __restrict__ out, "
                  "const float* __restrict__ lse, const float* __restrict__ delta, const unsigned* __restrict__ mask")
    elif mode == "fwd":
        params = ("const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmB1, const __grid_constant__ TMap tmB2, "
                  "const __grid_constant__ TMap tmO, float* __restrict__ out, float* __restrict__ lse, const unsigned* __restrict__ mask")
    else:
        params = ("const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmB1, const __grid_constant__ TMap tmB2, "
                  "const __grid_constant__ TMap tmO, float* __restrict__ out, const float* __restrict__ lse, const unsigned* __restrict__ mask")

    # ------------------------------------------------------------------ producer: streamed tiles of one step
    if mode == "ds":       # K block and V block, both K-major (64 rows x D)
        b2_loads = "\n                        ".join(
            f"tma_load_4d(st + {kk * 8192}u, &tmB2, fb, {kk * 32}, (int)r0, (int)hh, (int)bb);" for kk in range(kbq))
    else:                  # second tile MN-major: boxes of {32 columns, 32 rows}, k-block major
        b2_loads = "\n                        ".join(
            f"tma_load_4d(st + {(kb2 * kbq + c) * 4096}u, &tmB2, fb, {c * 32}, (int)(r0 + {kb2 * 32}u), (int)hh, (int)bb);"
            for kb2 in range(2) for c in range(kbq))
    b1_loads = "\n                        ".join(
        f"tma_load_4d(st + {kk * 8192}u, &tmB1, fb, {kk * 32}, (int)r0, (int)hh, (int)bb);" for kk in range(kbq))
    a_loads = "\n                ".join(
        f"tma_load_4d(smem + {kk * 16384}u, &tmA, qfull, {kk * 32}, (int)a0, (int)hh, (int)bb);" for kk in range(kbq))
    if mode == "ds":
        a_loads += "\n                " + "\n                ".join(
            f"tma_load_4d(smem + {(kbq + kk) * 16384}u, &tmA2, qfull, {kk * 32}, (int)a0, (int)hh, (int)bb);" for kk in range(kbq))

    # L2 prefetch of one streamed 128-row tile, box by box (the boxes of the tensor maps the loads use)
    pf = [f"tma_prefetch_4d(&tmB1, {kk * 32}, (int)(pr0 + {r * 64}u), (int)phh, (int)pbb);" for r in range(2) for kk in range(kbq)]
    if mode == "ds":
        pf += [f"tma_prefetch_4d(&tmB2, {kk * 32}, (int)(pr0 + {r * 64}u), (int)phh, (int)pbb);" for r in range(2) for kk in range(kbq)]
    else:
        pf += [f"tma_prefetch_4d(&tmB2, {c * 32}, (int)(pr0 + {r * 32}u), (int)phh, (int)pbb);" for r in range(4) for c in range(kbq)]
    pf_loads = "\n                ".join(pf)

    # ------------------------------------------------------------------ MMA bodies
    lo_off = tile_b // 16                                   # lo twin follows the raw tile inside a ring stage

    def s_mma(acc_col: str, a_col: int, b_off16: int) -> str:
        """acc[128 x 64] = A (TMEM at a_col, hi | lo per 32-wide k block) x B1/B2 tile (K-major) of stage `sdesc`."""
        return f"""
                        #pragma unroll
                        for (unsigned kk = 0; kk < {kbq}u; ++kk) {{
                            #pragma unroll
                            for (unsigned k = 0; k < 4u; ++k) {{
                                const unsigned a_hi = tmem_base + {a_col}u + kk * 64u + k * 8u, a_lo = a_hi + 32u;
                                const unsigned long long b_hi = sdesc + {b_off16}ull + (unsigned long long)(kk * 512u + k * 2u);
                                umma_tf32_ts({acc_col}, a_lo, b_hi, {idesc_s}u, (kk | k) != 0u);
                                umma_tf32_ts({acc_col}, a_hi, b_hi + {lo_off}ull, {idesc_s}u, 1u);
                                umma_tf32_ts({acc_col}, a_hi, b_hi, {idesc_s}u, 1u);
                            }}
                        }}"""
    o_mma = f"""
                        #pragma unroll
                        for (unsigned kb2 = 0; kb2 < 2u; ++kb2) {{
                            #pragma unroll
                            for (unsigned k = 0; k < 4u; ++k) {{
                                const unsigned a_hi = tmem_base + {PHI}u + kb2 * 32u + k * 8u, a_lo = tmem_base + {PLO}u + kb2 * 32u + k * 8u;
                                const unsigned long long b_hi = odesc + (unsigned long long)(kb2 * {kbq * 256}u + k * 64u);
                                umma_tf32_ts(tmem_base + {OACC}u, a_lo, b_hi, {idesc_o}u, (acc | kb2 | k) != 0u);
                                umma_tf32_ts(tmem_base + {OACC}u, a_hi, b_hi + {lo_off}ull, {idesc_o}u, 1u);
                                umma_tf32_ts(tmem_base + {OACC}u, a_hi, b_hi, {idesc_o}u, 1u);
                            }}
                        }}"""

    src = _HELPERS + _HELPERS_TS + _EXTRA + _sub(r"""
extern "C" __global__ void __launch_bounds__(512, 1)
$KNAME$($PARAMS$)
{
    extern __shared__ unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;        // swizzle atoms need 1024-byte alignment
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    const unsigned ring = smem + $A_BYTES$u;                            // ring 1: stages of (B1 raw | B1 lo); ring 2 behind it: (B2 raw | B2 lo)
    const unsigned epi = ring + $RING_BYTES$u;
    const unsigned bars = epi + $EPI_BYTES$u;
    // two rings, each with its own full / empty / split barriers: the first streamed tile of a step (K for fwd) is released
    // when the score MMAs have read it, the second (V) only when the second product has — with ONE ring both waited for the
    // later of the two, and the three 64 KiB stages could not cover TMA + split + MMA + softmax + MMA (~6000 clk per stage)
    const unsigned kfull = bars, kempty = bars + $STAGES$u * 8u, ksplit = bars + $STAGES$u * 16u;
    const unsigned vfull = bars + $STAGES$u * 24u, vempty = bars + $STAGES$u * 32u, vsplit = bars + $STAGES$u * 40u;
    const unsigned sfull = bars + $STAGES$u * 48u, sempty = sfull + 16u, pready = sempty + 16u, pvdone = pready + 8u;
    const unsigned qfull = pvdone + 8u, qsplit = qfull + 8u, qfree = qsplit + 8u, ofree = qfree + 8u;
    float* xch = reinterpret_cast<float*>(smem_gen + (bars - smem) + 256u);      // [8 row warps][32]
    __shared__ unsigned tmem_slot;
    __shared__ unsigned live_tab[32];
    $TRACE_DECL$
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    // streamed 128-row tiles that have at least one unmasked entry against resident tile `at` (cuda_kernel.mask_bits):
    // a tile the mask kills entirely is never loaded, multiplied or read (only for a -inf fill: P is exactly 0 there)
    auto live_tiles = [&](unsigned at) -> unsigned {
        $LIVE$
    };
    // work items (plane z, resident tile at) in the order of this CTA: heavy tiles first, CTAs taken in alternating
    // direction from round to round (cuda_tcgen05.balanced_items)
    auto item_at = [&](unsigned nn, unsigned& z, unsigned& at) -> bool {
        const unsigned p = nn * $GRID$u + ((nn & 1u) ? ($GRID$u - 1u - blockIdx.x) : blockIdx.x);
        if (p >= $ITEMS$u) return false;
        z = p % $NB$u;
        at = $TILE_OF$;
        return true;
    };

    if (threadIdx.x >= 64u && threadIdx.x < 64u + $NA$u) {
        unsigned lm = 0u;
        $LIVE_FILL$
        live_tab[threadIdx.x - 64u] = lm;
    }
    if (threadIdx.x == 0) {
        for (unsigned s = 0; s < $STAGES$u; ++s) {
            mbar_init(kfull + s * 8u, 1u);
            mbar_init(kempty + s * 8u, 1u);
            mbar_init(ksplit + s * 8u, 2u);
            mbar_init(vfull + s * 8u, 1u);
            mbar_init(vempty + s * 8u, 1u);
            mbar_init(vsplit + s * 8u, 2u);
        }
        for (unsigned b = 0; b < 2u; ++b) {
            mbar_init(sfull + b * 8u, $NISS$u);
            mbar_init(sempty + b * 8u, 8u);
        }
        mbar_init(pready, 8u);
        mbar_init(pvdone, 1u);
        mbar_init(qfull, 1u);
        mbar_init(qsplit, $NSPLIT$u);
        mbar_init(qfree, $NISS$u);
        mbar_init(ofree, $NSTORE$u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    TC_FENCE_BEFORE();
    __syncthreads();
    TC_FENCE_AFTER();
    const unsigned tmem_base = tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        unsigned it = 0, n = 0, z, at;
        auto load_a = [&](unsigned zz, unsigned aa) {
            const unsigned hh = zz % $H$u, bb = zz / $H$u, a0 = aa * 128u;
            if (elect_one()) {
                mbar_expect_tx(qfull, $A_BYTES$u);
                $A_LOADS$
            }
            __syncwarp();
        };
        // L2 prefetch cursor: the streamed tiles are requested from HBM $PFD$ live tiles before the TMA load that brings them
        // into shared memory (a load that misses L2 takes ~3000 clk; three 64 KiB stages cannot cover that, and the ring —
        // not the tensor pipe — paced the kernel: pipe active 52 %)
        unsigned pn = 0, pz = 0, pat = 0, plm = 0, pt = 0;
        bool pvalid = item_at(0u, pz, pat);
        if (pvalid) plm = live_tiles(pat);
        auto pf_next = [&]() {                 // prefetch the cursor's tile (if any) and move it to the next live tile
            while (pvalid && (pt >= $NS$u || !((plm >> pt) & 1u))) {
                if (pt >= $NS$u || (plm >> pt) == 0u) { pvalid = item_at(++pn, pz, pat); plm = pvalid ? live_tiles(pat) : 0u; pt = 0u; }
                else ++pt;
            }
            if (!pvalid) return;
            const unsigned phh = pz % $H$u, pbb = pz / $H$u, pr0 = pt * 128u;
            if (elect_one()) {
                $PF_LOADS$
            }
            __syncwarp();
            ++pt;
        };
        for (unsigned i = 0; i < $PFD$u; ++i) pf_next();
        if (item_at(0u, z, at)) load_a(z, at);
        for (; item_at(n, z, at); ++n) {
            const unsigned hh = z % $H$u, bb = z / $H$u;
            const unsigned lm = live_tiles(at);
            unsigned zn, an;
            bool a_pending = item_at(n + 1u, zn, an);           // the next item's A tile: once this one has left shared memory
            for (unsigned t = 0; t < $NS$u; ++t) {
                if (!((lm >> t) & 1u)) continue;
                pf_next();
                for (unsigned half = 0; half < 2u; ++half, ++it) {
                    if (a_pending && mbar_try(qsplit, n & 1u) && $OFREE_TRY$) { load_a(zn, an); a_pending = false; }
                    const unsigned s = it % $STAGES$u, ph = (it / $STAGES$u) & 1u;
                    mbar_wait(kempty + s * 8u, ph ^ 1u);
                    $TR16$
                    const unsigned st = ring + s * $RSTAGE$u, fb = kfull + s * 8u, r0 = t * 128u + half * 64u;
                    if (elect_one()) {
                        mbar_expect_tx(fb, $TILE_B$u);
                        $B1_LOADS$
                    }
                    __syncwarp();
                    $TR17$
                }
            }
            if (a_pending) { mbar_wait(qsplit, n & 1u); $OFREE_WAIT$ load_a(zn, an); }
        }
    } else if (warp == 1) {
        // ===== first MMA issuer: the score tiles S = A x B1^T (Q K^T; K Q^T for dv) into S[b] =====
        // Two warps issue MMAs, each into accumulators of its own (this one S, the other O / dV resp. dP), so the
        // accumulation order inside every accumulator is the program order of ONE thread: results stay bit-identical from
        // run to run.  (MMAs of two warps into the SAME accumulator are not ordered — cuda_tcgen05.matmul_tc_ts.)  One
        // thread issues an N = 64 MMA every ~50 clk, the tensor pipe needs 32: a single issuer left the pipe half idle.
        unsigned it = 0, g = 0, n = 0, z, at;
        const unsigned long long sdesc0 = make_desc(ring, 1u, 64u, 2u);                  // K-major tile: SBO 1024 B, SWIZZLE_128B
        for (; item_at(n, z, at); ++n) {
            const unsigned nsteps = 2u * __popc(live_tiles(at));
            $TR13$
            mbar_wait(qsplit, n & 1u);                            // A (hi | lo) of this item is in tensor memory
            $TR14$
            TC_FENCE_AFTER();
            if (nsteps == 0u) {
                if (elect_one()) umma_commit(qfree);
                __syncwarp();
                continue;
            }
            for (unsigned i = 0; i < nsteps; ++i, ++it, ++g) {
                const unsigned b = g & 1u, s = it % $STAGES$u, ph = (it / $STAGES$u) & 1u;
                mbar_wait(ksplit + s * 8u, ph);
                $TR0$
                mbar_wait(sempty + b * 8u, ((g >> 1) & 1u) ^ 1u);
                $TR1$
                TC_FENCE_AFTER();
                const unsigned long long sdesc = sdesc0 + (unsigned long long)(s * $PER_STAGE16$u);
                if (elect_one()) {
                    $S_MMA$
                    umma_commit(kempty + s * 8u);                      // the first streamed tile of this step may be refilled
                    umma_commit(sfull + b * 8u);
                    if (i + 1u == nsteps) umma_commit(qfree);          // the last read of A by this item
                }
                __syncwarp();
                $TR2$
            }
        }
    } else if (warp == 14u) {
        // ===== second MMA issuer: O += P V (fwd), dV += P^T dO (dv), dP = dO V^T (ds) =====
        unsigned it = 0, g = 0, n = 0, z, at;
        $ISSUER2$
    } else if (warp < 10u) {
        // ===== row warps: thread = row of the score tile = TMEM lane; warps 2-5 own score columns 0-31 of a step, 6-9 columns 32-63 =====
        const unsigned sq = warp & 3u;                            // the TMEM lane quarter this warp may access
        const unsigned wh = (warp >= 6u) ? 1u : 0u;               // which 32 of a step's 64 score columns this warp owns
        const unsigned lane_base = tmem_base + ((sq * 32u) << 16);
        unsigned g = 0, n = 0, z, at;
        auto split_a = [&](unsigned nn) {
            $TR9$
            mbar_wait(qfull, nn & 1u);                             // the tile has landed in shared memory
            $TR10$
            if (nn != 0u) { mbar_wait(qfree, (nn - 1u) & 1u); TC_FENCE_AFTER(); }    // and the previous item no longer reads A
            $TR11$
            if ($A_KBLOCKS$u < 2u && wh != 0u) return;             // a single k-block: the first warp of the pair splits it
            #pragma unroll 1
            for (unsigned kk = ($A_KBLOCKS$u < 2u ? 0u : wh); kk < $A_KBLOCKS$u; kk += 2u) {
                const unsigned char* arow = smem_gen + kk * 16384u + (sq * 32u + lane) * 128u;
                float x[32];
                #pragma unroll
                for (unsigned j = 0; j < 8u; ++j) {
                    const float4 v4 = *reinterpret_cast<const float4*>(arow + ((j ^ (lane & 7u)) << 4));
                    x[4 * j] = v4.x; x[4 * j + 1] = v4.y; x[4 * j + 2] = v4.z; x[4 * j + 3] = v4.w;
                }
                const unsigned ta = lane_base + $A0$u + kk * 64u;
                split_to_tmem(ta, ta + 32u, x);
            }
            TMEM_WAIT_ST();
            TC_FENCE_BEFORE();
            __syncwarp();
            if (lane == 0) mbar_arrive(qsplit);
            $TR12$
        };
        if (item_at(0u, z, at)) split_a(0u);
        for (; item_at(n, z, at); ++n) {
            const unsigned hh = z % $H$u, bb = z / $H$u;
            const unsigned row = at * 128u + sq * 32u + lane;      // this thread's row on the resident side
            const unsigned lm = live_tiles(at);
            const unsigned nsteps = 2u * __popc(lm);
            unsigned zn, an;
            const bool has_next = item_at(n + 1u, zn, an);
            $ROW_ITEM_BEGIN$
            $PF_DECL$
            if (lm != 0u) {
                const unsigned nc0 = (__ffs(lm) - 1u) * 128u;
                $PF_NEXT$
            }
            unsigned j = 0;
            for (unsigned t = 0; t < $NS$u; ++t) {
                if (!((lm >> t) & 1u)) continue;
                for (unsigned half = 0; half < 2u; ++half, ++j, ++g) {
                    const unsigned b = g & 1u;
                    const unsigned c0 = t * 128u + half * 64u;         // first streamed row (= score column) of this step
                    // the mask words (and, for dv, the row statistics) of a step are fetched one step ahead: a global load
                    // takes longer than the whole TMEM read it used to be issued next to
                    $PF_TAKE$
                    {
                        const unsigned rest = (t + 1u < $NS$u) ? (lm >> (t + 1u)) : 0u;
                        const unsigned nc0 = (half == 0u) ? c0 + 64u : (rest ? (t + __ffs(rest)) * 128u : c0);
                        $PF_NEXT$
                    }
                    mbar_wait(sfull + b * 8u, (g >> 1) & 1u);
                    $TR5$
                    TC_FENCE_AFTER();
                    $ROW_STEP$
                    $TR8$
                }
            }
            if (has_next) split_a(n + 1u);                         // next item's A while the last MMAs of this one run
            $ROW_ITEM_END$
            $TR15$
        }
        $ROW_EXIT$
    } else if (warp < 14u) {
        // ===== splitter warps: lo = x - trunc_tf32(x) into the twin buffer of the stage; warps 10-11 serve ring 1, 12-13 ring 2 =====
        const unsigned r2 = (warp >= 12u) ? 1u : 0u;
        const unsigned ctid = threadIdx.x - (r2 ? 384u : 320u);
        const unsigned fullb = r2 ? vfull : kfull, splitb = r2 ? vsplit : ksplit;
        unsigned char* rbase = smem_gen + $A_BYTES$u + (r2 ? $RING2$u : 0u);
        unsigned it = 0, n = 0, z, at;
        for (; item_at(n, z, at); ++n) {
            const unsigned nsteps = 2u * __popc(live_tiles(at));
            for (unsigned i2 = 0; i2 < nsteps; ++i2, ++it) {
                const unsigned s = it % $STAGES$u, ph = (it / $STAGES$u) & 1u;
                mbar_wait(fullb + s * 8u, ph);
                $TR18$
                const float4* bsrc = reinterpret_cast<const float4*>(rbase + s * $RSTAGE$u);
                float4* bdst = reinterpret_cast<float4*>(rbase + s * $RSTAGE$u + $TILE_B$u);
                #pragma unroll 8
                for (unsigned i = ctid; i < $TILE_V4$u; i += 64u) {
                    const float4 v4 = bsrc[i];
                    float4 l4;
                    l4.x = v4.x - __uint_as_float(__float_as_uint(v4.x) & 0xffffe000u);
                    l4.y = v4.y - __uint_as_float(__float_as_uint(v4.y) & 0xffffe000u);
                    l4.z = v4.z - __uint_as_float(__float_as_uint(v4.z) & 0xffffe000u);
                    l4.w = v4.w - __uint_as_float(__float_as_uint(v4.w) & 0xffffe000u);
                    bdst[i] = l4;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(splitb + s * 8u);
                $TR19$
            }
        }
    } else if (warp == 15u) {
        // ===== second TMA producer: ring 2 (V for fwd / ds, dO for dv) =====
        unsigned it = 0, n = 0, z, at;
        for (; item_at(n, z, at); ++n) {
            const unsigned hh = z % $H$u, bb = z / $H$u;
            const unsigned lm = live_tiles(at);
            for (unsigned t = 0; t < $NS$u; ++t) {
                if (!((lm >> t) & 1u)) continue;
                for (unsigned half = 0; half < 2u; ++half, ++it) {
                    const unsigned s = it % $STAGES$u, ph = (it / $STAGES$u) & 1u;
                    mbar_wait(vempty + s * 8u, ph ^ 1u);
                    const unsigned st = ring + $RING2$u + s * $RSTAGE$u, fb = vfull + s * 8u, r0 = t * 128u + half * 64u;
                    if (elect_one()) {
                        mbar_expect_tx(fb, $TILE_B$u);
                        $B2_LOADS$
                    }
                    __syncwarp();
                }
            }
        }
    }
    TC_FENCE_BEFORE();
    __syncthreads();
    $TRACE_DUMP$
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512));
    }
}
""",
        KNAME=KNAME, PARAMS=params, A_BYTES=a_bytes, RING_BYTES=stages * per_stage, EPI_BYTES=epi_bytes, STAGES=stages,
        TR0=TR(0, "g"), TR1=TR(1, "g"), TR2=TR(2, "g"), TR5=TR(5, "g", "threadIdx.x == 64u"), TR8=TR(8, "g", "threadIdx.x == 64u"),
        TRACE_DECL="", TRACE_DUMP="", TR16=TR(16, "it"), TR17=TR(17, "it"), TR18=TR(18, "it", "threadIdx.x == 320u"), TR19=TR(19, "it", "threadIdx.x == 320u"), TR9=TR(9, "nn", "threadIdx.x == 64u"), TR10=TR(10, "nn", "threadIdx.x == 64u"), TR11=TR(11, "nn", "threadIdx.x == 64u"),
        TR12=TR(12, "nn", "threadIdx.x == 64u"), TR13=TR(13, "n"), TR14=TR(14, "n"), TR15=TR(15, "n", "threadIdx.x == 64u"),
        OFREE_TRY="true" if mode == "ds" else "(n == 0u || mbar_try(ofree, (n - 1u) & 1u))",
        OFREE_WAIT="" if mode == "ds" else "if (n != 0u) mbar_wait(ofree, (n - 1u) & 1u);",
        NSPLIT=8 if a_tiles * kbq >= 2 else 4, NSTORE=8 if D == 64 else 4, NISS=2 if mode == "ds" else 1, PF_LOADS=pf_loads, PFD=int(os.environ.get("DLGRAD_FLASH_PF", "2")), LIVE=live, LIVE_FILL=live_fill.replace("threadIdx.x", "(threadIdx.x - 64u)"), NA=nA, GRID=grid, ITEMS=items, NB=NB, TILE_OF=tile_of, H=H, A_LOADS=a_loads, NS=nS, PER_STAGE=per_stage,
        RAW_B=raw_b, B1_LOADS=b1_loads, B2_LOADS=b2_loads, TILE_B=tile_b, PER_STAGE16=rstage // 16,
        S_MMA=s_mma(f"tmem_base + {S0}u + b * 64u", A0, 0),
        ISSUER2=_sub(r"""const unsigned long long sdesc0 = make_desc(ring + $RING2$u, 1u, 64u, 2u);      // ring 2: the V block, K-major
        for (; item_at(n, z, at); ++n) {
            const unsigned nsteps = 2u * __popc(live_tiles(at));
            mbar_wait(qsplit, n & 1u);                            // dO (hi | lo) of this item is in tensor memory
            TC_FENCE_AFTER();
            if (nsteps == 0u) {
                if (elect_one()) umma_commit(qfree);
                __syncwarp();
                continue;
            }
            for (unsigned i = 0; i < nsteps; ++i, ++it, ++g) {
                const unsigned b = g & 1u, s = it % $STAGES$u, ph = (it / $STAGES$u) & 1u;
                mbar_wait(vsplit + s * 8u, ph);
                mbar_wait(sempty + b * 8u, ((g >> 1) & 1u) ^ 1u);
                TC_FENCE_AFTER();
                const unsigned long long sdesc = sdesc0 + (unsigned long long)(s * $PER_STAGE16$u);
                if (elect_one()) {
                    $DP_MMA$
                    umma_commit(vempty + s * 8u);
                    umma_commit(sfull + b * 8u);
                    if (i + 1u == nsteps) umma_commit(qfree);
                }
                __syncwarp();
            }
        }""", STAGES=stages, PER_STAGE16=rstage // 16, RING2=ring2_off, DP_MMA=s_mma(f"tmem_base + {DP0}u + b * 64u", A1, 0))
        if mode == "ds" else _sub(r"""const unsigned long long odesc0 = make_desc(ring + $RING2$u, 256u, 32u, 1u);    // ring 2, MN-major tile: LBO 4096 B, SBO 512 B, 32-byte atoms
        for (; item_at(n, z, at); ++n) {
            const unsigned nsteps = 2u * __popc(live_tiles(at));
            for (unsigned j = 0; j < nsteps; ++j, ++it, ++g) {
                const unsigned s = it % $STAGES$u, ph = (it / $STAGES$u) & 1u;
                mbar_wait(vsplit + s * 8u, ph);                    // the second streamed tile of this step is split
                mbar_wait(pready, g & 1u);                         // P (hi | lo) of this step is in tensor memory
                $TR3$
                TC_FENCE_AFTER();
                const unsigned long long odesc = odesc0 + (unsigned long long)(s * $PER_STAGE16$u);
                const unsigned acc = j;                            // the first step of an item overwrites the accumulator
                if (elect_one()) {
                    $O_MMA$
                    umma_commit(vempty + s * 8u);
                    umma_commit(pvdone);
                }
                __syncwarp();
                $TR4$
            }
        }""", STAGES=stages, PER_STAGE16=rstage // 16, RING2=ring2_off, O_MMA=o_mma, TR3=TR(3, "g"), TR4=TR(4, "g")),
        A_KBLOCKS=a_tiles * kbq, A0=A0,
        ROW_ITEM_BEGIN=_row_item_begin(mode, Tq, Tk, H, delta_split, lazy_zero), PF_DECL=_pf(mode, Tq, Tk)[0], PF_NEXT=_pf(mode, Tq, Tk)[1], PF_TAKE=_pf(mode, Tq, Tk)[2],
        ROW_STEP=_row_step(mode, Tq, Tk, D, S0, DP0, PHI, PLO, OACC, sc2, fl2, _f32lit(float(np.float32(scale))), tau)
        .replace("TRACE6", TR(6, "g", "threadIdx.x == 64u")).replace("TRACE7", TR(7, "g", "threadIdx.x == 64u")),
        ROW_ITEM_END=_row_item_end(mode, Tq, Tk, D, OACC, sT, sH, sB, nS).replace("TRACE20", TR(20, "n", "threadIdx.x == 64u"))
        .replace("TRACE21", TR(21, "n", "threadIdx.x == 64u")).replace("TRACE22", TR(22, "n", "threadIdx.x == 64u"))
        .replace("TRACE23", TR(23, "n", "threadIdx.x == 64u")),
        ROW_EXIT='if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");',
        RSTAGE=rstage, RING2=ring2_off, TILE_V4=tile_b // 16)
    out_elems = {"fwd": NB * Tq * D, "dv": NB * Tk * D, "ds": NB * Tq * Tk}[mode]
    return TcKernel(f"flash_{mode}", src, (grid, 1, 1), 512, smem, out_elems, 64, stages)


def _pf(mode: str, Tq: int, Tk: int) -> tuple:
    """(declarations, prefetch for the step whose first streamed row is `nc0`, take-over at the top of a step)."""
    wpr = Tk // 32
    if mode == "fwd":
        return ("uint2 mw_n = make_uint2(0u, 0u);",
                f"mw_n = __ldg(reinterpret_cast<const uint2*>(mask + (size_t)row * {wpr}u + (nc0 >> 5)));",
                "const uint2 mw = mw_n;")
    if mode == "dv":
        return ("float lse_n = 0.f; unsigned mwe_n = 0u;",
                f"{{ const unsigned qe_ = nc0 + wh * 32u + lane; lse_n = __ldg(lse + (size_t)z * {Tq}u + qe_); "
                f"mwe_n = __ldg(mask + (size_t)qe_ * {wpr}u + (row >> 5)); }}",
                "const float lse_e = lse_n; const unsigned mw_e = mwe_n;")
    return ("unsigned bits_n = 0u;",
            f"bits_n = __ldg(mask + (size_t)row * {wpr}u + (nc0 >> 5) + wh);",
            "const unsigned bits = bits_n;")


def _row_item_begin(mode: str, Tq: int, Tk: int, H: int = 1, delta_split: bool = False, lazy_zero: bool = False) -> str:
    if mode == "fwd":
        return f"float m_run = {NEG_INF}, l_run = 0.f;      // l_run: this warp's 32 columns of every step only"
    if mode == "ds":
        return f"""const float lse_row = __ldg(lse + (size_t)z * {Tq}u + row);       // P = 2^(s c - lse) for this query row
            const float d_row = __ldg(delta + {"((size_t)bb * " + str(Tq) + "u + row) * " + str(H) + "u + hh" if delta_split else "(size_t)z * " + str(Tq) + "u + row"});       // rowsum(dO * O), in dO's storage order
            const unsigned my_epi = epi + (warp - 2u) * 4096u;
            unsigned char* sbuf_gen = smem_gen + (epi - smem) + (warp - 2u) * 4096u;
            // streamed tiles the mask kills entirely: dS is zero there whatever P and dP hold (eight lanes per 128-byte row;
            // each warp of a pair clears its 64 of the 128 columns)
            for (unsigned t = 0; t < {Tk // 128}u; ++t) {{
                if ((lm >> t) & 1u) continue;
                {"if (__ldg(mask + " + str(Tq * Tk // 32 + (Tq // 128) * (Tk // 128) + Tq) + "u) == " + str(Tq) + "u) break;      // every query row is live: the consumers skip dead tiles too" if lazy_zero else ""}
                float* dst = out + ((size_t)z * {Tq}u + at * 128u + sq * 32u + (lane >> 3)) * {Tk}u + t * 128u + wh * 64u + (lane & 7u) * 4u;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                #pragma unroll
                for (unsigned i = 0; i < 8u; ++i) {{
                    *reinterpret_cast<float4*>(dst + (size_t)(4u * i) * {Tk}u) = z4;
                    *reinterpret_cast<float4*>(dst + (size_t)(4u * i) * {Tk}u + 32u) = z4;
                }}
            }}"""
    return ""


def _row_step(mode, Tq, Tk, D, S0, DP0, PHI, PLO, OACC, sc2, fl2, sc, tau) -> str:
    wpr = Tk // 32
    if mode == "fwd":
        if D == 64:
            resc = f"""{{ unsigned o[32]; TMEM_LD32(lane_base + {OACC}u + wh * 32u, o); TMEM_WAIT_LD();
                          #pragma unroll
                          for (unsigned e = 0; e < 32u; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                          TMEM_ST32(lane_base + {OACC}u + wh * 32u, o); }}"""
        else:
            resc = f"""if (wh == 0u) {{ unsigned o[32]; TMEM_LD32(lane_base + {OACC}u, o); TMEM_WAIT_LD();
                          #pragma unroll
                          for (unsigned e = 0; e < 32u; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                          TMEM_ST32(lane_base + {OACC}u, o); }}"""
        return f"""// both warps of a pair read all 64 scores of the row for the maximum (no exchange between them), then each
                    // exponentiates, sums and splits its own 32 columns
                    unsigned v[32], u[32];
                    TMEM_LD32(lane_base + {S0}u + b * 64u + wh * 32u, v);             // mine
                    TMEM_LD32(lane_base + {S0}u + b * 64u + (wh ^ 1u) * 32u, u);      // the partner's
                    const unsigned mine_bits = wh ? mw.y : mw.x, other_bits = wh ? mw.x : mw.y;
                    TMEM_WAIT_LD();
                    TC_FENCE_BEFORE();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sempty + b * 8u);        // S[b] may be overwritten by the step after next
                    float cm;
                    if ((mine_bits | other_bits) == 0u) {{                // nothing masked in this row's 64 columns (scale > 0)
                        cm = fmaxf(__uint_as_float(v[0]), __uint_as_float(u[0]));
                        #pragma unroll
                        for (unsigned e = 1; e < 32u; ++e) cm = fmaxf(cm, fmaxf(__uint_as_float(v[e]), __uint_as_float(u[e])));
                        cm *= {sc2};
                    }} else {{
                        cm = {NEG_INF};
                        #pragma unroll
                        for (unsigned e = 0; e < 32u; ++e) {{
                            cm = fmaxf(cm, ((mine_bits >> e) & 1u) ? {fl2} : __uint_as_float(v[e]) * {sc2});
                            cm = fmaxf(cm, ((other_bits >> e) & 1u) ? {fl2} : __uint_as_float(u[e]) * {sc2});
                        }}
                    }}
                    // online softmax with a lazily updated maximum (see `tau`); identical in both warps of the pair
                    float alpha = 1.f;
                    const float m_new = fmaxf(m_run, cm);
                    const bool grow = m_new > m_run + {_f32lit(tau)};
                    if (grow) {{ alpha = ex2(m_run - m_new); l_run *= alpha; m_run = m_new; }}
                    const bool resc = __any_sync(0xffffffffu, grow) && j != 0u;
                    float x[32];
                    float add = 0.f;
                    if (!(m_run > {NEG_INF})) {{                          // nothing finite so far: P = 0 (an all-masked row ends NaN)
                        #pragma unroll
                        for (unsigned e = 0; e < 32u; ++e) x[e] = 0.f;
                    }} else if (mine_bits == 0u) {{
                        #pragma unroll
                        for (unsigned e = 0; e < 32u; ++e) {{ x[e] = ex2(fmaf(__uint_as_float(v[e]), {sc2}, -m_run)); add += x[e]; }}
                    }} else {{
                        #pragma unroll
                        for (unsigned e = 0; e < 32u; ++e) {{
                            x[e] = ex2((((mine_bits >> e) & 1u) ? {fl2} : __uint_as_float(v[e]) * {sc2}) - m_run);
                            add += x[e];
                        }}
                    }}
                    l_run += add;
                    TRACE6
                    if (g != 0u) {{ mbar_wait(pvdone, (g - 1u) & 1u); TC_FENCE_AFTER(); }}     // the MMA that read P (and wrote O) is done
                    TRACE7
                    if (resc) {{
                        {resc}
                    }}
                    split_to_tmem(lane_base + {PHI}u + wh * 32u, lane_base + {PLO}u + wh * 32u, x);
                    TMEM_WAIT_ST();
                    TC_FENCE_BEFORE();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pready);"""
    if mode == "dv":
        return f"""unsigned v[32];
                    TMEM_LD32(lane_base + {S0}u + b * 64u + wh * 32u, v);
                    // thread = key row; column e of this warp = query c0 + 32 wh + e.  Lane e fetches that query's lse and its
                    // mask word for this warp's 32 keys (coalesced); they reach the other lanes by shuffle.
                    const bool any_masked = __any_sync(0xffffffffu, mw_e != 0u);
                    TMEM_WAIT_LD();
                    TC_FENCE_BEFORE();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sempty + b * 8u);
                    float x[32];
                    if (!any_masked) {{
                        #pragma unroll
                        for (unsigned e = 0; e < 32u; ++e)
                            x[e] = ex2(fmaf(__uint_as_float(v[e]), {sc2}, -__shfl_sync(0xffffffffu, lse_e, e)));
                    }} else {{
                        #pragma unroll
                        for (unsigned e = 0; e < 32u; ++e) {{
                            const unsigned bit = (__shfl_sync(0xffffffffu, mw_e, e) >> lane) & 1u;
                            x[e] = ex2((bit ? {fl2} : __uint_as_float(v[e]) * {sc2}) - __shfl_sync(0xffffffffu, lse_e, e));
                        }}
                    }}
                    if (g != 0u) {{ mbar_wait(pvdone, (g - 1u) & 1u); TC_FENCE_AFTER(); }}
                    split_to_tmem(lane_base + {PHI}u + wh * 32u, lane_base + {PLO}u + wh * 32u, x);
                    TMEM_WAIT_ST();
                    TC_FENCE_BEFORE();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pready);"""
    # ds
    stage_writes = "\n                        ".join(
        f"*reinterpret_cast<float4*>(sbuf_gen + lane * 128u + (({jj}u ^ (lane & 7u)) << 4)) = "
        f"make_float4(x[{4 * jj}], x[{4 * jj + 1}], x[{4 * jj + 2}], x[{4 * jj + 3}]);" for jj in range(8))
    return f"""unsigned v[32], d[32];
                    TMEM_LD32(lane_base + {S0}u + b * 64u + wh * 32u, v);
                    TMEM_LD32(lane_base + {DP0}u + b * 64u + wh * 32u, d);
                    TMEM_WAIT_LD();
                    TC_FENCE_BEFORE();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sempty + b * 8u);
                    float x[32];
                    #pragma unroll
                    for (unsigned e = 0; e < 32u; ++e) {{
                        const float p = ex2(fmaf(__uint_as_float(v[e]), {sc2}, -lse_row));
                        const float tt = p * (__uint_as_float(d[e]) - d_row);
                        x[e] = ((bits >> e) & 1u) ? 0.f : tt * {sc};
                    }}
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous store has left the box
                    __syncwarp();
                    {stage_writes}
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {{
                        tma_store_3d(&tmOut, my_epi, (int)(c0 + wh * 32u), (int)(at * 128u + sq * 32u), (int)z);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }}"""


def _row_item_end(mode, Tq, Tk, D, OACC, sT, sH, sB, nS) -> str:
    if mode == "ds":
        return ""
    TA = Tk if mode == "dv" else Tq
    # the accumulator's D columns: 32 per warp of the pair at D = 64; at D = 32 the first warp of the pair takes them all
    mine = "true" if D == 64 else "wh == 0u"
    col0 = "wh * 32u" if D == 64 else "0u"
    # The tile leaves through shared memory and ONE TMA store per warp (a 32 x 32 box): a thread owns a row, so direct
    # stores put 32 rows x 16 bytes in every instruction — measured 1600 clk of store back-pressure per item plus another
    # 2000 before the next instruction got out, a quarter of the kernel.  The staging box is the warp's own 4 KiB of the A
    # tile's shared-memory region (exactly the rows / k-block this warp split into tensor memory a moment ago); the
    # producer does not refill that region before `ofree` says the stores have read it.
    stage_writes = "\n                    ".join(
        f"*reinterpret_cast<float4*>(obox_gen + lane * 128u + (({jj}u ^ (lane & 7u)) << 4)) = "
        f"make_float4(__uint_as_float(o[{4 * jj}]) * fac, __uint_as_float(o[{4 * jj + 1}]) * fac, "
        f"__uint_as_float(o[{4 * jj + 2}]) * fac, __uint_as_float(o[{4 * jj + 3}]) * fac);" for jj in range(8))
    kk_of = "wh" if D == 64 else "0u"
    store = f"""const unsigned obox = smem + {kk_of} * 16384u + sq * 4096u;
                    unsigned char* obox_gen = smem_gen + {kk_of} * 16384u + sq * 4096u;
                    {stage_writes}
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {{
                        tma_store_4d(&tmO, obox, (int)({col0}), (int)(at * 128u + sq * 32u), (int)hh, (int)bb);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the box has been read: it may be refilled
                        mbar_arrive(ofree);
                    }}
                    __syncwarp();"""
    if mode == "fwd":
        return f"""{{
                // the row sum: each warp of a pair holds the sum over its 32 columns of every step (same maximum in both)
                xch[(warp - 2u) * 32u + lane] = l_run;
                asm volatile("bar.sync %0, 64;" :: "r"(1u + sq) : "memory");
                const float l_tot = l_run + xch[((warp - 2u) ^ 4u) * 32u + lane];
                asm volatile("bar.sync %0, 64;" :: "r"(1u + sq) : "memory");
                TRACE20
                unsigned o[32];
                float fac = 1.f;
                if (nsteps == 0u) {{
                    // no live key tile at all: every row of this tile is fully masked -> NaN, like the unfused chain
                    #pragma unroll
                    for (unsigned e = 0; e < 32u; ++e) o[e] = 0x7fc00000u;
                }} else {{
                    mbar_wait(pvdone, (g - 1u) & 1u);
                    TRACE21
                    TC_FENCE_AFTER();
                    if ({mine}) {{ TMEM_LD32(lane_base + {OACC}u + {col0}, o); TMEM_WAIT_LD(); }}
                    TC_FENCE_BEFORE();
                    fac = 1.f / l_tot;                                 // l = 0 (a fully masked row): 0 * inf = NaN
                    TRACE22
                }}
                if ({mine}) {{
                    {store}
                }}
                TRACE23
                if (wh == 0u) lse[(size_t)z * {TA}u + row] = (l_tot > 0.f) ? m_run + log2f(l_tot) : {QNAN};
            }}"""
    return f"""{{
                unsigned o[32];
                const float fac = 1.f;
                if (nsteps == 0u) {{
                    #pragma unroll
                    for (unsigned e = 0; e < 32u; ++e) o[e] = 0u;           // no query attends to these keys
                }} else {{
                    mbar_wait(pvdone, (g - 1u) & 1u);
                    TC_FENCE_AFTER();
                    if ({mine}) {{ TMEM_LD32(lane_base + {OACC}u + {col0}, o); TMEM_WAIT_LD(); }}
                    TC_FENCE_BEFORE();
                }}
                if ({mine}) {{
                    {store}
                }}
            }}"""


def tmap4(T: int, D: int, H: int, B: int, strides: tuple, box_rows: int, mn: bool = False) -> dict:
    """Rank-4 tensor map (D, T, H, B) over a (.., T, .., D) operand with element strides (row, head, batch);
    K-major box {32, box_rows} with SWIZZLE_128B, or MN-major box {32 columns, 32 rows} with 32-byte atoms."""
    sT, sH, sB = strides
    return dict(rank=4, dims=(D, T, H, B), strides=(sT * 4, sH * 4, sB * 4), box=(32, 32 if mn else box_rows, 1, 1),
                swizzle=4 if mn else 3)
