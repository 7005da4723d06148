"""CUDA C generators for the bandwidth-bound ops of the dlgrad hot path (sm_100a).

Counterpart of dlgrad/codegen/cpu_kernel.py, built the other way round: the reference emits
ndim-generic scalar loops and passes shapes at run time (cpu_kernel.py:12-66); here every kernel
is *shape-specialised* — extents, broadcast strides, vector width, grid and unroll are literals in
the source — so index arithmetic is constant-folded (mul-shift instead of div), 128-bit accesses
are chosen statically, and the launch geometry is part of the plan.  One cubin per (op, shape),
cached by source hash (runtime/cuda/compiler.py).

Every kernel has the fixed plan signature
    (p0, p1, p2, p3, out, f0, f1, f2, f3)
(see include/dlgrad_cuda.h).  Generators return `Kernel(name, source, grid, block, smem, out_elems)`.
All generators are @cache'd: a given (op, shape) is generated once per process.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import cache

from dlgrad_b200.runtime.cuda.compiler import KNAME

SIG = ("(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2, "
       "const float* __restrict__ p3, float* __restrict__ out, float f0, float f1, float f2, float f3)")
# in-place kernels alias `out`/p* — no __restrict__
SIG_ALIAS = ("(float* p0, float* p1, float* p2, float* p3, float* out, "
             "float f0, float f1, float f2, float f3)")

NUM_SMS = 148


@dataclass(frozen=True)
class Kernel:
    name: str
    source: str
    grid: tuple
    block: tuple
    smem: int = 0
    out_elems: int = 0
    zero_out: bool = False


def _cdiv(a: int, b: int) -> int:
    return -(-a // b)


def _pow2_floor(x: int) -> int:
    return 1 << (max(1, x).bit_length() - 1)


def _pow2_ceil(x: int) -> int:
    return 1 << (max(1, x) - 1).bit_length()


# ------------------------------------------------------------------------------------------------
# elementwise map with broadcasting
# ------------------------------------------------------------------------------------------------
def collapse_dims(shape: tuple, strides_list: list[tuple]) -> tuple[tuple, list[tuple]]:
    """Drop size-1 dims and merge neighbours that are jointly contiguous for every operand."""
    dims = [(s, tuple(st[i] for st in strides_list)) for i, s in enumerate(shape) if s != 1]
    if not dims:
        return (1,), [(0,) for _ in strides_list]
    merged = [dims[0]]
    for size, strd in dims[1:]:
        psize, pstrd = merged[-1]
        if all(ps == s * size for ps, s in zip(pstrd, strd)):
            merged[-1] = (psize * size, strd)
        else:
            merged.append((size, strd))
    out_shape = tuple(s for s, _ in merged)
    out_strides = [tuple(m[1][j] for m in merged) for j in range(len(strides_list))]
    return out_shape, out_strides


@cache
def ewise(expr: str, out_shape: tuple, in_strides: tuple, aligned: bool = True, tag: str = "ew") -> Kernel:
    """out[idx] = expr(v0, v1, v2, v3, f0..f3) where v_j = p_j[<idx, in_strides[j]>].

    `in_strides[j]` are element strides of pointer operand j aligned to `out_shape`
    (0 = broadcast, as in get_broadcast_strides_2, dlgrad/helpers.py:276-308).
    Covers cpu_kernel.arithmetic (:12-66), utils (:246-319), clamp (:321-349), where (:536-587),
    masked_fill (:604-650), cmp_2d (:653-668), full (:758-770), arange (:743-755) and permute
    when the innermost axis stays innermost (:180-215).
    """
    numel = 1
    for s in out_shape:
        numel *= s
    n_in = len(in_strides)
    contiguous = []
    acc = 1
    for s in reversed(out_shape):
        contiguous.append(acc)
        acc *= s
    out_strides = tuple(reversed(contiguous))
    shape, strides = collapse_dims(tuple(out_shape), [out_strides, *in_strides])
    ostr, istr = strides[0], strides[1:]
    inner = shape[-1]
    vec = (aligned and inner % 4 == 0 and all(st[-1] in (0, 1) for st in istr)
           and all(all(s % 4 == 0 for s in st[:-1]) for st in istr))
    V = 4 if vec else 1
    nvec = numel // V
    big = max([numel] + [sum((d - 1) * abs(s) for d, s in zip(shape, st)) + 1 for st in istr]) >= 2**31
    idx_t = "unsigned long long" if big else "unsigned"
    BLOCK = 256
    U = 4 if nvec >= BLOCK * 4 * NUM_SMS else (2 if nvec >= BLOCK * 2 * NUM_SMS else 1)
    ntiles = max(1, _cdiv(nvec, BLOCK * U))
    # persistent grid: at most 8 resident CTAs on each of the 148 SMs, tiles walked grid-stride, so a
    # large tensor is not cut into 3.46 waves with a ragged tail
    grid = min(ntiles, NUM_SMS * 8)

    # index decomposition of the (vector) index into collapsed dims
    vshape = list(shape)
    vshape[-1] = inner // V
    lines = []
    nd = len(vshape)
    if nd == 1:
        lines.append(f"{idx_t} i0 = vi;")
    else:
        lines.append(f"{idx_t} t = vi;")
        for d in range(nd - 1, 0, -1):
            lines.append(f"{idx_t} i{d} = t % {vshape[d]}u; t /= {vshape[d]}u;")
        lines.append(f"{idx_t} i0 = t;")
    offs = []
    for j, st in enumerate(istr):
        terms = []
        for d, s in enumerate(st):
            if s == 0:
                continue
            mult = s * (V if d == nd - 1 else 1)
            terms.append(f"i{d}*{mult}u" if mult != 1 else f"i{d}")
        offs.append(" + ".join(terms) if terms else "0")
    decomp = "\n            ".join(lines)

    T = "float4" if vec else "float"
    loads = []
    for j, st in enumerate(istr):
        if vec and st[-1] == 0:
            loads.append(f"{{ float s = __ldg(p{j} + ({offs[j]})); a{j}[u] = make_float4(s, s, s, s); }}")
        elif vec:
            loads.append(f"a{j}[u] = __ldg(reinterpret_cast<const float4*>(p{j} + ({offs[j]})));")
        else:
            loads.append(f"a{j}[u] = __ldg(p{j} + ({offs[j]}));")
    load_block = "\n                ".join(loads)
    decl = "".join(f"    {T} a{j}[{U}];\n" for j in range(n_in))
    vargs = lambda comp: ", ".join([f"a{j}[u]{comp}" for j in range(n_in)] + ["f0", "f1", "f2", "f3"])  # noqa: E731
    params = ", ".join([f"float v{j}" for j in range(n_in)] + ["float f0", "float f1", "float f2", "float f3"])
    need_idx = "ei" in expr   # arange-style kernels use the flat element index
    if vec:
        if need_idx:
            compute = ("float4 r; " + " ".join(
                f"r.{c} = op({vargs('.' + c)}, (float)(vi*4u+{k}u));" for k, c in enumerate("xyzw")))
        else:
            compute = "float4 r; " + " ".join(f"r.{c} = op({vargs('.' + c)}, 0.f);" for c in "xyzw")
        store = "reinterpret_cast<float4*>(out)[vi] = r;"
    else:
        compute = f"float r = op({vargs('')}, {'(float)vi' if need_idx else '0.f'});"
        store = "out[vi] = r;"
    src = f"""
__device__ __forceinline__ float op({params}, float ei) {{ return {expr}; }}
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    constexpr {idx_t} NV = {nvec}u;
    for (unsigned tile = blockIdx.x; tile < {ntiles}u; tile += {grid}u) {{
    const {idx_t} base = ({idx_t})tile * {BLOCK * U}u + threadIdx.x;
{decl}    #pragma unroll
    for (int u = 0; u < {U}; ++u) {{
        const {idx_t} vi = base + u * {BLOCK}u;
        if (vi < NV) {{
            {decomp}
            {load_block}
        }}
    }}
    #pragma unroll
    for (int u = 0; u < {U}; ++u) {{
        const {idx_t} vi = base + u * {BLOCK}u;
        if (vi < NV) {{
            {compute}
            {store}
        }}
    }}
    }}
}}
"""
    if n_in == 0:
        src = src.replace("            \n", "")
    return Kernel(tag, src, (grid, 1, 1), (BLOCK, 1, 1), 0, numel)


# expressions (v0.. = operands in pointer-slot order, f0.. = scalar slots)
BINARY_EXPR = {
    "add": "{a} + {b}", "sub": "{a} - {b}", "mul": "{a} * {b}", "div": "{a} / {b}",
    "eq": "({a} == {b}) ? 1.f : 0.f", "gt": "({a} > {b}) ? 1.f : 0.f",
    "gte": "({a} >= {b}) ? 1.f : 0.f", "lte": "({a} <= {b}) ? 1.f : 0.f",
}


def pow_expr(n: int, v: str = "v0") -> str:
    """x ** int(p): the reference truncates the exponent (dlgrad/runtime/cpu.py:500,
    cpu_kernel.py:288-291: powf(x, (int)val)); gcc -ffast-math expands small powers to products."""
    if n == 0:
        return "1.f"
    if n == 1:
        return v
    if n == 2:
        return f"{v} * {v}"
    if n == 3:
        return f"{v} * {v} * {v}"
    if n == 4:
        return f"({v} * {v}) * ({v} * {v})"
    if n == -1:
        return f"1.f / {v}"
    if n == -2:
        return f"1.f / ({v} * {v})"
    return f"powf({v}, {float(n)}f)"


UNARY_EXPR = {
    "neg": "-v0", "exp": "expf(v0)", "log": "logf(v0)", "sqrt": "sqrtf(v0)",
    "rsqrt": "1.f / sqrtf(v0)", "copy": "v0",
}


# ------------------------------------------------------------------------------------------------
# reductions over one axis of a contiguous tensor viewed as (outer, R, inner)
# ------------------------------------------------------------------------------------------------
_RED = {
    "sum": ("0.f", "a + b", "acc"),
    "mean": ("0.f", "a + b", "acc / {R}.f"),
    "max": ("-3.402823466e+38f", "(b > a) ? b : a", "acc"),
}


@cache
def reduce_rows(op: str, outer: int, R: int, aligned: bool = True, scale_R: int | None = None) -> Kernel:
    """Last-axis reduce: out[o] = red_r x[o*R + r].  Sub-warp / warp / block per row, 128-bit loads, 8
    independent accumulators per thread, shuffle tree; rows are walked grid-stride by a grid of at most
    148 x 8 CTAs.  Restates reduce_source with the reduced axis innermost and reduce()
    (dlgrad/codegen/cpu_kernel.py:69-177); max keeps the reference's `-FLT_MAX` seed and strict `>`."""
    init, comb, fin = _RED[op]
    fin = fin.format(R=scale_R if scale_R is not None else R)
    V = 4 if (aligned and R % 4 == 0) else 1
    nv = R // V
    if nv >= 2048 and outer < NUM_SMS * 8:
        T = 256                       # block per row
    else:
        T = min(32, _pow2_ceil(_cdiv(nv, 4)))
    BLOCK = 256
    rpb = BLOCK // T
    nblocks = _cdiv(outer, rpb)
    grid = min(nblocks, NUM_SMS * 8)
    if V == 4:
        ld = ("const float4 v = __ldg(rowv + c), w = (c + {T}u < {nv}u) ? __ldg(rowv + c + {T}u) : make_float4({i}, {i}, {i}, {i}); "
              "a0 = red(a0, v.x); a1 = red(a1, v.y); a2 = red(a2, v.z); a3 = red(a3, v.w); "
              "a4 = red(a4, w.x); a5 = red(a5, w.y); a6 = red(a6, w.z); a7 = red(a7, w.w);").format(T=T, nv=nv, i=init)
        step = 2 * T
        rowdecl = "const float4* rowv = reinterpret_cast<const float4*>(p0 + (size_t)r * %du);" % R
    else:
        ld = "a0 = red(a0, __ldg(rowv + c));"
        step = T
        rowdecl = "const float* rowv = p0 + (size_t)r * %du;" % R
    if T <= 32:
        tail = f"""
        #pragma unroll
        for (int m = {T // 2}; m > 0; m >>= 1) acc = red(acc, __shfl_xor_sync(0xffffffffu, acc, m));
        if (lane == 0 && r < {outer}u) out[r] = {fin};"""
    else:
        tail = f"""
        #pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc = red(acc, __shfl_xor_sync(0xffffffffu, acc, m));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {{
            acc = threadIdx.x < {T // 32} ? part[threadIdx.x] : {init};
            #pragma unroll
            for (int m = {max(1, T // 64)}; m > 0; m >>= 1) acc = red(acc, __shfl_xor_sync(0xffffffffu, acc, m));
            if (threadIdx.x == 0) out[r] = {fin};
        }}"""
    src = f"""
__device__ __forceinline__ float red(float a, float b) {{ return {comb}; }}
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    {"__shared__ float part[%d];" % (T // 32) if T > 32 else ""}
    const unsigned lane = threadIdx.x % {T}u;
    for (unsigned blk = blockIdx.x; blk < {nblocks}u; blk += {grid}u) {{
        const unsigned r = blk * {rpb}u + threadIdx.x / {T}u;
        float a0 = {init}, a1 = {init}, a2 = {init}, a3 = {init}, a4 = {init}, a5 = {init}, a6 = {init}, a7 = {init};
        if (r < {outer}u) {{
            {rowdecl}
            #pragma unroll 4
            for (unsigned c = lane; c < {nv}u; c += {step}u) {{ {ld} }}
        }}
        float acc = red(red(red(a0, a1), red(a2, a3)), red(red(a4, a5), red(a6, a7)));{tail}
    }}
}}
"""
    return Kernel(f"red_{op}_rows", src, (grid, 1, 1), (BLOCK, 1, 1), 0, outer)


@cache
def reduce_cols(op: str, outer: int, R: int, inner: int, aligned: bool = True,
                splits: int = 1, scale_R: int | None = None, cluster: bool = False, prod3: bool = False) -> Kernel:
    """Strided reduce: out[o][i] = red_r x[(o*R + r)*inner + i].  Threads run along `inner` (coalesced
    128-bit loads), the 8 warps of a CTA take interleaved rows and meet in shared memory.  When
    outer*inner alone cannot fill 148 SMs, R is cut into `splits` slices along blockIdx.z:
      cluster=True  (splits <= 8): the slices of one column tile form a THREAD-BLOCK CLUSTER; every CTA
                    leaves its partial in its own shared memory, and after a cluster barrier rank 0 adds
                    the partials in rank order through distributed shared memory (mapa +
                    ld.shared::cluster) — one launch, no atomics, deterministic order;
      cluster=False the partials go to out[s][o][i] and a second reduce_cols pass over `splits` finishes.
    The reference walks this case with a strided inner loop (cpu_kernel.py:69-148), its slowest pattern
    (152 ms for 4096^2 axis 0).
    prod3: the reduced value is x[r][i] * (y[r][i] * t[r]) with y = p1 (same layout) and t = p2 (one value per
    row) — the RMSNorm weight gradient sum_rows g * (x * rstd) without writing the product first.  Explicit
    round-to-nearest multiplies, no FMA contraction with the running sum: bit-identical to multiply-then-sum."""
    init, comb, fin = _RED[op]
    final_here = splits == 1 or cluster
    fin = fin.format(R=scale_R if scale_R is not None else R) if final_here else "acc"
    V = 4 if (aligned and inner % 4 == 0) else 1
    ncolv = inner // V
    Y = 8
    rows_per_split = _cdiv(R, splits)
    T = "float4" if V == 4 else "float"
    comps = ["x", "y", "z", "w"] if V == 4 else [""]

    def each(fmt):
        return " ".join(fmt.format(c=("." + c) if c else "") for c in comps)
    initv = f"make_float4({init}, {init}, {init}, {init})" if V == 4 else init
    ldv = (f"__ldg(reinterpret_cast<const float4*>(base + (size_t)r * {inner}u) + cv)" if V == 4
           else f"__ldg(base + (size_t)r * {inner}u + cv)")
    fin_store = each("res{c} = " + fin.replace("acc", "acc{c}") + ";")
    if prod3:
        assert outer == 1 and not cluster
        ld2 = (f"__ldg(reinterpret_cast<const float4*>(p1 + (size_t)r * {inner}u) + cv)" if V == 4
               else f"__ldg(p1 + (size_t)r * {inner}u + cv)")
        ldv = f"mul3({ldv}, {ld2}, __ldg(p2 + r))"
    if cluster and splits > 1:
        ld_remote = ('asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(ra));'
                     if V == 4 else 'asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(o) : "r"(ra));')
        finish = f"""
    __shared__ {T} mine[32];
    unsigned crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (threadIdx.y == 0) mine[threadIdx.x] = acc;
    asm volatile("barrier.cluster.arrive.release.aligned;\\n\\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (crank == 0 && threadIdx.y == 0 && cv < {ncolv}u) {{
        const unsigned la = (unsigned)__cvta_generic_to_shared(&mine[threadIdx.x]);
        for (unsigned pr = 1; pr < {splits}u; ++pr) {{
            unsigned ra; {T} o;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(pr));
            {ld_remote}
            {each("acc{c} = red(acc{c}, o{c});")}
        }}
        float* dst = out + (size_t)o_ * {inner}u;
        {T} res; {fin_store}
        reinterpret_cast<{T}*>(dst)[cv] = res;
    }}
    asm volatile("barrier.cluster.arrive.release.aligned;\\n\\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");"""
        attr = f"__cluster_dims__(1, 1, {splits}) "
        out_elems = outer * inner
    else:
        finish = f"""
    if (threadIdx.y == 0 && cv < {ncolv}u) {{
        float* dst = out + ((size_t)s * {outer}u + o_) * {inner}u;
        {T} res; {fin_store}
        reinterpret_cast<{T}*>(dst)[cv] = res;
    }}"""
        attr = ""
        out_elems = splits * outer * inner
    src = f"""
__device__ __forceinline__ float red(float a, float b) {{ return {comb}; }}
__device__ __forceinline__ float mul3(float a, float b, float t) {{ return __fmul_rn(a, __fmul_rn(b, t)); }}
__device__ __forceinline__ float4 mul3(float4 a, float4 b, float t) {{
    return make_float4(mul3(a.x, b.x, t), mul3(a.y, b.y, t), mul3(a.z, b.z, t), mul3(a.w, b.w, t));
}}
extern "C" __global__ void {attr}__launch_bounds__({32 * Y}) {KNAME}{SIG} {{
    __shared__ {T} part[{Y}][33];
    const unsigned cv = blockIdx.x * 32u + threadIdx.x;
    const unsigned o_ = blockIdx.y, s = blockIdx.z;
    const unsigned r0 = s * {rows_per_split}u;
    const unsigned r1 = min(r0 + {rows_per_split}u, {R}u);
    const float* base = p0 + (size_t)o_ * {R * inner}u;
    {T} acc = {initv}, acc2 = {initv};
    if (cv < {ncolv}u) {{
        unsigned r = r0 + threadIdx.y;
        #pragma unroll 4
        for (; r + {Y}u < r1; r += {2 * Y}u) {{
            const {T} v = {ldv};
            const {T} w = {ldv.replace("(size_t)r", "(size_t)(r + %du)" % Y).replace("p2 + r)", "p2 + r + %du)" % Y)};
            {each("acc{c} = red(acc{c}, v{c});")} {each("acc2{c} = red(acc2{c}, w{c});")}
        }}
        if (r < r1) {{ const {T} v = {ldv}; {each("acc{c} = red(acc{c}, v{c});")} }}
        {each("acc{c} = red(acc{c}, acc2{c});")}
    }}
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0) {{
        #pragma unroll
        for (int y = 1; y < {Y}; ++y) {{ const {T} o = part[y][threadIdx.x]; {each("acc{c} = red(acc{c}, o{c});")} }}
    }}{finish}
}}
"""
    grid = (_cdiv(ncolv, 32), outer, splits)
    return Kernel(f"red_{op}_cols" + ("_prod3" if prod3 else ""), src, grid, (32, Y, 1), 0, out_elems)


@cache
def reduce_cols_slab(op: str, outer: int, R: int, inner: int, slabs: int, final: bool,
                     scale_R: int | None = None) -> Kernel:
    """Strided reduce for wide `inner` (128-bit aligned): out[s][o][i] = red_{r in slab s} x[(o*R+r)*inner+i].
    A CTA owns a slab of CONSECUTIVE rows and a span of up to 4096 contiguous columns, so every row it
    touches is read as one long contiguous run (DRAM-page friendly, unlike 512-byte column strips at a
    16 KiB pitch); each thread keeps up to 4 float4 column accumulators and two rows in flight.  With
    slabs > 1 a second pass over the `slabs` axis (same kernel, final=True) finishes."""
    init, comb, fin = _RED[op]
    fin = fin.format(R=scale_R if scale_R is not None else R) if final else "acc"
    ncolv = inner // 4
    BLOCK = 256
    TC = min(4, _cdiv(ncolv, BLOCK))
    span = BLOCK * TC                       # float4 columns per CTA
    nspans = _cdiv(ncolv, span)
    rows_per = _cdiv(R, slabs)
    accs = "\n    ".join(f"float4 a{j} = make_float4({init}, {init}, {init}, {init}), b{j} = a{j};" for j in range(TC))

    def upd(acc, v):
        return " ".join(f"{acc}.{c} = red({acc}.{c}, {v}.{c});" for c in "xyzw")
    body_a = "\n            ".join(
        f"if (c{j} < {ncolv}u) {{ const float4 v = __ldg(row0 + c{j}); {upd(f'a{j}', 'v')} }}" for j in range(TC))
    body_b = "\n            ".join(
        f"if (c{j} < {ncolv}u) {{ const float4 v = __ldg(row1 + c{j}); {upd(f'b{j}', 'v')} }}" for j in range(TC))
    cols = "\n    ".join(f"const unsigned c{j} = blockIdx.x * {span}u + threadIdx.x + {j * BLOCK}u;" for j in range(TC))
    stores = "\n    ".join(
        f"if (c{j} < {ncolv}u) {{ float4 acc; {' '.join(f'acc.{c} = red(a{j}.{c}, b{j}.{c});' for c in 'xyzw')} "
        f"float4 res; {' '.join('res.%s = %s;' % (c, fin.replace('acc', 'acc.' + c)) for c in 'xyzw')} dst[c{j}] = res; }}"
        for j in range(TC))
    src = f"""
__device__ __forceinline__ float red(float a, float b) {{ return {comb}; }}
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    const unsigned o_ = blockIdx.y, s = blockIdx.z;
    const unsigned r0 = s * {rows_per}u, r1 = min(r0 + {rows_per}u, {R}u);
    const float4* base = reinterpret_cast<const float4*>(p0 + (size_t)o_ * {R * inner}u);
    {cols}
    {accs}
    unsigned r = r0;
    #pragma unroll 2
    for (; r + 1u < r1; r += 2u) {{
        const float4* row0 = base + (size_t)r * {ncolv}u;
        const float4* row1 = row0 + {ncolv}u;
        {body_a}
        {body_b}
    }}
    if (r < r1) {{
        const float4* row0 = base + (size_t)r * {ncolv}u;
        {body_a}
    }}
    float4* dst = reinterpret_cast<float4*>(out + ((size_t)s * {outer}u + o_) * {inner}u);
    {stores}
}}
"""
    return Kernel(f"red_{op}_slab", src, (nspans, outer, slabs), (BLOCK, 1, 1), 0, slabs * outer * inner)


@cache
def reduce_all_partial(op: str, numel: int, aligned: bool = True) -> Kernel:
    """Stage 1 of a full reduce (cpu_kernel.reduce, :151-177): G blocks grid-stride over the tensor
    and each leaves one partial in out[blockIdx.x]; reduce_rows(op, 1, G) finishes."""
    init, comb, _ = _RED[op]
    V = 4 if (aligned and numel % 4 == 0) else 1
    nv = numel // V
    BLOCK = 256
    G = max(1, min(NUM_SMS * 4, _cdiv(nv, BLOCK * 4)))
    ld = ("float4 v = __ldg(reinterpret_cast<const float4*>(p0) + c); "
          "a0 = red(a0, v.x); a1 = red(a1, v.y); a2 = red(a2, v.z); a3 = red(a3, v.w);"
          if V == 4 else "a0 = red(a0, __ldg(p0 + c));")
    src = f"""
__device__ __forceinline__ float red(float a, float b) {{ return {comb}; }}
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    __shared__ float part[{BLOCK // 32}];
    float a0 = {init}, a1 = {init}, a2 = {init}, a3 = {init};
    #pragma unroll 4
    for (size_t c = (size_t)blockIdx.x * {BLOCK}u + threadIdx.x; c < {nv}ull; c += {G * BLOCK}ull) {{ {ld} }}
    float acc = red(red(a0, a1), red(a2, a3));
    #pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc = red(acc, __shfl_xor_sync(0xffffffffu, acc, m));
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {{
        acc = threadIdx.x < {BLOCK // 32} ? part[threadIdx.x] : {init};
        #pragma unroll
        for (int m = {BLOCK // 64}; m > 0; m >>= 1) acc = red(acc, __shfl_xor_sync(0xffffffffu, acc, m));
        if (threadIdx.x == 0) out[blockIdx.x] = acc;
    }}
}}
"""
    return Kernel(f"red_{op}_all", src, (G, 1, 1), (BLOCK, 1, 1), 0, G)


@cache
def argmax(rows: int, cols: int, dim: int) -> Kernel:
    """Index of the first maximum as float (cpu_kernel.argmax, :460-517): strict `>` scan so ties
    resolve to the lowest index; dim 0 / 1 of a 2-D tensor, dim -1 over everything.  One warp per
    output; lanes scan strided slices and merge on (value desc, index asc)."""
    if dim == 1:
        n_out, R, stride_r, stride_o = rows, cols, 1, cols
    elif dim == 0:
        n_out, R, stride_r, stride_o = cols, rows, cols, 1
    else:
        n_out, R, stride_r, stride_o = 1, rows * cols, 1, 0
    BLOCK = 256
    src = f"""
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    const unsigned o = blockIdx.x * {BLOCK // 32}u + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    if (o >= {n_out}u) return;
    const float* base = p0 + (size_t)o * {stride_o}u;
    float best = -3.402823466e+38f; unsigned bi = 0xffffffffu;
    for (unsigned r = lane; r < {R}u; r += 32u) {{
        float v = __ldg(base + (size_t)r * {stride_r}u);
        if (v > best || bi == 0xffffffffu) {{ if (v > best || bi == 0xffffffffu) {{ best = v; bi = r; }} }}
    }}
    #pragma unroll
    for (int m = 16; m > 0; m >>= 1) {{
        float ov = __shfl_xor_sync(0xffffffffu, best, m);
        unsigned oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (oi != 0xffffffffu && (bi == 0xffffffffu || ov > best || (ov == best && oi < bi))) {{ best = ov; bi = oi; }}
    }}
    if (lane == 0) out[o] = (float)bi;
}}
"""
    return Kernel("argmax", src, (_cdiv(n_out, BLOCK // 32), 1, 1), (BLOCK, 1, 1), 0, n_out)


# ------------------------------------------------------------------------------------------------
# permute
# ------------------------------------------------------------------------------------------------
@cache
def transpose_tiled(out_shape: tuple, in_strides: tuple, tdim: int) -> Kernel:
    """Materialised permutation whose innermost *output* axis is strided in the input, while output
    axis `tdim` is the input's unit-stride axis: classic 32x33 shared-memory tile so both the read
    (along tdim) and the write (along the last axis) are coalesced.  cpu_kernel.permute (:180-215)
    does the same gather with a scalar strided loop."""
    nd = len(out_shape)
    ostr = []
    acc = 1
    for s in reversed(out_shape):
        ostr.append(acc)
        acc *= s
    ostr = tuple(reversed(ostr))
    J, Tn = out_shape[-1], out_shape[tdim]
    sj = in_strides[-1]
    batch_dims = [d for d in range(nd - 1) if d != tdim]
    nbatch = 1
    for d in batch_dims:
        nbatch *= out_shape[d]
    lines = ["unsigned bz = blockIdx.z; size_t ib = 0, ob = 0;"]
    for d in reversed(batch_dims):
        lines.append(f"{{ unsigned q = bz % {out_shape[d]}u; bz /= {out_shape[d]}u; "
                     f"ib += (size_t)q * {in_strides[d]}u; ob += (size_t)q * {ostr[d]}u; }}")
    decomp = "\n    ".join(lines)
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG} {{
    __shared__ float tile[32][33];
    {decomp}
    const unsigned j0 = blockIdx.x * 32u, t0 = blockIdx.y * 32u;
    #pragma unroll
    for (int k = 0; k < 32; k += 8) {{
        const unsigned j = j0 + threadIdx.y + k, t = t0 + threadIdx.x;
        if (j < {J}u && t < {Tn}u) tile[threadIdx.y + k][threadIdx.x] = __ldg(p0 + ib + (size_t)j * {sj}u + t);
    }}
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < 32; k += 8) {{
        const unsigned t = t0 + threadIdx.y + k, j = j0 + threadIdx.x;
        if (j < {J}u && t < {Tn}u) out[ob + (size_t)t * {ostr[tdim]}u + j] = tile[threadIdx.x][threadIdx.y + k];
    }}
}}
"""
    numel = nbatch * J * Tn
    return Kernel("transpose", src, (_cdiv(J, 32), _cdiv(Tn, 32), nbatch), (32, 8, 1), 0, numel)


# ------------------------------------------------------------------------------------------------
# gather / scatter family (integer-valued floats index rows; results must be bit-exact)
# ------------------------------------------------------------------------------------------------
@cache
def embedding_fwd(n_idx: int, C: int, aligned: bool = True) -> Kernel:
    """out[i, :] = W[(int)idx[i], :]   (cpu_kernel.embedding, :218-229: row memcpy)."""
    V = 4 if (aligned and C % 4 == 0) else 1
    cv = C // V
    T = "float4" if V == 4 else "float"
    total = n_idx * cv
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG} {{
    const size_t g = (size_t)blockIdx.x * 256u + threadIdx.x;
    if (g >= {total}ull) return;
    const unsigned i = g / {cv}u, c = g % {cv}u;
    const int row = (int)__ldg(p1 + i);
    reinterpret_cast<{T}*>(out)[g] = __ldg(reinterpret_cast<const {T}*>(p0 + (size_t)row * {C}u) + c);
}}
"""
    return Kernel("emb_fwd", src, (_cdiv(total, 256), 1, 1), (256, 1, 1), 0, n_idx * C)


@cache
def embedding_bwd(n_rows: int, n_idx: int, C: int) -> Kernel:
    """dW[row, c] = sum over i with idx[i] == row of g[i, c], accumulated in increasing i — exactly
    the reference's sequential scatter-add order (cpu_kernel.embedding_backward, :232-243), so the
    fp32 result is bit-identical; atomics would not be.  One block per (row, 128-column slab): the
    index list streams through shared memory, matching tokens are compacted in order by a ballot
    scan, then every thread adds its column over the compacted list."""
    CB = 128
    CH = 1024
    src = f"""
extern "C" __global__ void __launch_bounds__({CB}) {KNAME}{SIG} {{
    // p0 = upstream grad (n_idx, C), p1 = idx (n_idx)
    __shared__ unsigned hits[{CH}];
    __shared__ unsigned warp_cnt[{CB // 32}];
    __shared__ unsigned n_hits;
    const unsigned row = blockIdx.x;
    const unsigned c = blockIdx.y * {CB}u + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    float acc = 0.f;
    for (unsigned base = 0; base < {n_idx}u; base += {CH}u) {{
        if (threadIdx.x == 0) n_hits = 0;
        __syncthreads();
        // ordered compaction of matching token positions in this chunk
        for (unsigned s = 0; s < {CH}u; s += {CB}u) {{
            const unsigned i = base + s + threadIdx.x;
            const bool hit = (i < {n_idx}u) && ((int)__ldg(p1 + i) == (int)row);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) warp_cnt[w] = __popc(m);
            __syncthreads();
            unsigned off = n_hits;
            for (unsigned k = 0; k < w; ++k) off += warp_cnt[k];
            if (hit) hits[off + __popc(m & ((1u << lane) - 1u))] = i;
            __syncthreads();
            if (threadIdx.x == 0) {{ unsigned t = 0; for (unsigned k = 0; k < {CB // 32}u; ++k) t += warp_cnt[k]; n_hits += t; }}
            __syncthreads();
        }}
        const unsigned nh = n_hits;
        if (c < {C}u) for (unsigned k = 0; k < nh; ++k) acc += __ldg(p0 + (size_t)hits[k] * {C}u + c);
        __syncthreads();
    }}
    if (c < {C}u) out[(size_t)row * {C}u + c] = acc;
}}
"""
    return Kernel("emb_bwd", src, (n_rows, _cdiv(C, CB), 1), (CB, 1, 1), 0, n_rows * C)


@cache
def ce_forward(n: int, stride0: int) -> Kernel:
    """out[i] = x[(int)t[i] + stride0*i]   (cpu_kernel.ce_forward, :431-441)."""
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG} {{
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i < {n}u) out[i] = __ldg(p0 + (size_t)i * {stride0}u + (int)__ldg(p1 + i));
}}
"""
    return Kernel("ce_fwd", src, (_cdiv(n, 256), 1, 1), (256, 1, 1), 0, n)


@cache
def ce_backward(n: int, stride0: int) -> Kernel:
    """In place: x[(int)t[i] + stride0*i] -= 1   (cpu_kernel.ce_backward, :444-457)."""
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG_ALIAS} {{
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i < {n}u) {{ float* q = p0 + (size_t)i * {stride0}u + (int)p1[i]; *q = *q - 1.f; }}
}}
"""
    return Kernel("ce_bwd", src, (_cdiv(n, 256), 1, 1), (256, 1, 1), 0, 0)


# ------------------------------------------------------------------------------------------------
# matmul, SIMT fallback (exact fp32 FMA chain, sequential in k like the reference's i-k-j loop)
# ------------------------------------------------------------------------------------------------
@cache
def matmul_simt(M: int, N: int, K: int, nb0: int, nb1: int, a_bs: tuple, b_bs: tuple,
                ta: bool = False, tb: bool = False, aligned: bool = True) -> Kernel:
    """C[z] = op(A[z]) @ op(B[z]) for z over (nb0, nb1) batch dims; batch strides of 0 broadcast an
    operand, as the reference's stride-zeroing does (dlgrad/runtime/cpu.py:507-575; kernels
    cpu_kernel.matmul_{2,3,4}d, :352-428).  `ta`/`tb` read A as (K, M) / B as (N, K) so backward
    needs no materialised transpose.  Shared-memory tiled, register-blocked FFMA; every output
    accumulates its K products in increasing k with fused multiply-adds, which is what gcc makes of
    the reference loop — the fp32 result is the same chain.  Used for shapes the tensor-core path
    cannot take (N or K not a multiple of 8, tiny problems)."""
    nb = nb0 * nb1
    # pick the largest tile that still gives every SM a block
    for BM, BN, TM, TN in ((128, 128, 8, 8), (64, 64, 4, 4), (32, 32, 2, 2)):
        if _cdiv(M, BM) * _cdiv(N, BN) * nb >= NUM_SMS or (BM == 32):
            break
    if M <= 32 and N <= 32:
        BM, BN, TM, TN = 32, 32, 2, 2
    BK = 16
    TX, TY = BN // TN, BM // TM
    NT = TX * TY
    lda = M if ta else K
    ldb = K if tb else N
    va = aligned and lda % 4 == 0 and all(s % 4 == 0 for s in a_bs)
    vb = aligned and ldb % 4 == 0 and all(s % 4 == 0 for s in b_bs)
    vc = aligned and N % 4 == 0 and TN % 4 == 0

    def loader(name, smem, rows_dim, R0, RT, ld, transposed_store, vec, Rmax):
        """Tile of `smem`[BK][RT] from a matrix whose tile is (RT x BK).  `transposed_store` means the
        matrix is stored with k contiguous (rows = RT index), else with the RT index contiguous."""
        out = []
        if transposed_store:        # stored [r][k], k contiguous
            if vec:
                n4 = RT * BK // 4
                out.append(f"for (unsigned e = tid; e < {n4}u; e += {NT}u) {{")
                out.append(f"  const unsigned r = e / {BK // 4}u, k4 = (e % {BK // 4}u) * 4u;")
                out.append(f"  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);")
                out.append(f"  if ({R0} + r < {Rmax}u && k0 + k4 < {K}u) v = __ldg(reinterpret_cast<const float4*>({name} + (size_t)({R0} + r) * {ld}u + k0 + k4));")
                out.append(f"  {smem}[k4][r] = v.x; {smem}[k4 + 1][r] = v.y; {smem}[k4 + 2][r] = v.z; {smem}[k4 + 3][r] = v.w;")
                out.append("}")
            else:
                out.append(f"for (unsigned e = tid; e < {RT * BK}u; e += {NT}u) {{")
                out.append(f"  const unsigned r = e / {BK}u, k = e % {BK}u;")
                out.append(f"  {smem}[k][r] = ({R0} + r < {Rmax}u && k0 + k < {K}u) ? __ldg({name} + (size_t)({R0} + r) * {ld}u + k0 + k) : 0.f;")
                out.append("}")
        else:                       # stored [k][r], r contiguous
            if vec:
                n4 = RT * BK // 4
                out.append(f"for (unsigned e = tid; e < {n4}u; e += {NT}u) {{")
                out.append(f"  const unsigned k = e / {RT // 4}u, r4 = (e % {RT // 4}u) * 4u;")
                out.append(f"  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);")
                out.append(f"  if (k0 + k < {K}u && {R0} + r4 < {Rmax}u) v = __ldg(reinterpret_cast<const float4*>({name} + (size_t)(k0 + k) * {ld}u + {R0} + r4));")
                out.append(f"  *reinterpret_cast<float4*>(&{smem}[k][r4]) = v;")
                out.append("}")
            else:
                out.append(f"for (unsigned e = tid; e < {RT * BK}u; e += {NT}u) {{")
                out.append(f"  const unsigned k = e / {RT}u, r = e % {RT}u;")
                out.append(f"  {smem}[k][r] = (k0 + k < {K}u && {R0} + r < {Rmax}u) ? __ldg({name} + (size_t)(k0 + k) * {ld}u + {R0} + r) : 0.f;")
                out.append("}")
        return "\n        ".join(out)

    a_load = loader("A", "As", "m", "m0", BM, lda, not ta, va, M)
    b_load = loader("B", "Bs", "n", "n0", BN, ldb, tb, vb, N)

    # register micro-tile: rows/cols split in two halves of 4 when the micro-tile is 8 wide so the
    # 128-bit shared loads of a quarter-warp hit distinct banks
    def frag(idx_var, T_, half):
        if T_ == 8:
            return [f"{idx_var} * 4u + {i}u" if i < 4 else f"{half}u + {idx_var} * 4u + {i - 4}u" for i in range(8)]
        return [f"{idx_var} * {T_}u + {i}u" for i in range(T_)]

    rows = frag("ty", TM, BM // 2)
    cols = frag("tx", TN, BN // 2)

    def frag_load(dst, smem, T_, idx_var, half):
        if T_ == 8:
            return (f"*reinterpret_cast<float4*>(&{dst}[0]) = *reinterpret_cast<const float4*>(&{smem}[kk][{idx_var} * 4u]); "
                    f"*reinterpret_cast<float4*>(&{dst}[4]) = *reinterpret_cast<const float4*>(&{smem}[kk][{half}u + {idx_var} * 4u]);")
        if T_ == 4:
            return f"*reinterpret_cast<float4*>(&{dst}[0]) = *reinterpret_cast<const float4*>(&{smem}[kk][{idx_var} * 4u]);"
        return " ".join(f"{dst}[{i}] = {smem}[kk][{idx_var} * {T_}u + {i}u];" for i in range(T_))

    stores = []
    for i in range(TM):
        r = f"(m0 + {rows[i]})"
        if vc:
            for j0 in range(0, TN, 4):
                c = f"(n0 + {cols[j0]})"
                stores.append(f"if ({r} < {M}u && {c} < {N}u) *reinterpret_cast<float4*>(C + (size_t){r} * {N}u + {c}) = "
                              f"make_float4(acc[{i}][{j0}], acc[{i}][{j0 + 1}], acc[{i}][{j0 + 2}], acc[{i}][{j0 + 3}]);")
        else:
            for j in range(TN):
                c = f"(n0 + {cols[j]})"
                stores.append(f"if ({r} < {M}u && {c} < {N}u) C[(size_t){r} * {N}u + {c}] = acc[{i}][{j}];")
    store_block = "\n    ".join(stores)

    src = f"""
extern "C" __global__ void __launch_bounds__({NT}) {KNAME}{SIG} {{
    __shared__ __align__(16) float As[{BK}][{BM + 4}];
    __shared__ __align__(16) float Bs[{BK}][{BN + 4}];
    const unsigned tid = threadIdx.x;
    const unsigned z = blockIdx.z, b0 = z / {nb1}u, b1 = z % {nb1}u;
    const float* A = p0 + (size_t)b0 * {a_bs[0]}u + (size_t)b1 * {a_bs[1]}u;
    const float* B = p1 + (size_t)b0 * {b_bs[0]}u + (size_t)b1 * {b_bs[1]}u;
    float* C = out + (size_t)z * {M * N}u;
    const unsigned m0 = blockIdx.y * {BM}u, n0 = blockIdx.x * {BN}u;
    const unsigned ty = tid / {TX}u, tx = tid % {TX}u;
    float acc[{TM}][{TN}];
    #pragma unroll
    for (int i = 0; i < {TM}; ++i)
        #pragma unroll
        for (int j = 0; j < {TN}; ++j) acc[i][j] = 0.f;
    for (unsigned k0 = 0; k0 < {K}u; k0 += {BK}u) {{
        {a_load}
        {b_load}
        __syncthreads();
        #pragma unroll
        for (int kk = 0; kk < {BK}; ++kk) {{
            __align__(16) float a[{TM}]; __align__(16) float b[{TN}];
            {frag_load("a", "As", TM, "ty", BM // 2)}
            {frag_load("b", "Bs", TN, "tx", BN // 2)}
            #pragma unroll
            for (int i = 0; i < {TM}; ++i)
                #pragma unroll
                for (int j = 0; j < {TN}; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }}
        __syncthreads();
    }}
    {store_block}
}}
"""
    grid = (_cdiv(N, BN), _cdiv(M, BM), nb)
    return Kernel("mm_simt", src, grid, (NT, 1, 1), 0, nb * M * N)


# ------------------------------------------------------------------------------------------------
# fused row ops: softmax / log-softmax forward and backward, cross-entropy gradient
# ------------------------------------------------------------------------------------------------
def _row_geometry(R: int, aligned: bool, outer: int):
    V = 4 if (aligned and R % 4 == 0) else 1
    nv = R // V
    T = min(32, _pow2_ceil(_cdiv(nv, 8)))
    if _cdiv(nv, T) > 16:
        T = 256 if _cdiv(nv, 256) <= 16 else 1024
    per = _cdiv(nv, T)
    return V, nv, T, per


def _row_reduce_code(T: int, var: str, comb: str, ident: str, tag: str) -> str:
    """reduce `var` across the T threads that share a row (in place, all threads get the result)."""
    if T <= 32:
        return (f"#pragma unroll\n    for (int m = {T // 2}; m > 0; m >>= 1) "
                f"{{ float o = __shfl_xor_sync(0xffffffffu, {var}, m); {var} = {comb.format(a=var, b='o')}; }}")
    nw = T // 32
    return f"""#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {{ float o = __shfl_xor_sync(0xffffffffu, {var}, m); {var} = {comb.format(a=var, b='o')}; }}
    __shared__ float sm_{tag}[{nw}];
    if ((threadIdx.x & 31) == 0) sm_{tag}[threadIdx.x >> 5] = {var};
    __syncthreads();
    {var} = {ident};
    #pragma unroll
    for (int w = 0; w < {nw}; ++w) {{ float o = sm_{tag}[w]; {var} = {comb.format(a=var, b='o')}; }}"""



def _chain_code(chain: tuple, var: str, comps: list, V: int) -> tuple[str, str]:
    """Code for a chain of fused pointwise ops applied to the row vector `var` (a float4 or float):
        ("scale", fslot)                                   var *= f<fslot>
        ("mfill", pslot, fslot, lead_shape, lead_strides)  var = mask > 0 ? f<fslot> : var, mask = p<pslot>
    Returns (per-row setup code computing the mask row offsets, per-vector code)."""
    setup, body = [], []

    def each(fmt):
        return " ".join(fmt.format(c=("." + c) if c else "") for c in comps)
    for k, op in enumerate(chain):
        if op[0] == "scale":
            body.append(each(var + "{c} = " + var + "{c} * f%d;" % op[1]))
        elif op[0] == "mfill":
            _, ps, fs, lead_shape, lead_strides = op
            lines = [f"size_t moff{k} = 0; {{ unsigned q_ = (live ? r : 0u);"]
            for d in range(len(lead_shape) - 1, -1, -1):
                if lead_shape[d] == 1:
                    continue
                lines.append(f" {{ const unsigned i_ = q_ % {lead_shape[d]}u; q_ /= {lead_shape[d]}u; moff{k} += (size_t)i_ * {lead_strides[d]}u; }}")
            lines.append("}")
            setup.append("".join(lines))
            VT = "float4" if V == 4 else "float"
            body.append(f"{{ const {VT} m_ = __ldg(reinterpret_cast<const {VT}*>(p{ps} + moff{k}) + c); "
                        + each("%s{c} = (m_{c} > 0.f) ? f%d : %s{c};" % (var, fs, var)) + " }")
        else:
            raise ValueError(op)
    return "\n    ".join(setup), " ".join(body)


@cache
def softmax_rows(kind: str, outer: int, R: int, aligned: bool = True, chain: tuple = ()) -> Kernel:
    """Fused softmax / log-softmax over the last axis: one HBM read and one write per element, the
    row lives in registers between the max, the exp-sum and the normalisation.  Replaces the
    reference's 5-kernel chain max / sub / exp / sum / div (dlgrad/tensor.py:634-669) and, for the
    loss, the first six kernels of ops.CrossEntropy.forward (dlgrad/ops.py:343-349).  Masked lanes
    holding -inf come out as exactly 0 (softmax) like the reference's.  `chain` folds pending pointwise
    producers of x into the load — attention's `(q @ k^T) * scale` and `.masked_fill(mask, -inf)`
    (examples/gpt.py:51-53) — so those two full-size tensors are never written."""
    V, nv, T, per = _row_geometry(R, aligned, outer)
    BLOCK = 256 if T <= 256 else T
    rpb = BLOCK // T
    VT = "float4" if V == 4 else "float"
    comps = ["x", "y", "z", "w"] if V == 4 else [""]
    pre_setup, pre_body = _chain_code(chain, "v[i]", comps, V)

    def each(fmt):
        return " ".join(fmt.format(c=("." + c) if c else "") for c in comps)
    load = "v[i] = __ldg(row + c); " + pre_body
    if chain and chain[-1][0] == "mfill":
        # The chain ends in a masked_fill: a vector whose lanes are ALL masked is the fill value whatever x
        # holds, so x is not read there.  Under a causal mask that is the upper triangle of (B,h,T,T): a
        # quarter of this kernel's HBM traffic.
        k_last, (_, ps, fs, _, _) = len(chain) - 1, chain[-1]
        allm = " && ".join(f"mk_{('.' + c) if c else ''} > 0.f" for c in comps)
        fillv = f"make_float4(f{fs}, f{fs}, f{fs}, f{fs})" if V == 4 else f"f{fs}"
        load = (f"const {VT} mk_ = __ldg(reinterpret_cast<const {VT}*>(p{ps} + moff{k_last}) + c); "
                f"if ({allm}) {{ v[i] = {fillv}; }} else {{ {load} }}")
    if kind == "softmax":
        # e * (1/s): one reciprocal per row instead of a division per element (<= 1 ulp from e / s)
        final = each("v[i]{c} = v[i]{c} * inv_s;")
    else:
        final = each("v[i]{c} = (xs[i]{c}) - ls;")
    keep_shift = kind != "softmax"
    src = f"""
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    const unsigned r = blockIdx.x * {rpb}u + threadIdx.x / {T}u;
    const unsigned lane = threadIdx.x % {T}u;
    const bool live = r < {outer}u;
    const {VT}* row = reinterpret_cast<const {VT}*>(p0 + (size_t)(live ? r : 0) * {R}u);
    {pre_setup}
    {VT} v[{per}];
    float mx = -3.402823466e+38f;
    #pragma unroll
    for (int i = 0; i < {per}; ++i) {{
        const unsigned c = lane + i * {T}u;
        if (c < {nv}u) {{ {load} {each("mx = (v[i]{c} > mx) ? v[i]{c} : mx;")} }}
    }}
    {_row_reduce_code(T, "mx", "({b} > {a}) ? {b} : {a}", "-3.402823466e+38f", "mx")}
    float s = 0.f;
    {VT} xs[{per if keep_shift else 1}];
    #pragma unroll
    for (int i = 0; i < {per}; ++i) {{
        const unsigned c = lane + i * {T}u;
        if (c < {nv}u) {{
            {each("v[i]{c} = v[i]{c} - mx;")}
            {"xs[i] = v[i];" if keep_shift else ""}
            {each("v[i]{c} = expf(v[i]{c}); s += v[i]{c};")}
        }}
    }}
    {_row_reduce_code(T, "s", "{a} + {b}", "0.f", "s")}
    {"const float ls = logf(s);" if keep_shift else "const float inv_s = 1.f / s;"}
    if (!live) return;
    {VT}* orow = reinterpret_cast<{VT}*>(out + (size_t)r * {R}u);
    #pragma unroll
    for (int i = 0; i < {per}; ++i) {{
        const unsigned c = lane + i * {T}u;
        if (c < {nv}u) {{ {final} orow[c] = v[i]; }}
    }}
}}
"""
    return Kernel(f"{kind}_rows", src, (_cdiv(outer, rpb), 1, 1), (BLOCK, 1, 1), 0, outer * R)


@cache
def softmax_bwd_rows(kind: str, outer: int, R: int, aligned: bool = True, chain: tuple = ()) -> Kernel:
    """Backward of the fused row ops.  softmax: dx = y * (g - sum(g*y)); log-softmax:
    dx = g - exp(y) * sum(g).  p0 = forward output y, p1 = upstream g.  Algebraically what the
    reference's autograd chain (Div/Exp/Sum/Sub/Max backward, ~17 kernels)
    evaluates; the Max branch there contributes only rounding noise because softmax is
    shift-invariant.  `chain` folds pending pointwise CONSUMERS of dx into the store (the backward of
    masked_fill and of the scalar scale in attention)."""
    V, nv, T, per = _row_geometry(R, aligned, outer)
    BLOCK = 256 if T <= 256 else T
    rpb = BLOCK // T
    VT = "float4" if V == 4 else "float"
    comps = ["x", "y", "z", "w"] if V == 4 else [""]
    post_setup, post_body = _chain_code(chain, "g[i]", comps, V)

    def each(fmt):
        return " ".join(fmt.format(c=("." + c) if c else "") for c in comps)
    gload = "g[i] = __ldg(grow + c);"
    if kind == "softmax":
        accum = each("d += g[i]{c} * y[i]{c};")
        final = each("g[i]{c} = y[i]{c} * (g[i]{c} - d);")
        # where y is exactly 0 in all lanes of a vector (masked attention scores: exp(-inf)), g only ever
        # meets a factor 0: it is not read.  Under a causal mask that is half of g.
        allz = " && ".join(f"y[i]{('.' + c) if c else ''} == 0.f" for c in comps)
        zero = "make_float4(0.f, 0.f, 0.f, 0.f)" if V == 4 else "0.f"
        gload = f"if ({allz}) {{ g[i] = {zero}; }} else {{ g[i] = __ldg(grow + c); }}"
    else:
        accum = each("d += g[i]{c};")
        final = each("g[i]{c} = g[i]{c} - expf(y[i]{c}) * d;")
    src = f"""
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    const unsigned r = blockIdx.x * {rpb}u + threadIdx.x / {T}u;
    const unsigned lane = threadIdx.x % {T}u;
    const bool live = r < {outer}u;
    const size_t ro = (size_t)(live ? r : 0) * {R}u;
    const {VT}* yrow = reinterpret_cast<const {VT}*>(p0 + ro);
    const {VT}* grow = reinterpret_cast<const {VT}*>(p1 + ro);
    {VT} y[{per}], g[{per}];
    float d = 0.f;
    #pragma unroll
    for (int i = 0; i < {per}; ++i) {{
        const unsigned c = lane + i * {T}u;
        if (c < {nv}u) {{ y[i] = __ldg(yrow + c); {gload} {accum} }}
    }}
    {_row_reduce_code(T, "d", "{a} + {b}", "0.f", "d")}
    if (!live) return;
    {post_setup}
    {VT}* orow = reinterpret_cast<{VT}*>(out + ro);
    #pragma unroll
    for (int i = 0; i < {per}; ++i) {{
        const unsigned c = lane + i * {T}u;
        if (c < {nv}u) {{ {final} {post_body} orow[c] = g[i]; }}
    }}
}}
"""
    return Kernel(f"{kind}_bwd_rows", src, (_cdiv(outer, rpb), 1, 1), (BLOCK, 1, 1), 0, outer * R)


@cache
def ce_grad(n: int, V: int, up_is_ptr: bool) -> Kernel:
    """(exp(logsm) - onehot(target)) * upstream / N in one pass — the four kernels of
    ops.CrossEntropy.backward (dlgrad/ops.py:359-370: exp, in-place -1 at the target, mul, div),
    same operation order per element.  p0 = log-softmax (n, V), p1 = target (n,), p2/f0 = upstream,
    f1 = N."""
    total = n * V
    up = "__ldg(p2)" if up_is_ptr else "f0"
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG} {{
    const size_t e = (size_t)blockIdx.x * 256u + threadIdx.x;
    if (e >= {total}ull) return;
    const unsigned i = e / {V}u, j = e % {V}u;
    float t = expf(__ldg(p0 + e));
    if ((int)__ldg(p1 + i) == (int)j) t = t - 1.f;
    out[e] = __fdiv_rn(__fmul_rn(t, {up}), f1);
}}
"""
    return Kernel("ce_grad", src, (_cdiv(total, 256), 1, 1), (256, 1, 1), 0, total)


@cache
def rowdot(rows: int, D: int) -> Kernel:
    """out[r] = sum_d a[r][d] * b[r][d] for D in {32, 64}: D/4 threads per row, float4 loads, shuffle tree.  Used for
    rowsum(dO * O), the per-row term of the softmax backward inside the fused attention backward.  p0 = a, p1 = b."""
    g = D // 4                                        # threads per row
    rpb = 256 // g
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG} {{
    const unsigned r = blockIdx.x * {rpb}u + threadIdx.x / {g}u, t = threadIdx.x % {g}u;
    float s = 0.f;
    if (r < {rows}u) {{
        const float4 a = __ldg(reinterpret_cast<const float4*>(p0) + (size_t)r * {g}u + t);
        const float4 b = __ldg(reinterpret_cast<const float4*>(p1) + (size_t)r * {g}u + t);
        s = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }}
    #pragma unroll
    for (unsigned o = {g // 2}u; o > 0u; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (t == 0u && r < {rows}u) out[r] = s;
}}
"""
    return Kernel("rowdot", src, (_cdiv(rows, rpb), 1, 1), (256, 1, 1), 0, rows)


@cache
def mask_bits(rows: int, cols: int) -> Kernel:
    """The attention mask, digested once per call for the fused attention kernels (the same (T, T) plane serves
    every head).  Output, as uint32 words behind the float* plan output:
        [rows * cols / 32]     bit e of word (r, w) set iff mask[r][32 w + e] > 0
        [rows/128 * cols/128]  non-zero iff that 128 x 128 tile has an UNMASKED entry ("live")
        [rows]                 non-zero iff that row has an unmasked entry
        [1]                    number of rows that have one
    A row's mask is then 2 words per key tile instead of 64 floats, dead tiles are skipped, and the k-sparse GEMMs
    that consume P / dS know whether some row is fully masked (its probabilities are NaN, nothing may be skipped).
    p0 = mask (rows, cols); rows, cols % 128 == 0."""
    wpr = cols // 32
    words = rows * wpr
    nk = cols // 128
    tiles = (rows // 128) * nk
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG} {{
    const unsigned w = blockIdx.x * 256u + threadIdx.x;
    if (w >= {words}u) return;
    const float4* m = reinterpret_cast<const float4*>(p0) + (size_t)w * 8u;
    unsigned bits = 0u;
    #pragma unroll
    for (unsigned j = 0; j < 8u; ++j) {{
        const float4 v = __ldg(m + j);
        bits |= ((v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u)) << (4u * j);
    }}
    unsigned* o = reinterpret_cast<unsigned*>(out);
    o[w] = bits;
    if (bits != 0xffffffffu) {{
        const unsigned r = w / {wpr}u;
        atomicOr(o + {words}u + (r >> 7) * {nk}u + ((w % {wpr}u) >> 2), 1u);
        if (atomicOr(o + {words + tiles}u + r, 1u) == 0u) atomicAdd(o + {words + tiles + rows}u, 1u);
    }}
}}
"""
    return Kernel("mask_bits", src, (_cdiv(words, 256), 1, 1), (256, 1, 1), 0, words + tiles + rows + 1, zero_out=True)


@cache
def cat2(outer: int, la: int, lb: int, aligned: bool = True) -> Kernel:
    """Concatenation of two tensors along one axis, seen as rows: out[o] = a[o] (la elements) ++ b[o] (lb
    elements) for `outer` rows.  No reference kernel: examples/gpt.py appends to its KV cache with
    np.concatenate on the host (:140-147), a D2H + H2D round trip per layer and token.  p0 = a, p1 = b."""
    L = la + lb
    total = outer * L
    V = 4 if (aligned and la % 4 == 0 and lb % 4 == 0) else 1
    T = "float4" if V == 4 else "float"
    nv, lav, lv = total // V, la // V, L // V
    grid = max(1, min(_cdiv(nv, 256), NUM_SMS * 8))
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG} {{
    const {T}* a = reinterpret_cast<const {T}*>(p0);
    const {T}* b = reinterpret_cast<const {T}*>(p1);
    {T}* o = reinterpret_cast<{T}*>(out);
    for (size_t i = (size_t)blockIdx.x * 256u + threadIdx.x; i < {nv}ull; i += (size_t)gridDim.x * 256u) {{
        const size_t r = i / {lv}u;
        const unsigned j = (unsigned)(i - r * {lv}u);
        o[i] = (j < {lav}u) ? __ldg(a + r * {lav}u + j) : __ldg(b + r * {lv - lav}u + (j - {lav}u));
    }}
}}
"""
    return Kernel("cat2", src, (grid, 1, 1), (256, 1, 1), 0, total)


# ------------------------------------------------------------------------------------------------
# optimizer steps (in place)
# ------------------------------------------------------------------------------------------------
def _flit(x: float) -> str:
    """exact float32 literal"""
    import struct
    bits = struct.unpack("<I", struct.pack("<f", x))[0]
    return f"__uint_as_float(0x{bits:08x}u)"


@cache
def adam_step(numel: int, beta1: float, beta2: float, eps: float, aligned: bool = True,
              scalars_in_memory: bool = False) -> Kernel:
    """One fused, in-place Adam update per parameter tensor: reads g, m, v, p and writes m, v, p
    (28 B/param) instead of the reference's ~14 elementwise dispatches with a fresh allocation
    each (dlgrad/nn/optim.py:59-70).  The float32 operation sequence per element is the
    reference's, op for op (no FMA contraction): m*b1 + g*(1-b1); v*b2 + (g*g)*(1-b2);
    m/(1-b1^t); v/(1-b2^t); p - (mh/(sqrt(vh)+eps))*lr.  f0 = 1-b1^t, f1 = 1-b2^t, f2 = lr,
    f3 = gradient scale (1.0 on one GPU — an exact no-op — or 1/world after a data-parallel sum).
    p0 = param, p1 = grad, p2 = m, p3 = v.  With `scalars_in_memory` f0..f2 come from out[0..2] (set_scalars): the
    form recorded into a captured graph."""
    V = 4 if (aligned and numel % 4 == 0) else 1
    nv = numel // V
    VT = "float4" if V == 4 else "float"
    comps = ["x", "y", "z", "w"] if V == 4 else [""]
    b1, b2, e_ = _flit(beta1), _flit(beta2), _flit(eps)
    omb1, omb2 = _flit(1.0 - beta1), _flit(1.0 - beta2)
    body = []
    for c in comps:
        s = ("." + c) if c else ""
        body.append(f"""
        {{ const float gg = __fmul_rn(g{s}, f3);
          const float mn = __fadd_rn(__fmul_rn(m{s}, {b1}), __fmul_rn(gg, {omb1}));
          const float vn = __fadd_rn(__fmul_rn(v{s}, {b2}), __fmul_rn(__fmul_rn(gg, gg), {omb2}));
          const float mh = __fdiv_rn(mn, f0), vh = __fdiv_rn(vn, f1);
          p{s} = __fsub_rn(p{s}, __fmul_rn(__fdiv_rn(mh, __fadd_rn(__fsqrt_rn(vh), {e_})), f2));
          m{s} = mn; v{s} = vn; }}""")
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG_ALIAS} {{
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= {nv}u) return;
    {"f0 = out[0]; f1 = out[1]; f2 = out[2];" if scalars_in_memory else ""}
    {VT} p = reinterpret_cast<{VT}*>(p0)[i];
    const {VT} g = reinterpret_cast<const {VT}*>(p1)[i];
    {VT} m = reinterpret_cast<{VT}*>(p2)[i];
    {VT} v = reinterpret_cast<{VT}*>(p3)[i];
    {''.join(body)}
    reinterpret_cast<{VT}*>(p0)[i] = p;
    reinterpret_cast<{VT}*>(p2)[i] = m;
    reinterpret_cast<{VT}*>(p3)[i] = v;
}}
"""
    return Kernel("adam", src, (_cdiv(nv, 256), 1, 1), (256, 1, 1), 0, 0)


@cache
def sgd_step(numel: int, aligned: bool = True) -> Kernel:
    """In-place p = p - lr*g (dlgrad/nn/optim.py:44: `param - self.lr * grad`, MUL then SUB).
    p0 = param, p1 = grad, f0 = lr, f3 = gradient scale (1.0 = exact no-op)."""
    V = 4 if (aligned and numel % 4 == 0) else 1
    nv = numel // V
    VT = "float4" if V == 4 else "float"
    comps = ["x", "y", "z", "w"] if V == 4 else [""]
    body = " ".join(f"p.{c} = __fsub_rn(p.{c}, __fmul_rn(f0, __fmul_rn(g.{c}, f3)));" if c else "p = __fsub_rn(p, __fmul_rn(f0, __fmul_rn(g, f3)));"
                    for c in comps)
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG_ALIAS} {{
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= {nv}u) return;
    {VT} p = reinterpret_cast<{VT}*>(p0)[i];
    const {VT} g = reinterpret_cast<const {VT}*>(p1)[i];
    {body}
    reinterpret_cast<{VT}*>(p0)[i] = p;
}}
"""
    return Kernel("sgd", src, (_cdiv(nv, 256), 1, 1), (256, 1, 1), 0, 0)


@cache
def sgd_momentum_step(numel: int, aligned: bool = True) -> Kernel:
    """In-place v = v*mu + g; p = p - lr*v — what the reference's momentum branch spells (dlgrad/nn/optim.py:38-44:
    `velocity = momentum * velocity + grad`, then `param - lr * velocity`), same op order as the eager chain
    MUL, ADD, MUL, SUB, with the velocity updated IN PLACE so that the launch arguments never change from step to
    step (a captured step replays it).  p0 = param, p1 = grad, p2 = velocity, f0 = lr, f1 = momentum, f3 = scale."""
    V = 4 if (aligned and numel % 4 == 0) else 1
    nv = numel // V
    VT = "float4" if V == 4 else "float"
    comps = [".x", ".y", ".z", ".w"] if V == 4 else [""]
    body = " ".join(f"v{c} = __fadd_rn(__fmul_rn(v{c}, f1), g{c}); p{c} = __fsub_rn(p{c}, __fmul_rn(f0, __fmul_rn(v{c}, f3)));"
                    for c in comps)
    src = f"""
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG_ALIAS} {{
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= {nv}u) return;
    {VT} p = reinterpret_cast<{VT}*>(p0)[i];
    const {VT} g = reinterpret_cast<const {VT}*>(p1)[i];
    {VT} v = reinterpret_cast<{VT}*>(p2)[i];
    {body}
    reinterpret_cast<{VT}*>(p2)[i] = v;
    reinterpret_cast<{VT}*>(p0)[i] = p;
}}
"""
    return Kernel("sgd_mom", src, (_cdiv(nv, 256), 1, 1), (256, 1, 1), 0, 0)


# ------------------------------------------------------------------------------------------------
# RMSNorm (fused forward)
# ------------------------------------------------------------------------------------------------
@cache
def rmsnorm_rows(outer: int, C: int, affine: bool, aligned: bool = True) -> Kernel:
    """y = (x * rsqrt(mean(x^2) + eps)) * w over the last axis in one pass; the per-row rsqrt is also
    written (to p2) for the backward.  Replaces the six kernels nn.RMSNorm composes (pow, mean, add eps,
    rsqrt, mul, mul — dlgrad/nn/__init__.py:28-41) with the same per-element operation order.
    p0 = x (outer, C), p1 = weight (C,), p2 = rstd out (outer,), f0 = eps."""
    V, nv, T, per = _row_geometry(C, aligned, outer)
    BLOCK = 256 if T <= 256 else T
    rpb = BLOCK // T
    VT = "float4" if V == 4 else "float"
    comps = ["x", "y", "z", "w"] if V == 4 else [""]

    def each(fmt):
        return " ".join(fmt.format(c=("." + c) if c else "") for c in comps)
    wload = f"const {VT} wv = __ldg(reinterpret_cast<const {VT}*>(p1) + c);" if affine else ""
    scale = each("v[i]{c} = __fmul_rn(__fmul_rn(v[i]{c}, t), wv{c});") if affine else each("v[i]{c} = __fmul_rn(v[i]{c}, t);")
    src = f"""
extern "C" __global__ void __launch_bounds__({BLOCK}) {KNAME}{SIG} {{
    const unsigned r = blockIdx.x * {rpb}u + threadIdx.x / {T}u;
    const unsigned lane = threadIdx.x % {T}u;
    const bool live = r < {outer}u;
    const {VT}* row = reinterpret_cast<const {VT}*>(p0 + (size_t)(live ? r : 0) * {C}u);
    {VT} v[{per}];
    float ss = 0.f;
    #pragma unroll
    for (int i = 0; i < {per}; ++i) {{
        const unsigned c = lane + i * {T}u;
        if (c < {nv}u) {{ v[i] = __ldg(row + c); {each("ss += v[i]{c} * v[i]{c};")} }}
    }}
    {_row_reduce_code(T, "ss", "{a} + {b}", "0.f", "ss")}
    const float t = 1.f / sqrtf(__fadd_rn(ss / {C}.f, f0));
    if (!live) return;
    if (lane == 0) const_cast<float*>(p2)[r] = t;
    {VT}* orow = reinterpret_cast<{VT}*>(out + (size_t)r * {C}u);
    #pragma unroll
    for (int i = 0; i < {per}; ++i) {{
        const unsigned c = lane + i * {T}u;
        if (c < {nv}u) {{ {wload} {scale} orow[c] = v[i]; }}
    }}
}}
"""
    return Kernel("rmsnorm_rows", src, (_cdiv(outer, rpb), 1, 1), (BLOCK, 1, 1), 0, outer * C)


@cache
def uniform_philox(numel: int, base_in_memory: bool = False) -> Kernel:
    """out[i] = f0 + (f1 - f0) * u_i, u_i in [0, 1) from Philox4x32-10 (Salmon et al., SC'11; the counter-based generator
    of cuRAND / Random123), replacing the reference's host PCG32 reseeded from /dev/random on every call
    (dlgrad/codegen/cpu_kernel.py:774-874).  Thread t encrypts the counter
    (t, 0, call_lo, call_hi) under the 64-bit key in p0[0..1] and emits elements 4t .. 4t+3 as (x >> 8) * 2^-24.
    The call number — distinct for every uniform() — is f2 + 2^24 * f3 (two exact 24-bit integers in floats); with
    `base_in_memory` (the form recorded into a captured graph) p0[2] is added to it, so that replays, which cannot
    change launch arguments, still draw fresh numbers: a hook rewrites p0[2] before every replay (set_call_base)."""
    nthreads = _cdiv(numel, 4)
    src = f"""
__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {{
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
}}
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG_ALIAS} {{
    const unsigned t = blockIdx.x * 256u + threadIdx.x;
    if (t >= {nthreads}u) return;
    const unsigned* state = reinterpret_cast<const unsigned*>(p0);
    unsigned k0 = state[0], k1 = state[1];
    unsigned long long call = (unsigned long long)f2 + ((unsigned long long)f3 << 24);
    {"call += state[2];" if base_in_memory else ""}
    unsigned c0 = t, c1 = 0u, c2 = (unsigned)call, c3 = (unsigned)(call >> 32);
    #pragma unroll
    for (int r = 0; r < 10; ++r) {{
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }}
    const float span = __fsub_rn(f1, f0);
    const float u[4] = {{ (float)(c0 >> 8) * 5.9604644775390625e-08f, (float)(c1 >> 8) * 5.9604644775390625e-08f,
                         (float)(c2 >> 8) * 5.9604644775390625e-08f, (float)(c3 >> 8) * 5.9604644775390625e-08f }};
    #pragma unroll
    for (unsigned j = 0; j < 4u; ++j) {{
        const unsigned i = 4u * t + j;
        if (i < {numel}u) out[i] = __fadd_rn(f0, __fmul_rn(span, u[j]));
    }}
}}
"""
    return Kernel("uniform_philox", src, (_cdiv(nthreads, 256), 1, 1), (256, 1, 1), 0, numel)


@cache
def sample_rows(rows: int, stride_rows: int, V: int) -> Kernel:
    """out[r] = index drawn from softmax(x[r, :] / f0): the sampling step of examples/gpt.py:262-266 (`logits / temperature`,
    subtract the max, exp, normalise, np.random.choice) as ONE block per row, so the V probabilities never travel to the
    host — only the chosen index does (or nothing at all, when it is fed straight back into the embedding).
    p0 = logits, row r starts at p0 + r * stride_rows (the last token of every sequence of a (B, T, V) tensor);
    p1 = one uniform number in [0, 1) per row (uniform_philox).  Inverse CDF like numpy's choice(): the first index whose
    running sum of e_i exceeds u * sum(e); thread t owns the contiguous chunk [t * per, (t + 1) * per), chunk sums are
    scanned in thread order, so the result is deterministic for a given u."""
    T = 256
    per = _cdiv(V, T)
    src = f"""
extern "C" __global__ void __launch_bounds__({T}) {KNAME}{SIG} {{
    __shared__ float red[{T}];
    __shared__ float pre[{T} + 1];
    const unsigned r = blockIdx.x, t = threadIdx.x;
    const float* x = p0 + (size_t)r * {stride_rows}u;
    const unsigned lo = t * {per}u, hi = min(lo + {per}u, {V}u);
    const float temp = f0;
    float m = __int_as_float(0xff800000);
    for (unsigned i = lo; i < hi; ++i) m = fmaxf(m, __fdiv_rn(x[i], temp));
    red[t] = m;
    __syncthreads();
    for (unsigned s = {T // 2}u; s > 0u; s >>= 1) {{ if (t < s) red[t] = fmaxf(red[t], red[t + s]); __syncthreads(); }}
    m = red[0];
    __syncthreads();
    float mine = 0.f;
    for (unsigned i = lo; i < hi; ++i) mine += expf(__fdiv_rn(x[i], temp) - m);
    red[t] = mine;
    __syncthreads();
    if (t == 0) {{
        float acc = 0.f;
        for (unsigned j = 0; j < {T}u; ++j) {{ pre[j] = acc; acc += red[j]; }}
        pre[{T}] = acc;
    }}
    __syncthreads();
    const float target = p1[r] * pre[{T}];
    // the chunk whose running sum first exceeds the target; the LAST non-empty chunk catches target == total (u -> 1)
    const bool mine_is_it = lo < hi && pre[t] <= target && (target < pre[t + 1] || hi == {V}u);
    if (mine_is_it) {{
        float acc = pre[t];
        unsigned pick = hi - 1u;
        for (unsigned i = lo; i < hi; ++i) {{
            acc += expf(__fdiv_rn(x[i], temp) - m);
            if (target < acc) {{ pick = i; break; }}
        }}
        out[r] = (float)pick;
    }}
}}
"""
    return Kernel("sample_rows", src, (rows, 1, 1), (T, 1, 1), 0, rows)


@cache
def set_call_base() -> Kernel:
    """word 2 of the RNG state at `out` = f0 + 2^24 * f1: the call number the next replay of a captured graph starts from."""
    src = f"""
extern "C" __global__ void __launch_bounds__(32) {KNAME}{SIG_ALIAS} {{
    if (threadIdx.x == 0) reinterpret_cast<unsigned*>(out)[2] = (unsigned)f0 + ((unsigned)f1 << 24);
}}
"""
    return Kernel("set_call_base", src, (1, 1, 1), (32, 1, 1), 0, 0)


@cache
def set_scalars() -> Kernel:
    """out[0..3] = f0..f3: launch arguments parked in device memory, for kernels inside a captured graph whose
    per-step scalars (Adam's bias corrections) must change between replays."""
    src = f"""
extern "C" __global__ void __launch_bounds__(32) {KNAME}{SIG_ALIAS} {{
    if (threadIdx.x == 0) {{ out[0] = f0; out[1] = f1; out[2] = f2; out[3] = f3; }}
}}
"""
    return Kernel("set_scalars", src, (1, 1, 1), (32, 1, 1), 0, 0)


def adam_multi(sizes: tuple, beta1: float, beta2: float, eps: float, scalars_in_memory: bool = False) -> Kernel:
    """The Adam update of EVERY parameter tensor in one launch (multi-tensor apply).  p0 is a device table of
    4 addresses per tensor (param, grad, m, v); the block -> tensor map is a literal prefix array, found by
    binary search.  Same per-element arithmetic as adam_step.  All tensors must be 16-byte aligned with
    numel % 4 == 0 (the caller falls back to per-tensor launches otherwise).  100 launches -> 1 for the
    benchmark GPT (85.9 M parameters in 100 tensors)."""
    CH = 1024                                       # float4 per block
    starts, acc = [0], 0
    for n in sizes:
        acc += _cdiv(n // 4, CH)
        starts.append(acc)
    b1, b2, e_ = _flit(beta1), _flit(beta2), _flit(eps)
    omb1, omb2 = _flit(1.0 - beta1), _flit(1.0 - beta2)
    body = []
    for c in "xyzw":
        body.append(f"""
            {{ const float gg = __fmul_rn(g.{c}, f3);
              const float mn = __fadd_rn(__fmul_rn(m.{c}, {b1}), __fmul_rn(gg, {omb1}));
              const float vn = __fadd_rn(__fmul_rn(v.{c}, {b2}), __fmul_rn(__fmul_rn(gg, gg), {omb2}));
              const float mh = __fdiv_rn(mn, f0), vh = __fdiv_rn(vn, f1);
              p.{c} = __fsub_rn(p.{c}, __fmul_rn(__fdiv_rn(mh, __fadd_rn(__fsqrt_rn(vh), {e_})), f2));
              m.{c} = mn; v.{c} = vn; }}""")
    src = f"""
__device__ const unsigned blk_start[{len(starts)}] = {{{", ".join(str(s) + "u" for s in starts)}}};
__device__ const unsigned vec_count[{len(sizes)}] = {{{", ".join(str(n // 4) + "u" for n in sizes)}}};
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG_ALIAS} {{
    const unsigned long long* tab = reinterpret_cast<const unsigned long long*>(p0);
    {"f0 = p1[0]; f1 = p1[1]; f2 = p1[2];      // bias corrections and learning rate of THIS step (set_scalars), not of the captured one" if scalars_in_memory else ""}
    unsigned lo = 0, hi = {len(sizes)}u;            // tensor t with blk_start[t] <= blockIdx.x < blk_start[t+1]
    while (hi - lo > 1u) {{ const unsigned mid = (lo + hi) >> 1; if (blk_start[mid] <= blockIdx.x) lo = mid; else hi = mid; }}
    const unsigned t = lo, nv = vec_count[t];
    float4* P = reinterpret_cast<float4*>(tab[4u * t]);
    const float4* G = reinterpret_cast<const float4*>(tab[4u * t + 1u]);
    float4* M = reinterpret_cast<float4*>(tab[4u * t + 2u]);
    float4* V = reinterpret_cast<float4*>(tab[4u * t + 3u]);
    const unsigned base = (blockIdx.x - blk_start[t]) * {CH}u;
    #pragma unroll
    for (unsigned k = 0; k < {CH // 256}u; ++k) {{
        const unsigned i = base + k * 256u + threadIdx.x;
        if (i < nv) {{
            float4 p = P[i]; const float4 g = G[i]; float4 m = M[i]; float4 v = V[i];
            {''.join(body)}
            P[i] = p; M[i] = m; V[i] = v;
        }}
    }}
}}
"""
    return Kernel("adam_multi", src, (acc, 1, 1), (256, 1, 1), 0, 0)


def pack_multi(sizes: tuple, offsets: tuple) -> Kernel:
    """Copy every gradient tensor into its slice of the flat bucket in ONE launch (instead of one device-to-device
    copy per tensor): p0 is a device table of source addresses, `out` the bucket, `offsets` the element offset of each
    slice (multiples of 4).  Sources must be 16-byte aligned; a size that is not a multiple of 4 ends in scalar copies."""
    CH = 2048                                       # float4 per block
    starts, acc = [0], 0
    for n in sizes:
        acc += max(1, _cdiv(n // 4, CH))
        starts.append(acc)
    src = f"""
__device__ const unsigned blk_start[{len(starts)}] = {{{", ".join(str(s) + "u" for s in starts)}}};
__device__ const unsigned elem_count[{len(sizes)}] = {{{", ".join(str(n) + "u" for n in sizes)}}};
__device__ const unsigned dst_vec[{len(sizes)}] = {{{", ".join(str(o // 4) + "u" for o in offsets)}}};
extern "C" __global__ void __launch_bounds__(256) {KNAME}{SIG_ALIAS} {{
    const unsigned long long* tab = reinterpret_cast<const unsigned long long*>(p0);
    unsigned lo = 0, hi = {len(sizes)}u;            // tensor t with blk_start[t] <= blockIdx.x < blk_start[t+1]
    while (hi - lo > 1u) {{ const unsigned mid = (lo + hi) >> 1; if (blk_start[mid] <= blockIdx.x) lo = mid; else hi = mid; }}
    const unsigned t = lo, n = elem_count[t], nv = n >> 2;
    const float4* S = reinterpret_cast<const float4*>(tab[t]);
    float4* D = reinterpret_cast<float4*>(out) + dst_vec[t];
    const unsigned base = (blockIdx.x - blk_start[t]) * {CH}u;
    float4 v[{CH // 256}];
    #pragma unroll
    for (unsigned k = 0; k < {CH // 256}u; ++k) {{
        const unsigned i = base + k * 256u + threadIdx.x;
        if (i < nv) v[k] = __ldcs(S + i);
    }}
    #pragma unroll
    for (unsigned k = 0; k < {CH // 256}u; ++k) {{
        const unsigned i = base + k * 256u + threadIdx.x;
        if (i < nv) D[i] = v[k];
    }}
    if (base == 0u && threadIdx.x < (n & 3u)) {{
        const unsigned i = (nv << 2) + threadIdx.x;
        reinterpret_cast<float*>(D)[i] = reinterpret_cast<const float*>(S)[i];
    }}
}}
"""
    return Kernel("pack_multi", src, (acc, 1, 1), (256, 1, 1), 0, 0)


def dp_adam(world: int, shard_vecs: int, beta1: float, beta2: float, eps: float, timeout_s: float = 20.0) -> Kernel:
    """Data-parallel optimizer step as ONE kernel per rank over NVLink peer memory: reduce-scatter of the gradients
    (P2P loads), Adam on this rank's 1/world shard, all-gather of the updated parameters (P2P stores).  It replaces
    ncclAllReduce(grads) + adam_multi ("fused with its collective").

      p0  table of 64-bit words: [world] gradient buckets, [world] parameter buckets, [world] flag arrays — every
          rank's, as mapped into THIS process (own entries are the local allocations) — then this rank's index
      p1, p2  Adam moments of the shard only (shard_vecs float4 each): optimizer state is sharded, not replicated
      p3  this rank's flag array (uint32): [0, world) "rank r's gradients are ready", [world, 2*world) "rank r has
          stored its shard of the new parameters everywhere", [2*world] last finished epoch, [2*world + 1] CTA counter
      f0..f3  bias corrections 1 - beta^t, learning rate, gradient scale (1/world)

    Element i of the shard is sum_{r = 0..world-1} grad_r[i] in RANK ORDER on every rank, and only the owner computes
    the update, so all replicas receive bit-identical parameters (NCCL's ring order differs per rank count).  The
    per-element arithmetic after the sum is adam_multi's.  Flags are epochs: they only grow, nothing is reset.  A rank
    that waits longer than `timeout_s` for a peer traps — a dead peer must fail loudly, not hang the node."""
    W = world
    b1, b2, e_ = _flit(beta1), _flit(beta2), _flit(eps)
    omb1, omb2 = _flit(1.0 - beta1), _flit(1.0 - beta2)
    body = []
    for c in "xyzw":
        body.append(f"""
            {{ const float gg = __fmul_rn(g.{c}, f3);
              const float mn = __fadd_rn(__fmul_rn(m.{c}, {b1}), __fmul_rn(gg, {omb1}));
              const float vn = __fadd_rn(__fmul_rn(v.{c}, {b2}), __fmul_rn(__fmul_rn(gg, gg), {omb2}));
              const float mh = __fdiv_rn(mn, f0), vh = __fdiv_rn(vn, f1);
              p.{c} = __fsub_rn(p.{c}, __fmul_rn(__fdiv_rn(mh, __fadd_rn(__fsqrt_rn(vh), {e_})), f2));
              m.{c} = mn; v.{c} = vn; }}""")
    THREADS = 512
    grid = max(1, min(2 * NUM_SMS, _cdiv(shard_vecs, THREADS)))
    src = f"""
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {{
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}}
__device__ __forceinline__ unsigned long long wall_ns() {{
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}}
__device__ __forceinline__ void wait_epoch(const unsigned* flag, unsigned e) {{
    const unsigned long long t0 = wall_ns();
    while ((int)(ld_acquire_sys(flag) - e) < 0) {{
        if (wall_ns() - t0 > {int(timeout_s * 1e9)}ULL) __trap();
        __nanosleep(200);
    }}
}}
extern "C" __global__ void __launch_bounds__({THREADS}) {KNAME}{SIG_ALIAS} {{
    const unsigned long long* tab = reinterpret_cast<const unsigned long long*>(p0);
    unsigned* sync = reinterpret_cast<unsigned*>(p3);
    const unsigned rank = (unsigned)tab[{3 * W}];
    const unsigned e = *reinterpret_cast<volatile unsigned*>(sync + {2 * W}) + 1u;     // this step's epoch
    __shared__ unsigned long long s_tab[{2 * W}];
    if (threadIdx.x < {2 * W}u) s_tab[threadIdx.x] = tab[threadIdx.x];

    // ---- every rank's gradients are complete (its kernel has started, i.e. its backward has finished) ----
    if (blockIdx.x == 0u && threadIdx.x < {W}u) {{
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned*>(tab[{2 * W}u + threadIdx.x]) + rank, e);
    }}
    if (threadIdx.x < {W}u) wait_epoch(sync + threadIdx.x, e);
    __syncthreads();

    // ---- reduce-scatter (peer loads) -> Adam on the shard -> all-gather (peer stores) ----
    float4* M = reinterpret_cast<float4*>(p1);
    float4* V = reinterpret_cast<float4*>(p2);
    const size_t base = (size_t)rank * {shard_vecs}u;
    for (unsigned i = blockIdx.x * {THREADS}u + threadIdx.x; i < {shard_vecs}u; i += gridDim.x * {THREADS}u) {{
        float4 gr[{W}];
        #pragma unroll
        for (unsigned r = 0; r < {W}u; ++r)
            gr[r] = __ldcg(reinterpret_cast<const float4*>(s_tab[r]) + base + i);
        float4 p = __ldcg(reinterpret_cast<const float4*>(s_tab[{W}u + rank]) + base + i);
        float4 m = M[i], v = V[i];
        float4 g = gr[0];
        #pragma unroll
        for (unsigned r = 1; r < {W}u; ++r) {{
            g.x = __fadd_rn(g.x, gr[r].x); g.y = __fadd_rn(g.y, gr[r].y);
            g.z = __fadd_rn(g.z, gr[r].z); g.w = __fadd_rn(g.w, gr[r].w);
        }}
        {''.join(body)}
        M[i] = m; V[i] = v;
        #pragma unroll
        for (unsigned r = 0; r < {W}u; ++r)
            reinterpret_cast<float4*>(s_tab[{W}u + r])[base + i] = p;
    }}

    // ---- this rank's shard is stored everywhere; leave only when every other shard has arrived here ----
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0u) {{
        __threadfence_system();
        if (atomicAdd(sync + {2 * W + 1}, 1u) == gridDim.x - 1u) {{          // last CTA of this rank
            sync[{2 * W + 1}] = 0u;
            for (unsigned r = 0; r < {W}u; ++r)
                st_release_sys(reinterpret_cast<unsigned*>(tab[{2 * W}u + r]) + {W}u + rank, e);
            for (unsigned r = 0; r < {W}u; ++r) wait_epoch(sync + {W}u + r, e);
            *reinterpret_cast<volatile unsigned*>(sync + {2 * W}) = e;
            __threadfence();
        }}
    }}
}}
"""
    return Kernel("dp_adam", src, (grid, 1, 1), (THREADS, 1, 1), 0, 0)


def _adam_body(beta1: float, beta2: float, eps: float) -> str:
    b1, b2, e_ = _flit(beta1), _flit(beta2), _flit(eps)
    omb1, omb2 = _flit(1.0 - beta1), _flit(1.0 - beta2)
    body = []
    for c in "xyzw":
        body.append(f"""
            {{ const float gg = __fmul_rn(g.{c}, f3);
              const float mn = __fadd_rn(__fmul_rn(m.{c}, {b1}), __fmul_rn(gg, {omb1}));
              const float vn = __fadd_rn(__fmul_rn(v.{c}, {b2}), __fmul_rn(__fmul_rn(gg, gg), {omb2}));
              const float mh = __fdiv_rn(mn, f0), vh = __fdiv_rn(vn, f1);
              p.{c} = __fsub_rn(p.{c}, __fmul_rn(__fdiv_rn(mh, __fadd_rn(__fsqrt_rn(vh), {e_})), f2));
              m.{c} = mn; v.{c} = vn; }}""")
    return "".join(body)


def dp_adam_bucket(world: int, shard_vecs: int, beta1: float, beta2: float, eps: float, threads: int = 128,
                   unroll: int = 0) -> Kernel:
    """dp_adam for ONE exchange bucket of the overlapped schedule (distributed.Adam): the same reduce-scatter (peer
    loads, summed in rank order) + Adam on the shard + all-gather (peer stores), same per-element arithmetic, without
    the in-kernel barriers — the cross-rank arrivals are stream operations around the launch (dp_signal + a stream
    wait on this rank's counter).  p0 is the BUCKET's table ([W] gradient bases, [W] parameter bases, [W] flag
    blocks, rank), p1 / p2 the bucket's slice of the sharded moments.

    Shaped to run NEXT TO the backward of earlier layers, whose persistent tensor-core kernels own one CTA per SM:
    many small short-lived CTAs (128 threads, `unroll` float4 per thread, no grid-stride loop) that fit into the
    registers a GEMM CTA leaves free, each with `world x unroll` 16-byte loads in flight per thread so that the
    ~150 resident CTAs keep NVLink busy; a kernel that needs the whole register file (flash attention) waits for at
    most one such CTA, a few microseconds, not for the exchange."""
    W = world
    U = unroll or max(1, 8 // W)
    chunk = threads * U
    grid = max(1, _cdiv(shard_vecs, chunk))
    src = f"""
extern "C" __global__ void __launch_bounds__({threads}) {KNAME}{SIG_ALIAS} {{
    const unsigned long long* tab = reinterpret_cast<const unsigned long long*>(p0);
    const unsigned rank = (unsigned)__ldg(tab + {3 * W});
    // no shared memory at all: a GEMM CTA of the backward leaves ~1.5 KB of the SM's 228 KB free, and a CTA of this kernel
    // must fit next to it (the 1 KB the system reserves per CTA is all it may take); the 2W table words come through L1
    unsigned long long s_tab[{2 * W}];
    #pragma unroll
    for (unsigned r = 0; r < {2 * W}u; ++r) s_tab[r] = __ldg(tab + r);
    const float4* own = reinterpret_cast<const float4*>(__ldg(tab + {W}u + rank));
    float4* M = reinterpret_cast<float4*>(p1);
    float4* V = reinterpret_cast<float4*>(p2);
    const size_t base = (size_t)rank * {shard_vecs}u;
    const unsigned i0 = blockIdx.x * {chunk}u + threadIdx.x;
    float4 gr[{U}][{W}];
    #pragma unroll
    for (unsigned u = 0; u < {U}u; ++u) {{
        const unsigned i = i0 + u * {threads}u;
        if (i < {shard_vecs}u) {{
            #pragma unroll
            for (unsigned r = 0; r < {W}u; ++r)
                gr[u][r] = __ldcg(reinterpret_cast<const float4*>(s_tab[r]) + base + i);
        }}
    }}
    #pragma unroll
    for (unsigned u = 0; u < {U}u; ++u) {{
        const unsigned i = i0 + u * {threads}u;
        if (i < {shard_vecs}u) {{
            float4 p = __ldcg(own + base + i);
            float4 m = M[i], v = V[i];
            float4 g = gr[u][0];
            #pragma unroll
            for (unsigned r = 1; r < {W}u; ++r) {{
                g.x = __fadd_rn(g.x, gr[u][r].x); g.y = __fadd_rn(g.y, gr[u][r].y);
                g.z = __fadd_rn(g.z, gr[u][r].z); g.w = __fadd_rn(g.w, gr[u][r].w);
            }}
            {_adam_body(beta1, beta2, eps)}
            M[i] = m; V[i] = v;
            #pragma unroll
            for (unsigned r = 0; r < {W}u; ++r)
                reinterpret_cast<float4*>(s_tab[{W}u + r])[base + i] = p;
        }}
    }}
}}
"""
    return Kernel("dp_adam_bucket", src, (grid, 1, 1), (threads, 1, 1), 0, 0)


def dp_signal(world: int) -> Kernel:
    """Cross-rank arrival for the overlapped exchange: add 1 to counter word `f0` of EVERY rank's flag block (p0: the
    table of dp_adam, flag blocks at [2W, 3W)).  Launched on the comm stream after the work it announces, so a
    system-scope fence + release makes that work (this rank's packed gradients, or its stores of new parameters into
    the peers' buckets) visible to whoever observes the count.  The matching wait is a stream memory operation
    (dlg_comm_wait_geq32: count >= epoch * world), which occupies no SM while a peer is late."""
    W = world
    src = f"""
extern "C" __global__ void __launch_bounds__(32) {KNAME}{SIG_ALIAS} {{
    const unsigned long long* tab = reinterpret_cast<const unsigned long long*>(p0);
    const unsigned word = (unsigned)f0;
    if (threadIdx.x < {W}u) {{
        __threadfence_system();
        unsigned* dst = reinterpret_cast<unsigned*>(tab[{2 * W}u + threadIdx.x]) + word;
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(dst), "r"(1u) : "memory");
    }}
}}
"""
    return Kernel("dp_signal", src, (1, 1, 1), (32, 1, 1), 0, 0)


def dp_wait(timeout_s: float = 600.0) -> Kernel:
    """Fallback of dlg_comm_wait_geq32 for drivers without stream memory operations: one thread spins on counter word
    `f0` of this rank's flag block (p0) until (int)(count - target) >= 0.  The kernel's scalar arguments are floats, so the
    32-bit target travels as two exact 16-bit halves, f1 | f2 << 16.  The thread holds an SM slot while a peer is late —
    which is why the stream operation is preferred — and traps after `timeout_s` (a dead peer must not hang the node)."""
    src = f"""
extern "C" __global__ void __launch_bounds__(32) {KNAME}{SIG_ALIAS} {{
    if (threadIdx.x != 0u) return;
    const unsigned* flag = reinterpret_cast<const unsigned*>(p0) + (unsigned)f0;
    const unsigned want = (unsigned)f1 | ((unsigned)f2 << 16);
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {{
        unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - want) >= 0) break;
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > {int(timeout_s * 1e9)}ULL) __trap();
        __nanosleep(500);
    }}
}}
"""
    return Kernel("dp_wait", src, (1, 1, 1), (32, 1, 1), 0, 0)
