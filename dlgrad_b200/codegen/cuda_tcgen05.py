"""Tensor-core matmul generator: tcgen05.mma (kind::tf32) + TMEM accumulator + TMA operand staging,
3xTF32 error-compensated so the result keeps fp32 accuracy (north star: <= 1e-4 relative).

Replaces cpu_kernel.matmul_{2,3,4}d (dlgrad/codegen/cpu_kernel.py:352-428) for shapes whose pitches are
16-byte multiples; everything else goes to cuda_kernel.matmul_simt.

Algorithm per CTA (one 128 x BN output tile, K walked in blocks of 32 fp32 = one 128-byte swizzle span):

    warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) of the fp32 A and B tiles -> full[s]
    warps 2-5   splitter       lo = x - (x & 0xffffe000) into a twin buffer with the SAME swizzled layout
                               (a purely linear pass, no layout knowledge needed) -> split[s].  The staged
                               fp32 tile itself serves as `hi`: tcgen05 kind::tf32 TRUNCATES its fp32
                               inputs to 10 mantissa bits (measured on B200: 1+2^-11+2^-12 -> 1.0,
                               scripts/tc_probe.py), so hi needs no store.
    warp 1      MMA issuer     one elected lane: per 8-wide k step three tcgen05.mma into one TMEM
                               accumulator — A_lo*B_hi, A_hi*B_lo, A_hi*B_hi (small terms first);
                               tcgen05.commit -> empty[s]; last commit -> acc_full
    warps 2-5   epilogue       tcgen05.ld (32 lanes x 32 columns per warp) TMEM -> registers -> C

Operand layouts (element = fp32, T = 4 elements per 16 bytes), after CUTLASS' canonical UMMA layouts:
    K-major  (stored [rows][K], K contiguous): tile = rows x 128 B, 8-row groups 1024 B apart (SBO);
             one TMA box {32, rows}; a k step of 8 advances the descriptor start by 32 B.
    MN-major (stored [K][cols], cols contiguous): tile = (cols/32) boxes of {32 cols, 32 k} = 4 KiB each;
             LBO = 4096 B between 32-column groups, SBO = 1024 B between 8-k groups; a k step of 8
             advances the start by 1024 B.
A is K-major for `x @ y`, MN-major for `x^T @ y`; B is MN-major for `x @ y`, K-major for `x @ y^T`.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import cache

from dlgrad_b200.runtime.cuda.compiler import KNAME

BM = 128
BK = 32


@dataclass(frozen=True)
class TcKernel:
    name: str
    source: str
    grid: tuple
    threads: int
    smem: int
    out_elems: int
    bn: int
    stages: int


def _cdiv(a, b):
    return -(-a // b)


def pick_bn(N: int) -> int:
    bn = min(256, _cdiv(N, 32) * 32)
    if bn > 128 and N % 256 != 0 and N % 128 == 0:
        bn = 128
    return bn


def _idesc(m: int, n: int, a_mn: bool, b_mn: bool) -> int:
    """tcgen05 instruction descriptor (upper 32 bits of the idesc operand): c=F32 [4,6), a/b format TF32=2
    at [7,10)/[10,13), a/b major at 15/16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)."""
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


@cache
def matmul_tc(M: int, N: int, K: int, batch: int, a_mn: bool, b_mn: bool, a_bcast: bool, b_bcast: bool,
              passes: int = 3, bn: int | None = None, stages: int | None = None, hi_inplace: bool = False) -> TcKernel:
    BN = bn or pick_bn(N)
    a_tile = BM * BK * 4                 # 16 KiB
    b_tile = BN * BK * 4
    per_stage = (a_tile + b_tile) * (2 if passes == 3 else 1)
    if stages is None:
        stages = max(2, min(6, (200 * 1024) // per_stage))
    kblocks = _cdiv(K, BK)
    stages = min(stages, max(2, kblocks))
    tmem_cols = 32
    while tmem_cols < BN:
        tmem_cols *= 2
    smem = stages * per_stage + 1024 + 256      # + alignment slack + barriers
    idesc = _idesc(BM, BN, a_mn, b_mn)
    # descriptor constants (units of 16 B): K-major: LBO unused(1), SBO = 1024 B; MN-major: LBO = 4096 B, SBO = 1024 B
    # MN-major fp32/tf32 operands exist only in the SWIZZLE_128B_BASE32B flavour (layout type 1: 32-byte
    # chunks XOR-ed with the row index, period 4 rows -> k-groups of 4 rows, SBO = 512 B); the TMA
    # counterpart is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  K-major uses plain SWIZZLE_128B (type 2).
    a_lbo, a_sbo, a_kstep, a_lt = (256, 32, 64, 1) if a_mn else (1, 64, 2, 2)
    b_lbo, b_sbo, b_kstep, b_lt = (256, 32, 64, 1) if b_mn else (1, 64, 2, 2)
    a_boxes = BM // 32 if a_mn else 1
    b_boxes = BN // 32 if b_mn else 1
    n_conv = 128
    a_v4, b_v4 = a_tile // 16, b_tile // 16

    def tma_issue(which, boxes, mn, tile_var, row0):
        """TMA loads for one stage of operand `which` ('A'/'B'); coordinates are {inner, outer, batch}."""
        tm = f"tm{which}"
        z = "0" if (a_bcast if which == "A" else b_bcast) else "z"
        if mn:      # stored [K][cols]: inner = cols (tile origin row0 + 32*j), outer = k
            return "\n                ".join(
                f"tma_load_3d({tile_var} + {j * 4096}u, &{tm}, full_bar, (int)({row0} + {j * 32}u), (int)k0, (int){z});"
                for j in range(boxes))
        return f"tma_load_3d({tile_var}, &{tm}, full_bar, (int)k0, (int){row0}, (int){z});"

    split_code = ""
    if passes == 3:
        hi_store = "src[i] = hi;" if hi_inplace else ""
        split_code = f"""
        // ===== splitter warps: hi/lo decomposition of the staged fp32 tiles =====
        for (unsigned kb = 0; kb < {kblocks}u; ++kb) {{
            const unsigned s = kb % {stages}u, ph = (kb / {stages}u) & 1u;
            mbar_wait(full + s * 8u, ph);
            float4* src = reinterpret_cast<float4*>(smem_gen + s * {per_stage}u);
            float4* dst = reinterpret_cast<float4*>(smem_gen + s * {per_stage}u + {a_tile + b_tile}u);
            #pragma unroll 4
            for (unsigned i = ctid; i < {a_v4 + b_v4}u; i += {n_conv}u) {{
                const float4 x = src[i];
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); lo.x = x.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); lo.y = x.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); lo.z = x.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); lo.w = x.w - hi.w;
                {hi_store}
                dst[i] = lo;
            }}
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to UMMA
            mbar_arrive(split + s * 8u);
        }}"""

    if passes == 3:
        mma_wait = "mbar_wait(split + s * 8u, ph);"
        mma_body = f"""
                    const unsigned long long a_hi = adesc + (unsigned long long)(k * {a_kstep}u);
                    const unsigned long long b_hi = bdesc + (unsigned long long)(k * {b_kstep}u);
                    const unsigned long long a_lo = a_hi + {(a_tile + b_tile) // 16}ull;
                    const unsigned long long b_lo = b_hi + {(a_tile + b_tile) // 16}ull;
                    umma_tf32(tmem_base, a_lo, b_hi, {idesc}u, (kb | k) != 0u);
                    umma_tf32(tmem_base, a_hi, b_lo, {idesc}u, 1u);
                    umma_tf32(tmem_base, a_hi, b_hi, {idesc}u, 1u);"""
    else:
        mma_wait = "mbar_wait(full + s * 8u, ph);"
        mma_body = f"""
                    umma_tf32(tmem_base, adesc + (unsigned long long)(k * {a_kstep}u),
                              bdesc + (unsigned long long)(k * {b_kstep}u), {idesc}u, (kb | k) != 0u);"""

    vec_store = N % 4 == 0
    if vec_store:
        store = f"""
                    #pragma unroll
                    for (int j = 0; j < 32; j += 4) {{
                        const unsigned col = n0 + c * 32u + j;
                        if (col < {N}u)
                            *reinterpret_cast<float4*>(crow + col) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    }}"""
    else:
        store = f"""
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) {{
                        const unsigned col = n0 + c * 32u + j;
                        if (col < {N}u) crow[col] = __uint_as_float(v[j]);
                    }}"""
    regs = ", ".join(f"%{i}" for i in range(32))
    outs = ", ".join(f'"=r"(v[{i}])' for i in range(32))

    src = f"""
struct alignas(64) TMap {{ unsigned long long opaque[16]; }};

__device__ __forceinline__ unsigned smem_u32(const void* p) {{ return (unsigned)__cvta_generic_to_shared(p); }}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {{
    // try_wait suspends for a HW-defined window; the watchdog turns a protocol bug into a trap
    // (CUDA error) instead of a hung GPU.
    const long long t0 = clock64();
    for (;;) {{
        unsigned ok;
        asm volatile("{{\\n\\t.reg .pred p;\\n\\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\\n\\t"
                     "selp.b32 %0, 1, 0, p;\\n\\t}}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }}
}}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const TMap* map, unsigned bar, int c0, int c1, int c2) {{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {{%3, %4, %5}}], [%2];"
        :: "r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {{
    asm volatile(
        "{{\\n\\t.reg .pred p;\\n\\t"
        "setp.ne.b32 p, %4, 0;\\n\\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\\n\\t}}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}}
__device__ __forceinline__ void umma_commit(unsigned bar) {{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}}
// shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type [61,64)
__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo16, unsigned sbo16, unsigned ltype) {{
    return (unsigned long long)((addr & 0x3ffffu) >> 4) | ((unsigned long long)lbo16 << 16) |
           ((unsigned long long)sbo16 << 32) | (1ull << 46) | ((unsigned long long)ltype << 61);
}}

extern "C" __global__ void __launch_bounds__(192, 1)
{KNAME}(const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmB, const __grid_constant__ TMap tmC, float* __restrict__ C)
{{
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte alignment
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));     // same place as a generic pointer
    const unsigned bars = smem + {stages * per_stage}u;
    const unsigned full = bars, empty = bars + {stages * 8}u, split = bars + {stages * 16}u, acc_full = bars + {stages * 24}u;
    __shared__ unsigned tmem_slot;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned n0 = blockIdx.x * {BN}u, m0 = blockIdx.y * {BM}u, z = blockIdx.z;

    if (threadIdx.x == 0) {{
        for (unsigned s = 0; s < {stages}u; ++s) {{
            mbar_init(full + s * 8u, 1u);
            mbar_init(empty + s * 8u, 1u);
            mbar_init(split + s * 8u, {n_conv}u);
        }}
        mbar_init(acc_full, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    if (warp == 1) {{
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"({tmem_cols}));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_slot;

    if (warp == 0) {{
        // ===== TMA producer =====
        if (lane == 0) {{
            for (unsigned kb = 0; kb < {kblocks}u; ++kb) {{
                const unsigned s = kb % {stages}u, ph = (kb / {stages}u) & 1u;
                mbar_wait(empty + s * 8u, ph ^ 1u);
                const unsigned full_bar = full + s * 8u;
                const unsigned a_s = smem + s * {per_stage}u, b_s = a_s + {a_tile}u;
                const unsigned k0 = kb * {BK}u;
                mbar_expect_tx(full_bar, {a_tile + b_tile}u);
                {tma_issue("A", a_boxes, a_mn, "a_s", "m0")}
                {tma_issue("B", b_boxes, b_mn, "b_s", "n0")}
            }}
        }}
    }} else if (warp == 1) {{
        // ===== MMA issuer (one elected lane) =====
        if (lane == 0) {{
            for (unsigned kb = 0; kb < {kblocks}u; ++kb) {{
                const unsigned s = kb % {stages}u, ph = (kb / {stages}u) & 1u;
                {mma_wait}
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned a_s = smem + s * {per_stage}u, b_s = a_s + {a_tile}u;
                const unsigned long long adesc = make_desc(a_s, {a_lbo}u, {a_sbo}u, {a_lt}u);
                const unsigned long long bdesc = make_desc(b_s, {b_lbo}u, {b_sbo}u, {b_lt}u);
                #pragma unroll
                for (unsigned k = 0; k < {BK // 8}u; ++k) {{{mma_body}
                }}
                umma_commit(empty + s * 8u);          // frees the smem slot when these MMAs retire
            }}
            umma_commit(acc_full);
        }}
    }} else {{
        const unsigned ctid = threadIdx.x - 64u;{split_code}
        // ===== epilogue: TMEM -> registers -> global =====
        mbar_wait(acc_full, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned q = warp & 3u;                       // TMEM lane quarter this warp may read
        const unsigned row = m0 + q * 32u + lane;
        float* crow = C + ((size_t)z * {M}u + row) * {N}u;
        #pragma unroll 1
        for (unsigned c = 0; c < {BN // 32}u; ++c) {{
            unsigned v[32];
            const unsigned taddr = tmem_base + ((q * 32u) << 16) + c * 32u;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {{{regs}}}, [%32];" : {outs} : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < {M}u) {{{store}
            }}
        }}
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {{
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"({tmem_cols}));
    }}
}}
"""
    grid = (_cdiv(N, BN), _cdiv(M, BM), batch)
    return TcKernel(f"mm_tc{passes}", src, grid, 192, smem, batch * M * N, BN, stages)


def tmap_descs(M: int, N: int, K: int, batch: int, a_mn: bool, b_mn: bool, a_bcast: bool, b_bcast: bool, BN: int,
               b_tile_rows: int | None = None):
    """(dims, strides_bytes, box) triples for A, B and C in the shim's dlg_tmap_desc layout."""
    def operand(rows, mn, bcast, tile_rows):
        nb = 1 if bcast else batch
        if mn:     # stored [K][rows]
            return dict(rank=3, dims=(rows, K, nb), strides=(rows * 4, rows * K * 4), box=(32, BK, 1), swizzle=4)
        return dict(rank=3, dims=(K, rows, nb), strides=(K * 4, rows * K * 4), box=(BK, tile_rows, 1), swizzle=3)
    a = operand(M, a_mn, a_bcast, BM)
    b = operand(N, b_mn, b_bcast, b_tile_rows or BN)
    c = dict(rank=3, dims=(N, M, batch), strides=(N * 4, M * N * 4), box=(32, 32, 1), swizzle=3)
    return a, b, c


# =====================================================================================================
# v2: persistent kernel — one CTA per SM walks output tiles; the TMA ring and the hi/lo splitter run
# continuously across tiles, the accumulator is double-buffered in TMEM so the epilogue of tile i
# (TMEM -> registers -> swizzled smem -> TMA store) overlaps the main loop of tile i+1.
# =====================================================================================================
_HELPERS = r"""
struct alignas(64) TMap { unsigned long long opaque[16]; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    // try_wait suspends for a HW-defined window; the watchdog turns a protocol bug into a trap
    // (a CUDA error) instead of a hung GPU.  The clock is only read once the first try has failed.
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const TMap* map, unsigned bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const TMap* map, unsigned src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"((unsigned long long)map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const TMap* map, unsigned src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"((unsigned long long)map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type [61,64)
__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo16, unsigned sbo16, unsigned ltype) {
    return (unsigned long long)((addr & 0x3ffffu) >> 4) | ((unsigned long long)lbo16 << 16) |
           ((unsigned long long)sbo16 << 32) | (1ull << 46) | ((unsigned long long)ltype << 61);
}
"""

NUM_SMS = 148


def pick_bn_persistent(M: int, N: int, batch: int, kblocks: int) -> int:
    """Column-tile width that minimises (tiles per CTA) x (tile cost) on 148 persistent CTAs.  Tile cost
    model: MMA time ~ BN per k-block plus a fixed per-tile part (epilogue / pipeline refill) that weighs more
    when K is short."""
    ntm = _cdiv(M, BM) * batch
    best, best_cost = None, None
    fixed = 32 + (256 // max(1, kblocks))
    for cand in (256, 192, 128, 96, 64, 32):
        if cand > _cdiv(N, 32) * 32 and cand != 32:
            continue
        tiles = ntm * _cdiv(N, cand)
        # wide tiles leave room for only 2 smem stages (hi + lo buffers): measured 207 (BN=256) vs 216
        # (BN=128, 3 stages) TFLOP/s at 4096^3 -> a 6 % handicap on the 2-stage widths when K is long
        depth_penalty = (1.12 if kblocks > 32 else 1.06) if (cand > 128 and kblocks > 8) else 1.0
        cost = _cdiv(tiles, NUM_SMS) * (cand + fixed) * depth_penalty
        if best_cost is None or cost < best_cost:
            best, best_cost = cand, cost
    return best


@cache
def matmul_tc_persistent(M: int, N: int, K: int, batch: int, a_mn: bool, b_mn: bool, a_bcast: bool, b_bcast: bool,
                         passes: int = 3, bn: int | None = None, stages: int | None = None,
                         kc_blocks: int | None = None, split_warps: int = 8) -> TcKernel:
    import os
    trace = os.environ.get("DLGRAD_TC_TRACE", "0") == "1"     # diagnostic: CTA 0 timestamps its first 32 k-blocks
    TR = (lambda kind, cond="true": f"if (blockIdx.x == 0u && it < 32u && {cond}) trace[{kind}][it] = clock64();") if trace else (lambda *a: "")
    kblocks = _cdiv(K, BK)
    # K is accumulated in TMEM in chunks of <= 1024: the tensor core adds into the fp32 accumulator with
    # truncation, a bias that grows with the accumulator's magnitude (8.5e-5 at K = 4096 on U[0,1) inputs,
    # measured).  Chunk partials are combined in global memory by TMA reduce-add (round-to-nearest fp32).
    kc = kc_blocks or (32 if kblocks > 64 else kblocks)
    nchunks = _cdiv(kblocks, kc)
    BN = bn or pick_bn_persistent(M, N, batch, kblocks)
    a_tile, b_tile = BM * BK * 4, BN * BK * 4
    ab = a_tile + b_tile
    per_stage = ab * (2 if passes == 3 else 1)
    EPI_BUFS = 2
    epi_bytes = 4 * EPI_BUFS * 4096                      # 4 epilogue warps x 2 staging boxes of 32x32 fp32
    budget = 227 * 1024 - 1024 - 512 - epi_bytes
    if stages is None:
        stages = max(2, min(6, budget // per_stage))
    ntm, ntn = _cdiv(M, BM), _cdiv(N, BN)
    ntiles = ntm * ntn * batch
    tmem_cols = 32
    while tmem_cols < 2 * BN:
        tmem_cols *= 2
    smem = stages * per_stage + epi_bytes + 1024 + 512
    idesc = _idesc(BM, BN, a_mn, b_mn)
    a_lbo, a_sbo, a_kstep, a_lt = (256, 32, 64, 1) if a_mn else (1, 64, 2, 2)
    b_lbo, b_sbo, b_kstep, b_lt = (256, 32, 64, 1) if b_mn else (1, 64, 2, 2)
    a_boxes = BM // 32 if a_mn else 1
    b_boxes = BN // 32 if b_mn else 1
    grid = min(ntiles, NUM_SMS)
    nsw = split_warps if passes == 3 else 4
    nthreads = 64 + 32 * nsw + 128
    epi_w0 = 2 + nsw                       # first epilogue warp (4 consecutive warps cover the 4 TMEM lane quarters)

    def tma_issue(which, boxes, mn, tile_var, row0):
        tm = f"tm{which}"
        z = "0" if (a_bcast if which == "A" else b_bcast) else "z"
        if mn:
            return "\n                    ".join(
                f"tma_load_3d({tile_var} + {j * 4096}u, &{tm}, full_bar, (int)({row0} + {j * 32}u), (int)k0, (int){z});"
                for j in range(boxes))
        return f"tma_load_3d({tile_var}, &{tm}, full_bar, (int)k0, (int){row0}, (int){z});"

    decode = f"""const unsigned z = t / {ntm * ntn}u, rem = t % {ntm * ntn}u;
            const unsigned m0 = (rem / {ntn}u) * {BM}u, n0 = (rem % {ntn}u) * {BN}u;"""
    if passes == 3:
        mma_wait = "mbar_wait(split + s * 8u, ph);"
        mma_body = f"""
                        const unsigned long long a_hi = adesc + (unsigned long long)(k * {a_kstep}u);
                        const unsigned long long b_hi = bdesc + (unsigned long long)(k * {b_kstep}u);
                        umma_tf32(tacc, a_hi + {ab // 16}ull, b_hi, {idesc}u, (kbl | k) != 0u);
                        umma_tf32(tacc, a_hi, b_hi + {ab // 16}ull, {idesc}u, 1u);
                        umma_tf32(tacc, a_hi, b_hi, {idesc}u, 1u);"""
        split_code = f"""
        // ===== splitter warps: lo = x - trunc_tf32(x) into the twin buffer (same swizzled layout) =====
        const unsigned ctid = threadIdx.x - 64u;
        unsigned it = 0;
        for (unsigned t = blockIdx.x; t < {ntiles}u; t += {grid}u) {{
            for (unsigned kb = 0; kb < {kblocks}u; ++kb, ++it) {{
                const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                mbar_wait(full + s * 8u, ph);
                {TR(1, "ctid == 0u")}
                const float4* src = reinterpret_cast<const float4*>(smem_gen + s * {per_stage}u);
                float4* dst = reinterpret_cast<float4*>(smem_gen + s * {per_stage}u + {ab}u);
                #pragma unroll 4
                for (unsigned i = ctid; i < {ab // 16}u; i += {32 * nsw}u) {{
                    const float4 x = src[i];
                    float4 lo;
                    lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                    lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                    lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                    lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                    dst[i] = lo;
                }}
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(split + s * 8u);
                {TR(2, "ctid == 0u")}
            }}
        }}"""
    else:
        mma_wait = "mbar_wait(full + s * 8u, ph);"
        mma_body = f"""
                        umma_tf32(tacc, adesc + (unsigned long long)(k * {a_kstep}u),
                                  bdesc + (unsigned long long)(k * {b_kstep}u), {idesc}u, (kbl | k) != 0u);"""
        split_code = ""

    regs = ", ".join(f"%{i}" for i in range(32))
    outs = ", ".join(f'"=r"(v[{i}])' for i in range(32))
    stage_writes = "\n                    ".join(
        f"*reinterpret_cast<float4*>(sbuf_gen + lane * 128u + (({j}u ^ (lane & 7u)) << 4)) = "
        f"make_float4(__uint_as_float(v[{4 * j}]), __uint_as_float(v[{4 * j + 1}]), __uint_as_float(v[{4 * j + 2}]), __uint_as_float(v[{4 * j + 3}]));"
        for j in range(8))

    src = _HELPERS + f"""
extern "C" __global__ void __launch_bounds__({nthreads}, 1)
{KNAME}(const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmB, const __grid_constant__ TMap tmC, float* __restrict__ C)
{{
    extern __shared__ unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;        // swizzle atoms need 1024-byte alignment
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    const unsigned epi = smem + {stages * per_stage}u;
    const unsigned bars = epi + {epi_bytes}u;
    const unsigned full = bars, empty = bars + {stages * 8}u, split = bars + {stages * 16}u;
    const unsigned tfull = bars + {stages * 24}u, tempty = tfull + 16u;
    __shared__ unsigned tmem_slot;
    {"__shared__ long long trace[5][32];" if trace else ""}
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    if (threadIdx.x == 0) {{
        for (unsigned s = 0; s < {stages}u; ++s) {{
            mbar_init(full + s * 8u, 1u);
            mbar_init(empty + s * 8u, 1u);
            mbar_init(split + s * 8u, {32 * nsw}u);
        }}
        for (unsigned b = 0; b < 2u; ++b) {{
            mbar_init(tfull + b * 8u, 1u);
            mbar_init(tempty + b * 8u, 4u);
        }}
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    if (warp == 1) {{
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"({tmem_cols}));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_slot;

    if (warp == 0) {{
        // ===== TMA producer =====
        if (lane == 0) {{
            unsigned it = 0;
            for (unsigned t = blockIdx.x; t < {ntiles}u; t += {grid}u) {{
                {decode}
                for (unsigned kb = 0; kb < {kblocks}u; ++kb, ++it) {{
                    const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                    mbar_wait(empty + s * 8u, ph ^ 1u);
                    {TR(0)}
                    const unsigned full_bar = full + s * 8u;
                    const unsigned a_s = smem + s * {per_stage}u, b_s = a_s + {a_tile}u;
                    const unsigned k0 = kb * {BK}u;
                    mbar_expect_tx(full_bar, {ab}u);
                    {tma_issue("A", a_boxes, a_mn, "a_s", "m0")}
                    {tma_issue("B", b_boxes, b_mn, "b_s", "n0")}
                }}
            }}
        }}
    }} else if (warp == 1) {{
        // ===== MMA issuer (one lane) =====
        if (lane == 0) {{
            unsigned it = 0, u = 0;
            for (unsigned t = blockIdx.x; t < {ntiles}u; t += {grid}u) {{
                for (unsigned ch = 0; ch < {nchunks}u; ++ch, ++u) {{
                    const unsigned b = u & 1u;
                    mbar_wait(tempty + b * 8u, ((u >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned tacc = tmem_base + b * {BN}u;
                    const unsigned kend = min((ch + 1u) * {kc}u, {kblocks}u);
                    for (unsigned kb = ch * {kc}u, kbl = 0; kb < kend; ++kb, ++kbl, ++it) {{
                        const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                        {mma_wait}
                        {TR(3)}
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned a_s = smem + s * {per_stage}u, b_s = a_s + {a_tile}u;
                        const unsigned long long adesc = make_desc(a_s, {a_lbo}u, {a_sbo}u, {a_lt}u);
                        const unsigned long long bdesc = make_desc(b_s, {b_lbo}u, {b_sbo}u, {b_lt}u);
                        #pragma unroll
                        for (unsigned k = 0; k < {BK // 8}u; ++k) {{{mma_body}
                        }}
                        umma_commit(empty + s * 8u);
                        {TR(4)}
                    }}
                    umma_commit(tfull + b * 8u);
                }}
            }}
        }}
    }} else if (warp < {epi_w0}u) {{{split_code}
    }} else {{
        // ===== epilogue warps: TMEM -> registers -> swizzled smem box -> TMA store =====
        const unsigned q = warp & 3u;                          // TMEM lane quarter this warp may read
        const unsigned my_epi = epi + (warp - {epi_w0}u) * {EPI_BUFS * 4096}u;
        unsigned char* my_epi_gen = smem_gen + {stages * per_stage}u + (warp - {epi_w0}u) * {EPI_BUFS * 4096}u;
        unsigned u = 0, nstore = 0;
        for (unsigned t = blockIdx.x; t < {ntiles}u; t += {grid}u) {{
            {decode}
            for (unsigned ch = 0; ch < {nchunks}u; ++ch, ++u) {{
                const unsigned b = u & 1u;
                mbar_wait(tfull + b * 8u, (u >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // a later K-chunk is ADDED to what the earlier chunk stored: that store must have landed
                if (ch != 0u && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                #pragma unroll 1
                for (unsigned c = 0; c < {BN // 32}u; ++c) {{
                    unsigned v[32];
                    const unsigned taddr = tmem_base + ((q * 32u) << 16) + b * {BN}u + c * 32u;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {{{regs}}}, [%32];" : {outs} : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (n0 + c * 32u < {N}u && m0 + q * 32u < {M}u) {{
                        const unsigned sb = nstore & {EPI_BUFS - 1}u;
                        // the staging box we are about to overwrite must have been read by its earlier TMA store
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read {EPI_BUFS - 1};" ::: "memory");
                        __syncwarp();
                        unsigned char* sbuf_gen = my_epi_gen + sb * 4096u;
                        {stage_writes}
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {{
                            if (ch == 0u) tma_store_3d(&tmC, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), (int)z);
                            else tma_reduce_add_3d(&tmC, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), (int)z);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }}
                        ++nstore;
                    }}
                }}
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + b * 8u);        // accumulator b may be overwritten
            }}
        }}
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    {"if (blockIdx.x == 0u && threadIdx.x < 160u) reinterpret_cast<long long*>(C)[threadIdx.x] = (&trace[0][0])[threadIdx.x];" if trace else ""}
    if (warp == 1) {{
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"({tmem_cols}));
    }}
}}
"""
    return TcKernel(f"mm_tcp{passes}", src, (grid, 1, 1), nthreads, smem, batch * M * N, BN, stages)


# =====================================================================================================
# v4: persistent kernel with the A operand in TENSOR MEMORY (tcgen05.mma "TS" form).
#
# Pipeline traces of v2 (scripts/tc_trace.py, profiles/r01_tc_pipeline_trace_v2.txt) show its k-block
# period equals the shared-memory traffic at 128 B/clk: the splitter reads and writes every staged byte
# once and the three MMAs of a k-step read every operand byte three times -> 5 x (A + B) bytes per k-block
# (BN=128: 160 KiB = 1280 clk against a 768 clk tensor-pipe floor).  Here the splitter threads own one A row
# each (thread = TMEM lane), read it once from the swizzled TMA tile and write A_hi and A_lo straight into
# TMEM with tcgen05.st; the MMAs take A from TMEM, so A costs 1 x and only B costs 5 x of its bytes
# (BN=128: 96 KiB = 768 clk, level with the tensor pipe), and a stage shrinks from 2(A+B) to A+2B bytes
# (4 stages instead of 3 at BN=128).
# =====================================================================================================
_HELPERS_TS = r"""
__device__ __forceinline__ void tma_load_4d(unsigned dst, const TMap* map, unsigned bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const TMap* map, unsigned src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"((unsigned long long)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// true in exactly one lane of a converged warp; ptxas recognises the elect.sync predicate and issues the
// guarded tcgen05 instructions once, without the per-thread serialisation loop an `if (lane == 0)` gets
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long bdesc,
                                             unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
"""


# k-block period and exposed per-tile epilogue of the TS kernel, in units of the BN=128 k-block period
# (fitted to scripts/tc_probe.py perf4 on B200: 4096^3, the GPT projection shapes and the attention GEMMs)
_TS_PERIOD = {32: 0.75, 64: 0.82, 96: 0.92, 128: 1.0, 192: 1.364, 256: 2.03}
_TS_EPILOGUE = {32: 0.3, 64: 0.5, 96: 0.9, 128: 1.2, 192: 2.5, 256: 3.5}


def wants_streamk(M: int, N: int, K: int, batch: int, bn: int) -> bool:
    """Stream-K pays when whole tiles leave a ragged last wave (2.59 waves of tiles cost 3 tile times) and K is
    short.  Measured (scripts/tc_probe.py sk): 8192x768x768 219 -> 231 TFLOP/s, 8192x768x3072 239 -> 248;
    but 4096^3 263 -> 247 and the batched attention GEMMs lose too — contiguous unit ranges spread the 148
    CTAs over distant tiles and cost L2 reuse.  Inside the GPT step the gain was within run-to-run noise,
    so the automatic choice is OFF unless DLGRAD_TC_STREAMK=1 (an explicit `streamk=True` always works)."""
    import os
    if os.environ.get("DLGRAD_TC_STREAMK", "0") != "1":
        return False
    tiles = _cdiv(M, BM) * _cdiv(N, bn) * batch
    kblocks = _cdiv(K, BK)
    if batch != 1 or tiles <= NUM_SMS or kblocks > 128 or kblocks < 8:
        return False
    waves = tiles / NUM_SMS
    return _cdiv(tiles, NUM_SMS) / waves >= 1.12


def pick_bn_ts(M: int, N: int, batch: int, kblocks: int) -> int:
    """Column-tile width minimising (tiles per persistent CTA) x (k-blocks x period + epilogues); a shape that
    will run stream-K recovers about half of its ragged last wave."""
    ntm = _cdiv(M, BM) * batch
    chunks = _cdiv(kblocks, 32) if kblocks > 64 else 1
    best, best_cost = 128, None
    for cand in (192, 128, 256, 96, 64, 32):
        if cand > _cdiv(N, 32) * 32 and cand != 32:
            continue
        tiles = ntm * _cdiv(N, cand)
        waves = float(_cdiv(tiles, NUM_SMS))
        if wants_streamk(M, N, kblocks * BK, batch, cand):
            waves = 0.5 * (waves + tiles / NUM_SMS)
        cost = waves * (kblocks * _TS_PERIOD[cand] + chunks * _TS_EPILOGUE[cand])
        if best_cost is None or cost < best_cost * 0.999:
            best, best_cost = cand, cost
    return best


def balanced_items(cta: int, grid: int, items: int, batch: int, ntile: int) -> list[int]:
    """Host mirror of the work order the attention kernels and the tile-skipping GEMMs use (`item_at` / `seg_open`):
    item = head * ntile + tile.  Positions run over tiles in DESCENDING tile index (heads inside), CTAs take positions
    round by round in alternating direction.  Under a causal mask tile i costs i + 1 (or ntile - i) units; this order
    keeps the busiest CTA within a few percent of the mean, plain round-robin does not (tests/test_host_logic.py)."""
    out, n = [], 0
    while True:
        p = n * grid + ((grid - 1 - cta) if n & 1 else cta)
        if p >= items:
            return out
        out.append((p % batch) * ntile + (ntile - 1 - p // batch))
        n += 1


@cache
def matmul_tc_ts(M: int, N: int, K: int, batch: int, a_mn: bool, b_mn: bool, a_bcast: bool, b_bcast: bool,
                 bn: int | None = None, stages: int | None = None, kc_blocks: int | None = None,
                 bias: bool = False, issuers: int | None = None, streamk: bool = False, ksparse: str | None = None,
                 a_relu: bool = False, b_relu: bool = False, splitk: int = 0,
                 relu_mask: bool = False, hsplit: int = 0, radd: bool = False, group: tuple | None = None) -> TcKernel:
    """Kernel parameters after C: `bias` (a length-N row the epilogue adds to every output row when `bias` is
    set — the Linear bias; it rides on the plain store) and `flags` (uint32[ntiles * 4], zero, used when
    `streamk` is set).

    streamk: instead of whole tiles round-robin, every CTA owns one CONTIGUOUS range of (tile, k-block) units
    of equal length, so 2.59 waves of tiles cost 2.59 — not 3 — tile times.  A tile cut by a range boundary
    is finished by two CTAs: the one holding its k-tail meets it FIRST in its range, stores the partial sum
    and raises flags[tile][quarter]; the one holding its k-head meets it LAST, waits for the flag (long
    since raised) and TMA-reduce-adds its partial.  Two addends in fixed roles -> the sum is deterministic.
    Requires ntiles >= grid (every range at least one tile long, so no tile has three owners).

    splitk = S > 1: the batch dimension holds S K-slices of ONE product (x^T @ g with K = batch * tokens: both operands
    are stored [K][cols], so slices are a batch).  C is then the scratch [S][M][N] of partial sums, `bias` carries the
    pointer of the real output and `flags` one counter per (output tile, row quarter): the slice that finishes a
    quarter LAST adds the S partials in slice order and writes the result — deterministic, no second kernel.

    relu_mask: `bias` points at a matrix X laid out like C; the epilogue writes X > 0 ? C : 0 — the backward of a relu
    whose output fed the Linear this GEMM differentiates (dX = g @ W masked by the pre-activation), so the separate
    `where(x > 0, g, 0)` pass over three (tokens, 4C) matrices disappears (ops.Linear.backward / ops.Relu.backward).

    a_relu / b_relu: the operand is relu(stored matrix) — a deferred `Tensor.relu()` (Buffer pending kind "relu") —
    applied where the operand passes through registers anyway: in the A splitter before the hi/lo split, in the B
    splitter (which then also writes the clamped value back as the `hi` tile).  The relu kernel and its 2 x 100 MB
    of traffic per MLP block disappear.

    radd: `bias` points at a matrix R laid out like C; the epilogue writes C + R (added once, by the chunk that stores).  This
    is the residual connection `x + proj(h)` of examples/gpt.py:88-89 and the gradient accumulation `grad + g @ W` of
    dlgrad/tensor.py:805 riding on the GEMM that produces the second operand: the separate broadcast-free ADD kernel (three
    passes over a (tokens, C) matrix) becomes one extra read inside the epilogue.

    hsplit = H > 0 (with ksparse): the B operand and C are the per-head matrices of a (B, T, H, D) tensor — the layout the
    q / k / v projections produce and consume — addressed through rank-4 tensor maps (D, T, H, B) with the batch index z
    split into (z % H, z / H): dq = dS @ k and dk = dS^T @ q then read k, q and write their results in place of the
    head split / merge permutes of examples/gpt.py:46-48.

    group = ("n", G): G sibling products  C_g = A @ B_g  that share the A operand — the q / k / v projections of one attention
    block (examples/gpt.py:43-45: three Linear layers applied to the same normed x) — as ONE persistent launch: called with
    batch = G and a_bcast, the batch index z of a tile selects one of G B tensor maps and one of G C tensor maps (separate
    allocations, nothing is copied or concatenated), so three launches of 2.59 tile waves each become one of 7.8.
    group = ("m", G) (with splitk = S): G split-K weight gradients  dW_g = dY_g^T @ x  that share the B operand x — called
    with batch = G x S work items: the batch index selects member g's A tensor map and the K slice, the partial sums go to
    one scratch [G x S][M][N], and the slice that finishes a member's output tile LAST adds that member's S partials in
    slice order and writes member g's own output (`bias`, `fin1`, `fin2`).  Three single-wave launches whose prologue and
    finish are a third of their time become one persistent launch of 2.9 waves.
    group = ("k", G): the sum of G products  C = sum_g A_g @ B_g  — the input gradient of those three layers,
    dq @ Wq + dk @ Wk + dv @ Wv — as ONE launch with the k loop running through the G operand pairs (K = G x K_g): the
    k-block index selects the A and B tensor maps, the accumulator in tensor memory carries the sum, and the two ADD passes
    (or add epilogues) disappear.

    ksparse ("nn" | "tn"): the A operand is an attention matrix (P or dS, query rows x key columns; "tn" = it is read
    transposed) whose 128 x 128 tiles that the mask kills entirely are exactly zero.  `flags` then points at the
    live-tile table of cuda_kernel.mask_bits ([query tiles][key tiles], then one word per query row, then the number
    of query rows that have a live entry): k-blocks inside dead tiles are neither loaded, split nor multiplied —
    under a causal mask 28 of 64 tile pairs.  If some query row has no live entry at all its probabilities are NaN,
    not zero, and nothing is skipped (the NaN must reach the output like in the dense product)."""
    import os
    tmode = os.environ.get("DLGRAD_TC_TRACE", "0")            # diagnostic: CTA 0 timestamps its first 32 k-blocks
    trace = tmode in ("1", "2", "3")                          # 1 = one stamp per pipeline role, 2 = inside the MMA warp,
                                                              # 3 = only kernel begin/end (clock64 + globaltimer -> SM clock)
    exp = os.environ.get("DLGRAD_TC_EXPERIMENT", "")          # diagnostic variants (wrong results): mma1, mma2, nobsplit

    def stamp(kind, cond):
        return f"if (blockIdx.x == 0u && it < 32u && {cond}) trace[{kind}][it] = clock64();"
    TR = (lambda kind, cond="true": stamp(kind, cond)) if tmode == "1" else (lambda *a: "")
    TM = (lambda kind, cond="lane == 0u": stamp(kind, cond)) if tmode == "2" else (lambda *a: "")
    kblocks = _cdiv(K, BK)
    gmode, G = group if group else (None, 1)
    if gmode:
        assert 2 <= G <= 3 and not (streamk or ksparse or relu_mask or hsplit or bias), group
        if gmode == "n":
            assert batch == G and a_bcast and not b_bcast and not radd and not splitk
        elif gmode == "m":
            # G split-K weight gradients  C_g = A_g^T-slices @ B-slices  that share the B operand (x): batch = G x S work items
            assert splitk > 1 and batch == G * splitk and not radd and not a_bcast and not b_bcast
        else:
            assert gmode == "k" and batch == 1 and K % (G * BK) == 0 and not splitk
    kbg = kblocks // G if gmode == "k" else kblocks           # k-blocks per operand pair

    def sel(name, idx):
        """Address of tensor map `name` number `idx` (a runtime 0..G-1): the maps are separate __grid_constant__ parameters."""
        return f"({idx} == 0u ? &tm{name} : " + (f"&tm{name}1)" if G == 2 else f"({idx} == 1u ? &tm{name}1 : &tm{name}2))")
    kc = kc_blocks or (32 if kblocks > 64 else kblocks)       # K chunks of <= 1024, see matmul_tc_persistent
    nchunks = _cdiv(kblocks, kc)
    BN = bn or pick_bn_ts(M, N, batch, kblocks)
    a_tile, b_tile = BM * BK * 4, BN * BK * 4
    per_stage = a_tile + 2 * b_tile                          # raw A | raw B (= B_hi) | B_lo
    EPI_BUFS = 2
    epi_bytes = 4 * EPI_BUFS * 4096
    budget = 227 * 1024 - 1024 - 512 - epi_bytes
    if stages is None:
        stages = max(2, min(6, budget // per_stage))
    # TMEM columns: accumulator(s) first, then 64 per stage (A_hi 32 | A_lo 32)
    acc_bufs = 2 if 2 * BN + 64 * min(stages, 3) <= 512 else 1
    stages = min(stages, (512 - acc_bufs * BN) // 64)
    assert stages >= 2, (BN, stages)
    a_tmem0 = acc_bufs * BN
    # MMA-issuing warps.  Two help the narrow tiles, whose k-block is issue-bound (BN=64: 390 -> 335 ns per
    # k-block); at BN >= 128 the k-block is bound by shared-memory traffic and a second issuer changes nothing
    # (BN=128: 569 vs 546 ns, BN=192: 794 vs 818 ns).
    # One MMA-issuing warp.  A second one taking alternate k-blocks (DLGRAD_TC_ISSUERS=2) is 14 % faster on narrow tiles
    # (BN <= 96 is issue-bound), but MMAs issued by two different warps into one accumulator are not guaranteed to
    # execute in issue order: the fp32 accumulation order — and with it the last bits of a training run — then varies
    # from run to run (measured: two 21-step runs of the GPT differ in the 7th digit of the loss with two issuers, are
    # identical with one).  Reproducibility wins; the reference's CPU path is deterministic.
    NI = issuers or int(os.environ.get("DLGRAD_TC_ISSUERS", "0")) or 1
    tmem_cols = 32
    while tmem_cols < a_tmem0 + 64 * stages:
        tmem_cols *= 2
    ntm, ntn = _cdiv(M, BM), _cdiv(N, BN)
    ntiles = ntm * ntn * batch
    smem = stages * per_stage + epi_bytes + 1024 + 512
    idesc = _idesc(BM, BN, False, b_mn)                      # A comes from TMEM: rows = lanes, k = columns
    b_lbo, b_sbo, b_kstep, b_lt = (256, 32, 64, 1) if b_mn else (1, 64, 2, 2)
    a_boxes = BM // 32 if a_mn else 1
    b_boxes = BN // 32 if b_mn else 1
    grid = min(ntiles, NUM_SMS)
    units = ntiles * kblocks
    if ksparse:
        assert not streamk and nchunks == 1 and M % 128 == 0 and K % 128 == 0 and K <= 4096, (M, K, nchunks)
        kt_n = K // 128                                      # tiles along K
        q_tiles, k_tiles = (M // 128, kt_n) if ksparse == "nn" else (kt_n, M // 128)
        q_rows = q_tiles * 128
        idx = f"mt_ * {k_tiles}u + kt" if ksparse == "nn" else f"kt * {k_tiles}u + mt_"
        sp_decl = f"""
            const unsigned mt_ = (t % {ntm * ntn}u) / {ntn}u;
            unsigned lk = {(1 << kt_n) - 1 if kt_n < 32 else 0xffffffff}u;
            if (sp_on) {{ lk = 0u; for (unsigned kt = 0; kt < {kt_n}u; ++kt) lk |= (__ldg(flags + {idx}) != 0u ? 1u : 0u) << kt; }}
            const unsigned last_kb = lk ? (31u - __clz(lk)) * 4u + 3u : 0xffffffffu;"""
        sp_on = f"const bool sp_on = __ldg(flags + {q_tiles * k_tiles + q_rows}u) == {q_rows}u;      // every query row has a live entry"
        dead = "if (!((lk >> (kb >> 2)) & 1u)) {{ --it; continue; }}"
        dead_mma = "if (!((lk >> (kb >> 2)) & 1u)) {{ --it; --kbl; continue; }}"
        last_kb = "kb == last_kb"
    else:
        sp_decl, sp_on, dead, dead_mma, last_kb = "", "", "", "", "kb + 1u == kend"
    if streamk:
        assert ntiles >= grid and units < (1 << 30), (ntiles, grid)
        seg_open = f"""for (unsigned u = u_begin, t, kb0, kb1; u < u_end; u += kb1 - kb0) {{
            t = u / {kblocks}u; kb0 = u - t * {kblocks}u; kb1 = min({kblocks}u, kb0 + (u_end - u));"""
    elif ksparse:
        # tiles cost what their row of the attention mask leaves alive (under a causal mask m-tile i reads i + 1 of the
        # k-tiles for "nn", ntm - i for "tn").  Plain round-robin hands a CTA the same two m-tiles again and again
        # ({grid} mod {ntm} = {grid % ntm}): the worst CTA does 36 tile-units against 23.4 on average.  So: m-tiles in
        # descending order, CTAs taken in alternating direction from round to round (worst CTA 24 units).
        per_m = batch * ntn
        seg_open = f"""for (unsigned jr = 0, t; ; ++jr) {{
            {{
                const unsigned pos = jr * {grid}u + ((jr & 1u) ? ({grid - 1}u - blockIdx.x) : blockIdx.x);
                if (pos >= {ntiles}u) break;
                const unsigned mt = {ntm - 1}u - pos / {per_m}u, idx = pos % {per_m}u;
                t = (idx / {ntn}u) * {ntm * ntn}u + mt * {ntn}u + idx % {ntn}u;
            }}
            const unsigned kb0 = 0u, kb1 = {kblocks}u;""" + sp_decl
    else:
        seg_open = f"""for (unsigned t = blockIdx.x; t < {ntiles}u; t += {grid}u) {{
            const unsigned kb0 = 0u, kb1 = {kblocks}u;""" + sp_decl
    nthreads = 64 + 128 + 128 + 128 + (32 if NI == 2 else 0)   # producer, MMA | 4 A-split | 4 B-split | 4 epilogue | 2nd MMA issuer
    epi_w0 = 10

    if hsplit:
        assert ksparse and not a_bcast and not b_bcast and N % 32 == 0

    def tma_issue(which, boxes, mn, tile_var, row0):
        tm = f"tm{which}"
        z = "0" if (a_bcast if which == "A" else b_bcast) else "z"
        if gmode == "m":
            # grouped split-K: the batch index z = g * S + slice; A comes from member g's map, B (shared) from the one map,
            # both at batch coordinate `slice`
            tmp = "tmAg" if which == "A" else "&tmB"
            if mn:
                return "\n                    ".join(
                    f"tma_load_3d({tile_var} + {j * 4096}u, {tmp}, full_bar, (int)({row0} + {j * 32}u), (int)k0, (int)sl);"
                    for j in range(boxes))
            return f"tma_load_3d({tile_var}, {tmp}, full_bar, (int)k0, (int){row0}, (int)sl);"
        if (gmode == "n" and which == "B") or gmode == "k":
            # grouped launch: the map was selected by the tile's batch index ("n") or the k-block's operand pair ("k")
            tmp = f"tm{which}g"
            if mn:
                return "\n                    ".join(
                    f"tma_load_3d({tile_var} + {j * 4096}u, {tmp}, full_bar, (int)({row0} + {j * 32}u), (int)k0, 0);"
                    for j in range(boxes))
            return f"tma_load_3d({tile_var}, {tmp}, full_bar, (int)k0, (int){row0}, 0);"
        if hsplit and which == "B":
            zz = f"(int)(z % {hsplit}u), (int)(z / {hsplit}u)"
            if mn:
                return "\n                    ".join(
                    f"tma_load_4d({tile_var} + {j * 4096}u, &{tm}, full_bar, (int)({row0} + {j * 32}u), (int)k0, {zz});"
                    for j in range(boxes))
            return f"tma_load_4d({tile_var}, &{tm}, full_bar, (int)k0, (int){row0}, {zz});"
        if mn:
            return "\n                    ".join(
                f"tma_load_3d({tile_var} + {j * 4096}u, &{tm}, full_bar, (int)({row0} + {j * 32}u), (int)k0, (int){z});"
                for j in range(boxes))
        return f"tma_load_3d({tile_var}, &{tm}, full_bar, (int)k0, (int){row0}, (int){z});"

    decode = f"""const unsigned z = t / {ntm * ntn}u, rem = t % {ntm * ntn}u;
            const unsigned m0 = (rem / {ntn}u) * {BM}u, n0 = (rem % {ntn}u) * {BN}u;"""
    if a_mn:
        # stored [K][M]: box `sq` = 32 k-rows x 32 m-columns (128 B rows, 32-byte chunks XOR-ed with k & 3)
        a_read = """
                const unsigned char* abox = stage_gen + sq * 4096u;
                #pragma unroll
                for (unsigned k = 0; k < 32u; ++k)
                    x[k] = *reinterpret_cast<const float*>(abox + k * 128u + ((((lane >> 3) ^ (k & 3u)) << 5) | ((lane & 7u) << 2)));"""
    else:
        # stored [M][K]: one 128-byte row per thread, 16-byte chunks XOR-ed with row & 7
        a_read = """
                const unsigned char* arow = stage_gen + (sq * 32u + lane) * 128u;
                #pragma unroll
                for (unsigned j = 0; j < 8u; ++j) {
                    const float4 v4 = *reinterpret_cast<const float4*>(arow + ((j ^ (lane & 7u)) << 4));
                    x[4 * j] = v4.x; x[4 * j + 1] = v4.y; x[4 * j + 2] = v4.z; x[4 * j + 3] = v4.w;
                }"""
    regs_in = ", ".join(f"%{i + 1}" for i in range(32))
    hi_ops = ", ".join(f'"r"(hi[{i}])' for i in range(32))
    lo_ops = ", ".join(f'"r"(lo[{i}])' for i in range(32))

    bias_code = "" if not bias else f"""
                        if (plain) {{
                            #pragma unroll
                            for (unsigned j = 0; j < 8u; ++j) {{
                                const unsigned col = n0 + c * 32u + j * 4u;
                                if (col < {N}u) {{                      // N % 4 == 0: a float4 is all in or all out
                                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + col));
                                    v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + bv.x);
                                    v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + bv.y);
                                    v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + bv.z);
                                    v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + bv.w);
                                }}
                            }}
                        }}"""
    mask_code = ""
    if relu_mask:
        assert not bias and not streamk and not splitk and N % 4 == 0
        sel = " ".join(
            f"v[4 * j + {i}] = xv.{c} > 0.f ? v[4 * j + {i}] : 0u;" for i, c in enumerate("xyzw"))
        mask_code = f"""
                        {{
                            // relu backward: keep the gradient where the pre-activation X is positive (NaN -> 0, like where()).
                            // The X tile comes in coalesced (4 rows x 128 B per instruction, 8 loads in flight per lane)
                            // through the staging box; each thread then reads the row it owns.
                            #pragma unroll
                            for (unsigned i = 0; i < 8u; ++i) {{
                                const unsigned r = 4u * i + (lane >> 3), cc = lane & 7u;
                                const unsigned row = m0 + q * 32u + r, col = n0 + c * 32u + cc * 4u;
                                float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (row < {M}u && col < {N}u)
                                    xv = __ldg(reinterpret_cast<const float4*>(bias + ((size_t)z * {M}u + row) * {N}u + col));
                                *reinterpret_cast<float4*>(sbuf_gen + r * 128u + ((cc ^ (r & 7u)) << 4)) = xv;
                            }}
                            __syncwarp();
                            #pragma unroll
                            for (unsigned j = 0; j < 8u; ++j) {{
                                const float4 xv = *reinterpret_cast<const float4*>(sbuf_gen + lane * 128u + ((j ^ (lane & 7u)) << 4));
                                {sel}
                            }}
                        }}"""
    if radd:
        assert not bias and not streamk and not splitk and not relu_mask and N % 4 == 0
        addv = " ".join(
            f"v[4 * j + {i}] = __float_as_uint(__fadd_rn(xv.{c}, __uint_as_float(v[4 * j + {i}])));" for i, c in enumerate("xyzw"))
        mask_code = f"""
                        if (plain) {{
                            // C + R: the R tile comes in coalesced (4 rows x 128 B per instruction, 8 loads in flight per lane)
                            // through the staging box; each thread then adds the row it owns (R + product: the operand order of
                            // the ADD kernel this replaces)
                            #pragma unroll
                            for (unsigned i = 0; i < 8u; ++i) {{
                                const unsigned r = 4u * i + (lane >> 3), cc = lane & 7u;
                                const unsigned row = m0 + q * 32u + r, col = n0 + c * 32u + cc * 4u;
                                float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (row < {M}u && col < {N}u)
                                    xv = __ldg(reinterpret_cast<const float4*>(bias + ((size_t)z * {M}u + row) * {N}u + col));
                                *reinterpret_cast<float4*>(sbuf_gen + r * 128u + ((cc ^ (r & 7u)) << 4)) = xv;
                            }}
                            __syncwarp();
                            #pragma unroll
                            for (unsigned j = 0; j < 8u; ++j) {{
                                const float4 xv = *reinterpret_cast<const float4*>(sbuf_gen + lane * 128u + ((j ^ (lane & 7u)) << 4));
                                {addv}
                            }}
                            __syncwarp();
                        }}"""
    if streamk:
        sk_seg_begin = """const bool seg_head = kb0 == 0u && kb1 < %du;      // k-head of a tile shared with the next CTA
            const bool seg_tail = kb0 != 0u;                      // k-tail of a tile shared with the previous CTA
            unsigned* flag = flags + t * 4u + q;
            if (seg_head) {
                // the tail's partial was stored by the next CTA at the START of its range; then we add ours
                if (lane == 0) {
                    const long long t0 = clock64();
                    unsigned f;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %%0, [%%1];" : "=r"(f) : "l"(flag) : "memory");
                        if (clock64() - t0 > 4000000000LL) __trap();
                    } while (f == 0u);
                    *flag = 0u;                                   // re-armed for the next launch
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                __syncwarp();
            }""" % kblocks
        sk_seg_end = """if (seg_tail && lane == 0) {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // our stores have landed
                asm volatile("fence.proxy.async;" ::: "memory");
                asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flag), "r"(1u) : "memory");
            }"""
    elif splitk > 1:
        assert nchunks == 1 and not bias and batch == (G if gmode == "m" else 1) * splitk and N % 4 == 0
        # rows of the quarter handled per pass: all R x S partial vectors are requested before the first add, so a
        # lane has ~16 independent 16-byte loads in flight (one row at a time is a chain of 32 L2 round trips, ~20 us)
        R = max(1, min(8, 16 // splitk))
        loads = "\n                            ".join(
            f"v[rr][{sl}] = ok ? __ldcg(reinterpret_cast<const float4*>(C + ((size_t)(gz + {sl}u) * {M}u + row) * {N}u + col)) : z4;"
            for sl in range(splitk))
        adds = " ".join(
            f"acc.x += v[rr][{sl}].x; acc.y += v[rr][{sl}].y; acc.z += v[rr][{sl}].z; acc.w += v[rr][{sl}].w;"
            for sl in range(1, splitk))
        sk_seg_begin = "const bool seg_head = false;"
        sk_seg_end = f"""{{
                // split-K: whichever slice finishes this quarter of the output tile LAST adds the {splitk} partial sums, in
                // slice order, and writes the result; the counter re-arms itself for the next launch
                unsigned last = 0u;
                {f"const unsigned g = z / {splitk}u, gz = g * {splitk}u;" if gmode == "m" else "const unsigned gz = 0u;"}
                unsigned* cnt = flags + ({f"g * {ntm * ntn}u + " if gmode == "m" else ""}rem) * 4u + q;
                if (lane == 0) {{
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // this slice's partial has landed
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                    last = (atomicAdd(cnt, 1u) == {splitk - 1}u) ? 1u : 0u;
                }}
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {{
                    __threadfence();
                    float* fin = {("(g == 0u ? const_cast<float*>(bias) : " + ("fin1)" if G == 2 else "(g == 1u ? fin1 : fin2))")) if gmode == "m" else "const_cast<float*>(bias)"};
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (unsigned cv = lane; cv < {BN // 4}u; cv += 32u) {{
                        const unsigned col = n0 + cv * 4u;
                        #pragma unroll 1
                        for (unsigned r0 = 0; r0 < 32u; r0 += {R}u) {{
                            float4 v[{R}][{splitk}];
                            #pragma unroll
                            for (unsigned rr = 0; rr < {R}u; ++rr) {{
                                const unsigned row = m0 + q * 32u + r0 + rr;
                                const bool ok = row < {M}u && col < {N}u;
                                {loads}
                            }}
                            #pragma unroll
                            for (unsigned rr = 0; rr < {R}u; ++rr) {{
                                const unsigned row = m0 + q * 32u + r0 + rr;
                                float4 acc = v[rr][0];
                                {adds}
                                if (row < {M}u && col < {N}u) *reinterpret_cast<float4*>(fin + (size_t)row * {N}u + col) = acc;
                            }}
                        }}
                    }}
                    if (lane == 0) *cnt = 0u;
                }}
            }}"""
    else:
        sk_seg_begin, sk_seg_end = "const bool seg_head = false;", ""
    CSTORE = (f"tma_store_4d(&tmC, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), (int)(z % {hsplit}u), (int)(z / {hsplit}u));"
              if hsplit else "tma_store_3d(tmCg, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), 0);" if gmode == "n"
              else "tma_store_3d(&tmC, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), (int)z);")
    CADD = ("tma_reduce_add_3d(tmCg, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), 0);" if gmode == "n"
            else "tma_reduce_add_3d(&tmC, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), (int)z);")
    zero_tile = "" if not ksparse else f"""if (lk == 0u) {{
                    // every k-block of this output tile lies in a dead attention tile: the product is exactly zero
                    #pragma unroll 1
                    for (unsigned c = 0; c < {BN // 32}u; ++c) {{
                        if (n0 + c * 32u < {N}u && m0 + q * 32u < {M}u) {{
                            const unsigned sb = nstore & {EPI_BUFS - 1}u;
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read {EPI_BUFS - 1};" ::: "memory");
                            __syncwarp();
                            float4* zb = reinterpret_cast<float4*>(my_epi_gen + sb * 4096u);
                            #pragma unroll
                            for (unsigned j = 0; j < 8u; ++j) zb[lane * 8u + j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {{
                                {CSTORE}
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }}
                            ++nstore;
                        }}
                    }}
                    --ua;
                    continue;
                }}"""
    regs = ", ".join(f"%{i}" for i in range(32))
    outs = ", ".join(f'"=r"(v[{i}])' for i in range(32))
    stage_writes = "\n                    ".join(
        f"*reinterpret_cast<float4*>(sbuf_gen + lane * 128u + (({j}u ^ (lane & 7u)) << 4)) = "
        f"make_float4(__uint_as_float(v[{4 * j}]), __uint_as_float(v[{4 * j + 1}]), __uint_as_float(v[{4 * j + 2}]), __uint_as_float(v[{4 * j + 3}]));"
        for j in range(8))

    if gmode == "m":
        extra = [f"tmA{g}" for g in range(1, G)]
    else:
        extra = [f"tmB{g}" for g in range(1, G)] + [f"tm{'C' if gmode == 'n' else 'A'}{g}" for g in range(1, G)] if gmode else []
    gparams = "".join(f"const __grid_constant__ TMap {n}, " for n in extra)
    gtail = "".join(f", float* __restrict__ fin{g}" for g in range(1, G)) if gmode == "m" else ""
    src = _HELPERS + _HELPERS_TS + f"""
extern "C" __global__ void __launch_bounds__({nthreads}, 1)
{KNAME}(const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmB, const __grid_constant__ TMap tmC, {gparams}float* __restrict__ C, const float* __restrict__ bias, unsigned* __restrict__ flags{gtail})
{{
    extern __shared__ unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;        // swizzle atoms need 1024-byte alignment
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    const unsigned epi = smem + {stages * per_stage}u;
    const unsigned bars = epi + {epi_bytes}u;
    const unsigned full = bars, empty = bars + {stages * 8}u, split = bars + {stages * 16}u;
    const unsigned tfull = bars + {stages * 24}u, tempty = tfull + 16u, turn = tempty + 16u;
    __shared__ unsigned tmem_slot;
    {"__shared__ long long trace[5][32];" if trace else ""}
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    {f"const unsigned u_begin = (unsigned)(((unsigned long long)blockIdx.x * {units}ull) / {grid}ull), u_end = (unsigned)(((unsigned long long)(blockIdx.x + 1u) * {units}ull) / {grid}ull);" if streamk else ""}
    {"if (blockIdx.x == 0u && threadIdx.x == 0u) { trace[0][0] = clock64(); unsigned long long g; asm volatile(\"mov.u64 %0, %%globaltimer;\" : \"=l\"(g)); trace[0][1] = (long long)g; }" if tmode == "3" else ""}

    if (threadIdx.x == 0) {{
        for (unsigned s = 0; s < {stages}u; ++s) {{
            mbar_init(full + s * 8u, 1u);
            mbar_init(empty + s * 8u, 1u);
            mbar_init(split + s * 8u, 8u);                 // one arrive per splitter warp
        }}
        for (unsigned b = 0; b < 2u; ++b) {{
            mbar_init(tfull + b * 8u, 1u);
            mbar_init(tempty + b * 8u, 4u);
            mbar_init(turn + b * 8u, 1u);
        }}
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    if (warp == 1) {{
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"({tmem_cols}));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_slot;
    {sp_on}

    if (warp == 0) {{
        // ===== TMA producer: the warp walks the ring together, one elected lane issues =====
        {{
            unsigned it = 0;
            {seg_open}
                {decode}
                for (unsigned kb = kb0; kb < kb1; ++kb, ++it) {{
                    {dead}
                    const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                    mbar_wait(empty + s * 8u, ph ^ 1u);
                    {TR(0, "lane == 0u")}
                    const unsigned full_bar = full + s * 8u;
                    const unsigned a_s = smem + s * {per_stage}u, b_s = a_s + {a_tile}u;
                    {f"const unsigned kg = kb / {kbg}u, k0 = (kb - kg * {kbg}u) * {BK}u; const TMap* tmAg = {sel('A', 'kg')}; const TMap* tmBg = {sel('B', 'kg')};" if gmode == "k" else f"const unsigned k0 = kb * {BK}u;"}
                    {f"const TMap* tmBg = {sel('B', 'z')};" if gmode == "n" else ""}
                    {f"const unsigned g = z / {splitk}u, sl = z - g * {splitk}u; const TMap* tmAg = {sel('A', 'g')};" if gmode == "m" else ""}
                    if (elect_one()) {{
                        mbar_expect_tx(full_bar, {a_tile + b_tile}u);
                        {tma_issue("A", a_boxes, a_mn, "a_s", "m0")}
                        {tma_issue("B", b_boxes, b_mn, "b_s", "n0")}
                    }}
                    __syncwarp();
                }}
            }}
        }}
    }} else if (warp == 1{" || warp == 14" if NI == 2 else ""}) {{
        // ===== MMA issuer(s): the whole warp walks the loop, one elected lane issues.  A_lo*B_hi, A_hi*B_lo, A_hi*B_hi
        //       per 8-wide k step, A from TMEM.  The issuing thread blocks at tcgen05.commit until the MMAs it queued
        //       have drained and then needs ~200 clk to re-arm (loop, try_wait, fence, descriptors); with two issuers
        //       taking alternate k-blocks — `turn` barriers keep the issue order, hence the accumulation order,
        //       fixed — one re-arms while the other's MMAs execute. =====
        {{
            const unsigned me = (warp == 1u) ? 0u : 1u;
            unsigned it = 0, ua = 0, mine = 0;
            const unsigned long long bdesc0 = make_desc(smem + {a_tile}u, {b_lbo}u, {b_sbo}u, {b_lt}u);   // stage 0; a stage is {per_stage // 16} 16-byte units
            const unsigned a_tmem = tmem_base + {a_tmem0}u;
            {seg_open}
                for (unsigned cs = kb0; cs < kb1; cs += {kc}u, ++ua) {{
                    {"if (lk == 0u) { --ua; continue; }      // no live k-block: the epilogue zero-fills, no accumulator is used" if ksparse else ""}
                    const unsigned b = ua % {acc_bufs}u;
                    const unsigned tacc = tmem_base + b * {BN}u;
                    const unsigned kend = min(cs + {kc}u, kb1);
                    for (unsigned kb = cs, kbl = 0; kb < kend; ++kb, ++kbl, ++it) {{
                        {dead_mma}
                        if ({NI}u == 2u && (it & 1u) != me) continue;
                        if (kbl == 0u) {{
                            mbar_wait(tempty + b * 8u, ((ua / {acc_bufs}u) & 1u) ^ 1u);      // epilogue has drained this accumulator
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }}
                        const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                        mbar_wait(split + s * 8u, ph);
                        {TR(3, "lane == 0u")}
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned long long bdesc = bdesc0 + (unsigned long long)(s * {per_stage // 16}u);
                        const unsigned a_hi = a_tmem + s * 64u, a_lo = a_hi + 32u;
                        {"mbar_wait(turn + me * 8u, (mine & 1u) ^ (me ^ 1u));      // the k-block before mine has been issued" if NI == 2 else ""}
                        if (elect_one()) {{
                            #pragma unroll
                            for (unsigned k = 0; k < {BK // 8}u; ++k) {{
                                const unsigned long long b_hi = bdesc + (unsigned long long)(k * {b_kstep}u);
                                umma_tf32_ts(tacc, a_lo + k * 8u, b_hi, {idesc}u, (kbl | k) != 0u);
                                {"" if exp in ("mma1",) else f"umma_tf32_ts(tacc, a_hi + k * 8u, b_hi + {b_tile // 16}ull, {idesc}u, 1u);"}
                                {"" if exp in ("mma1", "mma2") else f"umma_tf32_ts(tacc, a_hi + k * 8u, b_hi, {idesc}u, 1u);"}
                            }}
                            {"mbar_arrive(turn + (me ^ 1u) * 8u);" if NI == 2 else ""}
                            umma_commit(empty + s * 8u);
                            if ({last_kb}) umma_commit(tfull + b * 8u);      // in-order pipe: the other issuer's earlier MMAs are done too
                        }}
                        __syncwarp();
                        {TR(4, "lane == 0u")}
                        ++mine;
                    }}
                }}
            }}
        }}
    }} else if (warp < 6u) {{
        // ===== A-splitter warps: thread = A row = TMEM lane; A_hi and A_lo go straight to tensor memory =====
        const unsigned sq = warp & 3u;                         // the TMEM lane quarter this warp may write
        unsigned it = 0;
        {seg_open}
            for (unsigned kb = kb0; kb < kb1; ++kb, ++it) {{
                {dead}
                const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                mbar_wait(full + s * 8u, ph);
                {TR(1, "threadIdx.x == 64u")}
                const unsigned char* stage_gen = smem_gen + s * {per_stage}u;
                float x[32];{a_read if exp not in ("nosplit",) else "for (unsigned k = 0; k < 32u; ++k) x[k] = (float)k;"}
                unsigned hi[32], lo[32];
                #pragma unroll
                for (unsigned k = 0; k < 32u; ++k) {{
                    {"x[k] = fmaxf(x[k], 0.f);" if a_relu else ""}
                    hi[k] = __float_as_uint(x[k]) & 0xffffe000u;
                    lo[k] = __float_as_uint(x[k] - __uint_as_float(hi[k]));
                }}
                const unsigned ta = tmem_base + ((sq * 32u) << 16) + {a_tmem0}u + s * 64u;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {{{regs_in}}};" :: "r"(ta), {hi_ops} : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {{{regs_in}}};" :: "r"(ta + 32u), {lo_ops} : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(split + s * 8u);
                {TR(2, "threadIdx.x == 64u")}
            }}
        }}
    }} else if (warp < {epi_w0}u) {{
        // ===== B-splitter warps: B_lo = x - trunc_tf32(x) into the twin buffer (same swizzled layout, linear pass) =====
        const unsigned ctid = threadIdx.x - 192u;
        unsigned it = 0;
        {seg_open}
            for (unsigned kb = kb0; kb < kb1; ++kb, ++it) {{
                {dead}
                const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                mbar_wait(full + s * 8u, ph);
                {"float4*" if b_relu else "const float4*"} bsrc = reinterpret_cast<{"float4*" if b_relu else "const float4*"}>(smem_gen + s * {per_stage}u + {a_tile}u);
                float4* bdst = reinterpret_cast<float4*>(smem_gen + s * {per_stage}u + {a_tile + b_tile}u);
                #pragma unroll 8
                for (unsigned i = ctid; i < {0 if exp in ("nobsplit", "nosplit") else b_tile // 16}u; i += 128u) {{
                    {"float4 v4 = bsrc[i]; v4.x = fmaxf(v4.x, 0.f); v4.y = fmaxf(v4.y, 0.f); v4.z = fmaxf(v4.z, 0.f); v4.w = fmaxf(v4.w, 0.f); bsrc[i] = v4;" if b_relu else "const float4 v4 = bsrc[i];"}
                    float4 l4;
                    l4.x = v4.x - __uint_as_float(__float_as_uint(v4.x) & 0xffffe000u);
                    l4.y = v4.y - __uint_as_float(__float_as_uint(v4.y) & 0xffffe000u);
                    l4.z = v4.z - __uint_as_float(__float_as_uint(v4.z) & 0xffffe000u);
                    l4.w = v4.w - __uint_as_float(__float_as_uint(v4.w) & 0xffffe000u);
                    bdst[i] = l4;
                }}
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(split + s * 8u);
            }}
        }}
    }} else if (warp < 14u) {{
        // ===== epilogue warps: TMEM -> registers -> swizzled smem box -> TMA store =====
        const unsigned q = warp & 3u;                          // TMEM lane quarter this warp may read
        const unsigned my_epi = epi + (warp - {epi_w0}u) * {EPI_BUFS * 4096}u;
        unsigned char* my_epi_gen = smem_gen + {stages * per_stage}u + (warp - {epi_w0}u) * {EPI_BUFS * 4096}u;
        unsigned ua = 0, nstore = 0;
        {seg_open}
            {decode}
            {f"const TMap* tmCg = {sel('C', 'z')};" if gmode == "n" else ""}
            {sk_seg_begin}
            for (unsigned cs = kb0; cs < kb1; cs += {kc}u, ++ua) {{
                {zero_tile}
                const unsigned b = ua % {acc_bufs}u;
                const bool plain = (cs == kb0) && !seg_head;       // this chunk STORES; every other one reduce-adds
                mbar_wait(tfull + b * 8u, (ua / {acc_bufs}u) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // a later K-chunk is ADDED to what the earlier chunk stored: that store must have landed
                if (cs != kb0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                #pragma unroll 1
                for (unsigned c = 0; c < {BN // 32}u; ++c) {{
                    unsigned v[32];
                    const unsigned taddr = tmem_base + ((q * 32u) << 16) + b * {BN}u + c * 32u;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {{{regs}}}, [%32];" : {outs} : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (n0 + c * 32u < {N}u && m0 + q * 32u < {M}u) {{{bias_code}
                        const unsigned sb = nstore & {EPI_BUFS - 1}u;
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read {EPI_BUFS - 1};" ::: "memory");
                        __syncwarp();
                        unsigned char* sbuf_gen = my_epi_gen + sb * 4096u;{mask_code}
                        {stage_writes}
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {{
                            if (plain) {{ {CSTORE} }}
                            else {CADD}
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }}
                        ++nstore;
                    }}
                }}
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + b * 8u);        // accumulator b may be overwritten
            }}
            {sk_seg_end}
        }}
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    {"if (blockIdx.x == 0u && threadIdx.x == 0u) { trace[0][2] = clock64(); unsigned long long g; asm volatile(\"mov.u64 %0, %%globaltimer;\" : \"=l\"(g)); trace[0][3] = (long long)g; } __syncthreads();" if tmode == "3" else ""}
    {"if (blockIdx.x == 0u && threadIdx.x < 160u) reinterpret_cast<long long*>(C)[threadIdx.x] = (&trace[0][0])[threadIdx.x];" if trace else ""}
    if (warp == 1) {{
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"({tmem_cols}));
    }}
}}
"""
    return TcKernel("mm_tcts" + ("_b" if bias else "") + ("_sk" if streamk else "") + (f"_sp{ksparse}" if ksparse else "") + ("_ra" if a_relu else "") + ("_rb" if b_relu else "") + (f"_k{splitk}" if splitk > 1 else "") + ("_rm" if relu_mask else "") + ("_hs" if hsplit else "") + ("_ra2" if radd else "") + (f"_g{gmode}{G}" if gmode else ""), src, (grid, 1, 1), nthreads, smem, batch * M * N, BN, stages)


# =====================================================================================================
# Attention scores in one kernel:  P = softmax(masked_fill((q @ k^T) * scale, mask, fill), dim=-1)
#
# The eager chain writes S = q @ k^T (B,h,T,T) to HBM and the fused softmax reads it back: 2 x 403 MB per layer
# at C5.  Here a CTA owns 128 query rows of one head and walks the key tiles TWICE: pass 0 keeps only the
# running row maximum and sum (thread = row, so no cross-thread reduction), pass 1 recomputes the scores,
# normalises and stores P.  Q is split once per work item into TMEM (A operand, TS form, double buffered),
# K tiles stream through the same TMA ring / lo-splitter as matmul_tc_ts, S lives only in TMEM.
# =====================================================================================================
def _f32lit(x: float) -> str:
    import math
    import struct
    if math.isinf(x):
        return "__int_as_float(0xff800000)" if x < 0 else "__int_as_float(0x7f800000)"
    return "__uint_as_float(0x%08xu)" % struct.unpack("<I", struct.pack("<f", x))[0]


@cache
def attn_scores_tc(Tq: int, Tk: int, D: int, batch: int, scale: float, fill: float) -> TcKernel:
    """q (batch, Tq, D) K-major, k (batch, Tk, D) K-major, out (batch, Tq, Tk).  The mask arrives as bits
    (uint32 [Tq][Tk/32], bit set = masked; cuda_kernel.mask_bits).  Kernel parameters: (tmQ, tmK, tmP, P, bits, unused)."""
    import math
    import os
    trace = os.environ.get("DLGRAD_TC_TRACE", "0") == "1"      # diagnostic: CTA 0 timestamps its first 32 key tiles
    nk_ok = Tk // 128 <= 16 and os.environ.get("DLGRAD_ATTN_SKIP", "1") != "0"

    def TR(kind, idx, cond="lane == 0u"):
        return f"if (blockIdx.x == 0u && {idx} < 32u && {cond}) trace[{kind}][{idx}] = clock64();" if trace else ""
    skip = math.isinf(fill) and fill < 0 and nk_ok
    assert Tq % 128 == 0 and Tk % 128 == 0 and Tk <= 2048 and D in (32, 64) and scale > 0      # D <= 64: 2 Q buffers + 2 S buffers fill TMEM
    kbq = D // 32                                   # k-blocks per tile
    BN = 128
    q_bytes = kbq * 16384                           # one Q tile (128 x D fp32), K-major, 16 KiB per k-block
    k_tile = BN * BK * 4                            # 16 KiB: one k-block of a key tile
    per_stage = 2 * k_tile                          # raw (= hi) | lo
    epi_bytes = 8 * 2 * 4096             # two 32x32 staging boxes per row warp
    budget = 227 * 1024 - 1024 - 512 - epi_bytes - q_bytes       # Q: ONE shared-memory tile (it only lives until it is split into TMEM)
    stages = max(2, min(6, budget // per_stage))
    smem = q_bytes + stages * per_stage + epi_bytes + 1024 + 512
    q_tmem0 = 2 * BN                                # S accumulators first (double buffered), then Q hi/lo x 2 buffers
    tmem_need = q_tmem0 + 2 * kbq * 64
    assert tmem_need <= 512, tmem_need
    tmem_cols = 512
    idesc = _idesc(BM, BN, False, False)
    nq, nk = Tq // 128, Tk // 128
    items = batch * nq
    grid = min(items, NUM_SMS)
    nthreads = 64 + 256 + 128           # producer, MMA | 8 row warps (two per TMEM lane quarter, 64 columns each) | 4 K-splitter warps
    # the row warps work in the base-2 domain: softmax(x) = 2^((x - m) log2 e) / sum, one MUFU.EX2 per element
    # instead of expf's ~10 instructions (they, not the tensor pipe, were the bottleneck: 4 us per tile)
    import math
    import numpy as _np
    log2e = float(_np.float32(math.log2(math.e)))
    sc = _f32lit(float(_np.float32(scale) * _np.float32(log2e)))
    fl = _f32lit(fill if math.isinf(fill) else float(_np.float32(fill) * _np.float32(log2e)))
    regs = ", ".join(f"%{i}" for i in range(32))
    outs = ", ".join(f'"=r"(v[{i}])' for i in range(32))
    regs_in = ", ".join(f"%{i + 1}" for i in range(32))
    hi_ops = ", ".join(f'"r"(hi[{i}])' for i in range(32))
    lo_ops = ", ".join(f'"r"(lo[{i}])' for i in range(32))
    stage_writes = "\n                        ".join(
        f"*reinterpret_cast<float4*>(sbuf_gen + lane * 128u + (({j}u ^ (lane & 7u)) << 4)) = "
        f"make_float4(x[{4 * j}], x[{4 * j + 1}], x[{4 * j + 2}], x[{4 * j + 3}]);"
        for j in range(8))
    src = _HELPERS + _HELPERS_TS + """
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
""" + f"""
extern "C" __global__ void __launch_bounds__({nthreads}, 1)
{KNAME}(const __grid_constant__ TMap tmQ, const __grid_constant__ TMap tmK, const __grid_constant__ TMap tmP, float* __restrict__ P, const unsigned* __restrict__ mask, unsigned* __restrict__ unused)
{{
    extern __shared__ unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    const unsigned qs = smem;                                   // the Q tile being split
    const unsigned ring = smem + {q_bytes}u;                // K stages: raw | lo
    const unsigned epi = ring + {stages * per_stage}u;
    const unsigned bars = epi + {epi_bytes}u;
    const unsigned kfull = bars, kempty = bars + {stages * 8}u, ksplit = bars + {stages * 16}u;
    const unsigned sfull = bars + {stages * 24}u, sempty = sfull + 16u, qfull = sempty + 16u, qsplit = qfull + 16u, qfree = qsplit + 16u;
    __shared__ unsigned tmem_slot;
    {"__shared__ long long trace[5][32];" if trace else ""}
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    // key tiles with at least one unmasked entry, as a bit mask per query tile (from cuda_kernel.mask_bits); a tile that
    // is masked out entirely is never loaded, multiplied or read — under a causal mask that is 28 of 64 tile pairs.
    // Only when the fill value is -inf (a masked entry then contributes exactly nothing to the row sum).
    auto live_tiles = [&](unsigned ww) -> unsigned {{
        {"const unsigned* tf = mask + " + str(Tq * Tk // 32) + "u + (ww % " + str(nq) + "u) * " + str(nk) + "u; unsigned lm = 0u; for (unsigned j = 0; j < " + str(nk) + "u; ++j) lm |= (__ldg(tf + j) != 0u ? 1u : 0u) << j; return lm;" if skip else "return " + str((1 << nk) - 1) + "u;"}
    }};
    // work items in the order of CTA `blockIdx.x`: heavy query tiles first (under a causal mask tile q has q + 1 live key
    // tiles), CTAs taken in alternating direction from round to round.  Plain round-robin gives a CTA the same two
    // query tiles over and over ({grid} mod {nq} = {grid % nq}): 36 tile-units on the worst CTA against 23.4 on average.
    auto item_at = [&](unsigned nn) -> unsigned {{
        const unsigned p = nn * {grid}u + ((nn & 1u) ? ({grid - 1}u - blockIdx.x) : blockIdx.x);
        if (p >= {items}u) return {items}u;
        return (p % {batch}u) * {nq}u + ({nq - 1}u - p / {batch}u);
    }};

    if (threadIdx.x == 0) {{
        for (unsigned s = 0; s < {stages}u; ++s) {{
            mbar_init(kfull + s * 8u, 1u);
            mbar_init(kempty + s * 8u, 1u);
            mbar_init(ksplit + s * 8u, 4u);
        }}
        for (unsigned b = 0; b < 2u; ++b) {{
            mbar_init(sfull + b * 8u, 1u);
            mbar_init(sempty + b * 8u, 8u);
            mbar_init(qfull + b * 8u, 1u);
            mbar_init(qsplit + b * 8u, 4u);
            mbar_init(qfree + b * 8u, 1u);
        }}
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    if (warp == 1) {{
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"({tmem_cols}));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_slot;

    if (warp == 0) {{
        // ===== TMA producer: K tiles twice per item; the NEXT item's Q tile is requested between the two passes so
        //       that the row warps can split it while the MMAs run pass 1 =====
        unsigned it = 0, n = 0;
        auto load_q = [&](unsigned nn, unsigned ww) {{
            const unsigned zz = ww / {nq}u, qq0 = (ww % {nq}u) * 128u, qb = nn & 1u;
            mbar_wait(qfree + qb * 8u, ((nn >> 1) & 1u) ^ 1u);          // the MMAs of the item that used this TMEM Q buffer are done
            if (nn != 0u) mbar_wait(qsplit + ((nn - 1u) & 1u) * 8u, ((nn - 1u) >> 1) & 1u);   // the previous Q tile has left shared memory
            if (elect_one()) {{
                mbar_expect_tx(qfull + qb * 8u, {q_bytes}u);
                #pragma unroll
                for (unsigned kk = 0; kk < {kbq}u; ++kk)
                    tma_load_3d(qs + kk * 16384u, &tmQ, qfull + qb * 8u, (int)(kk * 32u), (int)qq0, (int)zz);
            }}
            __syncwarp();
        }};
        if (item_at(0u) < {items}u) load_q(0u, item_at(0u));
        for (unsigned w = item_at(n); w < {items}u; w = item_at(++n)) {{
            const unsigned z = w / {nq}u;
            const unsigned lm = live_tiles(w);
            for (unsigned pass = 0; pass < 2u; ++pass) {{
                if (pass == 1u && item_at(n + 1u) < {items}u) load_q(n + 1u, item_at(n + 1u));
                for (unsigned j = 0; j < {nk}u; ++j) {{
                    if (!((lm >> j) & 1u)) continue;
                    for (unsigned kk = 0; kk < {kbq}u; ++kk, ++it) {{
                        const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                        mbar_wait(kempty + s * 8u, ph ^ 1u);
                        {TR(4, f"(it / {kbq}u)", "lane == 0u && kk == 0u")}
                        if (elect_one()) {{
                            mbar_expect_tx(kfull + s * 8u, {k_tile}u);
                            tma_load_3d(ring + s * {per_stage}u, &tmK, kfull + s * 8u, (int)(kk * 32u), (int)(j * 128u), (int)z);
                        }}
                        __syncwarp();
                    }}
                }}
            }}
        }}
    }} else if (warp == 1) {{
        // ===== MMA issuer: S[b] = Q_hi/lo (TMEM) x K_hi/lo (smem), 3 MMAs per 8-wide k step =====
        unsigned it = 0, u = 0, n = 0;
        const unsigned long long bdesc0 = make_desc(ring, 1u, 64u, 2u);
        for (unsigned w = item_at(n); w < {items}u; w = item_at(++n)) {{
            const unsigned qb = n & 1u;
            mbar_wait(qsplit + qb * 8u, (n >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned q_t = tmem_base + {q_tmem0}u + qb * {kbq * 64}u;
            const unsigned lm = live_tiles(w);
            for (unsigned pj = 0; pj < {2 * nk}u; ++pj) {{
                if (!((lm >> (pj % {nk}u)) & 1u)) continue;
                const unsigned b = u & 1u;
                mbar_wait(sempty + b * 8u, ((u >> 1) & 1u) ^ 1u);
                {TR(0, "u")}
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tacc = tmem_base + b * {BN}u;
                for (unsigned kk = 0; kk < {kbq}u; ++kk, ++it) {{
                    const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                    mbar_wait(ksplit + s * 8u, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned long long bdesc = bdesc0 + (unsigned long long)(s * {per_stage // 16}u);
                    const unsigned a_hi = q_t + kk * 64u, a_lo = a_hi + 32u;
                    if (elect_one()) {{
                        #pragma unroll
                        for (unsigned k = 0; k < 4u; ++k) {{
                            const unsigned long long b_hi = bdesc + (unsigned long long)(k * 2u);
                            umma_tf32_ts(tacc, a_lo + k * 8u, b_hi, {idesc}u, (kk | k) != 0u);
                            umma_tf32_ts(tacc, a_hi + k * 8u, b_hi + {k_tile // 16}ull, {idesc}u, 1u);
                            umma_tf32_ts(tacc, a_hi + k * 8u, b_hi, {idesc}u, 1u);
                        }}
                        umma_commit(kempty + s * 8u);
                        if (kk + 1u == {kbq}u) umma_commit(sfull + b * 8u);
                    }}
                    __syncwarp();
                }}
                {TR(1, "u")}
                ++u;
            }}
            if (elect_one()) umma_commit(qfree + qb * 8u);
            __syncwarp();
        }}
    }} else if (warp < 10u) {{
        // ===== row warps: thread = query row = TMEM lane; warps 2-5 take columns 0-63 of every key tile, warps 6-9
        //       columns 64-127 (the row work — TMEM reads, exponentials, stores — is what paces this kernel).  Split Q into TMEM; pass 0: running max / sum over the
        //       key tiles; pass 1: normalise and store P =====
        const unsigned sq = warp & 3u;
        const unsigned half = (warp >= 6u) ? 1u : 0u;               // which 64 columns of a key tile
        const unsigned my_epi = epi + (warp - 2u) * 8192u;
        unsigned char* my_epi_gen = smem_gen + {q_bytes + stages * per_stage}u + (warp - 2u) * 8192u;
        unsigned nstore = 0;
        unsigned u = 0, n = 0;

        auto split_q = [&](unsigned nn) {{
            const unsigned qb = nn & 1u;
            mbar_wait(qfull + qb * 8u, (nn >> 1) & 1u);
            #pragma unroll 1
            for (unsigned kk = 0; kk < {kbq}u; ++kk) {{
                const unsigned char* arow = smem_gen + kk * 16384u + (sq * 32u + lane) * 128u;
                unsigned hi[32], lo[32];
                #pragma unroll
                for (unsigned j = 0; j < 8u; ++j) {{
                    const float4 v4 = *reinterpret_cast<const float4*>(arow + ((j ^ (lane & 7u)) << 4));
                    const float xv[4] = {{v4.x, v4.y, v4.z, v4.w}};
                    #pragma unroll
                    for (unsigned e = 0; e < 4u; ++e) {{
                        hi[4 * j + e] = __float_as_uint(xv[e]) & 0xffffe000u;
                        lo[4 * j + e] = __float_as_uint(xv[e] - __uint_as_float(hi[4 * j + e]));
                    }}
                }}
                const unsigned ta = tmem_base + ((sq * 32u) << 16) + {q_tmem0}u + qb * {kbq * 64}u + kk * 64u;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {{{regs_in}}};" :: "r"(ta), {hi_ops} : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {{{regs_in}}};" :: "r"(ta + 32u), {lo_ops} : "memory");
            }}
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(qsplit + qb * 8u);
        }};

        if (half == 0u && item_at(0u) < {items}u) split_q(0u);
        for (unsigned w = item_at(n); w < {items}u; w = item_at(++n)) {{
            const unsigned z = w / {nq}u, q0 = (w % {nq}u) * 128u;
            const unsigned row = q0 + sq * 32u + lane;
            // this row's mask bits for the 64 columns of every key tile this warp owns (2 words per tile)
            unsigned mbits[{Tk // 64}];
            {{
                const uint2* br = reinterpret_cast<const uint2*>(mask + (size_t)row * {Tk // 32}u + half * 2u);
                #pragma unroll
                for (unsigned j = 0; j < {nk}u; ++j) {{
                    const uint2 w2 = __ldg(br + j * 2u);
                    mbits[2 * j] = w2.x;
                    mbits[2 * j + 1] = w2.y;
                }}
            }}
            float m = __int_as_float(0xff800000), l = 0.f, inv_l = 0.f;
            const unsigned lm = live_tiles(w);
            for (unsigned pass = 0; pass < 2u; ++pass) {{
                if (pass == 1u) {{
                    // combine the two column halves of every row: (m, l) of the partner warp through shared memory
                    // (the staging boxes double as the exchange area: our last TMA store has finished reading ours)
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                    reinterpret_cast<float2*>(my_epi_gen)[lane] = make_float2(m, l);
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    const float2 o = reinterpret_cast<const float2*>(smem_gen + {q_bytes + stages * per_stage}u + ((warp - 2u) ^ 4u) * 8192u)[lane];
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    const float mn = (o.x > m) ? o.x : m;
                    if (mn > __int_as_float(0xff800000)) {{
                        l = l * ex2(m - mn) + o.y * ex2(o.x - mn);
                        m = mn;
                    }}
                    inv_l = 1.f / l;
                    if (half == 0u && item_at(n + 1u) < {items}u) split_q(n + 1u);   // next item's Q while the MMAs run pass 1
                }}
                for (unsigned j = 0; j < {nk}u; ++j) {{
                    if (!((lm >> j) & 1u)) {{
                        // a key tile that is masked out entirely: nothing for the row sums (fill = -inf); in pass 1 its
                        // probabilities are the per-row constant 2^(-inf - m) / l = 0 (NaN for a row with no live
                        // entry at all, like the unfused kernel), stored straight from registers, eight lanes per row
                        if (pass == 1u) {{
                            const float pv = ex2({fl} - m) * inv_l;
                            float* dst = P + ((size_t)z * {Tq}u + q0 + sq * 32u + (lane >> 3)) * {Tk}u + j * 128u + half * 64u + (lane & 7u) * 4u;
                            #pragma unroll
                            for (unsigned i = 0; i < 8u; ++i) {{
                                const float pr = __shfl_sync(0xffffffffu, pv, 4u * i + (lane >> 3));
                                const float4 o4 = make_float4(pr, pr, pr, pr);
                                *reinterpret_cast<float4*>(dst + (size_t)(4u * i) * {Tk}u) = o4;
                                *reinterpret_cast<float4*>(dst + (size_t)(4u * i) * {Tk}u + 32u) = o4;
                            }}
                        }}
                        continue;
                    }}
                    const unsigned b = u & 1u;
                    mbar_wait(sfull + b * 8u, (u >> 1) & 1u);
                    {TR(2, "u", "threadIdx.x == 64u")}
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    #pragma unroll 1
                    for (unsigned cc = 0; cc < 2u; ++cc) {{
                        const unsigned c = half * 2u + cc;
                        unsigned v[32];
                        const unsigned taddr = tmem_base + ((sq * 32u) << 16) + b * {BN}u + c * 32u;
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {{{regs}}}, [%32];" : {outs} : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        // The row work is instruction-issue bound (two row warps per scheduler), so the common cases get
                        // short paths: no lane masked (scale folded into one FMA with the running maximum; scale > 0, so
                        // the maximum can be taken on the raw scores) and every lane masked (a constant).
                        float x[32];
                        const unsigned bits = mbits[j * 2u + cc];
                        if (pass == 0u) {{
                            float cm;
                            if (bits == 0u) {{
                                cm = __uint_as_float(v[0]);
                                #pragma unroll
                                for (unsigned e = 1; e < 32u; ++e) cm = fmaxf(cm, __uint_as_float(v[e]));
                                cm *= {sc};
                            }} else if (bits == 0xffffffffu) {{
                                cm = {fl};
                            }} else {{
                                cm = __int_as_float(0xff800000);
                                #pragma unroll
                                for (unsigned e = 0; e < 32u; ++e) cm = fmaxf(cm, ((bits >> e) & 1u) ? {fl} : __uint_as_float(v[e]) * {sc});
                            }}
                            const float mn = fmaxf(cm, m);
                            if (mn > __int_as_float(0xff800000)) {{      // something finite so far (else l stays 0: an all -inf row ends NaN, like the unfused kernel)
                                float add = 0.f;
                                if (bits == 0u) {{
                                    #pragma unroll
                                    for (unsigned e = 0; e < 32u; ++e) add += ex2(fmaf(__uint_as_float(v[e]), {sc}, -mn));
                                }} else if (bits == 0xffffffffu) {{
                                    add = 32.f * ex2({fl} - mn);
                                }} else {{
                                    #pragma unroll
                                    for (unsigned e = 0; e < 32u; ++e) add += ex2((((bits >> e) & 1u) ? {fl} : __uint_as_float(v[e]) * {sc}) - mn);
                                }}
                                l = l * ex2(m - mn) + add;
                                m = mn;
                            }}
                        }} else {{
                            if (bits == 0u) {{
                                #pragma unroll
                                for (unsigned e = 0; e < 32u; ++e) x[e] = ex2(fmaf(__uint_as_float(v[e]), {sc}, -m)) * inv_l;
                            }} else if (bits == 0xffffffffu) {{
                                const float pv = ex2({fl} - m) * inv_l;
                                #pragma unroll
                                for (unsigned e = 0; e < 32u; ++e) x[e] = pv;
                            }} else {{
                                #pragma unroll
                                for (unsigned e = 0; e < 32u; ++e) x[e] = ex2((((bits >> e) & 1u) ? {fl} : __uint_as_float(v[e]) * {sc}) - m) * inv_l;
                            }}
                            const unsigned sb = nstore & 1u;
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            __syncwarp();
                            unsigned char* sbuf_gen = my_epi_gen + sb * 4096u;
                            {stage_writes}
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {{
                                tma_store_3d(&tmP, my_epi + sb * 4096u, (int)(j * 128u + c * 32u), (int)(q0 + sq * 32u), (int)z);
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }}
                            ++nstore;
                        }}
                    }}
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sempty + b * 8u);
                    {TR(3, "u", "threadIdx.x == 64u")}
                    ++u;
                }}
            }}
        }}
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }} else {{
        // ===== K splitter warps: K_lo = x - trunc_tf32(x) into the twin buffer =====
        const unsigned ctid = threadIdx.x - 320u;
        unsigned it = 0;
        for (unsigned nw = 0, w = item_at(0u); w < {items}u; w = item_at(++nw)) {{
            const unsigned nlive = __popc(live_tiles(w));
            for (unsigned i2 = 0; i2 < 2u * nlive * {kbq}u; ++i2, ++it) {{
                const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                mbar_wait(kfull + s * 8u, ph);
                const float4* bsrc = reinterpret_cast<const float4*>(smem_gen + {q_bytes}u + s * {per_stage}u);
                float4* bdst = reinterpret_cast<float4*>(smem_gen + {q_bytes}u + s * {per_stage}u + {k_tile}u);
                #pragma unroll 8
                for (unsigned i = ctid; i < {k_tile // 16}u; i += 128u) {{
                    const float4 v4 = bsrc[i];
                    float4 l4;
                    l4.x = v4.x - __uint_as_float(__float_as_uint(v4.x) & 0xffffe000u);
                    l4.y = v4.y - __uint_as_float(__float_as_uint(v4.y) & 0xffffe000u);
                    l4.z = v4.z - __uint_as_float(__float_as_uint(v4.z) & 0xffffe000u);
                    l4.w = v4.w - __uint_as_float(__float_as_uint(v4.w) & 0xffffe000u);
                    bdst[i] = l4;
                }}
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(ksplit + s * 8u);
            }}
        }}
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    {"if (blockIdx.x == 0u && threadIdx.x < 160u) reinterpret_cast<long long*>(P)[threadIdx.x] = (&trace[0][0])[threadIdx.x];" if trace else ""}
    if (warp == 1) {{
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"({tmem_cols}));
    }}
}}
"""
    return TcKernel("attn_scores", src, (grid, 1, 1), nthreads, smem, batch * Tq * Tk, BN, stages)


@cache
def attn_dscores_tc(Tq: int, Tk: int, D: int, batch: int, scale: float) -> TcKernel:
    """Backward twin of attn_scores_tc:  dS = masked_fill(P * (dP - rowsum(dP * P)), mask, 0) * scale  with
    dP = dO @ V^T computed tile by tile in TMEM and never written (the unfused chain writes dP, 403 MB per layer at
    C5, and reads it back in the softmax backward).  rowsum(dP * P) = rowsum(dO * O) — the flash-attention identity —
    arrives precomputed, so ONE pass over the key tiles suffices: read P once, write dS once.
    dO (batch, Tq, D) and V (batch, Tk, D) K-major.  Kernel parameters: (tmDO, tmV, tmDS, dS, P, aux) with
    aux = [rowsum(dO * O): batch * Tq floats | mask bits: Tq * Tk / 32 words]."""
    fill = 0.0
    import os
    trace = os.environ.get("DLGRAD_TC_TRACE", "0") == "1"      # diagnostic: CTA 0 timestamps its first 32 key tiles
    skip = Tk // 128 <= 16 and os.environ.get("DLGRAD_ATTN_SKIP", "1") != "0"

    def TR(kind, idx, cond="lane == 0u"):
        return f"if (blockIdx.x == 0u && {idx} < 32u && {cond}) trace[{kind}][{idx}] = clock64();" if trace else ""
    assert Tq % 128 == 0 and Tk % 128 == 0 and Tk <= 2048 and D in (32, 64)      # D <= 64: 2 Q buffers + 2 S buffers fill TMEM
    kbq = D // 32                                   # k-blocks per tile
    BN = 128
    q_bytes = kbq * 16384                           # one Q tile (128 x D fp32), K-major, 16 KiB per k-block
    k_tile = BN * BK * 4                            # 16 KiB: one k-block of a key tile
    per_stage = 2 * k_tile                          # raw (= hi) | lo
    NB, AHEAD = 4, 2                     # 32x32 boxes per row warp: P arrives in them by cp.async AHEAD chunks early, dS leaves from them
    epi_bytes = 8 * NB * 4096
    budget = 227 * 1024 - 1024 - 512 - epi_bytes - q_bytes       # Q: ONE shared-memory tile (it only lives until it is split into TMEM)
    stages = max(2, min(6, budget // per_stage))
    smem = q_bytes + stages * per_stage + epi_bytes + 1024 + 512
    q_tmem0 = 2 * BN                                # S accumulators first (double buffered), then Q hi/lo x 2 buffers
    tmem_need = q_tmem0 + 2 * kbq * 64
    assert tmem_need <= 512, tmem_need
    tmem_cols = 512
    idesc = _idesc(BM, BN, False, False)
    nq, nk = Tq // 128, Tk // 128
    items = batch * nq
    grid = min(items, NUM_SMS)
    nthreads = 64 + 256 + 128           # producer, MMA | 8 row warps (two per TMEM lane quarter, 64 columns each) | 4 K-splitter warps
    # the row warps work in the base-2 domain: softmax(x) = 2^((x - m) log2 e) / sum, one MUFU.EX2 per element
    # instead of expf's ~10 instructions (they, not the tensor pipe, were the bottleneck: 4 us per tile)
    import math
    import numpy as _np
    log2e = float(_np.float32(math.log2(math.e)))
    sc = _f32lit(float(_np.float32(scale)))
    fl = _f32lit(fill)
    regs = ", ".join(f"%{i}" for i in range(32))
    outs = ", ".join(f'"=r"(v[{i}])' for i in range(32))
    regs_in = ", ".join(f"%{i + 1}" for i in range(32))
    hi_ops = ", ".join(f'"r"(hi[{i}])' for i in range(32))
    lo_ops = ", ".join(f'"r"(lo[{i}])' for i in range(32))
    stage_writes = "\n                        ".join(
        f"*reinterpret_cast<float4*>(sbuf_gen + lane * 128u + (({j}u ^ (lane & 7u)) << 4)) = "
        f"make_float4(x[{4 * j}], x[{4 * j + 1}], x[{4 * j + 2}], x[{4 * j + 3}]);"
        for j in range(8))
    src = _HELPERS + _HELPERS_TS + """
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
""" + f"""
extern "C" __global__ void __launch_bounds__({nthreads}, 1)
{KNAME}(const __grid_constant__ TMap tmQ, const __grid_constant__ TMap tmK, const __grid_constant__ TMap tmP, float* __restrict__ P, const float* __restrict__ Pin, const float* __restrict__ aux)
{{
    extern __shared__ unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    const unsigned qs = smem;                                   // the Q tile being split
    const unsigned ring = smem + {q_bytes}u;                // K stages: raw | lo
    const unsigned epi = ring + {stages * per_stage}u;
    const unsigned bars = epi + {epi_bytes}u;
    const unsigned kfull = bars, kempty = bars + {stages * 8}u, ksplit = bars + {stages * 16}u;
    const unsigned sfull = bars + {stages * 24}u, sempty = sfull + 16u, qfull = sempty + 16u, qsplit = qfull + 16u, qfree = qsplit + 16u;
    __shared__ unsigned tmem_slot;
    {"__shared__ long long trace[5][32];" if trace else ""}
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    // key tiles with at least one unmasked entry, one bit per tile (cuda_kernel.mask_bits): dS of a tile that is masked out
    // entirely is zero whatever P and dP hold, so such a tile is neither loaded nor multiplied nor read — just zeroed
    auto live_tiles = [&](unsigned ww) -> unsigned {{
        {"const unsigned* tf = reinterpret_cast<const unsigned*>(aux + " + str(batch * Tq) + "u) + " + str(Tq * Tk // 32) + "u + (ww % " + str(nq) + "u) * " + str(nk) + "u; unsigned lm = 0u; for (unsigned j = 0; j < " + str(nk) + "u; ++j) lm |= (__ldg(tf + j) != 0u ? 1u : 0u) << j; return lm;" if skip else "return " + str((1 << nk) - 1) + "u;"}
    }};
    // work items in the order of CTA `blockIdx.x`: heavy query tiles first (under a causal mask tile q has q + 1 live key
    // tiles), CTAs taken in alternating direction from round to round.  Plain round-robin gives a CTA the same two
    // query tiles over and over ({grid} mod {nq} = {grid % nq}): 36 tile-units on the worst CTA against 23.4 on average.
    auto item_at = [&](unsigned nn) -> unsigned {{
        const unsigned p = nn * {grid}u + ((nn & 1u) ? ({grid - 1}u - blockIdx.x) : blockIdx.x);
        if (p >= {items}u) return {items}u;
        return (p % {batch}u) * {nq}u + ({nq - 1}u - p / {batch}u);
    }};

    if (threadIdx.x == 0) {{
        for (unsigned s = 0; s < {stages}u; ++s) {{
            mbar_init(kfull + s * 8u, 1u);
            mbar_init(kempty + s * 8u, 1u);
            mbar_init(ksplit + s * 8u, 4u);
        }}
        for (unsigned b = 0; b < 2u; ++b) {{
            mbar_init(sfull + b * 8u, 1u);
            mbar_init(sempty + b * 8u, 8u);
            mbar_init(qfull + b * 8u, 1u);
            mbar_init(qsplit + b * 8u, 4u);
            mbar_init(qfree + b * 8u, 1u);
        }}
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    if (warp == 1) {{
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"({tmem_cols}));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_slot;

    if (warp == 0) {{
        // ===== TMA producer: K tiles twice per item; the NEXT item's Q tile is requested between the two passes so
        //       that the row warps can split it while the MMAs run pass 1 =====
        unsigned it = 0, n = 0;
        auto load_q = [&](unsigned nn, unsigned ww) {{
            const unsigned zz = ww / {nq}u, qq0 = (ww % {nq}u) * 128u, qb = nn & 1u;
            mbar_wait(qfree + qb * 8u, ((nn >> 1) & 1u) ^ 1u);          // the MMAs of the item that used this TMEM Q buffer are done
            if (nn != 0u) mbar_wait(qsplit + ((nn - 1u) & 1u) * 8u, ((nn - 1u) >> 1) & 1u);   // the previous Q tile has left shared memory
            if (elect_one()) {{
                mbar_expect_tx(qfull + qb * 8u, {q_bytes}u);
                #pragma unroll
                for (unsigned kk = 0; kk < {kbq}u; ++kk)
                    tma_load_3d(qs + kk * 16384u, &tmQ, qfull + qb * 8u, (int)(kk * 32u), (int)qq0, (int)zz);
            }}
            __syncwarp();
        }};
        if (item_at(0u) < {items}u) load_q(0u, item_at(0u));
        for (unsigned w = item_at(n); w < {items}u; w = item_at(++n)) {{
            const unsigned z = w / {nq}u;
            {{
                const unsigned lm = live_tiles(w);
                if (item_at(n + 1u) < {items}u) load_q(n + 1u, item_at(n + 1u));       // (waits for the MMAs of item n - 1 only)
                for (unsigned j = 0; j < {nk}u; ++j) {{
                    if (!((lm >> j) & 1u)) continue;
                    for (unsigned kk = 0; kk < {kbq}u; ++kk, ++it) {{
                        const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                        mbar_wait(kempty + s * 8u, ph ^ 1u);
                        {TR(4, f"(it / {kbq}u)", "lane == 0u && kk == 0u")}
                        if (elect_one()) {{
                            mbar_expect_tx(kfull + s * 8u, {k_tile}u);
                            tma_load_3d(ring + s * {per_stage}u, &tmK, kfull + s * 8u, (int)(kk * 32u), (int)(j * 128u), (int)z);
                        }}
                        __syncwarp();
                    }}
                }}
            }}
        }}
    }} else if (warp == 1) {{
        // ===== MMA issuer: S[b] = Q_hi/lo (TMEM) x K_hi/lo (smem), 3 MMAs per 8-wide k step =====
        unsigned it = 0, u = 0, n = 0;
        const unsigned long long bdesc0 = make_desc(ring, 1u, 64u, 2u);
        for (unsigned w = item_at(n); w < {items}u; w = item_at(++n)) {{
            const unsigned qb = n & 1u;
            mbar_wait(qsplit + qb * 8u, (n >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned q_t = tmem_base + {q_tmem0}u + qb * {kbq * 64}u;
            const unsigned lm = live_tiles(w);
            for (unsigned pj = 0; pj < {nk}u; ++pj) {{
                if (!((lm >> pj) & 1u)) continue;
                const unsigned b = u & 1u;
                mbar_wait(sempty + b * 8u, ((u >> 1) & 1u) ^ 1u);
                {TR(0, "u")}
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tacc = tmem_base + b * {BN}u;
                for (unsigned kk = 0; kk < {kbq}u; ++kk, ++it) {{
                    const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                    mbar_wait(ksplit + s * 8u, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned long long bdesc = bdesc0 + (unsigned long long)(s * {per_stage // 16}u);
                    const unsigned a_hi = q_t + kk * 64u, a_lo = a_hi + 32u;
                    if (elect_one()) {{
                        #pragma unroll
                        for (unsigned k = 0; k < 4u; ++k) {{
                            const unsigned long long b_hi = bdesc + (unsigned long long)(k * 2u);
                            umma_tf32_ts(tacc, a_lo + k * 8u, b_hi, {idesc}u, (kk | k) != 0u);
                            umma_tf32_ts(tacc, a_hi + k * 8u, b_hi + {k_tile // 16}ull, {idesc}u, 1u);
                            umma_tf32_ts(tacc, a_hi + k * 8u, b_hi, {idesc}u, 1u);
                        }}
                        umma_commit(kempty + s * 8u);
                        if (kk + 1u == {kbq}u) umma_commit(sfull + b * 8u);
                    }}
                    __syncwarp();
                }}
                {TR(1, "u")}
                ++u;
            }}
            if (elect_one()) umma_commit(qfree + qb * 8u);
            __syncwarp();
        }}
    }} else if (warp < 10u) {{
        // ===== row warps: thread = query row = TMEM lane; warps 2-5 take columns 0-63 of every key tile, warps 6-9
        //       columns 64-127 (the row work — TMEM reads, exponentials, stores — is what paces this kernel).  Split Q into TMEM; pass 0: running max / sum over the
        //       key tiles; pass 1: normalise and store P =====
        const unsigned sq = warp & 3u;
        const unsigned half = (warp >= 6u) ? 1u : 0u;               // which 64 columns of a key tile
        const unsigned my_epi = epi + (warp - 2u) * {NB * 4096}u;
        unsigned char* my_epi_gen = smem_gen + {q_bytes + stages * per_stage}u + (warp - 2u) * {NB * 4096}u;
        unsigned nstore = 0, nfetch = 0, fw = 0, fn = 0, fg = 0, flm = 0;      // fetch cursor: item, live-chunk index, its live mask
        unsigned u = 0, n = 0;

        auto split_q = [&](unsigned nn) {{
            const unsigned qb = nn & 1u;
            mbar_wait(qfull + qb * 8u, (nn >> 1) & 1u);
            #pragma unroll 1
            for (unsigned kk = 0; kk < {kbq}u; ++kk) {{
                const unsigned char* arow = smem_gen + kk * 16384u + (sq * 32u + lane) * 128u;
                unsigned hi[32], lo[32];
                #pragma unroll
                for (unsigned j = 0; j < 8u; ++j) {{
                    const float4 v4 = *reinterpret_cast<const float4*>(arow + ((j ^ (lane & 7u)) << 4));
                    const float xv[4] = {{v4.x, v4.y, v4.z, v4.w}};
                    #pragma unroll
                    for (unsigned e = 0; e < 4u; ++e) {{
                        hi[4 * j + e] = __float_as_uint(xv[e]) & 0xffffe000u;
                        lo[4 * j + e] = __float_as_uint(xv[e] - __uint_as_float(hi[4 * j + e]));
                    }}
                }}
                const unsigned ta = tmem_base + ((sq * 32u) << 16) + {q_tmem0}u + qb * {kbq * 64}u + kk * 64u;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {{{regs_in}}};" :: "r"(ta), {hi_ops} : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {{{regs_in}}};" :: "r"(ta + 32u), {lo_ops} : "memory");
            }}
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(qsplit + qb * 8u);
        }};

        if (half == 0u && item_at(0u) < {items}u) split_q(0u);
        for (unsigned w = item_at(n); w < {items}u; w = item_at(++n)) {{
            const unsigned z = w / {nq}u, q0 = (w % {nq}u) * 128u;
            const unsigned row = q0 + sq * 32u + lane;
            // this row's mask bits for the 64 columns of every key tile this warp owns (2 words per tile)
            unsigned mbits[{Tk // 64}];
            {{
                const uint2* br = reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned*>(aux + {batch * Tq}u) + (size_t)row * {Tk // 32}u + half * 2u);
                #pragma unroll
                for (unsigned j = 0; j < {nk}u; ++j) {{
                    const uint2 w2 = __ldg(br + j * 2u);
                    mbits[2 * j] = w2.x;
                    mbits[2 * j + 1] = w2.y;
                }}
            }}
            const float drow = __ldg(aux + (size_t)z * {Tq}u + row);            // rowsum(dO * O) of this row
            // P travels global -> shared by cp.async (16 B per lane, eight lanes per 128-byte row, landing directly in the
            // swizzled position of a staging box), TWO 32-column chunks ahead (a box is refilled once the store before
            // last has left it): 8 KiB per warp, 64 KiB per SM in
            // flight.  One chunk ahead in registers left the loads latency-bound (6000-9000 clk per key tile), and a
            // thread reading its own row directly touches 32 cache lines per load.
            auto fetch_p = [&](unsigned ww, unsigned jt, unsigned cq) {{      // 32 columns cq (0/1) of this warp's half of key tile jt, item ww
                const unsigned zz = ww / {nq}u, qq0 = (ww % {nq}u) * 128u;
                const float* src = Pin + ((size_t)zz * {Tq}u + qq0 + sq * 32u + (lane >> 3)) * {Tk}u + jt * 128u + half * 64u + cq * 32u + (lane & 7u) * 4u;
                const unsigned box = my_epi + (nfetch % {NB}u) * 4096u;
                #pragma unroll
                for (unsigned i = 0; i < 8u; ++i) {{
                    const unsigned r = 4u * i + (lane >> 3);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(box + r * 128u + (((lane & 7u) ^ (r & 7u)) << 4)), "l"(src + (size_t)(4u * i) * {Tk}u) : "memory");
                }}
            }};
            // the fetch cursor runs {AHEAD} live chunks ahead of the chunk being processed, across work items
            auto fetch_next = [&]() {{
                while (fw < {items}u && fg >= 2u * __popc(flm)) {{ fw = item_at(++fn); fg = 0u; flm = (fw < {items}u) ? live_tiles(fw) : 0u; }}
                if (fw < {items}u) {{
                    unsigned t = flm, k = fg >> 1;
                    while (k--) t &= t - 1u;                         // drop the k lowest live tiles
                    fetch_p(fw, __ffs(t) - 1u, fg & 1u);
                    ++fg;
                }}
                asm volatile("cp.async.commit_group;" ::: "memory");   // (an empty group keeps the count in step)
                ++nfetch;
            }};
            const unsigned lm = live_tiles(w);
            if (n == 0u) {{
                fw = w; fn = 0u; fg = 0u; flm = lm;
                #pragma unroll 1
                for (unsigned g = 0; g < {AHEAD}u; ++g) fetch_next();
            }}
            // dead key tiles: zeros straight from registers, eight lanes per 128-byte row
            for (unsigned j = 0; j < {nk}u; ++j) {{
                if ((lm >> j) & 1u) continue;
                float* dst = P + ((size_t)z * {Tq}u + q0 + sq * 32u + (lane >> 3)) * {Tk}u + j * 128u + half * 64u + (lane & 7u) * 4u;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                #pragma unroll
                for (unsigned i = 0; i < 8u; ++i) {{
                    *reinterpret_cast<float4*>(dst + (size_t)(4u * i) * {Tk}u) = z4;
                    *reinterpret_cast<float4*>(dst + (size_t)(4u * i) * {Tk}u + 32u) = z4;
                }}
            }}
            const unsigned nlive = __popc(lm);
            unsigned kdone = 0;
            if (nlive == 0u && half == 0u && item_at(n + 1u) < {items}u) split_q(n + 1u);
            for (unsigned j = 0; j < {nk}u; ++j) {{
                if (!((lm >> j) & 1u)) continue;
                if (kdone++ == nlive / 2u && half == 0u && item_at(n + 1u) < {items}u) split_q(n + 1u);   // next item's dO, mid-item
                const unsigned b = u & 1u;
                mbar_wait(sfull + b * 8u, (u >> 1) & 1u);
                {TR(2, "u", "threadIdx.x == 64u")}
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                #pragma unroll 1
                for (unsigned cc = 0; cc < 2u; ++cc) {{
                    const unsigned c = half * 2u + cc;
                    unsigned v[32];
                    const unsigned taddr = tmem_base + ((sq * 32u) << 16) + b * {BN}u + c * 32u;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {{{regs}}}, [%32];" : {outs} : "r"(taddr));
                    {{   // fetch the live chunk {AHEAD} ahead (possibly the next item's) into the box the store before last has left
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read {NB - AHEAD - 1};" ::: "memory");
                        __syncwarp();
                        fetch_next();
                    }}
                    asm volatile("cp.async.wait_group {AHEAD};" ::: "memory");          // this chunk's P has landed (my copies)
                    __syncwarp();                                                          // ... and everybody else's
                    const unsigned sb = nstore % {NB}u;
                    unsigned char* sbuf_gen = my_epi_gen + sb * 4096u;
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float x[32];
                    const unsigned bits = mbits[j * 2u + cc];
                    #pragma unroll
                    for (unsigned jj = 0; jj < 8u; ++jj) {{
                        const float4 p4 = *reinterpret_cast<const float4*>(sbuf_gen + lane * 128u + ((jj ^ (lane & 7u)) << 4));
                        const float pe[4] = {{p4.x, p4.y, p4.z, p4.w}};
                        #pragma unroll
                        for (unsigned e = 0; e < 4u; ++e) {{
                            const float t = pe[e] * (__uint_as_float(v[4 * jj + e]) - drow);
                            x[4 * jj + e] = ((bits >> (4u * jj + e)) & 1u) ? 0.f : t * {sc};
                        }}
                    }}
                    {stage_writes}
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {{
                        tma_store_3d(&tmP, my_epi + sb * 4096u, (int)(j * 128u + c * 32u), (int)(q0 + sq * 32u), (int)z);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }}
                    ++nstore;
                }}
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(sempty + b * 8u);
                {TR(3, "u", "threadIdx.x == 64u")}
                ++u;
            }}
        }}
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }} else {{
        // ===== K splitter warps: K_lo = x - trunc_tf32(x) into the twin buffer =====
        const unsigned ctid = threadIdx.x - 320u;
        unsigned it = 0;
        for (unsigned nw = 0, w = item_at(0u); w < {items}u; w = item_at(++nw)) {{
            const unsigned nlive = __popc(live_tiles(w));
            for (unsigned i2 = 0; i2 < nlive * {kbq}u; ++i2, ++it) {{
                const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                mbar_wait(kfull + s * 8u, ph);
                const float4* bsrc = reinterpret_cast<const float4*>(smem_gen + {q_bytes}u + s * {per_stage}u);
                float4* bdst = reinterpret_cast<float4*>(smem_gen + {q_bytes}u + s * {per_stage}u + {k_tile}u);
                #pragma unroll 8
                for (unsigned i = ctid; i < {k_tile // 16}u; i += 128u) {{
                    const float4 v4 = bsrc[i];
                    float4 l4;
                    l4.x = v4.x - __uint_as_float(__float_as_uint(v4.x) & 0xffffe000u);
                    l4.y = v4.y - __uint_as_float(__float_as_uint(v4.y) & 0xffffe000u);
                    l4.z = v4.z - __uint_as_float(__float_as_uint(v4.z) & 0xffffe000u);
                    l4.w = v4.w - __uint_as_float(__float_as_uint(v4.w) & 0xffffe000u);
                    bdst[i] = l4;
                }}
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(ksplit + s * 8u);
            }}
        }}
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    {"if (blockIdx.x == 0u && threadIdx.x < 160u) reinterpret_cast<long long*>(P)[threadIdx.x] = (&trace[0][0])[threadIdx.x];" if trace else ""}
    if (warp == 1) {{
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"({tmem_cols}));
    }}
}}
"""
    return TcKernel("attn_dscores", src, (grid, 1, 1), nthreads, smem, batch * Tq * Tk, BN, stages)


# =====================================================================================================
# v3: CTA-pair kernel (cta_group::2).  Two CTAs of a cluster (one TPC) compute a 256 x BN tile: each
# stages ITS 128 rows of A and ITS half of the B columns, the leader issues M=256 MMAs that read both
# shared memories, so every SM moves, splits and feeds only half of B — the shared-memory traffic that
# bounds the 1-CTA kernel (TMA writes + hi/lo split + three operand reads per k-step) drops by a third.
# =====================================================================================================
_HELPERS_2CTA = r"""
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned map_to_cta(unsigned local_addr, unsigned rank) {
    unsigned ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank)); return ra;
}
__device__ __forceinline__ void mbar_arrive_cluster(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned bar, unsigned parity) {
    // acquire at CLUSTER scope: the arrivals come from the partner CTA's threads
    const long long t0 = clock64();
    for (;;) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void umma_tf32_2cta(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                               unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((unsigned short)3) : "memory");
}
"""


@cache
def matmul_tc_pair(M: int, N: int, K: int, batch: int, a_mn: bool, b_mn: bool, a_bcast: bool, b_bcast: bool,
                   bn: int | None = None, stages: int | None = None, kc_blocks: int | None = None) -> TcKernel:
    import os
    trace = os.environ.get("DLGRAD_TC_TRACE", "0") == "1"     # diagnostic: pair 0's leader timestamps its first 32 k-blocks
    TR = (lambda kind, cond="true": f"if (blockIdx.x == 0u && it < 32u && {cond}) trace[{kind}][it] = clock64();") if trace else (lambda *a: "")
    kblocks = _cdiv(K, BK)
    kc = kc_blocks or (32 if kblocks > 64 else kblocks)
    nchunks = _cdiv(kblocks, kc)
    BN = bn or 256
    BNH = BN // 2                                       # B columns staged by each CTA of the pair
    a_tile, b_tile = BM * BK * 4, BNH * BK * 4
    ab = a_tile + b_tile
    per_stage = 2 * ab
    EPI_BUFS = 2
    epi_bytes = 4 * EPI_BUFS * 4096
    budget = 227 * 1024 - 1024 - 512 - epi_bytes
    if stages is None:
        stages = max(2, min(6, budget // per_stage))
    ntm, ntn = _cdiv(M, 2 * BM), _cdiv(N, BN)
    ntiles = ntm * ntn * batch
    npairs = min(ntiles, NUM_SMS // 2)
    tmem_cols = 32
    while tmem_cols < 2 * BN:
        tmem_cols *= 2
    smem = stages * per_stage + epi_bytes + 1024 + 512
    idesc = _idesc(2 * BM, BN, a_mn, b_mn)
    a_lbo, a_sbo, a_kstep, a_lt = (256, 32, 64, 1) if a_mn else (1, 64, 2, 2)
    b_lbo, b_sbo, b_kstep, b_lt = (256, 32, 64, 1) if b_mn else (1, 64, 2, 2)
    a_boxes = BM // 32 if a_mn else 1
    b_boxes = BNH // 32 if b_mn else 1
    nsw = 4
    nthreads = 64 + 32 * nsw + 128
    epi_w0 = 2 + nsw

    def tma_issue(which, boxes, mn, tile_var, row0):
        tm = f"tm{which}"
        z = "0" if (a_bcast if which == "A" else b_bcast) else "z"
        if mn:
            return "\n                    ".join(
                f"tma_load_3d({tile_var} + {j * 4096}u, &{tm}, full_bar, (int)({row0} + {j * 32}u), (int)k0, (int){z});"
                for j in range(boxes))
        return f"tma_load_3d({tile_var}, &{tm}, full_bar, (int)k0, (int){row0}, (int){z});"

    decode = f"""const unsigned z = t / {ntm * ntn}u, rem = t % {ntm * ntn}u;
            const unsigned m0 = (rem / {ntn}u) * {2 * BM}u + crank * {BM}u, n0 = (rem % {ntn}u) * {BN}u;"""
    regs = ", ".join(f"%{i}" for i in range(32))
    outs = ", ".join(f'"=r"(v[{i}])' for i in range(32))
    stage_writes = "\n                        ".join(
        f"*reinterpret_cast<float4*>(sbuf_gen + lane * 128u + (({j}u ^ (lane & 7u)) << 4)) = "
        f"make_float4(__uint_as_float(v[{4 * j}]), __uint_as_float(v[{4 * j + 1}]), __uint_as_float(v[{4 * j + 2}]), __uint_as_float(v[{4 * j + 3}]));"
        for j in range(8))

    src = _HELPERS + _HELPERS_2CTA + f"""
extern "C" __global__ void __launch_bounds__({nthreads}, 1)   // launched with cluster dimension 2 (shim: cluster_x)
{KNAME}(const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmB, const __grid_constant__ TMap tmC, float* __restrict__ C)
{{
    extern __shared__ unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    const unsigned epi = smem + {stages * per_stage}u;
    const unsigned bars = epi + {epi_bytes}u;
    const unsigned full = bars, empty = bars + {stages * 8}u, split = bars + {stages * 16}u;
    const unsigned tfull = bars + {stages * 24}u, tempty = tfull + 16u;
    __shared__ unsigned tmem_slot;
    {"__shared__ long long trace[5][32];" if trace else ""}
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned crank = cluster_rank();               // 0 = leader (issues the MMAs)
    const unsigned pair = blockIdx.x >> 1;

    if (threadIdx.x == 0) {{
        for (unsigned s = 0; s < {stages}u; ++s) {{
            mbar_init(full + s * 8u, 1u);                // own TMA loads
            mbar_init(empty + s * 8u, 1u);               // multicast commit from the leader
            mbar_init(split + s * 8u, {2 * nsw}u);       // (leader's copy is used) one arrive per splitter warp of both CTAs
        }}
        for (unsigned b = 0; b < 2u; ++b) {{
            mbar_init(tfull + b * 8u, 1u);               // multicast commit from the leader
            mbar_init(tempty + b * 8u, 8u);              // (leader's copy) one arrive per epilogue warp of both CTAs
        }}
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    if (warp == 1) {{
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"({tmem_cols}));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' barriers are initialised before any remote arrive
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_slot;

    if (warp == 0) {{
        // ===== TMA producer (each CTA stages its own A rows and its own half of the B columns) =====
        if (lane == 0) {{
            unsigned it = 0;
            for (unsigned t = pair; t < {ntiles}u; t += {npairs}u) {{
                {decode}
                const unsigned nb0 = n0 + crank * {BNH}u;
                for (unsigned kb = 0; kb < {kblocks}u; ++kb, ++it) {{
                    const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                    mbar_wait(empty + s * 8u, ph ^ 1u);
                    {TR(0)}
                    const unsigned full_bar = full + s * 8u;
                    const unsigned a_s = smem + s * {per_stage}u, b_s = a_s + {a_tile}u;
                    const unsigned k0 = kb * {BK}u;
                    mbar_expect_tx(full_bar, {ab}u);
                    {tma_issue("A", a_boxes, a_mn, "a_s", "m0")}
                    {tma_issue("B", b_boxes, b_mn, "b_s", "nb0")}
                }}
            }}
        }}
    }} else if (warp == 1) {{
        // ===== MMA issuer: leader CTA only, one lane; M = 256 across the pair =====
        if (crank == 0 && lane == 0) {{
            unsigned it = 0, u = 0;
            for (unsigned t = pair; t < {ntiles}u; t += {npairs}u) {{
                for (unsigned ch = 0; ch < {nchunks}u; ++ch, ++u) {{
                    const unsigned b = u & 1u;
                    mbar_wait_cluster(tempty + b * 8u, ((u >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned tacc = tmem_base + b * {BN}u;
                    const unsigned kend = min((ch + 1u) * {kc}u, {kblocks}u);
                    for (unsigned kb = ch * {kc}u, kbl = 0; kb < kend; ++kb, ++kbl, ++it) {{
                        const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                        mbar_wait_cluster(split + s * 8u, ph);
                        {TR(3)}
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned a_s = smem + s * {per_stage}u, b_s = a_s + {a_tile}u;
                        const unsigned long long adesc = make_desc(a_s, {a_lbo}u, {a_sbo}u, {a_lt}u);
                        const unsigned long long bdesc = make_desc(b_s, {b_lbo}u, {b_sbo}u, {b_lt}u);
                        #pragma unroll
                        for (unsigned k = 0; k < {BK // 8}u; ++k) {{
                            const unsigned long long a_hi = adesc + (unsigned long long)(k * {a_kstep}u);
                            const unsigned long long b_hi = bdesc + (unsigned long long)(k * {b_kstep}u);
                            umma_tf32_2cta(tacc, a_hi + {ab // 16}ull, b_hi, {idesc}u, (kbl | k) != 0u);
                            umma_tf32_2cta(tacc, a_hi, b_hi + {ab // 16}ull, {idesc}u, 1u);
                            umma_tf32_2cta(tacc, a_hi, b_hi, {idesc}u, 1u);
                        }}
                        umma_commit_2cta(empty + s * 8u);          // frees stage s in BOTH CTAs
                        {TR(4)}
                    }}
                    umma_commit_2cta(tfull + b * 8u);               // accumulator b is complete in BOTH CTAs
                }}
            }}
        }}
    }} else if (warp < {epi_w0}u) {{
        // ===== splitter warps (both CTAs): lo = x - trunc_tf32(x); report to the LEADER's split barrier =====
        const unsigned ctid = threadIdx.x - 64u;
        unsigned it = 0;
        for (unsigned t = pair; t < {ntiles}u; t += {npairs}u) {{
            for (unsigned kb = 0; kb < {kblocks}u; ++kb, ++it) {{
                const unsigned s = it % {stages}u, ph = (it / {stages}u) & 1u;
                mbar_wait(full + s * 8u, ph);
                {TR(1, "ctid == 0u")}
                const float4* src = reinterpret_cast<const float4*>(smem_gen + s * {per_stage}u);
                float4* dst = reinterpret_cast<float4*>(smem_gen + s * {per_stage}u + {ab}u);
                #pragma unroll 4
                for (unsigned i = ctid; i < {ab // 16}u; i += {32 * nsw}u) {{
                    const float4 x = src[i];
                    float4 lo;
                    lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                    lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                    lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                    lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                    dst[i] = lo;
                }}
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(split + s * 8u, 0u));
                {TR(2, "ctid == 0u")}
            }}
        }}
    }} else {{
        // ===== epilogue warps (both CTAs drain their own 128 accumulator rows) =====
        const unsigned q = warp & 3u;
        const unsigned my_epi = epi + (warp - {epi_w0}u) * {EPI_BUFS * 4096}u;
        unsigned char* my_epi_gen = smem_gen + {stages * per_stage}u + (warp - {epi_w0}u) * {EPI_BUFS * 4096}u;
        unsigned u = 0, nstore = 0;
        for (unsigned t = pair; t < {ntiles}u; t += {npairs}u) {{
            {decode}
            for (unsigned ch = 0; ch < {nchunks}u; ++ch, ++u) {{
                const unsigned b = u & 1u;
                mbar_wait(tfull + b * 8u, (u >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (ch != 0u && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                #pragma unroll 1
                for (unsigned c = 0; c < {BN // 32}u; ++c) {{
                    unsigned v[32];
                    const unsigned taddr = tmem_base + ((q * 32u) << 16) + b * {BN}u + c * 32u;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {{{regs}}}, [%32];" : {outs} : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (n0 + c * 32u < {N}u && m0 + q * 32u < {M}u) {{
                        const unsigned sb = nstore & {EPI_BUFS - 1}u;
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read {EPI_BUFS - 1};" ::: "memory");
                        __syncwarp();
                        unsigned char* sbuf_gen = my_epi_gen + sb * 4096u;
                        {stage_writes}
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {{
                            if (ch == 0u) tma_store_3d(&tmC, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), (int)z);
                            else tma_reduce_add_3d(&tmC, my_epi + sb * 4096u, (int)(n0 + c * 32u), (int)(m0 + q * 32u), (int)z);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }}
                        ++nstore;
                    }}
                }}
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty + b * 8u, 0u));
            }}
        }}
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                  // no CTA leaves while its partner may still touch it
    {"if (blockIdx.x == 0u && threadIdx.x < 160u) reinterpret_cast<long long*>(C)[threadIdx.x] = (&trace[0][0])[threadIdx.x];" if trace else ""}
    if (warp == 1) {{
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"({tmem_cols}));
    }}
}}
"""
    return TcKernel("mm_tc2cta", src, (2 * npairs, 1, 1), nthreads, smem, batch * M * N, BN, stages)
