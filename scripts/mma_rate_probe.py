"""Tensor-pipe rate of kind::tf32 tcgen05.mma with the A operand in TMEM, with and without competing shared-memory
traffic, for one CTA (M=128) and for a CTA pair (cta_group::2, M=256).  Operands are whatever shared / tensor memory
holds; only the timing matters.  Answers: does a pair halve the B traffic each SM has to serve?

    python scripts/mma_rate_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dlgrad_b200.codegen import cuda_tcgen05 as tc  # noqa: E402
from dlgrad_b200.codegen.cuda_kernel import Kernel, SIG  # noqa: E402
from dlgrad_b200.runtime.cuda import ops as cops, shim, transfer  # noqa: E402
from dlgrad_b200.runtime.cuda.compiler import KNAME  # noqa: E402

ITERS = 200          # k-blocks; each = 4 k-steps x 3 MMAs


def kernel(pair: bool, N: int, traffic_warps: int, bytes_per_kblock: int) -> Kernel:
    cg = 2 if pair else 1
    M = 256 if pair else 128
    nrows = N // 2 if pair else N                 # B rows (columns of the product) staged by each CTA
    idesc = tc._idesc(M, N, False, False)        # A from TMEM, B K-major
    b_tile = nrows * 128                          # one k-block of B: nrows x 32 fp32
    mma = ("tcgen05.mma.cta_group::%d.kind::tf32 [%%0], [%%1], %%2, %%3, p;" % cg)
    commit = ("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
              if pair else "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];")
    commit_args = ':: "r"(bar), "h"((unsigned short)3) : "memory"' if pair else ':: "r"(bar) : "memory"'
    nthreads = 64 + 32 * traffic_warps
    src = tc._HELPERS + tc._HELPERS_2CTA + tc._HELPERS_TS + f"""
extern "C" __global__ void __launch_bounds__({nthreads}, 1) {KNAME}{SIG}
{{
    extern __shared__ unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    __shared__ unsigned tmem_slot;
    __shared__ unsigned long long barmem;
    __shared__ volatile unsigned stop;
    const unsigned bar = smem_u32(&barmem);
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned crank = {"cluster_rank()" if pair else "0u"};
    if (threadIdx.x == 0) {{ mbar_init(bar, 1u); stop = 0u; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }}
    if (warp == 0) {{
        asm volatile("tcgen05.alloc.cta_group::{cg}.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::{cg}.sync.aligned;");
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    {"cluster_sync_all();" if pair else ""}
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_slot;
    if (warp == 0) {{
        if (crank == 0u) {{
            const unsigned long long bdesc0 = make_desc(smem, 1u, 64u, 2u);
            const long long t0 = clock64();
            for (unsigned it = 0; it < {ITERS}u; ++it) {{
                const unsigned long long bdesc = bdesc0 + (unsigned long long)((it & 1u) * {2 * b_tile // 16}u);
                if (elect_one()) {{
                    #pragma unroll
                    for (unsigned k = 0; k < 4u; ++k) {{
                        const unsigned long long b_hi = bdesc + (unsigned long long)(k * 2u);
                        const unsigned a = tmem_base + {N}u + (it & 3u) * 64u + k * 8u;
                        asm volatile("{{\\n\\t.reg .pred p;\\n\\tsetp.ne.b32 p, %4, 0;\\n\\t{mma}\\n\\t}}" :: "r"(tmem_base), "r"(a + 32u), "l"(b_hi), "r"({idesc}u), "r"(1u) : "memory");
                        asm volatile("{{\\n\\t.reg .pred p;\\n\\tsetp.ne.b32 p, %4, 0;\\n\\t{mma}\\n\\t}}" :: "r"(tmem_base), "r"(a), "l"(b_hi + {b_tile // 16}ull), "r"({idesc}u), "r"(1u) : "memory");
                        asm volatile("{{\\n\\t.reg .pred p;\\n\\tsetp.ne.b32 p, %4, 0;\\n\\t{mma}\\n\\t}}" :: "r"(tmem_base), "r"(a), "l"(b_hi), "r"({idesc}u), "r"(1u) : "memory");
                    }}
                }}
                __syncwarp();
            }}
            if (elect_one()) asm volatile("{commit}" {commit_args});
            __syncwarp();
            mbar_wait(bar, 0u);
            const long long t1 = clock64();
            if (lane == 0) {{ out[blockIdx.x] = (float)(t1 - t0); }}
        }} else {{
            mbar_wait(bar, 0u);                     // the multicast commit releases the partner too
        }}
        stop = 1u;
    }} else if (warp >= 2u) {{
        // competing shared-memory traffic: what the A / B splitters of the real kernel move per k-block
        const unsigned tid = threadIdx.x - 64u;
        float4* src = reinterpret_cast<float4*>(smem_gen + 131072u);               // 32 KiB read window
        float4* dst = reinterpret_cast<float4*>(smem_gen + 131072u + 32768u);      // 32 KiB write window
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        while (stop == 0u) {{
            #pragma unroll 4
            for (unsigned i = tid; i < {bytes_per_kblock // 32}u; i += {32 * max(1, traffic_warps)}u) {{
                const float4 v = src[i & 2047u];
                acc.x += v.x;
                dst[i & 2047u] = make_float4(v.x - acc.x, v.y, v.z, v.w);
            }}
        }}
        if (acc.x == 123.456f) out[1000 + tid] = acc.x;
    }}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    {"cluster_sync_all();" if pair else ""}
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::{cg}.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512));
}}
"""
    grid = 148
    return Kernel("mma_rate", src, (grid, 1, 1), (nthreads, 1, 1), 200 * 1024, 4096)


shim.init()
print(f"{'config':44s} {'clk / MMA':>10s} {'clk / k-block (12 MMAs)':>24s}")
for pair, N in ((False, 64), (False, 128), (False, 256), (True, 128), (True, 256)):
    for traffic_warps, nbytes in ((0, 0), (8, 65536), (8, 131072)):
        k = kernel(pair, N, traffic_warps, nbytes)
        plan = cops.Plan(k, cluster=2 if pair else 1)
        out = plan.run()
        shim.synchronize()
        if shim.COMPILE_ONLY:
            continue
        res = transfer.download(out, 4096)[:148]
        res = res[res > 0]
        clk = float(np.median(res))
        what = f"{'pair M=256' if pair else '1 CTA M=128'} N={N} traffic={'none' if not traffic_warps else str(nbytes // 1024) + ' KiB r+w loop'}"
        print(f"{what:44s} {clk / (ITERS * 12):10.1f} {clk / ITERS:24.0f}", flush=True)
