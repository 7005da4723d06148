// libdlgrad_cuda.so — the Device.CUDA runtime shim (C ABI in include/dlgrad_cuda.h).
//
// Plain C++17 over the CUDA *driver* API and NCCL.  Both libraries are resolved with
// dlopen at dlg_init()/dlg_nccl_* time, so this object loads (and its symbol table can be
// checked) on a machine without a GPU driver, and fails loudly — not silently on a CPU
// path — the moment device work is requested there.
//
// Design notes
//   * one primary context, one compute stream, one comm stream (NCCL), one helper event;
//   * caching allocator: exact-size free lists keyed by the rounded request.  A train step
//     asks for the same sizes every iteration, so after the first step every dlg_alloc is
//     a vector pop.  All frees come from the single Python thread that also launches, and
//     all kernels run on the one compute stream, so reuse is stream-ordered by construction;
//   * launch plans freeze everything about a shape-specialised kernel; dlg_plan_run is
//     "allocate output + cuLaunchKernelEx" in one FFI crossing.
#include "../../include/dlgrad_cuda.h"

#include <cuda.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

// ------------------------------------------------------------------------------------
// driver entry points (resolved lazily; names go through cuda.h's _v2 macros)
// ------------------------------------------------------------------------------------
#define DLG_STR2(x) #x
#define DLG_STR(x) DLG_STR2(x)
#define DRV_FUNCS(X)                                                                     \
    X(cuInit) X(cuDeviceGet) X(cuDeviceGetCount) X(cuDeviceGetName) X(cuDeviceGetAttribute) \
    X(cuDevicePrimaryCtxRetain) X(cuCtxSetCurrent) X(cuCtxSynchronize) X(cuGetErrorString) \
    X(cuMemAlloc) X(cuMemFree) X(cuMemAllocHost) X(cuMemFreeHost) X(cuMemcpyHtoDAsync)     \
    X(cuMemcpyDtoHAsync) X(cuMemcpyDtoDAsync) X(cuMemsetD32Async) X(cuMemsetD8Async)      \
    X(cuStreamCreate) X(cuStreamCreateWithPriority) X(cuCtxGetStreamPriorityRange) X(cuStreamSynchronize) X(cuStreamWaitEvent) X(cuEventCreate)        \
    X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime) X(cuEventDestroy)        \
    X(cuModuleLoad) X(cuModuleLoadData) X(cuModuleGetFunction) X(cuFuncSetAttribute)      \
    X(cuLaunchKernel) X(cuLaunchKernelEx) X(cuTensorMapEncodeTiled)                       \
    X(cuStreamBeginCapture) X(cuStreamEndCapture) X(cuGraphInstantiate) X(cuGraphLaunch)  \
    X(cuGraphExecDestroy) X(cuGraphDestroy) X(cuMemGetInfo)                               \
    X(cuIpcGetMemHandle) X(cuIpcOpenMemHandle) X(cuIpcCloseMemHandle)

namespace {

struct Driver {
#define X(name) decltype(&name) p_##name = nullptr;
    DRV_FUNCS(X)
#undef X
    void *handle = nullptr;
};

Driver drv;
thread_local char g_err[512] = "";
thread_local int g_status = 0;

std::mutex g_mu;
bool g_ready = false;
int g_device = -1;
CUdevice g_cudev = 0;
CUcontext g_ctx = nullptr;
CUstream g_stream = nullptr;
CUstream g_comm_stream = nullptr;
CUevent g_ev_compute = nullptr;
CUevent g_ev_comm = nullptr;
char g_name[256] = "";
uint64_t g_launches = 0;

int fail(int code, const char *what) {
    const char *txt = nullptr;
    if (drv.p_cuGetErrorString && code > 0) drv.p_cuGetErrorString((CUresult)code, &txt);
    snprintf(g_err, sizeof g_err, "%s: %s (%d)", what, txt ? txt : "error", code);
    g_status = code;
    return code;
}

#define CU(call)                                             \
    do {                                                     \
        CUresult _r = drv.p_##call;                          \
        if (_r != CUDA_SUCCESS) return fail((int)_r, #call); \
    } while (0)

int load_driver() {
    if (drv.handle) return 0;
    drv.handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!drv.handle) {
        snprintf(g_err, sizeof g_err,
                 "libcuda.so.1 not found: Device.CUDA needs an NVIDIA driver (%s)", dlerror());
        g_status = -100;
        return -100;
    }
#define X(name)                                                                  \
    drv.p_##name = (decltype(&name))dlsym(drv.handle, DLG_STR(name));            \
    if (!drv.p_##name) {                                                         \
        snprintf(g_err, sizeof g_err, "driver symbol missing: %s", DLG_STR(name)); \
        g_status = -101;                                                         \
        return -101;                                                             \
    }
    DRV_FUNCS(X)
#undef X
    return 0;
}

// ------------------------------------------------------------------------------------
// caching allocator
// ------------------------------------------------------------------------------------
// A captured CUDA graph replays with the addresses it was captured with, so every block handed out while a
// graph is being captured belongs to that graph's private POOL: when Python frees it — during the capture or
// any time later — it goes back to the pool, never to the global cache, and only allocations of the same
// capture can get it again.  Eager allocations between replays therefore never alias memory the graph writes.
struct Pool {
    std::unordered_map<size_t, std::vector<void *>> cache;
    bool dead = false;                                         // its graph is gone: frees release to the driver
};

struct Allocator {
    struct Block { size_t size; int pool; };
    std::unordered_map<void *, Block> live;                   // handed out: ptr -> rounded size, owning pool (0 = none)
    std::unordered_map<size_t, std::vector<void *>> cache;    // rounded size -> free blocks
    std::unordered_map<int, Pool> pools;
    int cur_pool = 0, next_pool = 1;
    size_t live_bytes = 0, reserved_bytes = 0, peak_bytes = 0;

    static size_t round(size_t n) {
        if (n == 0) n = 1;
        if (n <= (1u << 20)) return (n + 511) & ~size_t(511);            // 512 B granules
        if (n <= (32u << 20)) return (n + 0xFFFF) & ~size_t(0xFFFF);      // 64 KiB granules
        return (n + 0x1FFFFF) & ~size_t(0x1FFFFF);                        // 2 MiB granules
    }

    void release_cached() {
        for (auto &kv : cache) {
            for (void *p : kv.second) {
                drv.p_cuMemFree((CUdeviceptr)p);
                reserved_bytes -= kv.first;
            }
            kv.second.clear();
        }
    }

    void *get(size_t nbytes) {
        size_t sz = round(nbytes);
        void *p = nullptr;
        if (cur_pool) {
            auto &pc = pools[cur_pool].cache;
            auto pit = pc.find(sz);
            if (pit != pc.end() && !pit->second.empty()) {
                p = pit->second.back();
                pit->second.pop_back();
            }
        }
        auto it = cache.find(sz);
        if (p) {
        } else if (it != cache.end() && !it->second.empty()) {
            p = it->second.back();
            it->second.pop_back();
        } else {
            CUdeviceptr d = 0;
            CUresult r = drv.p_cuMemAlloc(&d, sz);
            if (r == CUDA_ERROR_OUT_OF_MEMORY) {   // give cached blocks back and retry once
                drv.p_cuStreamSynchronize(g_stream);
                release_cached();
                r = drv.p_cuMemAlloc(&d, sz);
            }
            if (r != CUDA_SUCCESS) {
                fail((int)r, "cuMemAlloc");
                return nullptr;
            }
            reserved_bytes += sz;
            p = (void *)d;
        }
        live.emplace(p, Block{sz, cur_pool});
        live_bytes += sz;
        if (live_bytes > peak_bytes) peak_bytes = live_bytes;
        return p;
    }

    void put(void *p) {
        auto it = live.find(p);
        if (it == live.end()) return;   // not ours (a view into a block, or already freed)
        size_t sz = it->second.size;
        int pool = it->second.pool;
        live.erase(it);
        live_bytes -= sz;
        if (!pool) {
            cache[sz].push_back(p);
            return;
        }
        Pool &pl = pools[pool];
        if (pl.dead) {
            drv.p_cuMemFree((CUdeviceptr)p);
            reserved_bytes -= sz;
        } else {
            pl.cache[sz].push_back(p);
        }
    }

    void destroy_pool(int id) {
        auto it = pools.find(id);
        if (it == pools.end()) return;
        it->second.dead = true;                               // blocks still held by Python are released when freed
        for (auto &kv : it->second.cache) {
            for (void *p : kv.second) {
                drv.p_cuMemFree((CUdeviceptr)p);
                reserved_bytes -= kv.first;
            }
            kv.second.clear();
        }
    }
};

Allocator g_alloc;
bool g_shutdown = false;

struct AtExit {
    ~AtExit() { g_shutdown = true; }   // late ffi.gc destructors become no-ops
} g_atexit;

// ------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------
struct Plan {
    CUfunction fn;
    unsigned grid[3], block[3], smem;
    size_t out_bytes;
    int zero_out, cluster_x;
};

struct MMPlan {
    CUfunction fn;
    unsigned grid[3], threads, smem;
    int cluster_x;
    size_t out_bytes;
    dlg_tmap_desc a, b, c;
};

int set_smem_attr(CUfunction fn, unsigned smem) {
    if (smem > 48 * 1024)
        CU(cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    return 0;
}

// Every generated kernel starts with griddepcontrol.launch_dependents + griddepcontrol.wait
// (runtime/cuda/compiler.py), so launches may be programmatically serialised: the next kernel's CTAs become
// resident and run their prologue while this one drains, and block in griddepcontrol.wait until it has
// completed and flushed.  Measured: 2.4 us less GPU idle per launch than plain stream order.
static int g_pdl = -1;

int launch_ex(CUfunction fn, const unsigned g[3], const unsigned b[3], unsigned smem,
              int cluster_x, void **args) {
    g_launches++;
    if (g_pdl < 0) {
        const char *e = getenv("DLGRAD_PDL");
        g_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    if (cluster_x <= 1 && !g_pdl) {
        CU(cuLaunchKernel(fn, g[0], g[1], g[2], b[0], b[1], b[2], smem, g_stream, args, nullptr));
        return 0;
    }
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDimX = g[0]; cfg.gridDimY = g[1]; cfg.gridDimZ = g[2];
    cfg.blockDimX = b[0]; cfg.blockDimY = b[1]; cfg.blockDimZ = b[2];
    cfg.sharedMemBytes = smem;
    cfg.hStream = g_stream;
    CUlaunchAttribute at[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        at[n].id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
        at[n].value.clusterDim.x = (unsigned)cluster_x;
        at[n].value.clusterDim.y = 1;
        at[n].value.clusterDim.z = 1;
        n++;
    }
    if (g_pdl) {
        at[n].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        at[n].value.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    CU(cuLaunchKernelEx(&cfg, fn, args, nullptr));
    return 0;
}

int encode_tmap(CUtensorMap *out, const dlg_tmap_desc &d, const void *base) {
    cuuint64_t dims[3] = {d.dims[0], d.dims[1], d.dims[2]};
    cuuint64_t strides[2] = {d.strides_bytes[0], d.strides_bytes[1]};
    cuuint32_t box[3] = {d.box[0], d.box[1], d.box[2]};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMapSwizzle sw = d.swizzle == 4   ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                            : d.swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : d.swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : d.swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                             : CU_TENSOR_MAP_SWIZZLE_NONE;
    CU(cuTensorMapEncodeTiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.rank, const_cast<void *>(base),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    return 0;
}

// ------------------------------------------------------------------------------------
// NCCL (dlopen'd; the torch wheel bundles 2.28.9, the system has 2.27.3)
// ------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, void *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    ncclComm_t comm = nullptr;
} nccl;

int load_nccl() {
    if (nccl.handle) return 0;
    const char *env = getenv("DLGRAD_NCCL_LIB");
    const char *cands[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *c : cands) {
        if (!c) continue;
        nccl.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (nccl.handle) break;
    }
    if (!nccl.handle) {
        snprintf(g_err, sizeof g_err, "libnccl.so.2 not found (%s)", dlerror());
        return g_status = -110;
    }
#define N(field, sym)                                                     \
    *(void **)(&nccl.field) = dlsym(nccl.handle, sym);                    \
    if (!nccl.field) {                                                    \
        snprintf(g_err, sizeof g_err, "nccl symbol missing: %s", sym);    \
        return g_status = -111;                                           \
    }
    N(GetUniqueId, "ncclGetUniqueId")
    N(CommInitRank, "ncclCommInitRank")
    N(AllReduce, "ncclAllReduce")
    N(CommDestroy, "ncclCommDestroy")
    N(GetVersion, "ncclGetVersion")
    N(GetErrorString, "ncclGetErrorString")
#undef N
    return 0;
}

int nccl_fail(int r, const char *what) {
    snprintf(g_err, sizeof g_err, "%s: %s (%d)", what,
             nccl.GetErrorString ? nccl.GetErrorString(r) : "nccl error", r);
    return g_status = (r ? 1000 + r : 0);
}

#define REQUIRE_READY(ret)                                                       \
    if (!g_ready) {                                                              \
        snprintf(g_err, sizeof g_err, "dlg_init() has not succeeded; no CUDA device bound"); \
        g_status = -102;                                                         \
        return ret;                                                              \
    }

}  // namespace

// ====================================================================================
// C ABI
// ====================================================================================
extern "C" {

const char *dlg_last_error(void) { return g_err; }
int dlg_last_status(void) { return g_status; }
const char *dlg_device_name(void) { return g_name; }
uint64_t dlg_launch_count(void) { return g_launches; }

int dlg_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ready) {
        if (device == g_device) return 0;
        snprintf(g_err, sizeof g_err, "already bound to device %d (one process per GPU)", g_device);
        return g_status = -103;
    }
    if (int r = load_driver()) return r;
    CU(cuInit(0));
    CU(cuDeviceGet(&g_cudev, device));
    CU(cuDevicePrimaryCtxRetain(&g_ctx, g_cudev));
    CU(cuCtxSetCurrent(g_ctx));
    {
        // The compute stream outranks the comm stream.  The compute kernels are persistent — one CTA per SM for the whole
        // kernel — so a compute CTA that cannot become resident doubles its kernel's time; the exchange kernels of the
        // overlapped data-parallel step are many small CTAs that fill the registers a GEMM CTA leaves free.  With
        // this order a pending compute CTA always goes first and waits for at most one short exchange CTA.
        // (DLGRAD_COMM_PRIORITY=high reverses it, for measurements.)
        int least = 0, greatest = 0;
        if (drv.p_cuCtxGetStreamPriorityRange(&least, &greatest) != CUDA_SUCCESS) least = greatest = 0;
        const char *e = getenv("DLGRAD_COMM_PRIORITY");
        const bool comm_high = e && e[0] == 'h';
        CU(cuStreamCreateWithPriority(&g_stream, CU_STREAM_NON_BLOCKING, comm_high ? least : greatest));
        CU(cuStreamCreateWithPriority(&g_comm_stream, CU_STREAM_NON_BLOCKING, comm_high ? greatest : least));
    }
    CU(cuEventCreate(&g_ev_compute, CU_EVENT_DISABLE_TIMING));
    CU(cuEventCreate(&g_ev_comm, CU_EVENT_DISABLE_TIMING));
    drv.p_cuDeviceGetName(g_name, sizeof g_name, g_cudev);
    g_device = device;
    g_ready = true;
    g_status = 0;
    return 0;
}

int dlg_device_count(void) {
    if (load_driver()) return 0;
    if (drv.p_cuInit(0) != CUDA_SUCCESS) return 0;
    int n = 0;
    if (drv.p_cuDeviceGetCount(&n) != CUDA_SUCCESS) return 0;
    return n;
}

int dlg_device_attr(int attr) {
    REQUIRE_READY(-1)
    int v = -1;
    if (drv.p_cuDeviceGetAttribute(&v, (CUdevice_attribute)attr, g_cudev) != CUDA_SUCCESS) return -1;
    return v;
}

int dlg_sync(void) {
    REQUIRE_READY(g_status)
    drv.p_cuCtxSetCurrent(g_ctx);
    CU(cuStreamSynchronize(g_stream));
    CU(cuStreamSynchronize(g_comm_stream));
    return 0;
}

// ---- allocator ---------------------------------------------------------------------
void *dlg_alloc(size_t nbytes) {
    REQUIRE_READY(nullptr)
    std::lock_guard<std::mutex> lk(g_mu);
    return g_alloc.get(nbytes);
}

void *dlg_calloc(size_t nbytes) {
    void *p = dlg_alloc(nbytes);
    if (!p) return nullptr;
    size_t sz = Allocator::round(nbytes);
    if (drv.p_cuMemsetD8Async((CUdeviceptr)p, 0, sz, g_stream) != CUDA_SUCCESS) {
        fail(-1, "cuMemsetD8Async");
        dlg_free(p);
        return nullptr;
    }
    return p;
}

void dlg_free(void *dptr) {
    if (!dptr || !g_ready || g_shutdown) return;
    std::lock_guard<std::mutex> lk(g_mu);
    g_alloc.put(dptr);
}

int dlg_empty_cache(void) {
    REQUIRE_READY(g_status)
    std::lock_guard<std::mutex> lk(g_mu);
    drv.p_cuStreamSynchronize(g_stream);
    g_alloc.release_cached();
    return 0;
}

size_t dlg_mem_live(void) { return g_alloc.live_bytes; }
size_t dlg_mem_reserved(void) { return g_alloc.reserved_bytes; }
size_t dlg_mem_peak(void) { return g_alloc.peak_bytes; }
void dlg_mem_reset_peak(void) { g_alloc.peak_bytes = g_alloc.live_bytes; }

// ---- transfers ---------------------------------------------------------------------
int dlg_h2d(void *dst, const void *src, size_t n) {
    REQUIRE_READY(g_status)
    if (!n) return 0;
    CU(cuMemcpyHtoDAsync((CUdeviceptr)dst, src, n, g_stream));
    CU(cuStreamSynchronize(g_stream));   // the host buffer may be pageable and short-lived
    return 0;
}

int dlg_d2h(void *dst, const void *src, size_t n) {
    REQUIRE_READY(g_status)
    if (!n) return 0;
    CU(cuMemcpyDtoHAsync(dst, (CUdeviceptr)src, n, g_stream));
    CU(cuStreamSynchronize(g_stream));
    return 0;
}

int dlg_d2d(void *dst, const void *src, size_t n) {
    REQUIRE_READY(g_status)
    if (!n) return 0;
    CU(cuMemcpyDtoDAsync((CUdeviceptr)dst, (CUdeviceptr)src, n, g_stream));
    return 0;
}

int dlg_memset32(void *dst, uint32_t word, size_t nwords) {
    REQUIRE_READY(g_status)
    if (!nwords) return 0;
    CU(cuMemsetD32Async((CUdeviceptr)dst, word, nwords, g_stream));
    return 0;
}

void *dlg_host_alloc(size_t n) {
    REQUIRE_READY(nullptr)
    void *p = nullptr;
    if (drv.p_cuMemAllocHost(&p, n ? n : 1) != CUDA_SUCCESS) {
        fail(-1, "cuMemAllocHost");
        return nullptr;
    }
    return p;
}

void dlg_host_free(void *p) {
    if (p && g_ready && !g_shutdown) drv.p_cuMemFreeHost(p);
}

int dlg_h2d_async(void *dst, const void *src, size_t n) {
    REQUIRE_READY(g_status)
    if (!n) return 0;
    CU(cuMemcpyHtoDAsync((CUdeviceptr)dst, src, n, g_stream));
    return 0;
}

int dlg_d2h_async(void *dst, const void *src, size_t n) {
    REQUIRE_READY(g_status)
    if (!n) return 0;
    CU(cuMemcpyDtoHAsync(dst, (CUdeviceptr)src, n, g_stream));
    return 0;
}

// ---- modules / launches ------------------------------------------------------------
int dlg_module_load(const char *path, void **module_out) {
    REQUIRE_READY(g_status)
    CUmodule m = nullptr;
    CU(cuModuleLoad(&m, path));
    *module_out = (void *)m;
    return 0;
}

int dlg_module_get(void *module, const char *name, void **fn_out) {
    REQUIRE_READY(g_status)
    CUfunction f = nullptr;
    CU(cuModuleGetFunction(&f, (CUmodule)module, name));
    *fn_out = (void *)f;
    return 0;
}

int dlg_launch(void *fn, const unsigned grid[3], const unsigned block[3], unsigned smem,
               void **args) {
    REQUIRE_READY(g_status)
    if (int r = set_smem_attr((CUfunction)fn, smem)) return r;
    return launch_ex((CUfunction)fn, grid, block, smem, 1, args);
}

void *dlg_plan_create(void *fn, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                      unsigned bz, unsigned smem, size_t out_bytes, int zero_out, int cluster_x) {
    REQUIRE_READY(nullptr)
    if (set_smem_attr((CUfunction)fn, smem)) return nullptr;
    Plan *p = new Plan{(CUfunction)fn, {gx, gy, gz}, {bx, by, bz}, smem, out_bytes, zero_out, cluster_x};
    return p;
}

void dlg_plan_destroy(void *plan) { delete (Plan *)plan; }

int dlg_plan_run_into(void *plan, void *out, const void *p0, const void *p1, const void *p2,
                      const void *p3, float f0, float f1, float f2, float f3) {
    Plan *p = (Plan *)plan;
    if (p->zero_out && out && p->out_bytes)
        CU(cuMemsetD8Async((CUdeviceptr)out, 0, p->out_bytes, g_stream));
    void *args[9] = {&p0, &p1, &p2, &p3, &out, &f0, &f1, &f2, &f3};
    g_status = 0;
    return launch_ex(p->fn, p->grid, p->block, p->smem, p->cluster_x, args);
}

void *dlg_plan_run(void *plan, const void *p0, const void *p1, const void *p2, const void *p3,
                   float f0, float f1, float f2, float f3) {
    Plan *p = (Plan *)plan;
    void *out = nullptr;
    if (p->out_bytes) {
        out = dlg_alloc(p->out_bytes);
        if (!out) return nullptr;
    }
    if (dlg_plan_run_into(plan, out, p0, p1, p2, p3, f0, f1, f2, f3) != 0) {
        if (out) dlg_free(out);
        return nullptr;
    }
    return out;
}

// ---- tensor-core matmul plans ---------------------------------------------------------
void *dlg_mm_plan_create(void *fn, unsigned gx, unsigned gy, unsigned gz, unsigned threads,
                         unsigned smem, int cluster_x, size_t out_bytes, const dlg_tmap_desc *a,
                         const dlg_tmap_desc *b, const dlg_tmap_desc *c) {
    REQUIRE_READY(nullptr)
    if (set_smem_attr((CUfunction)fn, smem)) return nullptr;
    MMPlan *p = new MMPlan;
    p->fn = (CUfunction)fn;
    p->grid[0] = gx; p->grid[1] = gy; p->grid[2] = gz;
    p->threads = threads; p->smem = smem; p->cluster_x = cluster_x; p->out_bytes = out_bytes;
    p->a = *a; p->b = *b; p->c = *c;
    return p;
}

static int mm_launch(MMPlan *p, void *c, const void *a, const void *b, const void *aux0, void *aux1) {
    alignas(64) CUtensorMap ta, tb, tc;
    if (int r = encode_tmap(&ta, p->a, a)) return r;
    if (int r = encode_tmap(&tb, p->b, b)) return r;
    if (int r = encode_tmap(&tc, p->c, c)) return r;
    void *args[6] = {&ta, &tb, &tc, &c, &aux0, &aux1};      // kernels with four parameters ignore the rest
    unsigned block[3] = {p->threads, 1, 1};
    g_status = 0;
    return launch_ex(p->fn, p->grid, block, p->smem, p->cluster_x, args);
}

int dlg_mm_plan_run_into(void *plan, void *c, const void *a, const void *b) {
    return mm_launch((MMPlan *)plan, c, a, b, nullptr, nullptr);
}

void *dlg_mm_plan_run_aux(void *plan, const void *a, const void *b, const void *aux0, void *aux1) {
    MMPlan *p = (MMPlan *)plan;
    void *c = dlg_alloc(p->out_bytes);
    if (!c) return nullptr;
    if (mm_launch(p, c, a, b, aux0, aux1) != 0) {
        dlg_free(c);
        return nullptr;
    }
    return c;
}

void *dlg_mm_plan_run(void *plan, const void *a, const void *b) {
    return dlg_mm_plan_run_aux(plan, a, b, nullptr, nullptr);
}

// ---- fused tensor-core kernels with several (rank <= 5) tensor maps ---------------------
struct TCPlan {
    CUfunction fn;
    unsigned grid, threads, smem;
    int ntmaps, nptrs;
    dlg_tmap_desc5 maps[8];
};

static int encode_tmap5(CUtensorMap *out, const dlg_tmap_desc5 &d, const void *base) {
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 5; ++i) { dims[i] = d.dims[i]; box[i] = d.box[i]; }
    for (int i = 0; i < 4; ++i) strides[i] = d.strides_bytes[i];
    CUtensorMapSwizzle sw = d.swizzle == 4   ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                            : d.swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : d.swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : d.swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                             : CU_TENSOR_MAP_SWIZZLE_NONE;
    CU(cuTensorMapEncodeTiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.rank, const_cast<void *>(base),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    return 0;
}

void *dlg_tc_plan_create(void *fn, unsigned grid_x, unsigned threads, unsigned smem, int ntmaps,
                         const dlg_tmap_desc5 *maps, int nptrs) {
    REQUIRE_READY(nullptr)
    if (ntmaps < 0 || ntmaps > 8 || nptrs < 0 || nptrs > 8) {
        fail(-2, "dlg_tc_plan_create: at most 8 tensor maps and 8 pointers");
        return nullptr;
    }
    if (set_smem_attr((CUfunction)fn, smem)) return nullptr;
    TCPlan *p = new TCPlan;
    p->fn = (CUfunction)fn;
    p->grid = grid_x; p->threads = threads; p->smem = smem; p->ntmaps = ntmaps; p->nptrs = nptrs;
    for (int i = 0; i < ntmaps; ++i) p->maps[i] = maps[i];
    return p;
}

int dlg_tc_plan_run(void *plan, const void *const *bases, void *const *ptrs) {
    TCPlan *p = (TCPlan *)plan;
    alignas(64) CUtensorMap tm[8];
    void *pv[8];
    void *args[16];
    int n = 0;
    for (int i = 0; i < p->ntmaps; ++i) {
        if (int r = encode_tmap5(&tm[i], p->maps[i], bases[i])) return r;
        args[n++] = &tm[i];
    }
    for (int i = 0; i < p->nptrs; ++i) {
        pv[i] = ptrs[i];
        args[n++] = &pv[i];
    }
    unsigned grid[3] = {p->grid, 1, 1}, block[3] = {p->threads, 1, 1};
    g_status = 0;
    return launch_ex(p->fn, grid, block, p->smem, 1, args);
}

// ---- events / graphs -----------------------------------------------------------------
void *dlg_event_create(void) {
    REQUIRE_READY(nullptr)
    CUevent e = nullptr;
    if (drv.p_cuEventCreate(&e, CU_EVENT_DEFAULT) != CUDA_SUCCESS) return nullptr;
    return (void *)e;
}
int dlg_event_record(void *ev) {
    REQUIRE_READY(g_status)
    CU(cuEventRecord((CUevent)ev, g_stream));
    return 0;
}
int dlg_event_sync(void *ev) {
    REQUIRE_READY(g_status)
    CU(cuEventSynchronize((CUevent)ev));
    return 0;
}
float dlg_event_elapsed_ms(void *a, void *b) {
    float ms = -1.f;
    if (!g_ready) return ms;
    if (drv.p_cuEventElapsedTime(&ms, (CUevent)a, (CUevent)b) != CUDA_SUCCESS) return -1.f;
    return ms;
}
void dlg_event_destroy(void *ev) {
    if (ev && g_ready && !g_shutdown) drv.p_cuEventDestroy((CUevent)ev);
}

int dlg_graph_begin(void) {
    REQUIRE_READY(g_status)
    CU(cuStreamBeginCapture(g_stream, CU_STREAM_CAPTURE_MODE_RELAXED));
    return 0;
}
void *dlg_graph_end(void) {
    REQUIRE_READY(nullptr)
    CUgraph g = nullptr;
    if (drv.p_cuStreamEndCapture(g_stream, &g) != CUDA_SUCCESS || !g) {
        fail(-1, "cuStreamEndCapture");
        return nullptr;
    }
    CUgraphExec ge = nullptr;
    CUresult r = drv.p_cuGraphInstantiate(&ge, g, 0);
    drv.p_cuGraphDestroy(g);
    if (r != CUDA_SUCCESS) {
        fail((int)r, "cuGraphInstantiate");
        return nullptr;
    }
    return (void *)ge;
}
int dlg_pool_begin(void) {
    REQUIRE_READY(-1)
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_alloc.cur_pool) {
        snprintf(g_err, sizeof g_err, "a capture pool is already open");
        g_status = -120;
        return -1;
    }
    g_alloc.cur_pool = g_alloc.next_pool++;
    g_alloc.pools[g_alloc.cur_pool];
    return g_alloc.cur_pool;
}
int dlg_pool_end(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_alloc.cur_pool = 0;
    return 0;
}
int dlg_pool_destroy(int pool) {
    if (!g_ready || g_shutdown) return 0;
    drv.p_cuStreamSynchronize(g_stream);                      // no replay may still be writing the pool's blocks
    std::lock_guard<std::mutex> lk(g_mu);
    g_alloc.destroy_pool(pool);
    return 0;
}
int dlg_graph_launch(void *ge) {
    REQUIRE_READY(g_status)
    CU(cuGraphLaunch((CUgraphExec)ge, g_stream));
    return 0;
}
void dlg_graph_destroy(void *ge) {
    if (ge && g_ready && !g_shutdown) drv.p_cuGraphExecDestroy((CUgraphExec)ge);
}

// ---- NCCL ----------------------------------------------------------------------------
int dlg_nccl_version(void) {
    if (load_nccl()) return -1;
    int v = 0;
    nccl.GetVersion(&v);
    return v;
}

int dlg_nccl_unique_id(char out[128]) {
    if (int r = load_nccl()) return r;
    ncclUniqueId id;
    int r = nccl.GetUniqueId(&id);
    if (r) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
    return 0;
}

int dlg_nccl_init(int rank, int world, const char idbytes[128]) {
    REQUIRE_READY(g_status)
    if (int r = load_nccl()) return r;
    ncclUniqueId id;
    memcpy(id.internal, idbytes, 128);
    drv.p_cuCtxSetCurrent(g_ctx);
    int r = nccl.CommInitRank(&nccl.comm, world, id, rank);
    if (r) return nccl_fail(r, "ncclCommInitRank");
    return 0;
}

int dlg_nccl_allreduce_sum(void *dptr, size_t nfloats) {
    REQUIRE_READY(g_status)
    if (!nccl.comm) {
        snprintf(g_err, sizeof g_err, "dlg_nccl_init() has not been called");
        return g_status = -112;
    }
    // comm stream starts after everything queued so far on the compute stream
    CU(cuEventRecord(g_ev_compute, g_stream));
    CU(cuStreamWaitEvent(g_comm_stream, g_ev_compute, 0));
    int r = nccl.AllReduce(dptr, dptr, nfloats, /*ncclFloat32*/ 7, /*ncclSum*/ 0, nccl.comm,
                           (void *)g_comm_stream);
    if (r) return nccl_fail(r, "ncclAllReduce");
    return 0;
}

int dlg_nccl_wait(void) {
    REQUIRE_READY(g_status)
    CU(cuEventRecord(g_ev_comm, g_comm_stream));
    CU(cuStreamWaitEvent(g_stream, g_ev_comm, 0));
    return 0;
}

// ---- peer memory: allocations that other ranks of the node map into their own address space ----------------
// (cuMemAlloc'd directly: the caching allocator never recycles them, and an IPC handle covers exactly this block)
void *dlg_ipc_alloc(size_t nbytes) {
    REQUIRE_READY(nullptr)
    CUdeviceptr p = 0;
    size_t sz = (nbytes + 0x1FFFFF) & ~size_t(0x1FFFFF);
    CUresult r = drv.p_cuMemAlloc(&p, sz);
    if (r != CUDA_SUCCESS) { fail((int)r, "cuMemAlloc (peer memory)"); return nullptr; }
    r = drv.p_cuMemsetD8Async(p, 0, sz, g_stream);
    if (r == CUDA_SUCCESS) r = drv.p_cuStreamSynchronize(g_stream);
    if (r != CUDA_SUCCESS) { fail((int)r, "cuMemsetD8Async (peer memory)"); drv.p_cuMemFree(p); return nullptr; }
    return (void *)p;
}

int dlg_ipc_free(void *dptr) {
    if (!dptr || !g_ready || g_shutdown) return 0;
    CU(cuMemFree((CUdeviceptr)dptr));
    return 0;
}

int dlg_ipc_export(void *dptr, unsigned char handle[64]) {
    REQUIRE_READY(g_status)
    static_assert(sizeof(CUipcMemHandle) == 64, "CUipcMemHandle is 64 bytes");
    CUipcMemHandle h;
    CU(cuIpcGetMemHandle(&h, (CUdeviceptr)dptr));
    memcpy(handle, &h, 64);
    return 0;
}

void *dlg_ipc_open(const unsigned char handle[64]) {
    REQUIRE_READY(nullptr)
    CUipcMemHandle h;
    memcpy(&h, handle, 64);
    CUdeviceptr p = 0;
    CUresult r = drv.p_cuIpcOpenMemHandle(&p, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS);
    if (r != CUDA_SUCCESS) { fail((int)r, "cuIpcOpenMemHandle"); return nullptr; }
    return (void *)p;
}

int dlg_ipc_close(void *mapped) {
    if (!mapped || !g_ready || g_shutdown) return 0;
    CU(cuIpcCloseMemHandle((CUdeviceptr)mapped));
    return 0;
}

// ---- comm-stream services for the overlapped data-parallel exchange ------------------------------------------
int dlg_comm_after_compute(void) {
    REQUIRE_READY(g_status)
    CU(cuEventRecord(g_ev_compute, g_stream));
    CU(cuStreamWaitEvent(g_comm_stream, g_ev_compute, 0));
    return 0;
}

int dlg_compute_after_comm(void) { return dlg_nccl_wait(); }

int dlg_comm_plan_run_into(void *plan, void *out, const void *p0, const void *p1, const void *p2,
                           const void *p3, float f0, float f1, float f2, float f3) {
    REQUIRE_READY(g_status)
    Plan *p = (Plan *)plan;
    if (p->zero_out && out && p->out_bytes)
        CU(cuMemsetD8Async((CUdeviceptr)out, 0, p->out_bytes, g_comm_stream));
    void *args[9] = {&p0, &p1, &p2, &p3, &out, &f0, &f1, &f2, &f3};
    g_status = 0;
    g_launches++;
    // plain stream order (no programmatic serialisation): the kernel follows events and memory operations
    CU(cuLaunchKernel(p->fn, p->grid[0], p->grid[1], p->grid[2], p->block[0], p->block[1], p->block[2],
                      p->smem, g_comm_stream, args, nullptr));
    return 0;
}

// cuStreamWaitValue32 is resolved on demand: older drivers without stream memory operations still load the library
typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static WaitValue32Fn g_wait32 = nullptr;
static int g_wait32_state = -1;

int dlg_comm_memops_supported(void) {
    if (!g_ready) return 0;
    if (g_wait32_state < 0) {
        g_wait32 = (WaitValue32Fn)dlsym(drv.handle, "cuStreamWaitValue32_v2");
        if (!g_wait32) g_wait32 = (WaitValue32Fn)dlsym(drv.handle, "cuStreamWaitValue32");
        g_wait32_state = g_wait32 ? 1 : 0;
        const char *e = getenv("DLGRAD_DP_MEMOPS");
        if (e && e[0] == '0') g_wait32_state = 0;
    }
    return g_wait32_state;
}

int dlg_comm_wait_geq32(void *dptr, uint32_t value) {
    REQUIRE_READY(g_status)
    if (!dlg_comm_memops_supported()) {
        snprintf(g_err, sizeof g_err, "the driver has no stream memory operations (cuStreamWaitValue32)");
        return g_status = -120;
    }
    CUresult r = g_wait32(g_comm_stream, (CUdeviceptr)dptr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) return fail((int)r, "cuStreamWaitValue32");
    return 0;
}

int dlg_nccl_destroy(void) {
    if (nccl.comm && !g_shutdown) {
        nccl.CommDestroy(nccl.comm);
        nccl.comm = nullptr;
    }
    return 0;
}

}  // extern "C"
