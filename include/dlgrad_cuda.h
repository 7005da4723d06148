/*
 * dlgrad_cuda.h — C ABI of libdlgrad_cuda.so, the Device.CUDA runtime shim for dlgrad.
 *
 * This is the drop-in boundary.  dlgrad's CPU runtime reaches its kernels through cffi in
 * ABI mode (`ffi.dlopen`, reference dlgrad/runtime/cpu.py:64-73): one dlopen'd object per
 * kernel, raw `float *` in and out, ownership by `ffi.gc(ptr, free_ptr)`
 * (dlgrad/runtime/cpu.py:82-107).  A Device.CUDA runtime binds THIS library the same way
 * (cffi `ffi.dlopen`); there are no torch / C++ types in any signature.
 *
 * Conventions
 *   - every `int` return is 0 on success or a CUresult / ncclResult_t / negative shim code;
 *     nothing throws across the ABI.  dlg_last_error() returns a static, thread-local text.
 *   - all device work is issued on ONE compute stream owned by the shim (plus one comm stream
 *     for NCCL); calls are asynchronous unless documented otherwise.
 *   - device pointers are plain `void *` / `float *` holding CUdeviceptr values.
 */
#ifndef DLGRAD_CUDA_H
#define DLGRAD_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- context -------------------------------------------------------------------------
 * Replaces the implicit "host memory is the device" assumption of the reference
 * (dlgrad/runtime/cpu.py:30-37 picks a compiler at import; there is no device init). */
int         dlg_init(int device);            /* idempotent; retains the primary context      */
int         dlg_device_count(void);
int         dlg_device_attr(int attr);       /* CUdevice_attribute passthrough (SM count…)   */
const char *dlg_last_error(void);
const char *dlg_device_name(void);
int         dlg_sync(void);                  /* wait for the compute (and comm) stream       */

/* ---- caching device allocator --------------------------------------------------------
 * Replaces cpu_kernel.uninitialized_memory / initialized_memory / free_ptr
 * (dlgrad/codegen/cpu_kernel.py:671-740; callers dlgrad/runtime/cpu.py:82-191).
 * dlg_free has the `void (*)(void *)` shape cffi's ffi.gc() wants as a destructor. */
void  *dlg_alloc(size_t nbytes);             /* NULL on failure (caller raises MemoryError)  */
void  *dlg_calloc(size_t nbytes);            /* zero-filled on the compute stream            */
void   dlg_free(void *dptr);                 /* returns the block to the cache               */
int    dlg_empty_cache(void);                /* cuMemFree every cached block                 */
size_t dlg_mem_live(void);                   /* bytes handed out                             */
size_t dlg_mem_reserved(void);               /* bytes held from the driver                   */
size_t dlg_mem_peak(void);
void   dlg_mem_reset_peak(void);

/* ---- transfers -----------------------------------------------------------------------
 * Replace ffi.from_buffer / ffi.buffer / ffi.memmove on host pointers
 * (dlgrad/tensor.py:112-115,125-138,147-164; dlgrad/buffer.py:67-90). */
int   dlg_h2d(void *dst, const void *src, size_t nbytes);   /* returns after the copy        */
int   dlg_d2h(void *dst, const void *src, size_t nbytes);   /* returns after the copy        */
int   dlg_d2d(void *dst, const void *src, size_t nbytes);   /* async on the compute stream   */
int   dlg_memset32(void *dst, uint32_t word, size_t nwords);
void *dlg_host_alloc(size_t nbytes);         /* pinned host memory                           */
void  dlg_host_free(void *hptr);
int   dlg_h2d_async(void *dst, const void *pinned_src, size_t nbytes);
int   dlg_d2h_async(void *pinned_dst, const void *src, size_t nbytes);

/* ---- generated kernels ---------------------------------------------------------------
 * Replace CPU._build_shared_object / _get_handle / _ensure_sig + the cffi call
 * (dlgrad/runtime/cpu.py:47-78).  Kernels are emitted by dlgrad_b200/codegen, compiled by
 * nvcc to sm_100a cubins and loaded here through the driver API. */
int   dlg_module_load(const char *cubin_path, void **module_out);
int   dlg_module_get(void *module, const char *kernel_name, void **fn_out);
int   dlg_launch(void *fn, const unsigned grid[3], const unsigned block[3],
                 unsigned smem_bytes, void **kernel_args);

/* A launch plan freezes (fn, grid, block, smem, output size) for one shape-specialised
 * kernel so that one op = one FFI call.  Every generated kernel has the signature
 *   (const float *p0, const float *p1, const float *p2, const float *p3, float *out,
 *    float f0, float f1, float f2, float f3)
 * and ignores what it does not use. */
void *dlg_plan_create(void *fn, unsigned gx, unsigned gy, unsigned gz,
                      unsigned bx, unsigned by, unsigned bz, unsigned smem_bytes,
                      size_t out_bytes, int zero_out, int cluster_x);
void  dlg_plan_destroy(void *plan);
/* allocates out_bytes (if > 0), launches, returns the output pointer (NULL on error or when
 * out_bytes == 0; check dlg_last_status()). */
void *dlg_plan_run(void *plan, const void *p0, const void *p1, const void *p2, const void *p3,
                   float f0, float f1, float f2, float f3);
/* same, writing into caller-provided storage (in-place ops, optimizer, fused outputs) */
int   dlg_plan_run_into(void *plan, void *out, const void *p0, const void *p1, const void *p2,
                        const void *p3, float f0, float f1, float f2, float f3);
int   dlg_last_status(void);
uint64_t dlg_launch_count(void);             /* kernels launched by this library so far      */

/* ---- TMA-fed tensor-core matmul ------------------------------------------------------
 * Replaces cpu_kernel.matmul_{2,3,4}d (dlgrad/codegen/cpu_kernel.py:352-428; caller
 * dlgrad/runtime/cpu.py:505-602).  The kernel takes three CUtensorMap parameters that the
 * shim encodes per call from the operand addresses. */
typedef struct dlg_tmap_desc {
    uint32_t rank;                  /* 2 or 3                                               */
    uint64_t dims[3];               /* innermost first, elements                            */
    uint64_t strides_bytes[2];      /* strides of dims[1], dims[2]                          */
    uint32_t box[3];                /* box size per dim, elements                           */
    uint32_t swizzle;               /* 0 none, 1 32B, 2 64B, 3 128B, 4 128B with 32B atoms  */
} dlg_tmap_desc;

void *dlg_mm_plan_create(void *fn, unsigned gx, unsigned gy, unsigned gz, unsigned threads,
                         unsigned smem_bytes, int cluster_x, size_t out_bytes,
                         const dlg_tmap_desc *a, const dlg_tmap_desc *b, const dlg_tmap_desc *c);
void *dlg_mm_plan_run(void *plan, const void *a, const void *b);            /* returns C     */
int   dlg_mm_plan_run_into(void *plan, void *c, const void *a, const void *b);
/* same, with two auxiliary operands handed to the kernel as its 5th and 6th parameters: `aux0` = the bias
 * row that the epilogue of a Linear GEMM adds (replaces the separate broadcast ADD of
 * dlgrad/tensor.py:534), `aux1` = the zeroed flag workspace of a stream-K plan.  Either may be NULL. */
void *dlg_mm_plan_run_aux(void *plan, const void *a, const void *b, const void *aux0, void *aux1);

/* ---- fused tensor-core kernels with several tensor maps --------------------------------
 * The flash-attention kernels (dlgrad_b200/codegen/cuda_flash.py; they replace the MATMUL / MUL / MASKED_FILL /
 * softmax / MATMUL chain of examples/gpt.py:49-55, dlgrad/tensor.py:634-649) read q, k, v, dO through rank-4 tensor
 * maps (D, T, heads, batch) — so the (B, T, h, D) layout the projections produce needs no permute kernel — and take
 * several plain pointers.  Kernel signature: (CUtensorMap x ntmaps, pointer x nptrs). */
typedef struct dlg_tmap_desc5 {
    uint32_t rank;                  /* 2..5                                                 */
    uint64_t dims[5];               /* innermost first, elements                            */
    uint64_t strides_bytes[4];      /* strides of dims[1..4]                                */
    uint32_t box[5];
    uint32_t swizzle;               /* as dlg_tmap_desc                                     */
} dlg_tmap_desc5;
void *dlg_tc_plan_create(void *fn, unsigned grid_x, unsigned threads, unsigned smem_bytes,
                         int ntmaps, const dlg_tmap_desc5 *maps, int nptrs);
/* bases[ntmaps]: device address each tensor map is encoded over; ptrs[nptrs]: the pointer parameters */
int   dlg_tc_plan_run(void *plan, const void *const *bases, void *const *ptrs);

/* ---- timing / graphs -----------------------------------------------------------------*/
void *dlg_event_create(void);
int   dlg_event_record(void *ev);            /* on the compute stream                        */
int   dlg_event_sync(void *ev);
float dlg_event_elapsed_ms(void *start, void *stop);
void  dlg_event_destroy(void *ev);
int   dlg_graph_begin(void);                 /* stream capture on the compute stream         */
void *dlg_graph_end(void);                   /* instantiated executable graph                */
int   dlg_graph_launch(void *graph_exec);
void  dlg_graph_destroy(void *graph_exec);
/* memory pool of a captured step (dlgrad_b200/graph.py): every dlg_alloc between dlg_pool_begin and
 * dlg_pool_end belongs to the pool; dlg_free of such a block — then or later — returns it to the pool, never to
 * the global cache, so eager allocations cannot alias memory that replays of the graph write.
 * dlg_pool_destroy (after the graph is destroyed) releases the pool's blocks. */
int   dlg_pool_begin(void);                  /* pool id > 0, or -1                            */
int   dlg_pool_end(void);
int   dlg_pool_destroy(int pool);

/* ---- data-parallel gradient exchange (no reference counterpart) ----------*/
int   dlg_nccl_unique_id(char out[128]);
int   dlg_nccl_init(int rank, int world, const char id[128]);
int   dlg_nccl_allreduce_sum(void *dptr, size_t nfloats);   /* in place, comm stream,
                                                               ordered after compute stream */
int   dlg_nccl_wait(void);                   /* compute stream waits for the comm stream     */
int   dlg_nccl_destroy(void);
int   dlg_nccl_version(void);

/* ---- peer memory over NVLink (no reference counterpart; SURVEY §8e "fused with its collective") ----------
 * A rank allocates its gradient bucket, parameter bucket and flag words with dlg_ipc_alloc, exports a 64-byte
 * handle for each, and maps every peer's with dlg_ipc_open; the fused reduce-scatter + Adam + all-gather kernel
 * (dlgrad_b200/distributed.py) then loads and stores peer memory directly.  One process per GPU, one node. */
void *dlg_ipc_alloc(size_t nbytes);          /* zero-filled, never recycled by the caching allocator */
int   dlg_ipc_free(void *dptr);
int   dlg_ipc_export(void *dptr, unsigned char handle[64]);
void *dlg_ipc_open(const unsigned char handle[64]);      /* NULL on failure                  */
int   dlg_ipc_close(void *mapped);

/* ---- comm-stream services for the OVERLAPPED data-parallel exchange (dlgrad_b200/distributed.py) ----------
 * The gradient exchange of a parameter bucket runs on the comm stream while the compute stream is still in the
 * backward of earlier layers.  Cross-rank waits are stream memory operations (cuStreamWaitValue32 on a counter
 * in this rank's peer-mapped flag block), so a rank that waits for a peer occupies no SM. */
int   dlg_comm_after_compute(void);          /* later comm-stream work runs after everything enqueued so far on
                                                the compute stream                                       */
int   dlg_compute_after_comm(void);          /* and the reverse                                          */
int   dlg_comm_plan_run_into(void *plan, void *out, const void *p0, const void *p1, const void *p2,
                             const void *p3, float f0, float f1, float f2, float f3);
                                             /* dlg_plan_run_into on the comm stream                     */
int   dlg_comm_memops_supported(void);       /* 1 when the driver offers stream memory operations        */
int   dlg_comm_wait_geq32(void *dptr, uint32_t value);
                                             /* comm stream blocks until (int32)(*dptr - value) >= 0     */

#ifdef __cplusplus
}
#endif
#endif /* DLGRAD_CUDA_H */
