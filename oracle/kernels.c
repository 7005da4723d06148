/*
 * oracle/kernels.c — CPU restatement of dlgrad's generated C kernels.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under dlgrad_b200/ may import, link or call this file; it exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg have an independent statement of what
 * the reference computes.  Every function cites the reference generator it follows
 * (dlgrad/codegen/cpu_kernel.py, "ck:" below) and is compiled with the reference's own flags
 * (dlgrad/runtime/cpu.py:30: -O3 -march=native -ffast-math, -lm -lmvec) so that the compiler makes
 * the same value-changing choices (vector libm, FMA contraction, multi-accumulator sums).
 *
 * The reference emits one function per ndim with nested loops; here ranks are padded to 4 with
 * leading 1s, which leaves the visiting order (row-major over the output) and every arithmetic
 * expression unchanged.
 *
 * Pinned against: tests/golden/*.npz (outputs of the real reference run in the build container,
 * generator committed as tests/golden/make_golden.py) and, when present, oracle/_ref/*.so (the
 * reference's own generated C compiled by oracle/build.py).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOOP4(shape)                                   \
    for (int i0 = 0; i0 < (shape)[0]; i0++)            \
    for (int i1 = 0; i1 < (shape)[1]; i1++)            \
    for (int i2 = 0; i2 < (shape)[2]; i2++)            \
    for (int i3 = 0; i3 < (shape)[3]; i3++)
#define OFF(s) (i0 * (s)[0] + i1 * (s)[1] + i2 * (s)[2] + i3 * (s)[3])

/* ck:12-66 arithmetic(op, ndim): out[ptr++] = x[x_idx] OP y[y_idx], broadcast by 0-strides. */
#define DEF_BINARY(name, EXPR)                                                                  \
    void orc_##name(const float *x, const float *y, float *out, const int *shape, const int *xs, \
                    const int *ys) {                                                            \
        int ptr = 0;                                                                            \
        LOOP4(shape) {                                                                          \
            float a = x[OFF(xs)], b = y[OFF(ys)];                                               \
            out[ptr++] = EXPR;                                                                  \
        }                                                                                       \
    }
DEF_BINARY(add, a + b)
DEF_BINARY(sub, a - b)
DEF_BINARY(mul, a * b)
DEF_BINARY(div, a / b)
DEF_BINARY(eq, a == b)
DEF_BINARY(gt, a > b)
DEF_BINARY(gte, a >= b)
/* ck:653-668 cmp_2d("<="): x <= y ? 1.0 : 0.0 */
DEF_BINARY(lte, a <= b ? 1.0 : 0.0)

/* ck:246-319 utils(func).  The source calls the double exp/log on floats and `1.0/sqrtf`; under
 * -ffast-math gcc narrows them to vector float code, as it does for the reference. */
void orc_neg(const float *x, float *out, int n) { for (int i = 0; i < n; i++) out[i] = -1 * x[i]; }
void orc_exp(const float *x, float *out, int n) { for (int i = 0; i < n; i++) out[i] = exp(x[i]); }
void orc_log(const float *x, float *out, int n) { for (int i = 0; i < n; i++) out[i] = log(x[i]); }
void orc_pow(const float *x, float *out, int n, int val) { for (int i = 0; i < n; i++) out[i] = powf(x[i], val); }
void orc_sqrt(const float *x, float *out, int n) { for (int i = 0; i < n; i++) out[i] = sqrtf(x[i]); }
void orc_rsqrt(const float *x, float *out, int n) { for (int i = 0; i < n; i++) out[i] = 1.0 / sqrtf(x[i]); }

/* ck:520-533 relu */
void orc_relu(const float *x, float *out, int n) { for (int i = 0; i < n; i++) out[i] = x[i] <= 0 ? 0 : x[i]; }

/* ck:321-349 clamp(has_min, has_max) */
void orc_clamp(const float *x, float *out, int n, int has_min, float lo, int has_max, float hi) {
    for (int i = 0; i < n; i++) {
        float v = x[i];
        if (has_min && v < lo) v = lo;
        if (has_max && v > hi) v = hi;
        out[i] = v;
    }
}

/* ck:536-587 where(ndim): out = cond > 0 ? x : y, three broadcast stride sets */
void orc_where(const float *cond, const float *x, const float *y, float *out, const int *shape,
               const int *cs, const int *xs, const int *ys) {
    int ptr = 0;
    LOOP4(shape) { out[ptr++] = (cond[OFF(cs)] > 0) ? x[OFF(xs)] : y[OFF(ys)]; }
}

/* ck:604-650 masked_fill(ndim): out = mask > 0 ? fill : x */
void orc_masked_fill(const float *x, const float *mask, float *out, const int *shape, const int *xs,
                     const int *ms, float fill) {
    int ptr = 0;
    LOOP4(shape) { out[ptr++] = (mask[OFF(ms)] > 0) ? fill : x[OFF(xs)]; }
}

/* ck:69-148 reduce_source(op, ndim, dim): outer loops over kept dims, sequential inner loop over
 * the reduced dim; viewed here as (outer, R, inner).  max seeds with -FLT_MAX and uses strict >. */
void orc_reduce_sum(const float *x, float *out, int outer, int R, int inner) {
    for (int o = 0; o < outer; o++)
        for (int i = 0; i < inner; i++) {
            float val = 0;
            for (int r = 0; r < R; r++) val += x[((size_t)o * R + r) * inner + i];
            out[(size_t)o * inner + i] = val;
        }
}
void orc_reduce_mean(const float *x, float *out, int outer, int R, int inner) {
    for (int o = 0; o < outer; o++)
        for (int i = 0; i < inner; i++) {
            float val = 0;
            for (int r = 0; r < R; r++) val += x[((size_t)o * R + r) * inner + i];
            out[(size_t)o * inner + i] = val / R;
        }
}
void orc_reduce_max(const float *x, float *out, int outer, int R, int inner) {
    for (int o = 0; o < outer; o++)
        for (int i = 0; i < inner; i++) {
            float val = -FLT_MAX;
            for (int r = 0; r < R; r++) {
                float v = x[((size_t)o * R + r) * inner + i];
                if (v > val) val = v;
            }
            out[(size_t)o * inner + i] = val;
        }
}
/* ck:151-177 reduce(op): flat loop over numel */
void orc_reduce_all_sum(const float *x, float *out, int n) { float v = 0; for (int i = 0; i < n; i++) v += x[i]; out[0] = v; }
void orc_reduce_all_mean(const float *x, float *out, int n) { float v = 0; for (int i = 0; i < n; i++) v += x[i]; out[0] = v / n; }
void orc_reduce_all_max(const float *x, float *out, int n) {
    float v = -FLT_MAX;
    for (int i = 0; i < n; i++) if (x[i] > v) v = x[i];
    out[0] = v;
}

/* ck:460-517 argmax(ndim, dim): first strict maximum wins, index stored as float */
void orc_argmax(const float *x, float *out, int outer, int R, int inner) {
    for (int o = 0; o < outer; o++)
        for (int i = 0; i < inner; i++) {
            float max_val = -3.402823466e+38F;
            int best = 0;
            for (int r = 0; r < R; r++) {
                float v = x[((size_t)o * R + r) * inner + i];
                if (v > max_val) { max_val = v; best = r; }
            }
            out[(size_t)o * inner + i] = (float)best;
        }
}

/* ck:180-215 permute(ndim): out[contiguous] = x[permuted strides] */
void orc_permute(const float *x, float *out, const int *shape, const int *in_strides) {
    int ptr = 0;
    LOOP4(shape) { out[ptr++] = x[OFF(in_strides)]; }
}

/* ck:352-428 matmul_{2,3,4}d: i-k-j loops, out pre-zeroed by the caller, `+=` per product
 * (dlgrad/runtime/cpu.py:505-602 zeroes out and zeroes the batch stride of a broadcast operand).
 * bs* are the two batch strides (0 when that operand is broadcast over the batch dim). */
void orc_matmul(const float *x, const float *y, float *out, int B0, int B1, int M, int K, int N,
                int xb0, int xb1, int yb0, int yb1) {
    for (int b = 0; b < B0; b++)
        for (int c = 0; c < B1; c++) {
            const float *A = x + (size_t)b * xb0 + (size_t)c * xb1;
            const float *Bm = y + (size_t)b * yb0 + (size_t)c * yb1;
            float *O = out + ((size_t)b * B1 + c) * M * N;
            for (int i = 0; i < M; i++)
                for (int k = 0; k < K; k++) {
                    float a = A[(size_t)i * K + k];
                    const float *brow = Bm + (size_t)k * N;
                    float *orow = O + (size_t)i * N;
                    for (int j = 0; j < N; j++) orow[j] += a * brow[j];
                }
        }
}

/* ck:431-441 ce_forward / ck:444-457 ce_backward (in place) */
void orc_ce_forward(const float *x, const float *target, float *out, int n_rows, int stride0) {
    for (int i = 0; i < n_rows; i++) out[i] = x[(int)target[i] + (stride0 * i)];
}
void orc_ce_backward(float *x, const float *target, int n_rows, int stride0) {
    for (int i = 0; i < n_rows; i++) x[(int)target[i] + (stride0 * i)] -= 1;
}

/* ck:218-229 embedding (row memcpy) / ck:232-243 embedding_backward (sequential scatter-add
 * into a zeroed table) */
void orc_embedding(const float *x, const float *idx, float *out, int idx_numel, int width) {
    for (int i = 0; i < idx_numel; i++) {
        int x_idx = idx[i] * width;
        memcpy(&out[(size_t)i * width], &x[x_idx], width * sizeof(float));
    }
}
void orc_embedding_backward(float *out, const float *upstream, const float *idx, int idx_numel, int width) {
    for (int i = 0; i < idx_numel; i++)
        for (int j = 0; j < width; j++) {
            int x_idx = idx[i] * width + j;
            out[x_idx] += upstream[(size_t)i * width + j];
        }
}

/* ck:758-770 full: int parameter — the fill value is truncated (dlgrad/runtime/cpu.py:252) */
void orc_full(float *out, int n, int fill_value) { for (int i = 0; i < n; i++) out[i] = fill_value; }
/* ck:743-755 arange */
void orc_arange(float *out, int n) { for (int i = 0; i < n; i++) out[i] = i; }
