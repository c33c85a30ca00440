/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the reference's monotonic alignment search:
 *   mas_width1  — /root/reference/fs2/attn/alignment.py:48-74
 *   b_mas       — /root/reference/fs2/attn/alignment.py:77-85 (and the per-item loop in
 *                 VarianceAdaptor.binarize_attention, fs2/variance_adaptor.py:173-179)
 * The reference JIT-compiles these with numba; this file is the same arithmetic in the
 * same order (fp32 adds, `max`, `>=` tie → diagonal move).  Parity: PINNED against the
 * reference's numba functions by tests/test_oracle.py via tests/golden/kat_mas.npz.
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC (see intops.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* in: [n_mel][in_stride] log-probabilities (mel x text); out: [n_mel][out_stride] 0/1 (only the
 * n_mel x n_text window is written); scratch: n_mel*n_text floats (the in-place log_p of the
 * reference, which must not clobber the caller's input). */
void oracle_mas_width1(const float *in, float *out, float *scratch, int n_mel, int n_text,
                       long in_stride, long out_stride) {
    if (n_mel <= 0 || n_text <= 0) return;
    float *log_p = scratch;
    for (int i = 0; i < n_mel; ++i)
        memcpy(log_p + (long)i * n_text, in + (long)i * in_stride, sizeof(float) * n_text);
    for (int j = 1; j < n_text; ++j) log_p[j] = -INFINITY;            /* :54 */
    for (int i = 1; i < n_mel; ++i) {                                  /* :55-60 */
        float prev1 = -INFINITY;
        const float *up = log_p + (long)(i - 1) * n_text;
        float *cur = log_p + (long)i * n_text;
        for (int j = 0; j < n_text; ++j) {
            float prev2 = up[j];
            float m = prev2 > prev1 ? prev2 : prev1;
            cur[j] = cur[j] + m;
            prev1 = prev2;
        }
    }
    for (int i = 0; i < n_mel; ++i) memset(out + (long)i * out_stride, 0, sizeof(float) * n_text);
    int j = n_text - 1;
    for (int i = n_mel - 1; i > 0; --i) {                              /* :66-72 */
        out[(long)i * out_stride + j] = 1.0f;
        const float *up = log_p + (long)(i - 1) * n_text;
        int jm1 = j - 1 < 0 ? j - 1 + n_text : j - 1;                  /* numpy wrap-around */
        if (up[jm1] >= up[j]) {
            j -= 1;
            if (j == 0) {
                for (int r = 1; r < i; ++r) out[(long)r * out_stride] = 1.0f;
                break;
            }
            if (j < 0) j += n_text;                                    /* only reachable when n_text == 1 */
        }
    }
    out[j] = 1.0f;                                                     /* :73 */
}

/* in/out: [B][1][F][T]; zeros outside [:out_lens[b], :in_lens[b]]. threads<=0: all cores. */
void oracle_b_mas(const float *in, float *out, const int *in_lens, const int *out_lens, int B, int F,
                  int T, int threads) {
    memset(out, 0, sizeof(float) * (size_t)B * F * T);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int n_mel = out_lens[b], n_text = in_lens[b];
        if (n_mel <= 0 || n_text <= 0) continue;
        float *scratch = (float *)malloc(sizeof(float) * (size_t)n_mel * n_text);
        oracle_mas_width1(in + (size_t)b * F * T, out + (size_t)b * F * T, scratch, n_mel, n_text, T, T);
        free(scratch);
    }
}
