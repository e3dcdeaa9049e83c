/*
 * oracle/vq_oracle.c -- TEST INFRASTRUCTURE ONLY (checker, never the product path).
 *
 * Plain-C restatement of the nearest-codeword search + EMA statistics of the
 * reference quantizer (vqvae/layers.py:700-703 cdist+argmin+gather,
 * :638-643 one-hot sums) with the EXACT fp32 arithmetic order of the library
 * call the reference makes, torch.cdist(..., 'donot_use_mm_for_euclid_dist')
 * on the CPU (ATen 2.11, AVX2 and AVX512 dispatch, probed in this container):
 *
 *   agg = 0
 *   for d in [0, 4*floor(D/4)):  agg = fl(agg + fl(diff_d * diff_d))   (two roundings)
 *   for d in the D%4 tail:       agg = fma(diff_d, diff_d, agg)        (one rounding)
 *   dist = sqrtf(agg);  idx = first k minimising dist  (torch.argmin: first minimum)
 *
 * Parity pinning: tests/golden/quantizer_*.npz were produced by the imported
 * reference Quantizer (tests/golden/make_golden.py) and this file is checked
 * bit-for-bit against them in tests/test_oracle_golden.py.
 *
 * Build: see oracle/Makefile (must keep -ffp-contract=off so that the compiler
 * does not fuse the two-rounding part).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* squared distance in the reference's exact summation order */
static inline float ref_dist2(const float *x, const float *e, int D)
{
    const int nv = (D / 4) * 4;
    float agg = 0.0f;
    int d = 0;
    for (; d < nv; ++d) {
        volatile float diff = x[d] - e[d];
        volatile float sq = diff * diff; /* volatile: force the product to round on its own */
        agg = agg + sq;
    }
    for (; d < D; ++d) {
        float diff = x[d] - e[d];
        agg = fmaf(diff, diff, agg);
    }
    return agg;
}

typedef struct {
    const float *x, *embed;
    int64_t lo, hi;
    int D, K;
    int64_t *idx;
} assign_job;

static void *assign_rows(void *arg)
{
    assign_job *j = (assign_job *)arg;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        const float *xi = j->x + (size_t)i * j->D;
        float best = INFINITY;
        int64_t bi = 0;
        for (int k = 0; k < j->K; ++k) {
            float dist = sqrtf(ref_dist2(xi, j->embed + (size_t)k * j->D, j->D));
            if (dist < best) { /* strict: first minimum wins (torch.argmin) */
                best = dist;
                bi = k;
            }
        }
        j->idx[i] = bi;
    }
    return NULL;
}

/*
 * x:     [N, D] row-major fp32 (the reference's flat_input, layers.py:693)
 * embed: [K, D] row-major fp32
 * idx:   [N] int64 out
 * threads <= 0: all online cores.  Rows are independent, so the result does not
 * depend on the thread count.  Returns the number of threads used.
 */
int vq_oracle_assign(const float *x, const float *embed, int64_t N, int D, int K,
                     int64_t *idx, int threads)
{
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads > 256) threads = 256;
    if ((int64_t)threads > N) threads = N > 0 ? (int)N : 1;
    pthread_t tid[256];
    assign_job job[256];
    int64_t per = (N + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = (int64_t)t * per, hi = lo + per;
        if (lo > N) lo = N;
        if (hi > N) hi = N;
        job[t] = (assign_job){x, embed, lo, hi, D, K, idx};
        if (t > 0) pthread_create(&tid[t], NULL, assign_rows, &job[t]);
    }
    assign_rows(&job[0]);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
    return threads;
}

/* full N x K distance matrix, for pinning against torch.cdist itself */
void vq_oracle_cdist(const float *x, const float *embed, int64_t N, int D, int K, float *out)
{
    for (int64_t i = 0; i < N; ++i)
        for (int k = 0; k < K; ++k)
            out[(size_t)i * K + k] = sqrtf(ref_dist2(x + (size_t)i * D, embed + (size_t)k * D, D));
}

/*
 * EMA statistics (layers.py:638-643): n_k = #{i: idx_i = k}, dw_k = sum_{i: idx_i=k} x_i.
 * Accumulated in double so the checker is order-independent; compare with a
 * tolerance (the reference itself sums in fp32 via a matmul).
 */
void vq_oracle_stats(const float *x, const int64_t *idx, int64_t N, int D, int K,
                     double *n, double *dw)
{
    memset(n, 0, sizeof(double) * (size_t)K);
    memset(dw, 0, sizeof(double) * (size_t)K * D);
    for (int64_t i = 0; i < N; ++i) {
        int64_t k = idx[i];
        n[k] += 1.0;
        for (int d = 0; d < D; ++d) dw[(size_t)k * D + d] += (double)x[(size_t)i * D + d];
    }
}
