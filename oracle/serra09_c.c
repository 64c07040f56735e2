/* Plain-C restatement of the reference's Serra09 pair score (TEST INFRASTRUCTURE / CPU baseline).
 *
 * PARITY UNPINNED: this restates essentia ChromaCrossSimilarity + CoverSongSimilarity
 * (third-party, unpinned 'essentia' extra of the reference, /root/reference/setup.py:53, absent
 * from this image) as SURVEY.md Appendix A describes them.  Anchors in the reference:
 *   /root/reference/acoss/algorithms/rqa_serra09.py:60-67   call sites + parameters
 *   /root/reference/acoss/algorithms/algorithm_template.py:172-177  joblib fan-out over pair chunks
 * It is validated against oracle/serra09_np.py (tests/test_oracle_serra09.py), never shipped, and
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * The loops deliberately mirror essentia's scalar structure (three dot products per cell, one
 * sort per row and per column, full score matrix) because this file is also the timed
 * "reference CPU path" (kind = "port").  hoist_norms=1 computes a.a and b.b once per row/column
 * instead (same values, fewer flops) and is reported separately.
 *
 * Also restates the in-tree numba Smith-Waterman (alignment_tools.py:26-46) for C2 timing.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NB 12

typedef struct {
    int m, tau;
    float kappa;
    int oti, noti;
    float gamma_o, gamma_e;
    int integer_guard;   /* F1 alternative */
    int hoist_norms;
} oracle_params;

/* essentia dotProduct: float products, sequential double accumulator (F3), float return */
static float dotf(const float *a, const float *b, int n) {
    double acc = 0.0;
    for (int k = 0; k < n; k++) { float p = a[k] * b[k]; acc += (double)p; }
    return (float)acc;
}

static void global_chroma(const float *x, int n, float *g) {
    for (int b = 0; b < NB; b++) g[b] = 0.f;
    for (int i = 0; i < n; i++) for (int b = 0; b < NB; b++) g[b] = g[b] + x[i * NB + b];
    float mx = g[0];
    for (int b = 1; b < NB; b++) if (g[b] > mx) mx = g[b];
    if (mx != 0.f) for (int b = 0; b < NB; b++) g[b] = g[b] / mx;
}

int oracle_oti(const float *q, int nq, const float *r, int nr, int noti) {
    float gq[NB], gr[NB], rot[NB];
    global_chroma(q, nq, gq);
    global_chroma(r, nr, gr);
    int best = 0; float bv = 0.f;
    for (int s = 0; s <= noti; s++) {
        for (int b = 0; b < NB; b++) rot[b] = gr[((b - s) % NB + NB) % NB];
        float v = dotf(gq, rot, NB);
        if (s == 0 || v > bv) { bv = v; best = s; }
    }
    return best;
}

static int cmpf(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

static float percentile_sorted(const float *s, int L, float q, int guard) {
    float k = (L > 1) ? (float)(L - 1) * q : (float)L * q;
    float fk = floorf(k), ck = ceilf(k);
    if (guard && fk == ck) return s[(int)fk];
    float d0 = s[(int)fk] * (ck - k);
    float d1 = s[(int)ck] * (k - fk);
    return d0 + d1;
}

static float kappa_q(float kappa) {
    float q = kappa * 100.f;
    return (float)((double)q / 100.);
}

/* Qmax (serra09, symmetric).  Returns max cell. */
float oracle_qmax(const uint8_t *c, int M, int N, float go, float ge) {
    if (M <= 0 || N <= 0) return 0.f;
    float *Q = (float *)calloc((size_t)M * N, sizeof(float));
    float best = 0.f;
    for (int i = 2; i < M; i++) {
        for (int j = 2; j < N; j++) {
            float c1 = Q[(size_t)(i - 1) * N + j - 1], c2 = Q[(size_t)(i - 2) * N + j - 1],
                  c3 = Q[(size_t)(i - 1) * N + j - 2], v;
            if (c[(size_t)i * N + j] == 1) {
                v = fmaxf(fmaxf(c1, c2), c3) + 1.f;
            } else {
                c1 -= c[(size_t)(i - 1) * N + j - 1] ? go : ge;
                c2 -= c[(size_t)(i - 2) * N + j - 1] ? go : ge;
                c3 -= c[(size_t)(i - 1) * N + j - 2] ? go : ge;
                v = fmaxf(fmaxf(c1, c2), fmaxf(c3, 0.f));
            }
            Q[(size_t)i * N + j] = v;
            if (v > best) best = v;
        }
    }
    free(Q);
    return best;
}

/* Dmax (chen17, symmetric) -- the second score of ChenFusion.similarity (latefusion_chen.py:69-73).
 * Five predecessors, i, j >= 3; `bonus` (F10) adds Chen's bridging terms c[i-1][j], c[i][j-1], ... to the
 * skipped-cell predecessors.  float32, operations left to right.  UNPINNED restatement. */
float oracle_dmax(const uint8_t *c, int M, int N, float go, float ge, int bonus) {
    if (M <= 0 || N <= 0) return 0.f;
    float *D = (float *)calloc((size_t)M * N, sizeof(float));
    float best = 0.f;
#define CC(i, j) c[(size_t)(i) * N + (j)]
#define DD(i, j) D[(size_t)(i) * N + (j)]
    for (int i = 3; i < M; i++) {
        for (int j = 3; j < N; j++) {
            float P[5];
            P[0] = DD(i - 1, j - 1);
            P[1] = DD(i - 2, j - 1); P[2] = DD(i - 1, j - 2);
            P[3] = DD(i - 3, j - 1); P[4] = DD(i - 1, j - 3);
            if (bonus) {
                P[1] = P[1] + (float)CC(i - 1, j);
                P[2] = P[2] + (float)CC(i, j - 1);
                P[3] = (P[3] + (float)CC(i - 2, j)) + (float)CC(i - 1, j);
                P[4] = (P[4] + (float)CC(i, j - 2)) + (float)CC(i, j - 1);
            }
            float v;
            if (CC(i, j) == 1) {
                v = fmaxf(fmaxf(fmaxf(P[0], P[1]), fmaxf(P[2], P[3])), P[4]) + 1.f;
            } else {
                P[0] -= CC(i - 1, j - 1) ? go : ge;
                P[1] -= CC(i - 2, j - 1) ? go : ge;
                P[2] -= CC(i - 1, j - 2) ? go : ge;
                P[3] -= CC(i - 3, j - 1) ? go : ge;
                P[4] -= CC(i - 1, j - 3) ? go : ge;
                v = fmaxf(fmaxf(fmaxf(P[0], P[1]), fmaxf(P[2], P[3])), fmaxf(P[4], 0.f));
            }
            DD(i, j) = v;
            if (v > best) best = v;
        }
    }
#undef CC
#undef DD
    free(D);
    return best;
}

/* in-tree smith_waterman_constrained (alignment_tools.py:26-46), float64 like the reference */
double oracle_sw_constrained(const uint8_t *B, int M, int N) {
    double best = 0.0;
    if (N < 4 || M < 4) return best;
    double *S = (double *)calloc((size_t)M * N, sizeof(double));
#define BB(i, j) B[(size_t)(i) * N + (j)]
#define SS(i, j) S[(size_t)(i) * N + (j)]
    for (int i = 3; i < M; i++) {
        for (int j = 3; j < N; j++) {
            double mv = BB(i - 1, j - 1) == 1 ? 1.0 : -1.0;
            double d1 = SS(i - 1, j - 1) + mv + (BB(i - 2, j - 2) > 0 ? 0.0 : -0.7);
            double d2 = SS(i - 2, j - 1) + mv + (BB(i - 3, j - 2) > 0 ? 0.0 : -0.7);
            double d3 = SS(i - 1, j - 2) + mv + (BB(i - 2, j - 3) > 0 ? 0.0 : -0.7);
            double v = fmax(fmax(d1, d2), fmax(d3, 0.0));
            SS(i, j) = v;
            if (v > best) best = v;
        }
    }
#undef BB
#undef SS
    free(S);
    return best;
}

/* One pair.  Outputs (any may be NULL): oti, crp (M'*N' bytes), thr_q (M'), thr_r (N'),
 * d (M'*N' floats).  Returns 0 ok, -1 too-short/empty input, -2 NaN distance (F7). */
static int pair_impl(const float *q, int nq, const float *r_in, int nr, const oracle_params *p,
                     float *score, float *score_dmax, int *oti_out, uint8_t *crp_out, float *thr_q_out,
                     float *thr_r_out, float *d_out) {
    const int m = p->m, tau = p->tau, dim = m * NB, incr = m * tau;
    if (nq <= 0 || nr <= 0) return -1;
    if (m > 1 && (nq < incr + 1 || nr < incr + 1)) return -1;
    const int M = (m == 1) ? nq : nq - incr, N = (m == 1) ? nr : nr - incr;
    if (M < 2 || N < 2) return -1;   /* F9: essentia percentile() of a 1-element vector reads out of bounds */
    int s = 0;
    float *r = (float *)malloc((size_t)nr * NB * sizeof(float));
    if (p->oti) s = oracle_oti(q, nq, r_in, nr, p->noti);
    for (int j = 0; j < nr; j++)
        for (int b = 0; b < NB; b++) r[j * NB + b] = r_in[j * NB + ((b - s) % NB + NB) % NB];
    if (oti_out) *oti_out = s;
    /* stackChromaFrames */
    float *qs = (float *)malloc((size_t)M * dim * sizeof(float));
    float *rs = (float *)malloc((size_t)N * dim * sizeof(float));
    for (int i = 0; i < M; i++) for (int t = 0; t < m; t++) memcpy(qs + (size_t)i * dim + t * NB, q + (size_t)(i + t * tau) * NB, NB * sizeof(float));
    for (int j = 0; j < N; j++) for (int t = 0; t < m; t++) memcpy(rs + (size_t)j * dim + t * NB, r + (size_t)(j + t * tau) * NB, NB * sizeof(float));
    /* pairwiseDistance */
    float *d = (float *)malloc((size_t)M * N * sizeof(float));
    float *bbv = NULL;
    if (p->hoist_norms) {
        bbv = (float *)malloc((size_t)N * sizeof(float));
        for (int j = 0; j < N; j++) bbv[j] = dotf(rs + (size_t)j * dim, rs + (size_t)j * dim, dim);
    }
    int has_nan = 0;
    for (int i = 0; i < M; i++) {
        const float *a = qs + (size_t)i * dim;
        float aa_h = p->hoist_norms ? dotf(a, a, dim) : 0.f;
        for (int j = 0; j < N; j++) {
            const float *b = rs + (size_t)j * dim;
            float aa = p->hoist_norms ? aa_h : dotf(a, a, dim);
            float bb = p->hoist_norms ? bbv[j] : dotf(b, b, dim);
            float item = aa - 2 * dotf(a, b, dim) + bb;
            float v = sqrtf(item);
            if (v != v) has_nan = 1;
            d[(size_t)i * N + j] = v;
        }
    }
    if (d_out) memcpy(d_out, d, (size_t)M * N * sizeof(float));
    /* percentile thresholds: one sorted copy per row and per column */
    const float qq = kappa_q(p->kappa);
    float *thr_q = (float *)malloc((size_t)M * sizeof(float));
    float *thr_r = (float *)malloc((size_t)N * sizeof(float));
    float *tmp = (float *)malloc((size_t)(M > N ? M : N) * sizeof(float));
    for (int i = 0; i < M; i++) {
        memcpy(tmp, d + (size_t)i * N, (size_t)N * sizeof(float));
        qsort(tmp, N, sizeof(float), cmpf);
        thr_q[i] = percentile_sorted(tmp, N, qq, p->integer_guard);
    }
    for (int j = 0; j < N; j++) {
        for (int i = 0; i < M; i++) tmp[i] = d[(size_t)i * N + j];
        qsort(tmp, M, sizeof(float), cmpf);
        thr_r[j] = percentile_sorted(tmp, M, qq, p->integer_guard);
    }
    if (thr_q_out) memcpy(thr_q_out, thr_q, (size_t)M * sizeof(float));
    if (thr_r_out) memcpy(thr_r_out, thr_r, (size_t)N * sizeof(float));
    /* heaviside AND */
    uint8_t *c = (uint8_t *)malloc((size_t)M * N);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            float v = d[(size_t)i * N + j];
            c[(size_t)i * N + j] = (uint8_t)(((thr_q[i] - v) >= 0.f) && ((thr_r[j] - v) >= 0.f));
        }
    if (crp_out) memcpy(crp_out, c, (size_t)M * N);
    int rc = 0;
    if (has_nan) rc = -2;
    else {
        if (score) *score = oracle_qmax(c, M, N, p->gamma_o, p->gamma_e);
        if (score_dmax) *score_dmax = oracle_dmax(c, M, N, p->gamma_o, p->gamma_e, 1);
    }
    free(c); free(tmp); free(thr_r); free(thr_q); free(bbv); free(d); free(rs); free(qs); free(r);
    return rc;
}

int oracle_serra09_pair(const float *q, int nq, const float *r_in, int nr, const oracle_params *p,
                        float *score, int *oti_out, uint8_t *crp_out, float *thr_q_out,
                        float *thr_r_out, float *d_out) {
    return pair_impl(q, nq, r_in, nr, p, score, NULL, oti_out, crp_out, thr_q_out, thr_r_out, d_out);
}

/* ---- batched, multi-threaded driver (stands in for joblib.Parallel over pair chunks) ---- */
typedef struct {
    const float *frames; const int64_t *offsets; const int32_t *pairs; int64_t K;
    const oracle_params *p; float *scores; int *status; int64_t *next; pthread_mutex_t *mu;
    float *scores_dmax;
} job_t;

static void *worker(void *arg) {
    job_t *jb = (job_t *)arg;
    for (;;) {
        pthread_mutex_lock(jb->mu);
        int64_t k = (*jb->next)++;
        pthread_mutex_unlock(jb->mu);
        if (k >= jb->K) break;
        int32_t a = jb->pairs[2 * k], b = jb->pairs[2 * k + 1];
        const float *q = jb->frames + jb->offsets[a] * NB, *r = jb->frames + jb->offsets[b] * NB;
        float sc = 0.f, sd = 0.f;
        int rc = pair_impl(q, (int)(jb->offsets[a + 1] - jb->offsets[a]), r,
                           (int)(jb->offsets[b + 1] - jb->offsets[b]), jb->p, &sc, jb->scores_dmax ? &sd : NULL,
                           NULL, NULL, NULL, NULL, NULL);
        jb->scores[k] = sc;
        if (jb->scores_dmax) jb->scores_dmax[k] = sd;
        if (rc != 0) *jb->status = rc;
    }
    return NULL;
}

/* ChenFusion.similarity over a batch: qmax and dmax of the same CRP (latefusion_chen.py:58-73);
 * scores_dmax may be NULL (= Serra09.similarity). */
int oracle_chen_pairs(const float *frames, const int64_t *offsets, const int32_t *pairs, int64_t K,
                      const oracle_params *p, int nthreads, float *scores_qmax, float *scores_dmax) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t th[1024];
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int64_t next = 0; int status = 0;
    job_t jb = {frames, offsets, pairs, K, p, scores_qmax, &status, &next, &mu, scores_dmax};
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &jb);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return status;
}

int oracle_serra09_pairs(const float *frames, const int64_t *offsets, const int32_t *pairs, int64_t K,
                         const oracle_params *p, int nthreads, float *scores) {
    return oracle_chen_pairs(frames, offsets, pairs, K, p, nthreads, scores, NULL);
}

/* batched Smith-Waterman over byte matrices laid out back to back */
typedef struct { const uint8_t *B; const int64_t *off; const int32_t *shape; int64_t K; double *out; int64_t *next; pthread_mutex_t *mu; } swjob_t;
static void *swworker(void *arg) {
    swjob_t *jb = (swjob_t *)arg;
    for (;;) {
        pthread_mutex_lock(jb->mu); int64_t k = (*jb->next)++; pthread_mutex_unlock(jb->mu);
        if (k >= jb->K) break;
        jb->out[k] = oracle_sw_constrained(jb->B + jb->off[k], jb->shape[2 * k], jb->shape[2 * k + 1]);
    }
    return NULL;
}
int oracle_sw_batch(const uint8_t *B, const int64_t *off, const int32_t *shape, int64_t K, int nthreads, double *out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t th[1024]; pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER; int64_t next = 0;
    swjob_t jb = {B, off, shape, K, out, &next, &mu};
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, swworker, &jb);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
