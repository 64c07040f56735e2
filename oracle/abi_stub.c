/* TEST INFRASTRUCTURE — a CPU stand-in for the four entry points of include/acoss_b200.h that the INTEGRATION.md
 * binding (acoss_b200/integration.py) calls, backed by the plain-C oracle (serra09_c.c).
 *
 * Purpose: the build container has the reference (/root/reference) but no GPU, the GPU box has a GPU but no
 * reference.  tests/test_reference_class.py uses this stub to execute the binding under the reference's own,
 * unmodified CoverAlgorithm / Serra09 classes here, and commits the resulting score matrix and metrics as a golden
 * that the GPU library must reproduce on the GPU box.  Nothing in acoss_b200/ loads this library; it is not a
 * fallback (acoss_b200._lib binds libacoss_b200.so only and fails without a device). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/acoss_b200.h"

typedef struct {
    int m, tau;
    float kappa;
    int oti, noti;
    float gamma_o, gamma_e;
    int integer_guard, hoist_norms;
} oracle_params;
int oracle_serra09_pairs(const float *frames, const int64_t *offsets, const int32_t *pairs, int64_t K,
                         const oracle_params *p, int nthreads, float *out);

struct acoss_ctx {
    float *frames;
    int64_t *offsets;
    int32_t n;
};
static char g_err[256] = "";

void acoss_default_params(acoss_params *p) {
    memset(p, 0, sizeof(*p));
    p->m = 9; p->tau = 1; p->kappa = 0.095f; p->oti = 1; p->noti = 12; p->gamma_o = 0.5f; p->gamma_e = 0.5f;
}
const char *acoss_last_error(void) { return g_err; }
int acoss_create(acoss_ctx **ctx, int device) {
    (void)device;
    *ctx = (acoss_ctx *)calloc(1, sizeof(acoss_ctx));
    return *ctx ? ACOSS_OK : ACOSS_E_NOMEM;
}
int acoss_destroy(acoss_ctx *c) {
    if (c) { free(c->frames); free(c->offsets); free(c); }
    return ACOSS_OK;
}
int acoss_set_tracks(acoss_ctx *c, const float *frames, const int64_t *offsets, int32_t n, int on_device) {
    if (!c || !frames || !offsets || n <= 0 || on_device) { strcpy(g_err, "stub: bad arguments"); return ACOSS_E_INVALID; }
    free(c->frames); free(c->offsets);
    const int64_t total = offsets[n];
    c->frames = (float *)malloc((size_t)total * 12 * sizeof(float));
    c->offsets = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
    memcpy(c->frames, frames, (size_t)total * 12 * sizeof(float));
    memcpy(c->offsets, offsets, (size_t)(n + 1) * sizeof(int64_t));
    c->n = n;
    return ACOSS_OK;
}
int acoss_score_pairs(acoss_ctx *c, const int32_t *pairs, int64_t K, const acoss_params *p, float *scores) {
    if (!c || !c->frames) { strcpy(g_err, "stub: no tracks"); return ACOSS_E_INVALID; }
    if (p->align != ACOSS_ALIGN_QMAX || p->f2_strict || p->f3_float_acc || p->f4_keep_last || p->f5_asymmetric) {
        strcpy(g_err, "stub: only the default switches");
        return ACOSS_E_INVALID;
    }
    oracle_params op = {p->m, p->tau, p->kappa, p->oti, p->noti, p->gamma_o, p->gamma_e, p->integer_guard, 0};
    const int rc = oracle_serra09_pairs(c->frames, c->offsets, pairs, K, &op, 4, scores);
    if (rc != 0) { strcpy(g_err, "stub: oracle failed (too short / NaN)"); return rc == -2 ? ACOSS_E_NAN : ACOSS_E_TOO_SHORT; }
    return ACOSS_OK;
}
