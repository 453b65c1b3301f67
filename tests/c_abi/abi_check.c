/* Plain-C consumer of include/spamm_b200.h: shows that the boundary is a C ABI (no C++, no torch, no
 * Python in the signatures) and checks one small model end to end.
 *
 *   gcc -std=c99 -Wall -Werror -I include tests/c_abi/abi_check.c -o abi_check \
 *       -L spamm_b200 -lspamm_b200 -Wl,-rpath,$PWD/spamm_b200 -lm
 *
 * Without a CUDA device: checks the ABI version, the error path and that spamm_plan_create fails
 * loudly (SPAMM_ECUDA) -- prints "abi ok, no device".  With a device: a NuclearContinuum-only model
 * (NuclearContinuumComponent.flux, NC:176-201; Model.likelihood, Model.py:418-445; strict flat priors,
 * Model.py:449-470) evaluated through spamm_lnpost_batch_host and compared with the closed form
 * computed here in C -- prints "lnpost ok".  Test infrastructure only. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spamm_b200.h"

#define P 37           /* odd and smaller than one 64-pixel warp task */
#define W 5

int main(void) {
    if (spamm_abi_version() != SPAMM_ABI_VERSION) {
        fprintf(stderr, "ABI version %d, header %d\n", spamm_abi_version(), SPAMM_ABI_VERSION);
        return 1;
    }
    SpammPlan* plan = (SpammPlan*)0x1;
    if (spamm_plan_create(NULL, &plan) != SPAMM_EINVAL || strlen(spamm_last_error()) == 0) {
        fprintf(stderr, "null descriptor was not rejected\n");
        return 1;
    }

    double wl[P], ln_wl[P], flux[P], err[P], sum_ln = 0.0;
    const double pi = 3.141592653589793;
    for (int i = 0; i < P; ++i) {
        wl[i] = 1200.0 + 3.5 * i;
        ln_wl[i] = log(wl[i]);
    }
    const double wl0 = wl[P / 2];                                 /* np.median of an odd-length grid */
    for (int i = 0; i < P; ++i) {
        flux[i] = 4e-15 * pow(wl[i] / wl0, -1.7) * (1.0 + 0.01 * ((i * 7) % 5 - 2));
        err[i] = (i % 9 == 4 ? -0.05 : 0.05) * flux[i];           /* negative errors occur in real data */
        sum_ln += log(2.0 * pi * err[i] * err[i]);
    }
    const double lo[2] = {0.0, -3.0}, hi[2] = {1e-14, 3.0};       /* norm_PL, slope1 (NC:148-170) */
    /* walker 3 sits outside the box (-> -inf), walker 4 exactly on a bound (strict -> -inf) */
    const double params[W][2] = {{4e-15, 1.7}, {3.1e-15, 2.4}, {9.9e-15, -2.9}, {1.1e-14, 1.0}, {5e-15, 3.0}};

    SpammPlanDesc d;
    memset(&d, 0, sizeof d);
    d.abi_version = SPAMM_ABI_VERSION;
    d.n_pix = P;
    d.wl = wl; d.ln_wl = ln_wl; d.flux = flux; d.flux_error = err;
    d.sum_ln_2pi_err2 = sum_ln;
    d.n_components = 1;
    d.component_kind[0] = SPAMM_COMP_NUCLEAR;
    d.n_dim = 2;
    d.prior_lo = lo; d.prior_hi = hi;
    d.nc_broken_pl = 0;
    d.nc_norm_wl = wl0;
    d.max_walkers = W;

    plan = NULL;
    int rc = spamm_plan_create(&d, &plan);
    if (rc == SPAMM_ECUDA) {
        if (plan != NULL || strlen(spamm_last_error()) == 0) {
            fprintf(stderr, "failed plan left state behind\n");
            return 1;
        }
        printf("abi ok, no device (%s)\n", spamm_last_error());
        return 0;
    }
    if (rc != SPAMM_OK) {
        fprintf(stderr, "spamm_plan_create: %d %s\n", rc, spamm_last_error());
        return 1;
    }
    if (spamm_plan_n_dim(plan) != 2 || spamm_plan_n_pix(plan) != P) {
        fprintf(stderr, "plan shape\n");
        return 1;
    }
    double got[W];
    rc = spamm_lnpost_batch_host(plan, &params[0][0], W, got);
    if (rc != SPAMM_OK) {
        fprintf(stderr, "spamm_lnpost_batch_host: %d %s\n", rc, spamm_last_error());
        return 1;
    }
    int bad = 0;
    for (int w = 0; w < W; ++w) {
        double want;
        if (!(lo[0] < params[w][0] && params[w][0] < hi[0] && lo[1] < params[w][1] && params[w][1] < hi[1])) {
            want = -INFINITY;
        } else {
            double chi = 0.0;
            for (int i = 0; i < P; ++i) {
                double m = params[w][0] * pow(wl[i] / wl0, -params[w][1]);
                double r = (flux[i] - m) / err[i];
                chi += r * r;
            }
            want = -0.5 * (chi + sum_ln);
        }
        const int ok = isinf(want) ? (got[w] == want) : fabs(got[w] - want) <= 1e-10 * fabs(want);
        printf("walker %d: %.17g (C closed form %.17g) %s\n", w, got[w], want, ok ? "ok" : "MISMATCH");
        bad += !ok;
    }
    if (spamm_plan_launch_count(plan) < 1) { fprintf(stderr, "no kernel was launched\n"); bad++; }
    spamm_plan_destroy(plan);
    if (bad) return 1;
    printf("lnpost ok\n");
    return 0;
}
