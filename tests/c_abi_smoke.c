/*
 * Plain-C client of include/eml_b200.h: one qubit precessing under H = E sigmaz with sigmaz dephasing gamma,
 * checked against the closed form <sx>(t) = exp(-2 gamma t) cos(2 E t)   (SURVEY.md section 4 iii, iv).
 * Exit codes: 0 ok, 3 no CUDA device (what a CPU-only box must report), anything else = failure.
 * Built and run by tests/test_c_abi.py:  gcc -std=c99 -pedantic -Wall -Iinclude tests/c_abi_smoke.c -L... -leml_b200
 */
#include <math.h>
#include <stdio.h>

#include "eml_b200.h"

int main(void) {
    const double E = 0.8, gamma = 0.07;
    /* op table: sigmaz, sigmax (row-major 2x2 complex) */
    const double ops[2][8] = {{1, 0, 0, 0, 0, 0, -1, 0}, {0, 0, 1, 0, 1, 0, 0, 0}};
    const double rho0[8] = {0.5, 0, 0.5, 0, 0.5, 0, 0.5, 0}; /* 'plus' */
    const EmlTerm terms[2] = {{0, EML_TERM_ENERGY, 0, -1, 0, -1}, {1, EML_TERM_LINDBLAD, 0, -1, 0, -1}};
    double times[33], expect[2 * 1 * 33 * 2], loss[2], targets[33 * 2];
    const double params[2][2] = {{0.8, 0.07}, {0.8, 0.0}};
    EmlProblemDesc d;
    EmlPlan* plan = NULL;
    int i, rc, ndev = 0;

    if (eml_version() != EML_VERSION_MAJOR * 1000 + EML_VERSION_MINOR) return 10;
    eml_device_count(&ndev);
    for (i = 0; i < 33; ++i) {
        times[i] = 0.3 + 0.1 * i; /* the initial state sits at times[0] */
        targets[2 * i] = exp(-2 * gamma * (times[i] - times[0])) * cos(2 * E * (times[i] - times[0]));
        targets[2 * i + 1] = 0.0;
    }
    d.n_tls = 1; d.n_params = 2; d.n_terms = 2; d.terms = terms;
    d.n_ops = 2; d.op_table = &ops[0][0]; d.rho0 = rho0;
    d.obs_site = 0; d.n_obs = 1; d.obs = ops[1];
    d.n_times = 33; d.times = times; d.n_target_sets = 1; d.targets = targets;
    rc = eml_plan_create(&d, 0, NULL, &plan);
    if (rc == EML_ERR_NO_DEVICE) {
        printf("no device: %s\n", eml_last_error_string());
        return ndev == 0 ? 3 : 11;
    }
    if (rc != EML_OK) { printf("plan: %s\n", eml_last_error_string()); return 12; }
    rc = eml_eval_batch_host(plan, 2, &params[0][0], NULL, expect, loss);
    if (rc != EML_OK) { printf("eval: %s\n", eml_last_error_string()); return 13; }
    for (i = 0; i < 33; ++i) {
        const double t = times[i] - times[0];
        if (fabs(expect[2 * i] - exp(-2 * gamma * t) * cos(2 * E * t)) > 1e-9) return 14;   /* model 0 */
        if (fabs(expect[2 * (33 + i)] - cos(2 * E * t)) > 1e-9) return 15;                  /* model 1: no dephasing */
        if (fabs(expect[2 * i + 1]) > 1e-12) return 16;
    }
    if (!(loss[0] < 1e-18) || !(loss[1] > 1e-4)) return 17;
    d.n_tls = 9;
    if (eml_plan_create(&d, 0, NULL, &plan) != EML_ERR_UNSUPPORTED) return 18;
    printf("c abi ok: kernel %s, loss %.3e %.3e, launches %lld\n", eml_plan_kernel_name(plan), loss[0], loss[1],
           (long long)eml_launch_count());
    return eml_plan_destroy(plan) == EML_OK ? 0 : 19;
}
