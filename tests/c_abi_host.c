/* A plain-C host of the drop-in boundary: solves an SDPA file through include/loraine_b200.h ONLY (no Python, no torch).
 * The control flow is the reference's predictor-corrector loop (src/Solvers.jl:304-361, :448-478; src/predictor_corrector.jl:
 * 5-246, regularisation retry :53-97, sigma_update :148-179) with every array expression replaced by one ABI call -- the same
 * text a Julia host keeps.  Usage: c_abi_host file.dat-s [eDIMACS] ; prints "objective <value> iterations <k> status <s>". */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "loraine_b200.h"

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int32_t rc__ = (call);                                                                       \
        if (rc__ < 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, lrn_last_error(h)); return 2; } \
    } while (0)

static double vmin(const double* a, int64_t n, double extra) {
    double m = extra;
    for (int64_t i = 0; i < n; i++) if (a[i] < m) m = a[i];
    return m;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.dat-s [eDIMACS]\n", argv[0]); return 1; }
    const double eDIMACS = argc > 2 ? atof(argv[2]) : 1e-6;
    lrn_options_t opt;
    lrn_default_options(&opt);
    opt.kit = 0; opt.datarank = 0; opt.datasparsity = 8;
    lrn_handle_t h = NULL;
    int32_t rc = lrn_load_sdpa(&h, argv[1], &opt, 1);
    if (rc != LRN_OK) { fprintf(stderr, "lrn_load_sdpa -> %d: %s\n", rc, h ? lrn_last_error(h) : "no usable sm_100 device (no CPU fallback)"); return 2; }
    int64_t n_var, nlmi, nlin;
    CHECK(lrn_get_dims(h, &n_var, &nlmi, &nlin, NULL));
    int64_t* ms = (int64_t*)calloc(nlmi > 0 ? nlmi : 1, sizeof(int64_t));
    CHECK(lrn_get_dims(h, NULL, NULL, NULL, ms));
    double summ = (double)nlin;
    for (int64_t i = 0; i < nlmi; i++) summ += (double)ms[i];
    double* alpha = (double*)calloc(nlmi > 0 ? nlmi : 1, sizeof(double));
    double* beta = (double*)calloc(nlmi > 0 ? nlmi : 1, sizeof(double));
    CHECK(lrn_initial_point(h, 1));                                   /* initpoint = 1, examples/solve_sdpa.jl:51 */
    double sigma = 3.0, tau = 0.95, expon = 3.0, err[6], by = 0, trCX = 0, dx = 0, D = 1.0, mu = 0;
    int status = 0, iter = 0, regcount = 0;
    while (status == 0) {
        if (++iter > 100) status = 4;
        int32_t st4 = 0;
        CHECK(lrn_find_mu(h, &mu));
        CHECK(lrn_prepare_W(h, &st4));
        if (st4) status = 4;
        /* predictor */
        CHECK(lrn_residuals(h));
        CHECK(lrn_schur_assemble(h));
        CHECK(lrn_rhs_predictor(h));
        int32_t which = 3;
        rc = lrn_schur_factor(h);
        if (rc < 0) { fprintf(stderr, "lrn_schur_factor: %s\n", lrn_last_error(h)); return 2; }
        if (rc > 0) {                                                  /* PosDefException: regularise like the reference */
            int icount = 0;
            if (++regcount > 5) { status = 3; break; }
            do {
                CHECK(lrn_schur_shift(h, 1e-4));
                rc = lrn_schur_factor(h);
                if (rc < 0) return 2;
            } while (rc > 0 && ++icount <= 1000);
            if (rc > 0) { status = 3; break; }
            which = 6;
        }
        CHECK(lrn_schur_solve(h, which));
        double al = 1.0, bl = 1.0;
        CHECK(lrn_find_step(h, 1, sigma, mu, tau, alpha, beta, &al, &bl));
        /* sigma_update */
        double step = fmin(vmin(alpha, nlmi, al), vmin(beta, nlmi, bl));
        double ex = (mu > 1e-6) ? (step < 1.0 / sqrt(3.0) ? 1.0 : fmax(expon, 3.0 * step * step))
                                : fmax(1.0, fmin(expon, 3.0 * step * step));
        double tr = 0, dl = 0;
        CHECK(lrn_sigma_trace(h, &tr, &dl));
        if (tr < 0) sigma = 0.8;
        else sigma = fmin(1.0, pow(((nlmi > 0 ? tr : 0.0) + (nlin > 0 ? dl : 0.0)) / summ / mu, ex));
        /* corrector */
        CHECK(lrn_rhs_corrector(h, sigma, mu));
        CHECK(lrn_schur_solve(h, which));
        CHECK(lrn_find_step(h, 0, sigma, mu, tau, alpha, beta, &al, &bl));
        /* check_convergence */
        CHECK(lrn_dimacs(h, err, &by, &trCX, &dx));
        D = (nlmi > 0 ? err[0] : 0.0) + err[1] + err[2] + err[3] + fabs(err[4]) + err[5];
        if (D < eDIMACS) status = 1;
        if (D > 1e55) status = 2;
        else if (fabs(by) > 1e55) status = 3;
    }
    printf("objective %.10f iterations %d status %d dimacs %.3e dual %.10f launches %lld\n", -by, iter, status, D, -trCX - dx,
           (long long)lrn_kernel_launches());
    free(ms); free(alpha); free(beta);
    lrn_destroy(h);
    return status == 1 ? 0 : 3;
}
