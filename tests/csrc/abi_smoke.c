/* abi_smoke.c -- include/varpro_b200.h consumed from plain C (C99, gcc): proves the drop-in boundary is a C ABI and
 * not a ctypes artefact. Builds the reference's double-exponential + offset problem (tests/integration_tests/main.rs:
 * 399-463 shape: m = 20 samples, S = 2 right-hand sides, grid of shared_test_code/src/lib.rs:20-34), fits it through
 * vp_fit and checks tau = (1, 3) and the six linear coefficients to 1e-8.
 * Exit code: 0 = ok, 77 = no CUDA device (the library has no CPU fallback), anything else = failure.
 * Test infrastructure only (tests/test_abi_c_consumer.py compiles and runs it). */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/varpro_b200.h"

#define M 20
#define S 2

static int fail(vp_ctx *ctx, const char *what, int status)
{
    fprintf(stderr, "abi_smoke: %s failed with status %d (%s): %s\n", what, status, vp_status_string(status), vp_last_error(ctx));
    return 1;
}

int main(void)
{
    if (vp_abi_version() != VP_ABI_VERSION) {
        fprintf(stderr, "abi_smoke: header says ABI %d, library says %d\n", VP_ABI_VERSION, vp_abi_version());
        return 2;
    }
    vp_ctx *ctx = NULL;
    int st = vp_ctx_create(0, &ctx);
    if (st == VP_ERR_CUDA) {
        printf("abi_smoke: no CUDA device: %s\n", vp_last_error(NULL));
        return 77;
    }
    if (st != VP_OK) return fail(NULL, "vp_ctx_create", st);

    /* x_i = first + (first - last) / (count - 1) * i: the reference's (bug-compatible) linspace(0, 12.5, 20) */
    double x[M], Y[M * S];
    const double coef[S][3] = {{2.0, 4.0, 0.2}, {5.0, 1.0, 9.0}};
    for (int i = 0; i < M; ++i) x[i] = 0.0 + (0.0 - 12.5) / (M - 1) * i;
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < M; ++i) Y[s * M + i] = coef[s][0] * exp(-x[i] / 1.0) + coef[s][1] * exp(-x[i] / 3.0) + coef[s][2];

    vp_basis_desc basis[3];
    memset(basis, 0, sizeof(basis));
    basis[0].kind = VP_BASIS_EXP_DECAY; basis[0].n_params = 1; basis[0].param_idx[0] = 0;
    basis[1].kind = VP_BASIS_EXP_DECAY; basis[1].n_params = 1; basis[1].param_idx[0] = 1;
    basis[2].kind = VP_BASIS_CONSTANT;  basis[2].n_params = 0;
    vp_model *model = NULL;
    st = vp_model_create(ctx, VP_F64, M, x, 2, 3, basis, &model);
    if (st != VP_OK) return fail(ctx, "vp_model_create", st);

    const double alpha0[2] = {2.5, 6.5};
    vp_problem *problem = NULL;
    st = vp_problem_create(ctx, model, S, Y, M, NULL, -1.0, alpha0, &problem);
    if (st != VP_OK) return fail(ctx, "vp_problem_create", st);

    vp_reduced red;
    st = vp_reduce(problem, &red);
    if (st != VP_OK || !red.finite || red.q != 2) return fail(ctx, "vp_reduce", st);

    vp_lm_options opt = {-1.0, -1.0, -1.0, -1.0, -1, -1}; /* crate defaults */
    vp_fit_report rep;
    st = vp_fit(problem, &opt, &rep);
    if (st != VP_OK) return fail(ctx, "vp_fit", st);
    double alpha[2], C[3 * S];
    if ((st = vp_params(problem, alpha)) != VP_OK) return fail(ctx, "vp_params", st);
    if ((st = vp_linear_coefficients(problem, C)) != VP_OK) return fail(ctx, "vp_linear_coefficients", st);
    const int swapped = alpha[0] > alpha[1];
    const double t1 = swapped ? alpha[1] : alpha[0], t2 = swapped ? alpha[0] : alpha[1];
    double worst = fmax(fabs(t1 - 1.0), fabs(t2 - 3.0) / 3.0);
    for (int s = 0; s < S; ++s) {
        const double c1 = C[s * 3 + (swapped ? 1 : 0)], c2 = C[s * 3 + (swapped ? 0 : 1)], c3 = C[s * 3 + 2];
        worst = fmax(worst, fmax(fabs(c1 - coef[s][0]), fmax(fabs(c2 - coef[s][1]), fabs(c3 - coef[s][2]))));
    }
    printf("abi_smoke: termination %d, %d evaluations, tau = (%.12g, %.12g), worst error %.2e, %lld kernels launched\n",
           rep.termination, rep.number_of_evaluations, t1, t2, worst, (long long)vp_ctx_kernel_launches(ctx));
    vp_problem_destroy(problem);
    vp_model_destroy(model);
    vp_ctx_destroy(ctx);
    if (!rep.successful || worst > 1e-8) return 3;
    return 0;
}
