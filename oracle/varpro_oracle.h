/*
 * varpro_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99, fp64) of the hot path of geo-ant/varpro v0.13.3:
 * set_params / residuals / jacobian of `impl LeastSquaresProblem for
 * SeparableProblem` (reference: src/solvers/levmar/mod.rs:42-201) and of the
 * MINPACK-lmder trust-region loop the reference delegates to
 * (`levenberg-marquardt` 0.14, call site src/solvers/levmar/mod.rs:247).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library. The product library
 * (varpro_b200/csrc) never links or calls it.
 *
 * Parity pinning: every golden vector the reference's own tests hold for this
 * path is replayed against this oracle in
 * tests/test_oracle_goldens.py. The Rust crate itself cannot be built here (no
 * cargo/rustc), and nalgebra 0.33 / levenberg-marquardt 0.14 are not vendored,
 * so SVD and the LM loop follow their published algorithms (one-sided Jacobi
 * SVD; MINPACK lmder/lmpar/qrfac/qrsolv), cross-checked against
 * scipy.optimize.leastsq (the original Fortran lmder).
 */
#ifndef VARPRO_ORACLE_H
#define VARPRO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Built-in basis-function kinds. */
enum {
    VO_BASIS_EXP_DECAY = 0,    /* exp(-x/tau)            shared_test_code/src/lib.rs:101-114 */
    VO_BASIS_CONSTANT = 1,     /* 1                      shared_test_code/src/lib.rs:123     */
    VO_BASIS_EXP_RATE_COS = 2, /* exp(-a x) cos(b x)     shared_test_code/src/models.rs:321  */
    VO_BASIS_SIN_PHASE = 3,    /* sin(omega x + phi)     src/test_helpers/mod.rs:27-51       */
    VO_BASIS_LINEAR_X = 4      /* scale * x (invariant)  src/model/builder/test.rs:97,101    */
};

#define VO_MAX_BASIS_PARAMS 4

typedef struct {
    int kind;
    int n_params;
    int param_idx[VO_MAX_BASIS_PARAMS]; /* indices into the model parameter vector */
    double scale;                       /* only VO_BASIS_LINEAR_X */
} vo_basis;

/* TerminationReason of the levenberg-marquardt crate. */
enum {
    VO_TERM_USER = 0,
    VO_TERM_NUMERICAL = 1,
    VO_TERM_RESIDUALS_ZERO = 2,
    VO_TERM_ORTHOGONAL = 3,
    VO_TERM_CONVERGED_FTOL = 4,
    VO_TERM_CONVERGED_XTOL = 5,
    VO_TERM_CONVERGED_BOTH = 6,
    VO_TERM_NO_IMPROVEMENT_POSSIBLE = 7,
    VO_TERM_LOST_PATIENCE = 8,
    VO_TERM_NO_PARAMETERS = 9,
    VO_TERM_NO_RESIDUALS = 10,
    VO_TERM_WRONG_DIMENSIONS = 11
};

typedef struct {
    double ftol, xtol, gtol; /* <=0 selects the crate default 30*eps */
    double stepbound;        /* <=0 selects 100 */
    int patience;            /* <=0 selects 100 ; maxfev = patience*(q+1) */
    int scale_diag;          /* <0 selects true */
} vo_lm_opts;

typedef struct {
    int termination;
    int number_of_evaluations; /* residual evaluations (nfev) */
    int number_of_jacobians;
    double objective_function; /* 0.5*||r||^2 */
    int successful;
} vo_report;

typedef struct vo_problem vo_problem;

/* Mirrors SeparableProblemBuilder::build (src/problem/builder.rs:278-324):
 * validates, forms Y_w = W*Y, runs the first set_params at alpha0.
 * Y is column-major m x S (ld = m). w may be NULL (unit weights).
 * Returns NULL on invalid sizes. */
vo_problem *vo_problem_new(int m, int n, int q, const vo_basis *basis, const double *x,
                           int S, const double *Y, const double *w, double svd_eps,
                           const double *alpha0);
void vo_problem_free(vo_problem *p);
void vo_set_threads(int nthreads); /* OpenMP threads for the O(m*S) loops; 1 = reference behaviour */

/* src/solvers/levmar/mod.rs:42-73. Returns 0 if the cache is valid afterwards. */
int vo_set_params(vo_problem *p, const double *alpha);
void vo_params(const vo_problem *p, double *alpha_out);
/* :91-95 -> vec(R), length m*S, RHS-major. Returns 0 or -1 (cache is None). */
int vo_residuals(const vo_problem *p, double *out);
/* :101-201 -> (m*S) x q column-major. */
int vo_jacobian(const vo_problem *p, double *out);
/* n x S column-major. */
int vo_linear_coefficients(const vo_problem *p, double *out);
/* model.eval() (unweighted Phi, m x n col-major) and eval_partial_deriv(k). */
int vo_model_eval(const vo_problem *p, double *phi_out);
int vo_model_eval_partial_deriv(const vo_problem *p, int k, double *d_out);
/* FitResult::best_fit (src/fit.rs:55-59): Phi*C, m x S (unweighted Phi). */
int vo_best_fit(const vo_problem *p, double *out);
/* LevMarSolver::fit (src/solvers/levmar/mod.rs:238-254). */
int vo_fit(vo_problem *p, const vo_lm_opts *opts, vo_report *report);

/* FitStatistics::try_calculate (src/statistics/mod.rs:352-441), applied to RHS
 * column `s` with the shared alpha. cov: (n+q)^2 row/col symmetric, ordering
 * (c..., alpha...). conf_sigma: m values (unscaled confidence sigma) or NULL. */
int vo_statistics(const vo_problem *p, int s, double *cov, double *reduced_chi2,
                  double *weighted_residuals, double *conf_sigma);

#ifdef __cplusplus
}
#endif
#endif
