/*
 * varpro_oracle.c -- TEST INFRASTRUCTURE ONLY (see varpro_oracle.h).
 *
 * Literal CPU restatement of the reference algorithm: thin SVD of Phi_w,
 * C = V Sigma^+ U^T Y_w, materialised residual matrix R and Kaufman Jacobian J,
 * MINPACK lmder on the explicit (m*S) x q Jacobian. It deliberately keeps the
 * reference's cost structure (several O(m*S) sweeps per evaluation), because it
 * doubles as the timed CPU baseline ("port").
 *
 * Parity: pinned -- see tests/test_oracle_goldens.py for the golden vectors of
 * the reference's own tests that this file reproduces.
 */
#include "varpro_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct vo_problem {
    int m, n, q, S;
    vo_basis *basis;
    double *x;       /* m */
    double *w;       /* m or NULL (Weights::Unit, src/util/weights.rs:10-99) */
    double *Yw;      /* m x S, weighted once at build (src/problem/builder.rs:307) */
    double svd_eps;  /* absolute singular value threshold (builder.rs:246-251,282) */
    double *alpha;   /* q : model.params() */
    /* CachedCalculations (src/problem.rs:88-107) */
    int cached;
    double *R;       /* m x S current_residuals */
    double *U;       /* m x n */
    double *sigma;   /* n */
    double *V;       /* n x n (V, not V^T) */
    double *C;       /* n x S linear_coefficients */
    double *Phiw;    /* m x n scratch: weighted model matrix */
};

static int g_threads = 1;
void vo_set_threads(int nthreads) { g_threads = nthreads < 1 ? 1 : nthreads; }

/* ------------------------------------------------------------------------- */
/* model evaluation: src/model/mod.rs:441-512 (column j = basis function j;   */
/* d/dalpha_k is zero except in the columns whose function depends on k).     */
/* Formulas exactly as the reference writes them.      */
/* ------------------------------------------------------------------------- */
static void basis_eval(const vo_basis *b, const double *alpha, const double *x, int m, double *col)
{
    int i;
    switch (b->kind) {
    case VO_BASIS_EXP_DECAY: {
        double tau = alpha[b->param_idx[0]];
        for (i = 0; i < m; ++i) col[i] = exp(-x[i] / tau);
        break;
    }
    case VO_BASIS_CONSTANT:
        for (i = 0; i < m; ++i) col[i] = 1.0;
        break;
    case VO_BASIS_EXP_RATE_COS: {
        double a = alpha[b->param_idx[0]], c = alpha[b->param_idx[1]];
        for (i = 0; i < m; ++i) col[i] = exp(-a * x[i]) * cos(c * x[i]);
        break;
    }
    case VO_BASIS_SIN_PHASE: {
        double om = alpha[b->param_idx[0]], ph = alpha[b->param_idx[1]];
        for (i = 0; i < m; ++i) col[i] = sin(om * x[i] + ph);
        break;
    }
    case VO_BASIS_LINEAR_X:
        for (i = 0; i < m; ++i) col[i] = b->scale * x[i];
        break;
    default:
        for (i = 0; i < m; ++i) col[i] = NAN;
    }
}

/* derivative of basis function b with respect to its local parameter slot `slot` */
static void basis_deriv(const vo_basis *b, int slot, const double *alpha, const double *x, int m,
                        double *col)
{
    int i;
    switch (b->kind) {
    case VO_BASIS_EXP_DECAY: {
        /* shared_test_code/src/lib.rs:108-114: exp(-t/tau) * t / (tau*tau) */
        double tau = alpha[b->param_idx[0]];
        for (i = 0; i < m; ++i) col[i] = exp(-x[i] / tau) * x[i] / (tau * tau);
        break;
    }
    case VO_BASIS_EXP_RATE_COS: {
        /* shared_test_code/src/models.rs:362-385 */
        double a = alpha[b->param_idx[0]], c = alpha[b->param_idx[1]];
        if (slot == 0)
            for (i = 0; i < m; ++i) col[i] = -x[i] * (exp(-a * x[i]) * cos(c * x[i]));
        else
            for (i = 0; i < m; ++i) col[i] = -x[i] * exp(-a * x[i]) * sin(c * x[i]);
        break;
    }
    case VO_BASIS_SIN_PHASE: {
        /* src/test_helpers/mod.rs:36-51 */
        double om = alpha[b->param_idx[0]], ph = alpha[b->param_idx[1]];
        if (slot == 0)
            for (i = 0; i < m; ++i) col[i] = x[i] * cos(om * x[i] + ph);
        else
            for (i = 0; i < m; ++i) col[i] = cos(om * x[i] + ph);
        break;
    }
    default:
        for (i = 0; i < m; ++i) col[i] = 0.0;
    }
}

int vo_model_eval(const vo_problem *p, double *phi)
{
    for (int j = 0; j < p->n; ++j)
        basis_eval(&p->basis[j], p->alpha, p->x, p->m, phi + (size_t)j * p->m);
    return 0;
}

int vo_model_eval_partial_deriv(const vo_problem *p, int k, double *d)
{
    if (k < 0 || k >= p->q) return -1;
    memset(d, 0, sizeof(double) * (size_t)p->m * p->n);
    for (int j = 0; j < p->n; ++j) {
        const vo_basis *b = &p->basis[j];
        for (int s = 0; s < b->n_params; ++s)
            if (b->param_idx[s] == k)
                basis_deriv(b, s, p->alpha, p->x, p->m, d + (size_t)j * p->m);
    }
    return 0;
}

/* &Weights * M : row scaling (src/util/weights.rs:82-99, src/util/mod.rs:76-96) */
static void weight_rows(const double *w, int m, int ncols, double *M)
{
    if (!w) return;
    for (int j = 0; j < ncols; ++j)
        for (int i = 0; i < m; ++i) M[(size_t)j * m + i] *= w[i];
}

/* ------------------------------------------------------------------------- */
/* thin SVD A = U diag(sigma) V^T by one-sided Jacobi (Hestenes). Stands in   */
/* for nalgebra's `svd(true,true)` (src/solvers/levmar/mod.rs:51); the result */
/* is determined up to rounding and column signs.                             */
/* A (m x n, col-major) is overwritten by U.                                  */
/* ------------------------------------------------------------------------- */
static void svd_jacobi(int m, int n, double *A, double *sigma, double *V)
{
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) V[(size_t)j * n + i] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double *ap = A + (size_t)p * m, *aq = A + (size_t)q * m;
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < m; ++i) {
                    alpha += ap[i] * ap[i];
                    beta += aq[i] * aq[i];
                    gamma += ap[i] * aq[i];
                }
                if (gamma == 0.0 || fabs(gamma) <= DBL_EPSILON * sqrt(alpha * beta)) continue;
                rotated = 1;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < m; ++i) {
                    double tp = ap[i], tq = aq[i];
                    ap[i] = c * tp - s * tq;
                    aq[i] = s * tp + c * tq;
                }
                double *vp = V + (size_t)p * n, *vq = V + (size_t)q * n;
                for (int i = 0; i < n; ++i) {
                    double tp = vp[i], tq = vq[i];
                    vp[i] = c * tp - s * tq;
                    vq[i] = s * tp + c * tq;
                }
            }
        }
        if (!rotated) break;
    }
    for (int j = 0; j < n; ++j) {
        double *aj = A + (size_t)j * m, s2 = 0;
        for (int i = 0; i < m; ++i) s2 += aj[i] * aj[i];
        sigma[j] = sqrt(s2);
        if (sigma[j] > 0)
            for (int i = 0; i < m; ++i) aj[i] /= sigma[j];
    }
}

/* ------------------------------------------------------------------------- */
/* set_params: src/solvers/levmar/mod.rs:42-73                                */
/* ------------------------------------------------------------------------- */
int vo_set_params(vo_problem *p, const double *alpha)
{
    const int m = p->m, n = p->n, S = p->S;
    memcpy(p->alpha, alpha, sizeof(double) * p->q);
    p->cached = 0;
    /* Phi_w = W * model.eval()  (:47) */
    vo_model_eval(p, p->Phiw);
    weight_rows(p->w, m, n, p->Phiw);
    for (size_t i = 0; i < (size_t)m * n; ++i)
        if (!isfinite(p->Phiw[i])) return -1;
    /* svd of a clone of Phi_w (:51) */
    memcpy(p->U, p->Phiw, sizeof(double) * (size_t)m * n);
    svd_jacobi(m, n, p->U, p->sigma, p->V);
    if (p->svd_eps < 0) return -1; /* svd.solve errs for eps<0 */
    /* C = V diag(sigma_i > eps ? 1/sigma_i : 0) U^T Y_w  (:52-54) */
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int s = 0; s < S; ++s) {
        const double *y = p->Yw + (size_t)s * m;
        double uty[64];
        for (int j = 0; j < n; ++j) {
            const double *u = p->U + (size_t)j * m;
            double acc = 0;
            for (int i = 0; i < m; ++i) acc += u[i] * y[i];
            uty[j] = (p->sigma[j] > p->svd_eps) ? acc / p->sigma[j] : 0.0;
        }
        double *c = p->C + (size_t)s * n;
        for (int i = 0; i < n; ++i) {
            double acc = 0;
            for (int j = 0; j < n; ++j) acc += p->V[(size_t)j * n + i] * uty[j];
            c[i] = acc;
        }
        /* R = Y_w - Phi_w * C  (:57-59) */
        double *r = p->R + (size_t)s * m;
        for (int i = 0; i < m; ++i) r[i] = y[i];
        for (int j = 0; j < n; ++j) {
            const double *f = p->Phiw + (size_t)j * m;
            double cj = c[j];
            for (int i = 0; i < m; ++i) r[i] -= f[i] * cj;
        }
    }
    p->cached = 1;
    return 0;
}

void vo_params(const vo_problem *p, double *out) { memcpy(out, p->alpha, sizeof(double) * p->q); }

/* residuals: :91-95 ; to_vector = column stacking (src/util/mod.rs:101-106) */
int vo_residuals(const vo_problem *p, double *out)
{
    if (!p->cached) return -1;
    memcpy(out, p->R, sizeof(double) * (size_t)p->m * p->S);
    return 0;
}

int vo_linear_coefficients(const vo_problem *p, double *out)
{
    if (!p->cached) return -1;
    memcpy(out, p->C, sizeof(double) * (size_t)p->n * p->S);
    return 0;
}

/* jacobian: src/solvers/levmar/mod.rs:101-201 (Kaufman approximation).
 * J[:,k] = vec(U U^T D_k C - D_k C), two orderings selected by S <= q. */
int vo_jacobian(const vo_problem *p, double *J)
{
    if (!p->cached) return -1;
    const int m = p->m, n = p->n, S = p->S, q = p->q;
    double *Dk = (double *)malloc(sizeof(double) * (size_t)m * n);
    double *UtD = (double *)malloc(sizeof(double) * (size_t)n * n);
    if (!Dk || !UtD) { free(Dk); free(UtD); return -1; }
    for (int k = 0; k < q; ++k) {
        double *Jk = J + (size_t)k * m * S;
        /* Dk = W * eval_partial_deriv(k)  (:141) */
        vo_model_eval_partial_deriv(p, k, Dk);
        weight_rows(p->w, m, n, Dk);
        if (S <= q) {
            /* j_k = vec(U*(U^T*(Dk*C)) - Dk*C)   (:156-171) */
#pragma omp parallel for num_threads(g_threads) schedule(static)
            for (int s = 0; s < S; ++s) {
                double *t = Jk + (size_t)s * m;
                const double *c = p->C + (size_t)s * n;
                double ut[64];
                for (int i = 0; i < m; ++i) t[i] = 0;
                for (int j = 0; j < n; ++j) {
                    const double *d = Dk + (size_t)j * m;
                    for (int i = 0; i < m; ++i) t[i] += d[i] * c[j];
                }
                for (int j = 0; j < n; ++j) {
                    const double *u = p->U + (size_t)j * m;
                    double acc = 0;
                    for (int i = 0; i < m; ++i) acc += u[i] * t[i];
                    ut[j] = acc;
                }
                /* gemm(one, U, Ut_DkC, -one): t = U*ut - t */
                for (int i = 0; i < m; ++i) t[i] = -t[i];
                for (int j = 0; j < n; ++j) {
                    const double *u = p->U + (size_t)j * m;
                    for (int i = 0; i < m; ++i) t[i] += u[i] * ut[j];
                }
            }
        } else {
            /* Dk <- U*(U^T*Dk) - Dk ; j_k = vec(Dk*C)   (:172-186) */
            for (int c2 = 0; c2 < n; ++c2)
                for (int j = 0; j < n; ++j) {
                    const double *u = p->U + (size_t)j * m, *d = Dk + (size_t)c2 * m;
                    double acc = 0;
                    for (int i = 0; i < m; ++i) acc += u[i] * d[i];
                    UtD[(size_t)c2 * n + j] = acc;
                }
            for (int c2 = 0; c2 < n; ++c2) {
                double *d = Dk + (size_t)c2 * m;
                for (int i = 0; i < m; ++i) d[i] = -d[i];
                for (int j = 0; j < n; ++j) {
                    const double *u = p->U + (size_t)j * m;
                    double f = UtD[(size_t)c2 * n + j];
                    for (int i = 0; i < m; ++i) d[i] += u[i] * f;
                }
            }
#pragma omp parallel for num_threads(g_threads) schedule(static)
            for (int s = 0; s < S; ++s) {
                double *t = Jk + (size_t)s * m;
                const double *c = p->C + (size_t)s * n;
                for (int i = 0; i < m; ++i) t[i] = 0;
                for (int j = 0; j < n; ++j) {
                    const double *d = Dk + (size_t)j * m;
                    double cj = c[j];
                    for (int i = 0; i < m; ++i) t[i] += d[i] * cj;
                }
            }
        }
    }
    free(Dk);
    free(UtD);
    return 0;
}

/* FitResult::best_fit: model.eval() * C with the UNWEIGHTED Phi (src/fit.rs:55-59) */
int vo_best_fit(const vo_problem *p, double *out)
{
    if (!p->cached) return -1;
    const int m = p->m, n = p->n, S = p->S;
    double *phi = (double *)malloc(sizeof(double) * (size_t)m * n);
    if (!phi) return -1;
    vo_model_eval(p, phi);
    for (int s = 0; s < S; ++s) {
        double *o = out + (size_t)s * m;
        const double *c = p->C + (size_t)s * n;
        for (int i = 0; i < m; ++i) o[i] = 0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) o[i] += phi[(size_t)j * m + i] * c[j];
    }
    free(phi);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* problem construction: src/problem/builder.rs:278-324                       */
/* ------------------------------------------------------------------------- */
vo_problem *vo_problem_new(int m, int n, int q, const vo_basis *basis, const double *x, int S,
                           const double *Y, const double *w, double svd_eps, const double *alpha0)
{
    if (m <= 0 || S <= 0 || n <= 0 || n > 64 || q < 0) return NULL; /* ZeroLengthVector etc. */
    vo_problem *p = (vo_problem *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->m = m; p->n = n; p->q = q; p->S = S;
    p->svd_eps = fabs(svd_eps); /* builder.rs:248 */
    p->basis = (vo_basis *)malloc(sizeof(vo_basis) * n);
    memcpy(p->basis, basis, sizeof(vo_basis) * n);
    p->x = (double *)malloc(sizeof(double) * m);
    memcpy(p->x, x, sizeof(double) * m);
    if (w) {
        p->w = (double *)malloc(sizeof(double) * m);
        memcpy(p->w, w, sizeof(double) * m);
    }
    p->Yw = (double *)malloc(sizeof(double) * (size_t)m * S);
    memcpy(p->Yw, Y, sizeof(double) * (size_t)m * S);
    weight_rows(p->w, m, S, p->Yw); /* Y_w = &weights * Y (builder.rs:307) */
    p->alpha = (double *)calloc(q > 0 ? q : 1, sizeof(double));
    p->R = (double *)malloc(sizeof(double) * (size_t)m * S);
    p->U = (double *)malloc(sizeof(double) * (size_t)m * n);
    p->Phiw = (double *)malloc(sizeof(double) * (size_t)m * n);
    p->sigma = (double *)malloc(sizeof(double) * n);
    p->V = (double *)malloc(sizeof(double) * (size_t)n * n);
    p->C = (double *)malloc(sizeof(double) * (size_t)n * S);
    vo_set_params(p, alpha0); /* builder.rs:321 */
    return p;
}

void vo_problem_free(vo_problem *p)
{
    if (!p) return;
    free(p->basis); free(p->x); free(p->w); free(p->Yw); free(p->alpha);
    free(p->R); free(p->U); free(p->Phiw); free(p->sigma); free(p->V); free(p->C);
    free(p);
}

/* ------------------------------------------------------------------------- */
/* MINPACK lmder (levenberg-marquardt 0.14 is documented as a port of it;     */
/* SURVEY.md Appendix A). Jacobian a is column-major M x n, M = m*S.          */
/* ------------------------------------------------------------------------- */
static double enorm(size_t n, const double *x)
{
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}

/* Column-pivoted Householder QR (MINPACK qrfac). */
static void qrfac(size_t M, int n, double *a, int *ipvt, double *rdiag, double *acnorm, double *wa)
{
    const double epsmch = DBL_EPSILON;
    for (int j = 0; j < n; ++j) {
        acnorm[j] = enorm(M, a + (size_t)j * M);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
    int minmn = (M < (size_t)n) ? (int)M : n;
    for (int j = 0; j < minmn; ++j) {
        int kmax = j;
        for (int k = j; k < n; ++k)
            if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            double *cj = a + (size_t)j * M, *ck = a + (size_t)kmax * M;
            for (size_t i = 0; i < M; ++i) { double t = cj[i]; cj[i] = ck[i]; ck[i] = t; }
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            int t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
        }
        double *aj = a + (size_t)j * M;
        double ajnorm = enorm(M - j, aj + j);
        if (ajnorm != 0.0) {
            if (aj[j] < 0.0) ajnorm = -ajnorm;
            for (size_t i = j; i < M; ++i) aj[i] /= ajnorm;
            aj[j] += 1.0;
            for (int k = j + 1; k < n; ++k) {
                double *ak = a + (size_t)k * M;
                double sum = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : sum) schedule(static) if (M > 65536)
                for (long long i = j; i < (long long)M; ++i) sum += aj[i] * ak[i];
                double temp = sum / aj[j];
#pragma omp parallel for num_threads(g_threads) schedule(static) if (M > 65536)
                for (long long i = j; i < (long long)M; ++i) ak[i] -= temp * aj[i];
                if (rdiag[k] != 0.0) {
                    temp = ak[j] / rdiag[k];
                    double d = 1.0 - temp * temp;
                    rdiag[k] *= sqrt(d > 0 ? d : 0);
                    double r = rdiag[k] / wa[k];
                    if (0.05 * r * r <= epsmch) {
                        rdiag[k] = enorm(M - j - 1, ak + j + 1);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

/* MINPACK qrsolv on the n x n upper triangle r (column-major, ldr). */
static void qrsolv(int n, double *r, int ldr, const int *ipvt, const double *diag,
                   const double *qtb, double *x, double *sdiag, double *wa)
{
#define RR(i, j) r[(size_t)(j) * ldr + (i)]
    for (int j = 0; j < n; ++j) {
        for (int i = j; i < n; ++i) RR(i, j) = RR(j, i);
        x[j] = RR(j, j);
        wa[j] = qtb[j];
    }
    for (int j = 0; j < n; ++j) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < n; ++k) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < n; ++k) {
                if (sdiag[k] == 0.0) continue;
                double c, s;
                if (fabs(RR(k, k)) < fabs(sdiag[k])) {
                    double cotan = RR(k, k) / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
                    c = s * cotan;
                } else {
                    double t = sdiag[k] / RR(k, k);
                    c = 0.5 / sqrt(0.25 + 0.25 * t * t);
                    s = c * t;
                }
                RR(k, k) = c * RR(k, k) + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < n; ++i) {
                    temp = c * RR(i, k) + s * sdiag[i];
                    sdiag[i] = -s * RR(i, k) + c * sdiag[i];
                    RR(i, k) = temp;
                }
            }
        }
        sdiag[j] = RR(j, j);
        RR(j, j) = x[j];
    }
    int nsing = n;
    for (int j = 0; j < n; ++j) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        int j = nsing - k;
        double sum = 0;
        for (int i = j + 1; i < nsing; ++i) sum += RR(i, j) * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
#undef RR
}

/* MINPACK lmpar. */
static void lmpar(int n, double *r, int ldr, const int *ipvt, const double *diag,
                  const double *qtb, double delta, double *par, double *x, double *sdiag,
                  double *wa1, double *wa2)
{
#define RR(i, j) r[(size_t)(j) * ldr + (i)]
    const double dwarf = DBL_MIN;
    int nsing = n;
    for (int j = 0; j < n; ++j) {
        wa1[j] = qtb[j];
        if (RR(j, j) == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        int j = nsing - k;
        wa1[j] /= RR(j, j);
        double temp = wa1[j];
        for (int i = 0; i < j; ++i) wa1[i] -= RR(i, j) * temp;
    }
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *par = 0.0; return; }
    double parl = 0.0;
    if (nsing >= n) {
        for (int j = 0; j < n; ++j) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; ++j) {
            double sum = 0;
            for (int i = 0; i < j; ++i) sum += RR(i, j) * wa1[i];
            wa1[j] = (wa1[j] - sum) / RR(j, j);
        }
        double temp = enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < n; ++j) {
        double sum = 0;
        for (int i = 0; i <= j; ++i) sum += RR(i, j) * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
    *par = fmax(*par, parl);
    *par = fmin(*par, paru);
    if (*par == 0.0) *par = gnorm / dxnorm;
    for (;;) {
        ++iter;
        if (*par == 0.0) *par = fmax(dwarf, 0.001 * paru);
        double temp = sqrt(*par);
        for (int j = 0; j < n; ++j) wa1[j] = temp * diag[j];
        qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
        for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10)
            break;
        for (int j = 0; j < n; ++j) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; ++j) {
            wa1[j] /= sdiag[j];
            temp = wa1[j];
            for (int i = j + 1; i < n; ++i) wa1[i] -= RR(i, j) * temp;
        }
        temp = enorm(n, wa1);
        double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, *par);
        if (fp < 0.0) paru = fmin(paru, *par);
        *par = fmax(parl, *par + parc);
    }
#undef RR
}

/* LevMarSolver::fit -> LevenbergMarquardt::minimize (src/solvers/levmar/mod.rs:238-254) */
int vo_fit(vo_problem *p, const vo_lm_opts *o, vo_report *rep)
{
    const int n = p->q; /* MINPACK's n = number of nonlinear parameters */
    const size_t M = (size_t)p->m * p->S;
    const double epsmch = DBL_EPSILON;
    double ftol = (o && o->ftol > 0) ? o->ftol : 30.0 * epsmch;
    double xtol = (o && o->xtol > 0) ? o->xtol : 30.0 * epsmch;
    double gtol = (o && o->gtol > 0) ? o->gtol : 30.0 * epsmch;
    double factor = (o && o->stepbound > 0) ? o->stepbound : 100.0;
    int patience = (o && o->patience > 0) ? o->patience : 100;
    int scale_diag = (o && o->scale_diag >= 0) ? o->scale_diag : 1;
    int maxfev = patience * (n + 1);

    memset(rep, 0, sizeof(*rep));
    if (n == 0) { rep->termination = VO_TERM_NO_PARAMETERS; return 0; }
    if (M == 0) { rep->termination = VO_TERM_NO_RESIDUALS; return 0; }

    double *x = (double *)malloc(sizeof(double) * n);
    double *fvec = (double *)malloc(sizeof(double) * M);
    double *wa4 = (double *)malloc(sizeof(double) * M);
    double *fjac = (double *)malloc(sizeof(double) * M * n);
    double diag[64], qtf[64], wa1[64], wa2[64], wa3[64], rdiag[64], acnorm[64], sdiag[64];
    double rsmall[64 * 64];
    int ipvt[64];
    int term = -1, nfev = 0, njev = 0, last_rejected = 0;
    double fnorm = 0, par = 0, delta = 0, xnorm = 0, gnorm = 0;
    int iter = 1;

    vo_params(p, x);
    /* evaluate the function at the starting point (the builder already did
     * set_params(alpha0); lmder re-evaluates, the crate calls residuals()). */
    if (!p->cached || vo_residuals(p, fvec) != 0) { term = VO_TERM_USER; goto done; }
    nfev = 1;
    fnorm = enorm(M, fvec);
    if (!isfinite(fnorm)) { term = VO_TERM_NUMERICAL; goto done; }
    if (fnorm == 0.0) { term = VO_TERM_RESIDUALS_ZERO; goto done; }

    for (;;) {
        if (vo_jacobian(p, fjac) != 0) { term = VO_TERM_USER; goto done; }
        ++njev;
        qrfac(M, n, fjac, ipvt, rdiag, acnorm, wa3);
        if (iter == 1) {
            for (int j = 0; j < n; ++j) {
                diag[j] = scale_diag ? acnorm[j] : 1.0;
                if (scale_diag && acnorm[j] == 0.0) diag[j] = 1.0;
            }
            for (int j = 0; j < n; ++j) wa3[j] = diag[j] * x[j];
            xnorm = enorm(n, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        /* qtf = first n components of Q^T fvec */
        memcpy(wa4, fvec, sizeof(double) * M);
        for (int j = 0; j < n; ++j) {
            double *aj = fjac + (size_t)j * M;
            if (aj[j] != 0.0) {
                double sum = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : sum) schedule(static) if (M > 65536)
                for (long long i = j; i < (long long)M; ++i) sum += aj[i] * wa4[i];
                double temp = -sum / aj[j];
#pragma omp parallel for num_threads(g_threads) schedule(static) if (M > 65536)
                for (long long i = j; i < (long long)M; ++i) wa4[i] += aj[i] * temp;
            }
            aj[j] = rdiag[j];
            qtf[j] = wa4[j];
        }
        /* small copy of R for lmpar/qrsolv */
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i)
                rsmall[(size_t)j * n + i] = (i <= j) ? fjac[(size_t)j * M + i] : 0.0;
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < n; ++j) {
                int l = ipvt[j];
                if (acnorm[l] != 0.0) {
                    double sum = 0;
                    for (int i = 0; i <= j; ++i) sum += rsmall[(size_t)j * n + i] * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / acnorm[l]));
                }
            }
        }
        if (!isfinite(gnorm)) { term = VO_TERM_NUMERICAL; goto done; }
        if (gnorm <= gtol) { term = VO_TERM_ORTHOGONAL; goto done; }
        if (scale_diag)
            for (int j = 0; j < n; ++j) diag[j] = fmax(diag[j], acnorm[j]);

        for (;;) {
            lmpar(n, rsmall, n, ipvt, diag, qtf, delta, &par, wa1, sdiag, wa2, wa3);
            for (int j = 0; j < n; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            double pnorm = enorm(n, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            vo_set_params(p, wa2);
            if (vo_residuals(p, wa4) != 0) { term = VO_TERM_USER; goto done; }
            ++nfev;
            double fnorm1 = enorm(M, wa4);
            if (!isfinite(fnorm1)) { term = VO_TERM_NUMERICAL; goto done; }
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            for (int j = 0; j < n; ++j) {
                wa3[j] = 0.0;
                double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; ++i) wa3[i] += rsmall[(size_t)j * n + i] * temp;
            }
            double temp1 = enorm(n, wa3) / fnorm;
            double temp2 = (sqrt(par) * pnorm) / fnorm;
            double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            double dirder = -(temp1 * temp1 + temp2 * temp2);
            double ratio = (prered != 0.0) ? actred / prered : 0.0;
            if (ratio <= 0.25) {
                double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < n; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                memcpy(fvec, wa4, sizeof(double) * M);
                xnorm = enorm(n, wa2);
                fnorm = fnorm1;
                ++iter;
                last_rejected = 0;
            } else {
                last_rejected = 1;
            }
            if (fnorm == 0.0) { term = VO_TERM_RESIDUALS_ZERO; goto done; }
            int f_ok = (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0);
            int x_ok = (delta <= xtol * xnorm);
            if (f_ok && x_ok) { term = VO_TERM_CONVERGED_BOTH; goto done; }
            if (f_ok) { term = VO_TERM_CONVERGED_FTOL; goto done; }
            if (x_ok) { term = VO_TERM_CONVERGED_XTOL; goto done; }
            if (nfev >= maxfev) { term = VO_TERM_LOST_PATIENCE; goto done; }
            if ((fabs(actred) <= epsmch && prered <= epsmch && 0.5 * ratio <= 1.0) ||
                delta <= epsmch * xnorm || gnorm <= epsmch) {
                term = VO_TERM_NO_IMPROVEMENT_POSSIBLE;
                goto done;
            }
            if (ratio >= 1e-4) break;
        }
    }
done:
    /* leave the problem consistent with the last accepted x (src/fit.rs:113,45-47
     * read params and coefficients from the returned problem) */
    if (last_rejected) vo_set_params(p, x);
    rep->termination = term;
    rep->number_of_evaluations = nfev;
    rep->number_of_jacobians = njev;
    rep->objective_function = 0.5 * fnorm * fnorm;
    rep->successful = (term == VO_TERM_RESIDUALS_ZERO || term == VO_TERM_ORTHOGONAL ||
                       term == VO_TERM_CONVERGED_FTOL || term == VO_TERM_CONVERGED_XTOL ||
                       term == VO_TERM_CONVERGED_BOTH);
    free(x); free(fvec); free(wa4); free(fjac);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* FitStatistics::try_calculate: src/statistics/mod.rs:352-441 for column s.  */
/* ------------------------------------------------------------------------- */
static int invert_spd_general(int n, double *A, double *Ainv)
{
    /* Gauss-Jordan with partial pivoting (stands in for nalgebra try_inverse,
     * src/statistics/mod.rs:397-399). A is destroyed. */
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) Ainv[(size_t)j * n + i] = (i == j);
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int i = c + 1; i < n; ++i)
            if (fabs(A[(size_t)c * n + i]) > fabs(A[(size_t)c * n + piv])) piv = i;
        if (A[(size_t)c * n + piv] == 0.0) return -1;
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                double t = A[(size_t)j * n + c]; A[(size_t)j * n + c] = A[(size_t)j * n + piv]; A[(size_t)j * n + piv] = t;
                t = Ainv[(size_t)j * n + c]; Ainv[(size_t)j * n + c] = Ainv[(size_t)j * n + piv]; Ainv[(size_t)j * n + piv] = t;
            }
        double d = 1.0 / A[(size_t)c * n + c];
        for (int j = 0; j < n; ++j) { A[(size_t)j * n + c] *= d; Ainv[(size_t)j * n + c] *= d; }
        for (int i = 0; i < n; ++i) {
            if (i == c) continue;
            double f = A[(size_t)c * n + i];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) {
                A[(size_t)j * n + i] -= f * A[(size_t)j * n + c];
                Ainv[(size_t)j * n + i] -= f * Ainv[(size_t)j * n + c];
            }
        }
    }
    return 0;
}

int vo_statistics(const vo_problem *p, int s, double *cov, double *reduced_chi2,
                  double *weighted_residuals, double *conf_sigma)
{
    if (!p->cached || s < 0 || s >= p->S) return -1;
    const int m = p->m, n = p->n, q = p->q, t = n + q;
    if (m <= t) return -2; /* Error::Underdetermined (:377-379) */
    const double *c = p->C + (size_t)s * n;
    double *J = (double *)calloc((size_t)m * t, sizeof(double));
    double *H = (double *)malloc(sizeof(double) * (size_t)m * t);
    double *Dk = (double *)malloc(sizeof(double) * (size_t)m * n);
    double *HtH = (double *)malloc(sizeof(double) * t * t);
    /* J = [Phi, (dPhi/dalpha_k) c]  (model_function_jacobian :486-511) */
    vo_model_eval(p, J);
    for (int k = 0; k < q; ++k) {
        vo_model_eval_partial_deriv(p, k, Dk);
        double *col = J + (size_t)(n + k) * m;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) col[i] += Dk[(size_t)j * m + i] * c[j];
    }
    memcpy(H, J, sizeof(double) * (size_t)m * t);
    weight_rows(p->w, m, t, H); /* H = W J (:373) */
    /* weighted residuals = y_w - W Phi c (:374) */
    double rn2 = 0;
    for (int i = 0; i < m; ++i) {
        double r = p->Yw[(size_t)s * m + i];
        for (int j = 0; j < n; ++j) r -= H[(size_t)j * m + i] * c[j];
        if (weighted_residuals) weighted_residuals[i] = r;
        rn2 += r * r;
    }
    double chi2 = rn2 / (double)(m - t);
    if (reduced_chi2) *reduced_chi2 = chi2;
    for (int a = 0; a < t; ++a)
        for (int b = 0; b < t; ++b) {
            double acc = 0;
            for (int i = 0; i < m; ++i) acc += H[(size_t)a * m + i] * H[(size_t)b * m + i];
            HtH[(size_t)b * t + a] = acc;
        }
    int rc = invert_spd_general(t, HtH, cov);
    if (rc == 0) {
        for (int i = 0; i < t * t; ++i) cov[i] *= chi2; /* sigma^2 (HtH)^-1 (:397-400) */
        if (conf_sigma) {
            /* sqrt(j^T Cov j) with j = rows of the UNWEIGHTED J (:415-430) */
            for (int i = 0; i < m; ++i) {
                double acc = 0;
                for (int a = 0; a < t; ++a) {
                    double inner = 0;
                    for (int b = 0; b < t; ++b) inner += cov[(size_t)b * t + a] * J[(size_t)b * m + i];
                    acc += J[(size_t)a * m + i] * inner;
                }
                conf_sigma[i] = sqrt(acc);
            }
        }
    }
    free(J); free(H); free(Dk); free(HtH);
    return rc == 0 ? 0 : -3; /* Error::MatrixInversion */
}
