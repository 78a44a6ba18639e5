// lm_step.cuh -- MINPACK-lmder trust-region update restated on the q x q system.
//
// Replaces the external `levenberg-marquardt` 0.14 loop the reference calls at
// src/solvers/levmar/mod.rs:247 (LevenbergMarquardt::minimize; defaults at
// :307-315). The reference hands that loop the (m*S) x q Jacobian J and the
// m*S residual vector r; lmder only ever uses the R factor of the pivoted QR
// of J, the first q entries of Q^T r, the column norms of J and ||r||. All of
// these are functions of H = J^T J, g = J^T r and ||r|| (SURVEY.md Appendix A,
// "small-system equivalence"):
//   P^T H P = R^T R  (greedy diagonal pivoting == qrfac's column-norm pivoting)
//   qtf = R^-T P^T g,  acnorm_j = sqrt(H_jj),  ||J d||^2 = d^T H d.
// The streaming kernel produces (||r||^2, g, H) in one pass over Y, so the LM
// step never touches an O(m*S) object.
//
// The code is __host__ __device__: the host-driven vp_fit loop and the
// device-resident LM step of the graph path share it.
#pragma once

#include <math.h>
#include <float.h>

#ifndef VP_HD
#ifdef __CUDACC__
#define VP_HD __host__ __device__ __forceinline__
#else
#define VP_HD inline
#endif
#endif

#define VP_LM_MAXQ 8

namespace vp {

enum Termination : int {
    TERM_RUNNING = -1,
    TERM_USER = 0,
    TERM_NUMERICAL = 1,
    TERM_RESIDUALS_ZERO = 2,
    TERM_ORTHOGONAL = 3,
    TERM_CONVERGED_FTOL = 4,
    TERM_CONVERGED_XTOL = 5,
    TERM_CONVERGED_FTOL_XTOL = 6,
    TERM_NO_IMPROVEMENT_POSSIBLE = 7,
    TERM_LOST_PATIENCE = 8,
    TERM_NO_PARAMETERS = 9,
    TERM_NO_RESIDUALS = 10,
    TERM_WRONG_DIMENSIONS = 11
};

struct LmConfig {
    double ftol, xtol, gtol, stepbound, epsmch;
    int maxfev;
    int scale_diag;
};

// One evaluation of the projected functional at a parameter vector.
struct LmEval {
    double rnorm2;                     // ||r||^2
    double g[VP_LM_MAXQ];              // J^T r
    double H[VP_LM_MAXQ * VP_LM_MAXQ]; // J^T J, column-major q x q
    int finite;
};

struct LmState {
    int q;
    int phase;        // 0: waiting for the evaluation at x0, 1: waiting for a trial evaluation
    int termination;  // Termination
    int nfev, iter;
    int last_accepted; // 1 if the most recent trial was accepted (trial buffers hold the state of x)
    double x[VP_LM_MAXQ];       // last accepted parameters
    double x_trial[VP_LM_MAXQ]; // parameters the next evaluation must be made at
    double step[VP_LM_MAXQ];    // x_trial - x
    double fnorm, xnorm, gnorm, delta, par, pnorm;
    double diag[VP_LM_MAXQ];
    double R[VP_LM_MAXQ * VP_LM_MAXQ]; // upper triangle of the pivoted Cholesky factor, col-major
    double qtf[VP_LM_MAXQ];
    double acnorm[VP_LM_MAXQ];
    int ipvt[VP_LM_MAXQ];
};

VP_HD double lm_enorm(int n, const double *v)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i] * v[i];
    return sqrt(s);
}

// Pivoted Cholesky of H standing in for MINPACK qrfac(J): R upper triangular,
// P^T H P = R^T R, pivot = largest remaining diagonal of the Schur complement
// (= largest remaining column norm of J, qrfac's rule). Rank deficiency gives
// zero rows, which lmpar treats like qrfac's zero diagonals.
VP_HD void lm_pivoted_cholesky(int n, const double *H, double *R, int *ipvt, double *acnorm)
{
    double A[VP_LM_MAXQ * VP_LM_MAXQ];
    for (int j = 0; j < n; ++j) {
        ipvt[j] = j;
        acnorm[j] = sqrt(fmax(H[j * n + j], 0.0));
        for (int i = 0; i < n; ++i) {
            A[j * n + i] = H[j * n + i];
            R[j * n + i] = 0.0;
        }
    }
    for (int j = 0; j < n; ++j) {
        int kmax = j;
        for (int k = j + 1; k < n; ++k)
            if (A[k * n + k] > A[kmax * n + kmax]) kmax = k;
        if (kmax != j) {
            // symmetric swap of rows/columns j and kmax of the working matrix,
            // and of the already computed columns of R
            for (int i = 0; i < n; ++i) { double t = A[j * n + i]; A[j * n + i] = A[kmax * n + i]; A[kmax * n + i] = t; }
            for (int i = 0; i < n; ++i) { double t = A[i * n + j]; A[i * n + j] = A[i * n + kmax]; A[i * n + kmax] = t; }
            for (int i = 0; i < j; ++i) { double t = R[j * n + i]; R[j * n + i] = R[kmax * n + i]; R[kmax * n + i] = t; }
            int t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
        }
        double d = A[j * n + j];
        if (!(d > 0.0)) {
            // remaining block is (numerically) zero: rank deficient
            for (int k = j; k < n; ++k) R[k * n + j] = 0.0;
            for (int k = j + 1; k < n; ++k) A[k * n + k] = 0.0;
            continue;
        }
        double rjj = sqrt(d);
        R[j * n + j] = rjj;
        for (int k = j + 1; k < n; ++k) R[k * n + j] = A[k * n + j] / rjj;
        for (int k = j + 1; k < n; ++k)
            for (int i = j + 1; i <= k; ++i) {
                A[k * n + i] -= R[i * n + j] * R[k * n + j];
                A[i * n + k] = A[k * n + i];
            }
    }
}

// MINPACK qrsolv on a private copy S of R (R itself is left untouched).
VP_HD void lm_qrsolv(int n, const double *R, const int *ipvt, const double *diag,
                     const double *qtb, double *x, double *sdiag, double *S)
{
    double wa[VP_LM_MAXQ];
    // S holds r^T in its strict lower triangle during the sweep (MINPACK stores
    // it inside r); S(i,j) i>j.
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) S[j * n + i] = R[j * n + i];
    }
    for (int j = 0; j < n; ++j) {
        for (int i = j; i < n; ++i) S[j * n + i] = S[i * n + j];
        x[j] = S[j * n + j];
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
                double rkk = S[k * n + k];
                if (fabs(rkk) < fabs(sdiag[k])) {
                    double cotan = rkk / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
                    c = s * cotan;
                } else {
                    double t = sdiag[k] / rkk;
                    c = 0.5 / sqrt(0.25 + 0.25 * t * t);
                    s = c * t;
                }
                S[k * n + k] = c * rkk + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < n; ++i) {
                    temp = c * S[k * n + i] + s * sdiag[i];
                    sdiag[i] = -s * S[k * n + i] + c * sdiag[i];
                    S[k * n + i] = temp;
                }
            }
        }
        sdiag[j] = S[j * n + j];
        S[j * n + j] = x[j];
    }
    int nsing = n;
    for (int j = 0; j < n; ++j) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; ++i) sum += S[j * n + i] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
}

// MINPACK lmpar: determines par such that ||D x|| ~ delta. x receives the
// solution of (J^T J + par D^2) x = J^T r.
VP_HD void lm_lmpar(int n, const double *R, const int *ipvt, const double *diag,
                    const double *qtb, double delta, double *par, double *x)
{
    const double dwarf = DBL_MIN;
    double wa1[VP_LM_MAXQ], wa2[VP_LM_MAXQ], sdiag[VP_LM_MAXQ];
    double S[VP_LM_MAXQ * VP_LM_MAXQ];
    int nsing = n;
    for (int j = 0; j < n; ++j) {
        wa1[j] = qtb[j];
        if (R[j * n + j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        int j = nsing - k;
        wa1[j] /= R[j * n + j];
        double temp = wa1[j];
        for (int i = 0; i < j; ++i) wa1[i] -= R[j * n + i] * temp;
    }
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = lm_enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *par = 0.0; return; }
    double parl = 0.0;
    if (nsing >= n) {
        for (int j = 0; j < n; ++j) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            for (int i = 0; i < j; ++i) sum += R[j * n + i] * wa1[i];
            wa1[j] = (wa1[j] - sum) / R[j * n + j];
        }
        double temp = lm_enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < n; ++j) {
        double sum = 0.0;
        for (int i = 0; i <= j; ++i) sum += R[j * n + i] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = lm_enorm(n, wa1);
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
        lm_qrsolv(n, R, ipvt, wa1, qtb, x, sdiag, S);
        for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = lm_enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < n; ++j) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; ++j) {
            wa1[j] /= sdiag[j];
            temp = wa1[j];
            for (int i = j + 1; i < n; ++i) wa1[i] -= S[j * n + i] * temp;
        }
        temp = lm_enorm(n, wa1);
        double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, *par);
        if (fp < 0.0) paru = fmin(paru, *par);
        *par = fmax(parl, *par + parc);
    }
}

VP_HD void lm_init(LmState &st, int q, const double *x0)
{
    st.q = q;
    st.phase = 0;
    st.termination = TERM_RUNNING;
    st.nfev = 0;
    st.iter = 1;
    st.last_accepted = 1;
    st.fnorm = st.xnorm = st.gnorm = st.delta = st.par = st.pnorm = 0.0;
    for (int j = 0; j < VP_LM_MAXQ; ++j) {
        st.x[j] = j < q ? x0[j] : 0.0;
        st.x_trial[j] = st.x[j];
        st.step[j] = 0.0;
        st.diag[j] = 1.0;
    }
}

// Start of an lmder outer iteration: factor H at the accepted point, form qtf,
// the scaled gradient norm and the diag rescale. Returns false on termination.
VP_HD bool lm_outer_prepare(LmState &st, const LmConfig &cfg, const LmEval &ev)
{
    const int n = st.q;
    lm_pivoted_cholesky(n, ev.H, st.R, st.ipvt, st.acnorm);
    if (st.iter == 1) {
        for (int j = 0; j < n; ++j) {
            st.diag[j] = cfg.scale_diag ? st.acnorm[j] : 1.0;
            if (cfg.scale_diag && st.acnorm[j] == 0.0) st.diag[j] = 1.0;
        }
        double wa3[VP_LM_MAXQ];
        for (int j = 0; j < n; ++j) wa3[j] = st.diag[j] * st.x[j];
        st.xnorm = lm_enorm(n, wa3);
        st.delta = cfg.stepbound * st.xnorm;
        if (st.delta == 0.0) st.delta = cfg.stepbound;
    }
    // qtf = R^-T P^T g (forward substitution; zero for the rank-deficient tail)
    for (int j = 0; j < n; ++j) {
        double sum = ev.g[st.ipvt[j]];
        for (int i = 0; i < j; ++i) sum -= st.R[j * n + i] * st.qtf[i];
        st.qtf[j] = (st.R[j * n + j] != 0.0) ? sum / st.R[j * n + j] : 0.0;
    }
    st.gnorm = 0.0;
    if (st.fnorm != 0.0) {
        for (int j = 0; j < n; ++j) {
            int l = st.ipvt[j];
            if (st.acnorm[l] != 0.0) {
                double sum = 0.0;
                for (int i = 0; i <= j; ++i) sum += st.R[j * n + i] * (st.qtf[i] / st.fnorm);
                st.gnorm = fmax(st.gnorm, fabs(sum / st.acnorm[l]));
            }
        }
    }
    if (!isfinite(st.gnorm)) { st.termination = TERM_NUMERICAL; return false; }
    if (st.gnorm <= cfg.gtol) { st.termination = TERM_ORTHOGONAL; return false; }
    if (cfg.scale_diag)
        for (int j = 0; j < n; ++j) st.diag[j] = fmax(st.diag[j], st.acnorm[j]);
    return true;
}

// lmder inner loop head: lmpar -> trial point.
VP_HD void lm_inner_propose(LmState &st)
{
    const int n = st.q;
    double p[VP_LM_MAXQ], wa3[VP_LM_MAXQ];
    lm_lmpar(n, st.R, st.ipvt, st.diag, st.qtf, st.delta, &st.par, p);
    for (int j = 0; j < n; ++j) {
        st.step[j] = -p[j];
        st.x_trial[j] = st.x[j] + st.step[j];
        wa3[j] = st.diag[j] * st.step[j];
    }
    st.pnorm = lm_enorm(n, wa3);
    if (st.iter == 1) st.delta = fmin(st.delta, st.pnorm);
}

// Feed the evaluation made at st.x_trial. Returns true if another evaluation
// (at the new st.x_trial) is required, false when the fit has terminated.
// `H_acc`/`g_acc` semantics: when a trial is accepted the caller's evaluation
// becomes the accepted one (st.last_accepted = 1) and its H, g seed the next
// outer iteration -- the Jacobian products come out of the same streaming pass
// as the residual norm, so no second pass is needed.
VP_HD bool lm_advance(LmState &st, const LmConfig &cfg, const LmEval &ev)
{
    const int n = st.q;
    if (st.termination != TERM_RUNNING) return false;
    if (st.phase == 0) {
        st.nfev = 1;
        st.last_accepted = 1;
        if (n == 0) { st.termination = TERM_NO_PARAMETERS; return false; }
        double fn = sqrt(ev.rnorm2);
        if (!ev.finite || !isfinite(fn)) { st.fnorm = fn; st.termination = TERM_NUMERICAL; return false; }
        st.fnorm = fn;
        if (fn == 0.0) { st.termination = TERM_RESIDUALS_ZERO; return false; }
        st.par = 0.0;
        st.iter = 1;
        if (!lm_outer_prepare(st, cfg, ev)) return false;
        lm_inner_propose(st);
        st.phase = 1;
        return true;
    }
    // trial evaluation
    st.nfev += 1;
    double fnorm1 = sqrt(ev.rnorm2);
    if (!ev.finite || !isfinite(fnorm1)) { st.last_accepted = 0; st.termination = TERM_NUMERICAL; return false; }
    double actred = -1.0;
    if (0.1 * fnorm1 < st.fnorm) actred = 1.0 - (fnorm1 / st.fnorm) * (fnorm1 / st.fnorm);
    double wa3[VP_LM_MAXQ];
    for (int j = 0; j < n; ++j) wa3[j] = 0.0;
    for (int j = 0; j < n; ++j) {
        double temp = st.step[st.ipvt[j]];
        for (int i = 0; i <= j; ++i) wa3[i] += st.R[j * n + i] * temp;
    }
    double temp1 = lm_enorm(n, wa3) / st.fnorm;
    double temp2 = (sqrt(st.par) * st.pnorm) / st.fnorm;
    double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
    double dirder = -(temp1 * temp1 + temp2 * temp2);
    double ratio = (prered != 0.0) ? actred / prered : 0.0;
    if (ratio <= 0.25) {
        double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= st.fnorm || temp < 0.1) temp = 0.1;
        st.delta = temp * fmin(st.delta, st.pnorm / 0.1);
        st.par /= temp;
    } else if (st.par == 0.0 || ratio >= 0.75) {
        st.delta = st.pnorm / 0.5;
        st.par *= 0.5;
    }
    const bool accepted = ratio >= 1e-4;
    if (accepted) {
        for (int j = 0; j < n; ++j) { st.x[j] = st.x_trial[j]; wa3[j] = st.diag[j] * st.x[j]; }
        st.xnorm = lm_enorm(n, wa3);
        st.fnorm = fnorm1;
        st.iter += 1;
    }
    st.last_accepted = accepted ? 1 : 0;
    if (st.fnorm == 0.0) { st.termination = TERM_RESIDUALS_ZERO; return false; }
    const bool f_ok = fabs(actred) <= cfg.ftol && prered <= cfg.ftol && 0.5 * ratio <= 1.0;
    const bool x_ok = st.delta <= cfg.xtol * st.xnorm;
    if (f_ok && x_ok) { st.termination = TERM_CONVERGED_FTOL_XTOL; return false; }
    if (f_ok) { st.termination = TERM_CONVERGED_FTOL; return false; }
    if (x_ok) { st.termination = TERM_CONVERGED_XTOL; return false; }
    if (st.nfev >= cfg.maxfev) { st.termination = TERM_LOST_PATIENCE; return false; }
    if ((fabs(actred) <= cfg.epsmch && prered <= cfg.epsmch && 0.5 * ratio <= 1.0) ||
        st.delta <= cfg.epsmch * st.xnorm || st.gnorm <= cfg.epsmch) {
        st.termination = TERM_NO_IMPROVEMENT_POSSIBLE;
        return false;
    }
    if (accepted) {
        if (!lm_outer_prepare(st, cfg, ev)) return false;
    }
    lm_inner_propose(st);
    return true;
}

VP_HD bool lm_successful(int termination)
{
    return termination == TERM_RESIDUALS_ZERO || termination == TERM_ORTHOGONAL ||
           termination == TERM_CONVERGED_FTOL || termination == TERM_CONVERGED_XTOL ||
           termination == TERM_CONVERGED_FTOL_XTOL;
}

} // namespace vp
