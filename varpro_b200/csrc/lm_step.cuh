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
// The step is the serial link between two evaluations of a fit (one thread, a chain of
// dependent fp64 divisions and square roots), so it is written for latency: every routine
// is a template on QT, the number of nonlinear parameters known at compile time
// (QT = 1..VP_LM_SPECIALISED; QT = 0: any q <= VP_LM_MAXQ at run time). For QT > 0 all loops
// have constant trip counts and every array index is a compile-time constant -- the pivot
// permutation is applied through select chains (lm_get / lm_put) -- so the whole state lives
// in registers and the only long-latency operations left are the divisions and square roots
// of the algorithm itself. The first version ran lmder on run-time-indexed arrays in local
// memory: ~31 us per step on B200 (profiles/r02v_queue_phase_breakdown.txt), more than
// streaming the observations six times.
// The specialised and the generic instantiation execute the same floating-point operations
// in the same order (tests/test_lm_state_machine.py checks bitwise equality).
//
// The code is __host__ __device__: the host-driven vp_fit loop, the in-kernel LM step and
// the CPU test harness (tests/csrc/lm_harness.cpp) share it.
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
#ifdef __CUDACC__
#define VP_LM_NOINLINE static __host__ __device__ __noinline__
// full unrolling for the compile-time-q instantiations only (the generic one keeps its loops rolled)
#define VP_LM_UNROLL _Pragma("unroll (QT > 0 ? 64 : 1)")
#else
#define VP_LM_NOINLINE static inline
#define VP_LM_UNROLL
#endif

#define VP_LM_MAXQ 8
#define VP_LM_SPECIALISED 4 // q = 1..4 run the register-resident instantiations

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
    double H[VP_LM_MAXQ * VP_LM_MAXQ]; // J^T J, column-major q x q (leading dimension q)
    int finite; // bit 0: the residual is valid (Phi_w and ||r||^2 finite: the reference's cache is Some); bit 1: g and H are finite
};
#define VP_EVAL_RESIDUAL_OK 1
#define VP_EVAL_DERIVS_OK 2
#define VP_EVAL_ALL_OK 3

// The lmder state between two evaluations. NQ = capacity of the arrays; matrices are packed
// with leading dimension q, so LmStateT<q> is a prefix-compatible view of LmStateT<VP_LM_MAXQ>.
template <int NQ>
struct LmStateT {
    int q;
    int phase;        // 0: waiting for the evaluation at x0, 1: waiting for a trial evaluation
    int termination;  // Termination
    int nfev, iter;
    int last_accepted; // 1 if the most recent trial was accepted (trial buffers hold the state of x)
    double x[NQ];       // last accepted parameters
    double x_trial[NQ]; // parameters the next evaluation must be made at
    double step[NQ];    // x_trial - x
    double fnorm, xnorm, gnorm, delta, par, pnorm;
    double diag[NQ];
    double R[NQ * NQ]; // upper triangle of the pivoted Cholesky factor, col-major, leading dimension q
    double qtf[NQ];
    double acnorm[NQ];
    int ipvt[NQ];
};
typedef LmStateT<VP_LM_MAXQ> LmState;

// ---- indexing with a run-time index into a register-resident array --------------------------
template <int QT, typename T, int NA>
VP_HD T lm_get(const T (&a)[NA], const int idx)
{
    if constexpr (QT > 0) {
        T v = a[0];
        VP_LM_UNROLL
        for (int k = 1; k < QT; ++k) v = (idx == k) ? a[k] : v;
        return v;
    } else {
        return a[idx];
    }
}
template <int QT, typename T, int NA>
VP_HD void lm_put(T (&a)[NA], const int idx, const T v)
{
    if constexpr (QT > 0) {
        VP_LM_UNROLL
        for (int k = 0; k < QT; ++k) a[k] = (idx == k) ? v : a[k];
    } else {
        a[idx] = v;
    }
}

template <int QT, int NA>
VP_HD double lm_enorm(const int n, const double (&v)[NA])
{
    double s = 0.0;
    VP_LM_UNROLL
    for (int i = 0; i < (QT > 0 ? QT : n); ++i) s += v[i] * v[i];
    return sqrt(s);
}

// Pivoted Cholesky of H standing in for MINPACK qrfac(J): R upper triangular,
// P^T H P = R^T R, pivot = largest remaining diagonal of the Schur complement
// (= largest remaining column norm of J, qrfac's rule). Rank deficiency gives
// zero rows, which lmpar treats like qrfac's zero diagonals.
template <int QT, int NQ>
VP_HD void lm_pivoted_cholesky(const int nrt, const double *H, double (&R)[NQ * NQ], int (&ipvt)[NQ], double (&acnorm)[NQ])
{
    const int n = QT > 0 ? QT : nrt;
    double A[NQ * NQ];
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        ipvt[j] = j;
        acnorm[j] = sqrt(fmax(H[j * n + j], 0.0));
        VP_LM_UNROLL
        for (int i = 0; i < n; ++i) {
            A[j * n + i] = H[j * n + i];
            R[j * n + i] = 0.0;
        }
    }
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        int kmax = j;
        double dmax = A[j * n + j];
        VP_LM_UNROLL
        for (int k = j + 1; k < n; ++k)
            if (A[k * n + k] > dmax) { kmax = k; dmax = A[k * n + k]; }
        // symmetric swap of rows/columns j and kmax of the working matrix, and of the already
        // computed columns of R (the loop over k finds kmax with compile-time indices)
        VP_LM_UNROLL
        for (int k = j + 1; k < n; ++k) {
            if (k != kmax) continue;
            VP_LM_UNROLL
            for (int i = 0; i < n; ++i) { const double t = A[j * n + i]; A[j * n + i] = A[k * n + i]; A[k * n + i] = t; }
            VP_LM_UNROLL
            for (int i = 0; i < n; ++i) { const double t = A[i * n + j]; A[i * n + j] = A[i * n + k]; A[i * n + k] = t; }
            VP_LM_UNROLL
            for (int i = 0; i < j; ++i) { const double t = R[j * n + i]; R[j * n + i] = R[k * n + i]; R[k * n + i] = t; }
            const int t = ipvt[j]; ipvt[j] = ipvt[k]; ipvt[k] = t;
        }
        const double d = A[j * n + j];
        if (!(d > 0.0)) {
            // remaining block is (numerically) zero: rank deficient
            VP_LM_UNROLL
            for (int k = j; k < n; ++k) R[k * n + j] = 0.0;
            VP_LM_UNROLL
            for (int k = j + 1; k < n; ++k) A[k * n + k] = 0.0;
            continue;
        }
        const double rjj = sqrt(d);
        R[j * n + j] = rjj;
        VP_LM_UNROLL
        for (int k = j + 1; k < n; ++k) R[k * n + j] = A[k * n + j] / rjj;
        VP_LM_UNROLL
        for (int k = j + 1; k < n; ++k) {
            VP_LM_UNROLL
            for (int i = j + 1; i <= k; ++i) {
                A[k * n + i] -= R[i * n + j] * R[k * n + j];
                A[i * n + k] = A[k * n + i];
            }
        }
    }
}

// MINPACK qrsolv on a private copy S of R (R itself is left untouched).
template <int QT, int NQ>
VP_HD void lm_qrsolv(const int nrt, const double (&R)[NQ * NQ], const int (&ipvt)[NQ], const double (&diag)[NQ],
                     const double (&qtb)[NQ], double (&x)[NQ], double (&sdiag)[NQ], double (&S)[NQ * NQ])
{
    const int n = QT > 0 ? QT : nrt;
    double wa[NQ];
    // S holds r^T in its strict lower triangle during the sweep (MINPACK stores
    // it inside r); S(i,j) i>j.
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        VP_LM_UNROLL
        for (int i = 0; i < n; ++i) S[j * n + i] = R[j * n + i];
    }
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        VP_LM_UNROLL
        for (int i = j; i < n; ++i) S[j * n + i] = S[i * n + j];
        x[j] = S[j * n + j];
        wa[j] = qtb[j];
    }
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        const double dl = lm_get<QT>(diag, ipvt[j]);
        if (dl != 0.0) {
            VP_LM_UNROLL
            for (int k = j; k < n; ++k) sdiag[k] = 0.0;
            sdiag[j] = dl;
            double qtbpj = 0.0;
            VP_LM_UNROLL
            for (int k = j; k < n; ++k) {
                if (sdiag[k] == 0.0) continue;
                double c, s;
                const double rkk = S[k * n + k];
                if (fabs(rkk) < fabs(sdiag[k])) {
                    const double cotan = rkk / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
                    c = s * cotan;
                } else {
                    const double t = sdiag[k] / rkk;
                    c = 0.5 / sqrt(0.25 + 0.25 * t * t);
                    s = c * t;
                }
                S[k * n + k] = c * rkk + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                VP_LM_UNROLL
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
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    // back substitution over j = nsing-1 .. 0 (constant trip count, predicated)
    VP_LM_UNROLL
    for (int j = n - 1; j >= 0; --j) {
        if (j >= nsing) continue;
        double sum = 0.0;
        VP_LM_UNROLL
        for (int i = j + 1; i < n; ++i)
            if (i < nsing) sum += S[j * n + i] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) lm_put<QT>(x, ipvt[j], wa[j]);
}

// MINPACK lmpar: determines par such that ||D x|| ~ delta. x receives the
// solution of (J^T J + par D^2) x = J^T r.
template <int QT, int NQ>
VP_HD void lm_lmpar(const int nrt, const double (&R)[NQ * NQ], const int (&ipvt)[NQ], const double (&diag)[NQ],
                    const double (&qtb)[NQ], const double delta, double &par, double (&x)[NQ])
{
    const int n = QT > 0 ? QT : nrt;
    const double dwarf = DBL_MIN;
    double wa1[NQ], wa2[NQ], sdiag[NQ];
    double S[NQ * NQ];
    int nsing = n;
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        wa1[j] = qtb[j];
        if (R[j * n + j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    VP_LM_UNROLL
    for (int j = n - 1; j >= 0; --j) {
        if (j >= nsing) continue;
        wa1[j] /= R[j * n + j];
        const double temp = wa1[j];
        VP_LM_UNROLL
        for (int i = 0; i < j; ++i) wa1[i] -= R[j * n + i] * temp;
    }
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) lm_put<QT>(x, ipvt[j], wa1[j]);
    int iter = 0;
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = lm_enorm<QT>(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { par = 0.0; return; }
    double parl = 0.0;
    if (nsing >= n) {
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) {
            const int l = ipvt[j];
            wa1[j] = lm_get<QT>(diag, l) * (lm_get<QT>(wa2, l) / dxnorm);
        }
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            VP_LM_UNROLL
            for (int i = 0; i < j; ++i) sum += R[j * n + i] * wa1[i];
            wa1[j] = (wa1[j] - sum) / R[j * n + j];
        }
        const double temp = lm_enorm<QT>(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        double sum = 0.0;
        VP_LM_UNROLL
        for (int i = 0; i <= j; ++i) sum += R[j * n + i] * qtb[i];
        wa1[j] = sum / lm_get<QT>(diag, ipvt[j]);
    }
    const double gnorm = lm_enorm<QT>(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    for (;;) {
        ++iter;
        if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
        double temp = sqrt(par);
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) wa1[j] = temp * diag[j];
        lm_qrsolv<QT, NQ>(n, R, ipvt, wa1, qtb, x, sdiag, S);
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = lm_enorm<QT>(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) {
            const int l = ipvt[j];
            wa1[j] = lm_get<QT>(diag, l) * (lm_get<QT>(wa2, l) / dxnorm);
        }
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) {
            wa1[j] /= sdiag[j];
            temp = wa1[j];
            VP_LM_UNROLL
            for (int i = j + 1; i < n; ++i) wa1[i] -= S[j * n + i] * temp;
        }
        temp = lm_enorm<QT>(n, wa1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
}

template <int NQ>
VP_HD void lm_init(LmStateT<NQ> &st, const int q, const double *x0)
{
    st.q = q;
    st.phase = 0;
    st.termination = TERM_RUNNING;
    st.nfev = 0;
    st.iter = 1;
    st.last_accepted = 1;
    st.fnorm = st.xnorm = st.gnorm = st.delta = st.par = st.pnorm = 0.0;
    for (int j = 0; j < NQ; ++j) {
        st.x[j] = j < q ? x0[j] : 0.0;
        st.x_trial[j] = st.x[j];
        st.step[j] = 0.0;
        st.diag[j] = 1.0;
        st.qtf[j] = 0.0;
        st.acnorm[j] = 0.0;
        st.ipvt[j] = j;
    }
    for (int j = 0; j < NQ * NQ; ++j) st.R[j] = 0.0;
}

// Start of an lmder outer iteration: factor H at the accepted point, form qtf,
// the scaled gradient norm and the diag rescale. Returns false on termination.
template <int QT, int NQ>
VP_HD bool lm_outer_prepare(LmStateT<NQ> &st, const LmConfig &cfg, const double *Hev, const double *gev)
{
    const int n = QT > 0 ? QT : st.q;
    lm_pivoted_cholesky<QT, NQ>(n, Hev, st.R, st.ipvt, st.acnorm);
    if (st.iter == 1) {
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) {
            st.diag[j] = cfg.scale_diag ? st.acnorm[j] : 1.0;
            if (cfg.scale_diag && st.acnorm[j] == 0.0) st.diag[j] = 1.0;
        }
        double wa3[NQ];
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) wa3[j] = st.diag[j] * st.x[j];
        st.xnorm = lm_enorm<QT>(n, wa3);
        st.delta = cfg.stepbound * st.xnorm;
        if (st.delta == 0.0) st.delta = cfg.stepbound;
    }
    // qtf = R^-T P^T g (forward substitution; zero for the rank-deficient tail)
    double gl[NQ];
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) gl[j] = gev[j];
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        double sum = lm_get<QT>(gl, st.ipvt[j]);
        VP_LM_UNROLL
        for (int i = 0; i < j; ++i) sum -= st.R[j * n + i] * st.qtf[i];
        st.qtf[j] = (st.R[j * n + j] != 0.0) ? sum / st.R[j * n + j] : 0.0;
    }
    st.gnorm = 0.0;
    if (st.fnorm != 0.0) {
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) {
            const double acl = lm_get<QT>(st.acnorm, st.ipvt[j]);
            if (acl != 0.0) {
                double sum = 0.0;
                VP_LM_UNROLL
                for (int i = 0; i <= j; ++i) sum += st.R[j * n + i] * (st.qtf[i] / st.fnorm);
                st.gnorm = fmax(st.gnorm, fabs(sum / acl));
            }
        }
    }
    if (!isfinite(st.gnorm)) { st.termination = TERM_NUMERICAL; return false; }
    if (st.gnorm <= cfg.gtol) { st.termination = TERM_ORTHOGONAL; return false; }
    if (cfg.scale_diag) {
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) st.diag[j] = fmax(st.diag[j], st.acnorm[j]);
    }
    return true;
}

// lmder inner loop head: lmpar -> trial point.
template <int QT, int NQ>
VP_HD void lm_inner_propose(LmStateT<NQ> &st)
{
    const int n = QT > 0 ? QT : st.q;
    double p[NQ], wa3[NQ];
    lm_lmpar<QT, NQ>(n, st.R, st.ipvt, st.diag, st.qtf, st.delta, st.par, p);
    VP_LM_UNROLL
    for (int j = 0; j < n; ++j) {
        st.step[j] = -p[j];
        st.x_trial[j] = st.x[j] + st.step[j];
        wa3[j] = st.diag[j] * st.step[j];
    }
    st.pnorm = lm_enorm<QT>(n, wa3);
    if (st.iter == 1) st.delta = fmin(st.delta, st.pnorm);
}

// The state machine proper on a state with compile-time capacity (see lm_advance below).
template <int QT, int NQ>
VP_HD bool lm_advance_core(LmStateT<NQ> &st, const LmConfig &cfg, const double rnorm2, const int finite, const double *gev,
                           const double *Hev)
{
    const int n = QT > 0 ? QT : st.q;
    if (st.termination != TERM_RUNNING) return false;
    bool prepare; // a new outer iteration starts: refactor H at the (new) accepted point
    if (st.phase == 0) {
        st.nfev = 1;
        st.last_accepted = 1;
        if (n == 0) { st.termination = TERM_NO_PARAMETERS; return false; }
        const double fn = sqrt(rnorm2);
        // residuals() is None -> the crate stops with User; a non-finite norm of an existing residual -> Numerical
        if (!(finite & VP_EVAL_RESIDUAL_OK)) { st.fnorm = fn; st.termination = TERM_USER; return false; }
        if (!isfinite(fn)) { st.fnorm = fn; st.termination = TERM_NUMERICAL; return false; }
        st.fnorm = fn;
        if (fn == 0.0) { st.termination = TERM_RESIDUALS_ZERO; return false; }
        st.par = 0.0;
        st.iter = 1;
        st.phase = 1;
        prepare = true;
    } else {
        // trial evaluation
        st.nfev += 1;
        const double fnorm1 = sqrt(rnorm2);
        if (!(finite & VP_EVAL_RESIDUAL_OK)) { st.last_accepted = 0; st.termination = TERM_USER; return false; }
        if (!isfinite(fnorm1)) { st.last_accepted = 0; st.termination = TERM_NUMERICAL; return false; }
        double actred = -1.0;
        if (0.1 * fnorm1 < st.fnorm) actred = 1.0 - (fnorm1 / st.fnorm) * (fnorm1 / st.fnorm);
        double wa3[NQ];
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) wa3[j] = 0.0;
        VP_LM_UNROLL
        for (int j = 0; j < n; ++j) {
            const double temp = lm_get<QT>(st.step, st.ipvt[j]);
            VP_LM_UNROLL
            for (int i = 0; i <= j; ++i) wa3[i] += st.R[j * n + i] * temp;
        }
        const double temp1 = lm_enorm<QT>(n, wa3) / st.fnorm;
        const double temp2 = (sqrt(st.par) * st.pnorm) / st.fnorm;
        const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
        const double dirder = -(temp1 * temp1 + temp2 * temp2);
        const double ratio = (prered != 0.0) ? actred / prered : 0.0;
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
            VP_LM_UNROLL
            for (int j = 0; j < n; ++j) { st.x[j] = st.x_trial[j]; wa3[j] = st.diag[j] * st.x[j]; }
            st.xnorm = lm_enorm<QT>(n, wa3);
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
        prepare = accepted;
    }
    // one copy of the factorisation / lmpar code for both entry points (the step is straight-line
    // code of tens of KB once unrolled; the instruction cache is cold every time it runs)
    if (prepare) {
        // the Jacobian of the (new) accepted point is needed now: non-finite derivatives end the fit here, a
        // rejected trial with a valid residual never looks at them (like lmder, which evaluates J only after acceptance)
        if (!(finite & VP_EVAL_DERIVS_OK)) { st.termination = TERM_NUMERICAL; return false; }
        if (!lm_outer_prepare<QT, NQ>(st, cfg, Hev, gev)) return false;
    }
    lm_inner_propose<QT, NQ>(st);
    return true;
}

// Specialised step: copy the q-sized prefix of the state into a register-resident LmStateT<QT>,
// advance, copy back. `st` may live in shared, global or host memory.
template <int QT>
VP_LM_NOINLINE bool lm_advance_q(LmState &st, const LmConfig &cfg, const LmEval &ev)
{
    LmStateT<QT> s;
    s.q = QT; s.phase = st.phase; s.termination = st.termination; s.nfev = st.nfev; s.iter = st.iter;
    s.last_accepted = st.last_accepted;
    s.fnorm = st.fnorm; s.xnorm = st.xnorm; s.gnorm = st.gnorm; s.delta = st.delta; s.par = st.par; s.pnorm = st.pnorm;
    double g[QT], H[QT * QT];
    VP_LM_UNROLL
    for (int j = 0; j < QT; ++j) {
        s.x[j] = st.x[j]; s.x_trial[j] = st.x_trial[j]; s.step[j] = st.step[j]; s.diag[j] = st.diag[j];
        s.qtf[j] = st.qtf[j]; s.acnorm[j] = st.acnorm[j]; s.ipvt[j] = st.ipvt[j];
        g[j] = ev.g[j];
    }
    VP_LM_UNROLL
    for (int j = 0; j < QT * QT; ++j) { s.R[j] = st.R[j]; H[j] = ev.H[j]; }
    const bool more = lm_advance_core<QT, QT>(s, cfg, ev.rnorm2, ev.finite, g, H);
    st.phase = s.phase; st.termination = s.termination; st.nfev = s.nfev; st.iter = s.iter;
    st.last_accepted = s.last_accepted;
    st.fnorm = s.fnorm; st.xnorm = s.xnorm; st.gnorm = s.gnorm; st.delta = s.delta; st.par = s.par; st.pnorm = s.pnorm;
    VP_LM_UNROLL
    for (int j = 0; j < QT; ++j) {
        st.x[j] = s.x[j]; st.x_trial[j] = s.x_trial[j]; st.step[j] = s.step[j]; st.diag[j] = s.diag[j];
        st.qtf[j] = s.qtf[j]; st.acnorm[j] = s.acnorm[j]; st.ipvt[j] = s.ipvt[j];
    }
    VP_LM_UNROLL
    for (int j = 0; j < QT * QT; ++j) st.R[j] = s.R[j];
    return more;
}

// Any q <= VP_LM_MAXQ with run-time loop bounds (the arrays are indexed dynamically).
VP_LM_NOINLINE bool lm_advance_generic(LmState &st, const LmConfig &cfg, const LmEval &ev)
{
    return lm_advance_core<0, VP_LM_MAXQ>(st, cfg, ev.rnorm2, ev.finite, ev.g, ev.H);
}

// Feed the evaluation made at st.x_trial. Returns true if another evaluation
// (at the new st.x_trial) is required, false when the fit has terminated.
// `H_acc`/`g_acc` semantics: when a trial is accepted the caller's evaluation
// becomes the accepted one (st.last_accepted = 1) and its H, g seed the next
// outer iteration -- the Jacobian products come out of the same streaming pass
// as the residual norm, so no second pass is needed.
VP_HD bool lm_advance(LmState &st, const LmConfig &cfg, const LmEval &ev)
{
    switch (st.q) {
    case 1: return lm_advance_q<1>(st, cfg, ev);
    case 2: return lm_advance_q<2>(st, cfg, ev);
    case 3: return lm_advance_q<3>(st, cfg, ev);
    case 4: return lm_advance_q<4>(st, cfg, ev);
    default: return lm_advance_generic(st, cfg, ev);
    }
}

VP_HD bool lm_successful(int termination)
{
    return termination == TERM_RESIDUALS_ZERO || termination == TERM_ORTHOGONAL ||
           termination == TERM_CONVERGED_FTOL || termination == TERM_CONVERGED_XTOL ||
           termination == TERM_CONVERGED_FTOL_XTOL;
}

} // namespace vp
