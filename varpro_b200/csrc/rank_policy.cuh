// rank_policy.cuh -- which singular values of Phi_w count as zero in the inner linear solve.
//
// The reference solves C = V diag(sigma_i > eps ? 1/sigma_i : 0) U^T Y_w with the ABSOLUTE threshold eps of
// SeparableProblemBuilder::epsilon (src/solvers/levmar/mod.rs:52-54, src/problem/builder.rs:246-251,282), forms the
// residual with that C (:57-59) and keeps the UNtruncated U in the Jacobian's projector (:123-124). The original
// MATLAB code uses the RELATIVE rule sigma_i > m * eps_machine * sigma_1 (matlab/varpro.m:642-643).
//
// Here the panel is factored by QR (Phi_w = Q R1, backward stable), so the singular values of Phi_w are those of
// the n x n triangle R1 = Ur Sigma V^T and Phi_w = (Q Ur) Sigma V^T. The policy therefore costs nothing on a
// full-rank panel: with the bounds sigma_min >= 1 / ||R1^-1||_F and sigma_max <= ||R1||_F a cheap test proves
// "nothing is truncated" (both rules), and the triangular R1^-1 is used as before. Only if that test fails is the
// tiny SVD computed (one-sided Jacobi on R1, one thread); if singular values are then truncated the panel is
// re-expressed in the singular basis with the truncated directions removed from the SOLVE but not from the
// projector:
//     Q'' = Q Ur diag(keep),   Rinv_eff = V diag(keep_j / sigma_j)
//     b'' = Q''^T y,  c = Rinv_eff b''  (= V Sigma^+ U^T y),   r = y - Q'' b''  (= y - Phi_w c)
//     E = (I - Q Q^T) D  is computed with the FULL Q beforehand (the reference's untruncated projector).
// The streaming kernels need no change: they already multiply b by a general n x n Rinv and subtract Q b.
//
// tol encodes the rule: tol >= 0: absolute threshold eps (the reference; default = machine epsilon of the scalar);
// tol < 0: relative threshold |tol| * sigma_1 (the library passes -m * eps_machine for the MATLAB rule).
//
// Exactly rank-deficient panels (a column that is exactly zero after the previous Householder steps) are handled
// by the same path: R1 gets a zero diagonal, the cheap test fails, sigma = 0 is truncated under either rule.
#pragma once

#include <math.h>
#include <float.h>

#include "../../include/varpro_b200.h"

#ifndef __CUDACC__ // the CPU test harness (tests/csrc/lm_harness.cpp) compiles this header with g++
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace vp {

// A weighted basis column whose sum of squares overflows (entries beyond ~1e154, e.g. exp(-x/tau) at a slightly
// negative trial tau) cannot be factored; the reference's restatement ends up with a zero singular vector and a
// zero coefficient for it (sigma = inf => U_j = a_j / inf = 0, c_j = 0) and carries on with a finite residual, which
// the LM loop then simply rejects. The panel evaluators reproduce that: such a column -- and its derivative
// columns -- are zeroed before the factorisation (exact zero column => sigma = 0 => truncated), instead of letting
// inf * 0 poison the evaluation. Bit 0 of the evaluators' flag word = a non-finite entry of Phi_w (cache = None,
// src/solvers/levmar/mod.rs:43-72); bit 1 + j = column j overflows.
constexpr double RANK_HUGE_ENTRY = 1e154;

#ifdef __CUDACC__
// Bitwise OR of the evaluators' flag words over the CTA. __syncthreads_or only tells whether ANY thread had a
// non-zero word, which is the common (all clear) answer after one barrier; the per-bit reductions run only then.
__device__ __forceinline__ int block_or_flags(int flags, int nbits)
{
    if (!__syncthreads_or(flags)) return 0;
    int out = 0;
    for (int b = 0; b < nbits; ++b)
        if (__syncthreads_or((flags >> b) & 1)) out |= 1 << b;
    return out;
}
#endif

struct SmallSvd {
    int truncated;                      // 1: at least one singular value was truncated
    double Urot[VP_MAX_N * VP_MAX_N];   // Ur diag(keep), column-major, leading dimension n
    double RinvEff[VP_MAX_N * VP_MAX_N]; // V diag(keep/sigma), column-major, leading dimension n
};

// threshold below which a singular value is truncated, given sigma_max (or an upper bound of it)
__host__ __device__ inline double rank_threshold(double tol, double sigma_max)
{
    return tol >= 0.0 ? tol : -tol * sigma_max;
}

// Cheap sufficient test for "no singular value of the upper triangular R (n x n, column-major, ld ldr) is
// truncated": sigma_min >= 1/||R^-1||_F must exceed the threshold at sigma_max <= ||R||_F. Rinv: the triangular
// inverse the caller has already computed (column-major, ld ldi).
__host__ __device__ inline bool rank_surely_full(int n, const double *R, int ldr, const double *Rinv, int ldi, double tol)
{
    double fr = 0.0, fi = 0.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i <= j; ++i) {
            fr += R[j * ldr + i] * R[j * ldr + i];
            fi += Rinv[j * ldi + i] * Rinv[j * ldi + i];
        }
    if (!(fi > 0.0) || !isfinite(fi) || !isfinite(fr)) return false;
    return 1.0 / sqrt(fi) > rank_threshold(tol, sqrt(fr));
}

// One-sided Jacobi SVD of the n x n matrix R (column-major, ld ldr; upper triangular on entry, any matrix works)
// and the truncation. Serial: called by ONE thread, rarely. out->truncated = 0 means every singular value passed
// and the caller keeps its triangular factor.
__host__ __device__ inline void rank_policy_svd(int n, const double *R, int ldr, double tol, SmallSvd *out)
{
    double A[VP_MAX_N * VP_MAX_N], V[VP_MAX_N * VP_MAX_N], sigma[VP_MAX_N];
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            A[j * n + i] = (i <= j) ? R[j * ldr + i] : 0.0;
            V[j * n + i] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (int i = 0; i < n; ++i) {
                    alpha += A[p * n + i] * A[p * n + i];
                    beta += A[q * n + i] * A[q * n + i];
                    gamma += A[p * n + i] * A[q * n + i];
                }
                if (gamma == 0.0 || fabs(gamma) <= DBL_EPSILON * sqrt(alpha * beta)) continue;
                rotated = 1;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < n; ++i) {
                    const double ap = A[p * n + i], aq = A[q * n + i];
                    A[p * n + i] = c * ap - s * aq;
                    A[q * n + i] = s * ap + c * aq;
                    const double vp = V[p * n + i], vq = V[q * n + i];
                    V[p * n + i] = c * vp - s * vq;
                    V[q * n + i] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double smax = 0.0;
    for (int j = 0; j < n; ++j) {
        double s2 = 0.0;
        for (int i = 0; i < n; ++i) s2 += A[j * n + i] * A[j * n + i];
        sigma[j] = sqrt(s2);
        smax = fmax(smax, sigma[j]);
    }
    const double thr = rank_threshold(tol, smax);
    int truncated = 0;
    for (int j = 0; j < n; ++j) {
        const bool keep = sigma[j] > thr && isfinite(sigma[j]);
        if (!keep) truncated = 1;
        for (int i = 0; i < n; ++i) {
            out->Urot[j * n + i] = keep ? A[j * n + i] / sigma[j] : 0.0;   // column j of Ur, or zero
            out->RinvEff[j * n + i] = keep ? V[j * n + i] / sigma[j] : 0.0; // column j of V / sigma_j, or zero
        }
    }
    out->truncated = truncated;
}

} // namespace vp
