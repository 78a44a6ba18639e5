// batch_fit_kernel.cuh -- K4: independent-batch evaluator + fit (BASELINE config 3).
//
// P independent problems share the model structure, the independent variable x and the weights,
// but each has its OWN observations y_p, nonlinear parameters alpha_p and linear coefficients c_p:
// what a user of the reference writes as a loop over SingleRhs problems,
//   for p in 0..P { LevMarSolver::default().fit(SeparableProblemBuilder::new(model_p).observations(y_p).build()) }
// (src/problem/builder.rs:116-324, src/solvers/levmar/mod.rs:238-254 per problem).
//
// One CTA fits one problem at a time, start to finish, and then takes the next one from a global
// counter (fits need different numbers of evaluations): y_p is read from HBM ONCE per fit and stays
// in registers; per evaluation the CTA
//   1. regenerates Phi_w(alpha_p), D(alpha_p) from x into a shared-memory working matrix (m x (n+p))
//      -- materialising Phi for 65 536 problems would take 6.4 GB --
//   2. runs n Householder steps on [Phi_w | D | y] (y carried along in registers): one fused block
//      reduction per step (panel_kernel_hh.cuh explains the identities),
//   3. one more reduction over the rows >= n of the rotated system gives everything the LM step
//      needs:  ||r||^2 = sum y~_i^2,  u_e = D~_e . y~,  M_ef = D~_e . D~_f   and the top rows give
//      c = R1^-1 y~[0:n];  then  g_k = -sum_{e in k} c_j(e) u_e,  H_kl = sum M_ef c_j(e) c_j(f)
//      (the S = 1 case of the formulas in stream_kernel.cuh; no explicit Q or E is formed),
//   4. thread 0 advances the lmder state machine (lm_step.cuh) in shared memory.
// Bound: fp64 ALU / exp and block-reduction latency, not HBM (32 KB of y per fit against ~15
// evaluations of ~0.5 MFLOP each).
#pragma once

#include "device_common.cuh"
#include "lm_step.cuh"
#include "panel_kernel_hh.cuh"

namespace vp {

// Sum K per-thread values over the CTA; every thread gets the totals. Two-level: per-warp shuffle
// fold -> shared, warp 0 folds the NW per-warp partials (one lane per value), everybody reads the K
// totals back (with 16 warps and K ~ 10 the one-level version makes every thread read NW*K values).
// buf: NW*K + K doubles; two __syncthreads; buf must not be reused by the next call.
template <int K, int NW>
__device__ __forceinline__ void block_sum_two_level(double (&v)[K], double *buf)
{
    static_assert(K <= 32, "one lane of warp 0 per value");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) buf[warp * K + k] = t;
    }
    __syncthreads();
    if (warp == 0 && lane < K) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += buf[w * K + lane];
        buf[NW * K + lane] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = buf[NW * K + k];
}

struct BatchArgs {
    ModelDesc md;
    const double *x;      // m
    const double *w;      // m or nullptr
    const double *Y;      // ld x P observations (unweighted), column p = problem p
    long long ld;
    long long P;
    double svd_eps;
    LmConfig cfg;
    const double *alpha0; // q x P
    double *alpha_out;    // q x P
    double *C_out;        // n x P
    double *obj_out;      // P: 0.5 * ||r_w||^2
    int *term_out;        // P: Termination
    int *nfev_out;        // P
    unsigned long long *next; // work counter (zeroed by the host)
    int mpad;             // rows of the shared-memory working matrix (>= m, multiple of 2)
};

// What one evaluation leaves behind for the LM phase (per problem slot, shared memory).
template <int N, int P, int KMAX>
struct BatchTail {
    double tv[KMAX];          // ||r||^2, u_e, M_ef (upper triangle)
    double top[N][N + P + 1]; // rows 0..n-1 of the rotated system: R, Q^T D, Q^T y
    double rdiag[N];
    int dropped, bad;
};

// G problems are in flight per CTA ("slots"). The serial LM step of a problem costs about as much as
// its evaluation (ncu: 47 % of all samples were the other 511 threads waiting for thread 0), so the
// CTA evaluates its G problems one after the other with all threads and then runs the G LM steps
// CONCURRENTLY, each on lane 0 of a different warp: the LM latency is paid once per G evaluations.
// y_p is re-read from global memory for every evaluation (first touch from HBM, then L2): keeping
// G columns in registers is not possible at 128 registers per thread.
template <int N, int P, int RPT, int THREADS, int G = 4>
__global__ void __launch_bounds__(THREADS, 1)
batch_fit_kernel(const BatchArgs a)
{
    constexpr int NPV = N + P;
    constexpr int NW = THREADS / 32;
    constexpr int NTAIL = 1 + P + P * (P + 1) / 2; // ||r||^2, u_e, M_ef (upper)
    constexpr int KMAX = (NTAIL > NPV + 1) ? NTAIL : NPV + 1;
    static_assert(G <= NW, "one warp per slot in the LM phase");
    extern __shared__ __align__(16) double colm[]; // NPV columns of mpad doubles
    __shared__ double red[2][NW * KMAX + KMAX];
    __shared__ double alpha_s[VP_MAX_Q];
    __shared__ LmState st_s[G];
    __shared__ BatchTail<N, P, KMAX> tail_s[G];
    __shared__ double coef_acc[G][N];
    __shared__ long long prob_s[G]; // problem in the slot, -1 = empty
    __shared__ int exhausted_s, nactive_s;

    const int tid = threadIdx.x;
    const int m = a.md.m, mpad = a.mpad, q = a.md.q;

    // the thread's rows of x and w (shared by all problems)
    double xi[RPT], wi[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
        const bool in = i < m;
        xi[r] = in ? a.x[i] : 0.0;
        wi[r] = in ? (a.w ? a.w[i] : 1.0) : 0.0;
    }
    if (tid < G) prob_s[tid] = -1;
    if (tid == 0) exhausted_s = 0;
    __syncthreads();

    for (;;) {
        // ---- refill empty slots from the global work counter -------------------------------------
        if (tid == 0) {
            int nact = 0;
            for (int g = 0; g < G; ++g) {
                if (prob_s[g] < 0 && !exhausted_s) {
                    const long long pnew = (long long)atomicAdd(a.next, 1ull);
                    if (pnew < a.P) {
                        double x0[VP_MAX_Q];
                        for (int k = 0; k < VP_MAX_Q; ++k) x0[k] = k < q ? a.alpha0[(size_t)pnew * q + k] : 0.0;
                        lm_init(st_s[g], q, x0);
                        prob_s[g] = pnew;
                    } else {
                        exhausted_s = 1;
                    }
                }
                nact += prob_s[g] >= 0;
            }
            nactive_s = nact;
        }
        __syncthreads();
        if (nactive_s == 0) break;

        // ---- evaluation phase: the active slots one after the other, all threads -----------------
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const long long prob = prob_s[g];
            if (prob < 0) continue; // uniform
            BatchTail<N, P, KMAX> &tl = tail_s[g];
            if (tid < VP_MAX_Q) alpha_s[tid] = tid < q ? st_s[g].x_trial[tid] : 0.0;
            // y_p (weighted like builder.rs:307): issued first, consumed after the basis evaluation
            double yt[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int i = tid + r * THREADS;
                yt[r] = (i < m) ? wi[r] * __ldg(&a.Y[(size_t)prob * a.ld + i]) : 0.0;
            }
            __syncthreads();

            // 1. Phi_w, D into the working matrix (rolled over basis functions, rows unrolled)
            int bad = 0;
            {
                int e = 0;
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    const int kind = a.md.kind[j], np = a.md.npar[j];
                    const double a0 = np > 0 ? alpha_s[a.md.pidx[j][0]] : 0.0;
                    const double a1 = np > 1 ? alpha_s[a.md.pidx[j][1]] : 0.0;
                    const double scale = a.md.scale[j];
                    const double inv0 = (kind == VP_BASIS_EXP_DECAY) ? 1.0 / a0 : 0.0;
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const int i = tid + r * THREADS;
                        double v, da = 0.0, db = 0.0;
                        if (kind == VP_BASIS_EXP_DECAY) {
                            const double t = xi[r] * inv0, ex = exp(-t); // one division per basis function, see basis_eval_all
                            v = ex; da = ex * t * inv0;
                        } else if (kind == VP_BASIS_EXP_RATE_COS) {
                            const double ex = exp(-a0 * xi[r]);
                            double sn, cs;
                            sincos(a1 * xi[r], &sn, &cs);
                            v = ex * cs; da = -xi[r] * (ex * cs); db = -xi[r] * ex * sn;
                        } else if (kind == VP_BASIS_SIN_PHASE) {
                            double sn, cs;
                            sincos(a0 * xi[r] + a1, &sn, &cs);
                            v = sn; da = xi[r] * cs; db = cs;
                        } else {
                            v = kind == VP_BASIS_CONSTANT ? 1.0 : (kind == VP_BASIS_LINEAR_X ? scale * xi[r] : nan(""));
                        }
                        if (i < m) {
                            const double pv = wi[r] * v, pa = wi[r] * da, pb = wi[r] * db;
                            bad |= (!isfinite(pv) ? 1 : 0) | (fabs(pv) > RANK_HUGE_ENTRY ? (2 << j) : 0); // flag word of rank_policy.cuh
                            colm[(size_t)j * mpad + i] = pv;
                            if (np > 0) colm[(size_t)(N + e) * mpad + i] = pa;
                            if (np > 1) colm[(size_t)(N + e + 1) * mpad + i] = pb;
                        }
                    }
                    e += np;
                }
            }
            bad = block_or_flags(bad, 1 + N);
            if (bad >> 1) { // overflowing basis columns: zero them and their derivative columns
                for (int idx = tid; idx < m * NPV; idx += THREADS) {
                    const int c = idx / m, i = idx - c * m;
                    const int j = c < N ? c : a.md.e_basis[c - N];
                    if ((bad >> (1 + j)) & 1) colm[(size_t)c * mpad + i] = 0.0;
                }
                __syncthreads();
            }
            bad &= 1;

            // 2. Householder steps on [Phi_w | D | y] (y in registers)
            int dropped = 0;
#pragma unroll
            for (int J = 0; J < N; ++J) {
                const int K = NPV - J + 1; // sigma, dots with the columns to the right, dot with y
                double s[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) s[k] = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int i = tid + r * THREADS;
                    if (i >= J && i < m) {
                        const double aj = colm[(size_t)J * mpad + i];
#pragma unroll
                        for (int k = 0; k < NPV - J; ++k) s[k] = fma(aj, colm[(size_t)(J + k) * mpad + i], s[k]);
                        s[NPV - J] = fma(aj, yt[r], s[NPV - J]);
                    }
                }
                if (tid == J) { // row J is owned by thread J (r = 0): publish it
#pragma unroll
                    for (int k = 0; k < NPV; ++k) tl.top[J][k] = colm[(size_t)k * mpad + J];
                    tl.top[J][NPV] = yt[0];
                }
                (void)K;
                block_sum_two_level<KMAX, NW>(s, red[J & 1]);
                const double sigma = s[0];
                const double ajj = tl.top[J][J];
                const double nrm = sqrt(sigma);
                const bool keep = isfinite(nrm) && nrm > 0.0; // near-dependence: rank policy on R1 in the LM phase
                const double al = (ajj >= 0.0) ? -nrm : nrm;
                const double vnorm2 = 2.0 * (sigma - ajj * al);
                const double bt = (keep && vnorm2 > 0.0) ? 2.0 / vnorm2 : 0.0;
                if (tid == 0) tl.rdiag[J] = keep ? al : 0.0;
                if (!keep) dropped |= 1 << J;
                const double vjj = ajj - al;
                double tau[NPV + 1];
#pragma unroll
                for (int k = 1; k < NPV - J; ++k) tau[k] = bt * (s[k] - al * tl.top[J][J + k]);
                const double tau_y = bt * (s[NPV - J] - al * tl.top[J][NPV]);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int i = tid + r * THREADS;
                    if (i >= J && i < m) {
                        const double vi = (i > J) ? colm[(size_t)J * mpad + i] : vjj;
#pragma unroll
                        for (int k = 1; k < NPV - J; ++k)
                            colm[(size_t)(J + k) * mpad + i] = fma(-tau[k], vi, colm[(size_t)(J + k) * mpad + i]);
                        yt[r] = fma(-tau_y, vi, yt[r]);
                    }
                }
                __syncthreads(); // everyone has read top[J]; its owner rewrites it with the final row
                if (tid == J) {
#pragma unroll
                    for (int k = J + 1; k < NPV; ++k) tl.top[J][k] = colm[(size_t)k * mpad + J]; // R[J][k], (Q^T D)[J][:]
                    tl.top[J][NPV] = yt[0];                                                        // (Q^T y)[J]
                }
            }

            // 3. tail reduction: rows >= n (the untruncated projector; truncated directions are added by the rank policy)
            {
                double tv[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) tv[k] = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int i = tid + r * THREADS;
                    const bool tail = (i < m) && (i >= N);
                    if (tail) {
                        const double yv = yt[r];
                        tv[0] = fma(yv, yv, tv[0]);
                        double de[P > 0 ? P : 1];
#pragma unroll
                        for (int e = 0; e < P; ++e) {
                            de[e] = colm[(size_t)(N + e) * mpad + i];
                            tv[1 + e] = fma(de[e], yv, tv[1 + e]);
                        }
                        int t = 1 + P;
#pragma unroll
                        for (int e = 0; e < P; ++e)
#pragma unroll
                            for (int f2 = e; f2 < P; ++f2) { tv[t] = fma(de[e], de[f2], tv[t]); ++t; }
                    }
                }
                block_sum_two_level<KMAX, NW>(tv, red[N & 1]);
                if (tid == 0) {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) tl.tv[k] = tv[k];
                    tl.dropped = dropped;
                    tl.bad = bad;
                }
            }
            __syncthreads(); // the working matrix and the reduction buffers are free for the next slot
        }

        // ---- LM phase: lane 0 of warp g advances slot g; the G steps run concurrently ----------------
        if ((tid & 31) == 0 && (tid >> 5) < G && prob_s[tid >> 5] >= 0) {
            const int g = tid >> 5;
            const BatchTail<N, P, KMAX> &tl = tail_s[g];
            // inner solve c = R1^-1 (Q^T y) under the rank policy (rank_policy.cuh): triangular inverse, cheap
            // full-rank test, and -- rarely -- the SVD of R1 with the truncated directions moved into the residual
            double Rm[N * N], Ri[N * N], coef[N];
#pragma unroll
            for (int c = 0; c < N; ++c)
#pragma unroll
                for (int i = 0; i < N; ++i) { Rm[c * N + i] = (i < c) ? tl.top[i][c] : ((i == c) ? tl.rdiag[c] : 0.0); Ri[c * N + i] = 0.0; }
#pragma unroll
            for (int c = 0; c < N; ++c)
#pragma unroll
                for (int i = N - 1; i >= 0; --i) {
                    if (i > c) continue;
                    double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < N; ++k)
                        if (k > i && k <= c) sacc -= Rm[k * N + i] * Ri[c * N + k];
                    Ri[c * N + i] = sacc / Rm[i * N + i];
                }
            double rn2_extra = 0.0;
            if (rank_surely_full(N, Rm, N, Ri, N, a.svd_eps)) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    double sacc = 0.0;
#pragma unroll
                    for (int c = i; c < N; ++c) sacc += Ri[c * N + i] * tl.top[c][NPV];
                    coef[i] = sacc;
                }
            } else {
                SmallSvd sv;
                rank_policy_svd(N, Rm, N, a.svd_eps, &sv);
                double bb[N], b2 = 0.0, bb2 = 0.0;
                for (int c = 0; c < N; ++c) {
                    double sacc = 0.0;
                    for (int k = 0; k < N; ++k) sacc += sv.Urot[c * N + k] * tl.top[k][NPV];
                    bb[c] = sacc;
                    bb2 += sacc * sacc;
                    b2 += tl.top[c][NPV] * tl.top[c][NPV];
                }
                for (int i = 0; i < N; ++i) {
                    double sacc = 0.0;
                    for (int c = 0; c < N; ++c) sacc += sv.RinvEff[c * N + i] * bb[c];
                    coef[i] = sacc;
                }
                rn2_extra = fmax(b2 - bb2, 0.0); // the truncated components of Q^T y stay in the residual
            }
            LmEval ev;
            ev.rnorm2 = tl.tv[0] + rn2_extra;
            const int resid_ok = !tl.bad && isfinite(tl.tv[0] + rn2_extra);
            int finite = 1; // derivatives
            for (int k = 0; k < VP_LM_MAXQ; ++k) ev.g[k] = 0.0;
            for (int k = 0; k < VP_LM_MAXQ * VP_LM_MAXQ; ++k) ev.H[k] = 0.0;
            double Mm[P > 0 ? P : 1][P > 0 ? P : 1];
            {
                int t = 1 + P;
#pragma unroll
                for (int e = 0; e < P; ++e)
#pragma unroll
                    for (int f2 = e; f2 < P; ++f2) { Mm[e][f2] = tl.tv[t]; Mm[f2][e] = tl.tv[t]; ++t; }
            }
#pragma unroll
            for (int e = 0; e < P; ++e) {
                double ce = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) ce = (a.md.e_basis[e] == r) ? coef[r] : ce;
                const int ke = a.md.e_param[e];
                ev.g[ke] -= ce * tl.tv[1 + e];
#pragma unroll
                for (int f2 = 0; f2 < P; ++f2) {
                    double cf = 0.0;
#pragma unroll
                    for (int r = 0; r < N; ++r) cf = (a.md.e_basis[f2] == r) ? coef[r] : cf;
                    ev.H[a.md.e_param[f2] * q + ke] += Mm[e][f2] * ce * cf;
                }
            }
            for (int k = 0; k < q; ++k) finite = finite && isfinite(ev.g[k]);
            for (int k = 0; k < q * q; ++k) finite = finite && isfinite(ev.H[k]);
            ev.finite = (resid_ok ? VP_EVAL_RESIDUAL_OK : 0) | (finite ? VP_EVAL_DERIVS_OK : 0);
            LmState &st = st_s[g];
            const bool more = lm_advance(st, a.cfg, ev);
            if (st.last_accepted) {
#pragma unroll
                for (int r = 0; r < N; ++r) coef_acc[g][r] = coef[r];
            }
            if (!more) { // results of this problem; the slot is refilled in the next round
                const long long prob = prob_s[g];
                for (int k = 0; k < q; ++k) a.alpha_out[(size_t)prob * q + k] = st.x[k];
                for (int r = 0; r < N; ++r) a.C_out[(size_t)prob * N + r] = coef_acc[g][r];
                a.obj_out[prob] = 0.5 * st.fnorm * st.fnorm;
                a.term_out[prob] = st.termination;
                a.nfev_out[prob] = st.nfev;
                prob_s[g] = -1;
            }
        }
        __syncthreads();
    }
}

} // namespace vp
