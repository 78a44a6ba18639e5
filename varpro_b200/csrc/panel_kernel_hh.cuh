// panel_kernel_hh.cuh -- K1, fast path: the panel of one evaluation by Householder QR with
// ONE block reduction per column, everything else in registers.
//
// Same outputs as panel_kernel (panel_kernel.cuh; see there for the mapping to the
// reference: model eval src/model/mod.rs:441-512, weighting and the projector
// src/solvers/levmar/mod.rs:47,51,123-124,141):  [Q | E | 0] in the problem dtype,
// R1^-1, M = E^T E, flags.
//
// Algorithm. The augmented matrix A = [Phi_w | D] (m x (n+p)) is reduced column by
// column. Step j needs only sums over rows i >= j with column j:
//     sigma = sum a_j[i]^2,   d_k = sum a_j[i] a_k[i]  (k > j)
// -- one fused block reduction -- because with alpha = -sign(a_jj) sqrt(sigma),
// v = a_j[j:] - alpha e_j:   v^T v = 2 (sigma - a_jj alpha),   v^T a_k = d_k - alpha a_k[j].
// After the n steps the rows >= n of the D columns hold the projected derivatives in
// the rotated basis, so M = sum_{i>=n} D~[i,e] D~[i,f]; it is reduced together with
// V^T V, which gives the compact-WY factor T (T^-1 = striu(V^T V) + diag(v_j^T v_j / 2)).
// Explicit Q = I - V T V^T[:, :n] and E = D - Q (Q^T D) (Q^T D = top n rows of D~) are
// then row-local. n + 1 reductions in total (the CGS2 interpreter needs 3n + 2p + 1).
// Householder QR is unconditionally backward stable; Q is orthonormal to working
// precision whatever the conditioning of Phi_w.
//
// Rank policy (rank_policy.cuh, identical in panel_kernel): the reference's rule -- singular values <= eps are
// truncated in the solve, the projector stays untruncated -- or MATLAB's relative rule, decided on the n x n
// triangle R1; a cheap bound proves the common full-rank case without an SVD. A column that is EXACTLY zero
// when its Householder step starts is skipped (q_j = 0).
#pragma once

#include "panel_kernel.cuh"
#include "rank_policy.cuh"

namespace vp {

// sum K per-thread values over the CTA with one __syncthreads; every thread gets the totals.
// buf: NW*K doubles, must not be reused by the next call (double-buffer across rounds).
template <int K, int NW>
__device__ __forceinline__ void block_sum_once(double (&v)[K], double *buf)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) buf[warp * K + k] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += buf[w * K + k];
        v[k] = t;
    }
}

// Householder step J (compile-time) on the register-resident working matrix, then recurse.
template <int J, int N, int P, int RPT, int THREADS>
__device__ __forceinline__ void hh_steps(double (&a)[RPT][N + P], double (&beta)[N], double (&rdiag)[N], int &dropped,
                                         double (*top)[N + P], double (*red)[(THREADS / 32) * ((N + P > 8) ? N + P : 8)],
                                         const double svd_eps)
{
    if constexpr (J < N) {
        constexpr int NPV = N + P;
        constexpr int NW = THREADS / 32;
        constexpr int K = NPV - J; // sigma and the dots with the NPV-J-1 columns to the right
        const int tid = threadIdx.x;
        double s[K];
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] = 0.0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = tid + r * THREADS;
            if (i >= J) {
                const double aj = a[r][J];
#pragma unroll
                for (int k = 0; k < K; ++k) s[k] = fma(aj, a[r][J + k], s[k]); // k = 0: sigma
            }
        }
        // row J of the working matrix is owned by thread J (r = 0): publish it
        if (tid == J) {
#pragma unroll
            for (int k = 0; k < NPV; ++k) top[J][k] = a[0][k];
        }
        block_sum_once<K, NW>(s, red[J & 1]);
        const double sigma = s[0];
        const double ajj = top[J][J];
        const double nrm = sqrt(sigma);
        const bool keep = isfinite(nrm) && nrm > 0.0; // near-dependence is the rank policy's business (singular values of R1)
        const double al = (ajj >= 0.0) ? -nrm : nrm;
        const double vnorm2 = 2.0 * (sigma - ajj * al);
        const double bt = (keep && vnorm2 > 0.0) ? 2.0 / vnorm2 : 0.0;
        beta[J] = bt;
        rdiag[J] = keep ? al : 0.0;
        if (!keep) dropped |= 1 << J;
        const double vjj = ajj - al; // v[J]
#pragma unroll
        for (int k = 1; k < K; ++k) {
            const double tau = bt * (s[k] - al * top[J][J + k]);
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int i = tid + r * THREADS;
                const double vi = (i > J) ? a[r][J] : ((i == J) ? vjj : 0.0);
                a[r][J + k] = fma(-tau, vi, a[r][J + k]);
            }
        }
        // column J now stores v_J (rows >= J); rows < J keep their R entries
        if (tid == J) a[0][J] = vjj;
        __syncthreads(); // everyone has read top[J]; its owner rewrites it with the final row
        if (tid == J) {
#pragma unroll
            for (int k = 0; k < NPV; ++k) top[J][k] = a[0][k]; // v_JJ, R[J][k>J], (Q^T D)[J][:]
        }
        hh_steps<J + 1, N, P, RPT, THREADS>(a, beta, rdiag, dropped, top, red, svd_eps);
    }
}

// Evaluate the weighted basis functions and derivative columns for this thread's rows
// (straight-line code, static register indices).
template <int N, int P, int RPT, int THREADS>
__device__ __forceinline__ int panel_hh_eval(const ModelDesc &md, const double (&xi_r)[RPT], const double (&wi_r)[RPT],
                                             const double *alpha_s, double (&a)[RPT][N + P], double (&d0)[RPT][P > 0 ? P : 1])
{
    const int tid = threadIdx.x;
    const int m = md.m;
    int bad = 0;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
        const bool in = i < m;
        const double xi = xi_r[r];
        const double wi = wi_r[r];
        int e = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int np = md.npar[j];
            const double a0 = np > 0 ? alpha_s[md.pidx[j][0]] : 0.0;
            const double a1 = np > 1 ? alpha_s[md.pidx[j][1]] : 0.0;
            const BasisVals bv = basis_eval_all(md.kind[j], xi, a0, a1, md.scale[j]);
            const double v = in ? wi * bv.v : 0.0;
            bad |= (!isfinite(v) ? 1 : 0) | (fabs(v) > RANK_HUGE_ENTRY ? (2 << j) : 0); // flag word of rank_policy.cuh
            a[r][j] = v;
            // derivative columns are ordered by (basis function, slot); e is uniform across threads
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s < np) {
                    const double dv = in ? wi * (s == 0 ? bv.d0 : bv.d1) : 0.0;
#pragma unroll
                    for (int ee = 0; ee < P; ++ee)
                        if (ee == e) { a[r][N + ee] = dv; d0[r][ee] = dv; }
                    ++e;
                }
            }
        }
    }
    return bad;
}


// The panel computation for one parameter vector, executed by all THREADS threads of a CTA.
// xi/wi: the thread's rows i = tid + r*THREADS of the independent variable and the weights
// (wi = 0 for i >= m). alpha_s: the q parameters (shared memory). Pq (ldp rows per column) and
// `small` may point to global OR shared memory (generic addressing): panel_kernel_hh publishes
// them to HBM for the streaming kernel, fit_kernel_dmma keeps them in the CTA.
// red/top: shared scratch. The caller synchronises before anybody reads Pq / small.
template <typename T, int N, int P, int RPT, int THREADS>
__device__ __forceinline__ void panel_hh_factor(const ModelDesc &md, double (&a)[RPT][N + P], double (&d0)[RPT][P > 0 ? P : 1],
                                                int bad, const double *alpha_s, const double svd_eps, const int ldp, T *Pq,
                                                PanelSmall *small,
                                                double (*red)[(THREADS / 32) * ((N + P > 8) ? N + P : 8)],
                                                double (*top)[N + P], unsigned long long *dbg, SmallSvd *svd_s)
{
    constexpr int NPV = N + P;
    constexpr int NW = THREADS / 32;
    constexpr int NVV = N * (N - 1) / 2; // strict upper triangle of V^T V
    constexpr int NMM = P * (P + 1) / 2; // upper triangle of M
    constexpr int KMAX = (NPV > 8) ? NPV : 8;
    const int tid = threadIdx.x;
    const int m = md.m;

    bad = block_or_flags(bad, 1 + N);
    if (bad >> 1) { // overflowing basis columns (rank_policy.cuh): zero them and their derivative columns
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
#pragma unroll
            for (int j = 0; j < N; ++j)
                if ((bad >> (1 + j)) & 1) a[r][j] = 0.0;
#pragma unroll
            for (int e = 0; e < P; ++e)
                if ((bad >> (1 + md.e_basis[e])) & 1) { a[r][N + e] = 0.0; d0[r][e] = 0.0; }
        }
    }
    bad &= 1;
    dbg_mark(dbg, 1);

    // ---- 2. Householder steps ------------------------------------------------------------
    double beta[N], rdiag[N];
    int dropped = 0;
    hh_steps<0, N, P, RPT, THREADS>(a, beta, rdiag, dropped, top, red, svd_eps);
    __syncthreads();

    // ---- 3. V^T V (strict upper) and M = D~[n:]^T D~[n:] in one or more reductions --------
    double vtv[NVV > 0 ? NVV : 1], Mm[NMM > 0 ? NMM : 1];
    {
        constexpr int TOT = NVV + NMM;
        double tot[TOT > 0 ? TOT : 1];
#pragma unroll
        for (int t = 0; t < TOT; ++t) tot[t] = 0.0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = tid + r * THREADS;
            int t = 0;
#pragma unroll
            for (int aa = 0; aa < N; ++aa)
#pragma unroll
                for (int bb = aa + 1; bb < N; ++bb) {
                    if (i >= bb) tot[t] = fma(a[r][aa], a[r][bb], tot[t]); // v_a, v_b both live on rows >= bb
                    ++t;
                }
            if (i >= N) {
#pragma unroll
                for (int e = 0; e < P; ++e)
#pragma unroll
                    for (int f = e; f < P; ++f) { tot[t] = fma(a[r][N + e], a[r][N + f], tot[t]); ++t; }
            }
        }
        if constexpr (TOT > 0) {
            if constexpr (TOT <= KMAX) {
                block_sum_once<TOT, NW>(tot, red[N & 1]);
            } else {
#pragma unroll
                for (int base = 0; base < TOT; base += KMAX) {
                    double s[KMAX];
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) s[k] = (base + k < TOT) ? tot[base + k] : 0.0;
                    if (base > 0) __syncthreads();
                    block_sum_once<KMAX, NW>(s, red[(N + base / KMAX) & 1]);
#pragma unroll
                    for (int k = 0; k < KMAX; ++k)
                        if (base + k < TOT) tot[base + k] = s[k];
                }
            }
        }
#pragma unroll
        for (int t = 0; t < NVV; ++t) vtv[t] = tot[t];
#pragma unroll
        for (int t = 0; t < NMM; ++t) Mm[t] = tot[NVV + t];
    }
    dbg_mark(dbg, 2);

    // ---- 4. small matrices, redundantly in every thread (registers, static indices) -------
    // Tinv = striu(V^T V) + diag(1/beta);  Tm = Tinv^-1 (upper triangular)
    double Tm[N][N];
    {
        double Ti[N][N];
        int t = 0;
#pragma unroll
        for (int aa = 0; aa < N; ++aa)
#pragma unroll
            for (int bb = 0; bb < N; ++bb) Ti[aa][bb] = 0.0;
#pragma unroll
        for (int aa = 0; aa < N; ++aa)
#pragma unroll
            for (int bb = aa + 1; bb < N; ++bb) Ti[aa][bb] = vtv[t++];
        // invert: columns c, back substitution; diagonal of Tm is beta
#pragma unroll
        for (int c = 0; c < N; ++c) {
#pragma unroll
            for (int i = N - 1; i >= 0; --i) {
                if (i > c) { Tm[i][c] = 0.0; continue; }
                double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = i + 1; k <= c; ++k) sacc -= Ti[i][k] * Tm[k][c];
                Tm[i][c] = sacc * beta[i]; // divide by Tinv[i][i] = 1/beta[i]
            }
        }
    }
    // W[a][c] = sum_b Tm[a][b] * V[c][b],  V[c][b] = v_b[c] (rows c >= b), from `top`
    double W[N][N];
#pragma unroll
    for (int aa = 0; aa < N; ++aa)
#pragma unroll
        for (int c = 0; c < N; ++c) {
            double sacc = 0.0;
#pragma unroll
            for (int bb = aa; bb < N; ++bb)
                if (c >= bb) sacc = fma(Tm[aa][bb], top[c][bb], sacc);
            W[aa][c] = sacc;
        }

    // ---- 4b. rank policy on R1 (thread 0; the others go on with Q and E and meet it at the next barrier) ----
    if (tid == 0) {
        double Rm[N * N], Ri[N * N]; // column-major, ld N: R1 and its triangular inverse
#pragma unroll
        for (int c = 0; c < N; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) { Rm[c * N + i] = (i < c) ? top[i][c] : ((i == c) ? rdiag[c] : 0.0); Ri[c * N + i] = 0.0; }
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
        svd_s->truncated = 0;
        if (!rank_surely_full(N, Rm, N, Ri, N, svd_eps)) rank_policy_svd(N, Rm, N, svd_eps, svd_s);
    }

    // ---- 5. explicit Q and E for my rows -------------------------------------------------------
    // E = D - Q (Q^T D) leaves Q^T E = O(eps) ||D||; a second projection pass (one more
    // reduction, n*p values) brings it down to O(eps) ||E||, which sets the noise floor of the
    // gradient J^T r near a zero-residual solution.
    double qrow[RPT][N], erow[RPT][P > 0 ? P : 1];
    {
        constexpr int NQE = N * P;
        double qe[NQE > 0 ? NQE : 1];
#pragma unroll
        for (int t = 0; t < NQE; ++t) qe[t] = 0.0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = tid + r * THREADS;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int aa = 0; aa < N; ++aa)
                    if (i >= aa) sacc = fma(-a[r][aa], W[aa][c], sacc); // V[i][aa] = a[r][aa] on rows >= aa
                qrow[r][c] = (i >= m) ? 0.0 : sacc; // (a skipped step has beta = 0: its column is H_0..H_{c-1} e_c, still orthonormal)
            }
#pragma unroll
            for (int e = 0; e < P; ++e) {
                double ev = d0[r][e];
#pragma unroll
                for (int c = 0; c < N; ++c) ev = fma(-qrow[r][c], top[c][N + e], ev); // (Q^T D)[c][e]
                erow[r][e] = (i < m) ? ev : 0.0;
#pragma unroll
                for (int c = 0; c < N; ++c) qe[c * P + e] = fma(qrow[r][c], erow[r][e], qe[c * P + e]);
            }
        }
        if constexpr (NQE > 0) {
            static_assert(NQE <= 32, "Q^T E reduction chunking not implemented beyond 32 values");
            __syncthreads();
            if constexpr (NQE <= KMAX) {
                block_sum_once<NQE, NW>(qe, red[0]);
            } else {
#pragma unroll
                for (int base = 0; base < NQE; base += KMAX) {
                    double s2[KMAX];
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) s2[k] = (base + k < NQE) ? qe[base + k] : 0.0;
                    if (base > 0) __syncthreads();
                    block_sum_once<KMAX, NW>(s2, red[(base / KMAX) & 1]);
#pragma unroll
                    for (int k = 0; k < KMAX; ++k)
                        if (base + k < NQE) qe[base + k] = s2[k];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int e = 0; e < P; ++e)
#pragma unroll
                for (int c = 0; c < N; ++c) erow[r][e] = fma(-qrow[r][c], qe[c * P + e], erow[r][e]);
    }
    if constexpr (N * P == 0) __syncthreads(); // (the Q^T E reduction above has the barrier otherwise)
    const int truncated = svd_s->truncated;
    if (truncated) { // rare: Q'' = Q Ur diag(keep); E above was formed with the full Q (untruncated projector)
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            double qn[N];
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) sacc = fma(qrow[r][k], svd_s->Urot[c * N + k], sacc);
                qn[c] = sacc;
            }
#pragma unroll
            for (int c = 0; c < N; ++c) qrow[r][c] = qn[c];
        }
    }
    // publish [Q | E | 0] in the problem dtype
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
        if (i < ldp) {
#pragma unroll
            for (int c = 0; c < N; ++c) Pq[(size_t)c * ldp + i] = (T)qrow[r][c];
#pragma unroll
            for (int e = 0; e < P; ++e) Pq[(size_t)(N + e) * ldp + i] = (T)erow[r][e];
            Pq[(size_t)NPV * ldp + i] = (T)0;
        }
    }
    // rows beyond RPT*THREADS up to ldp (zero padding of the panel)
    for (int i = tid + RPT * THREADS; i < ldp; i += THREADS) {
#pragma unroll
        for (int c = 0; c <= NPV; ++c) Pq[(size_t)c * ldp + i] = (T)0;
    }
    dbg_mark(dbg, 3);

    // ---- 6. small outputs -----------------------------------------------------------------
    if (tid < N) {
        // thread c solves R1 x = e_c restricted to the kept columns
        const int c = tid;
        double xcol[N];
#pragma unroll
        for (int i = 0; i < N; ++i) xcol[i] = 0.0;
        if (truncated) { // V diag(keep / sigma): a full matrix
#pragma unroll
            for (int i = 0; i < N; ++i) xcol[i] = svd_s->RinvEff[c * N + i];
        } else if (!((dropped >> c) & 1)) {
#pragma unroll
            for (int i = N - 1; i >= 0; --i) {
                if (i > c || ((dropped >> i) & 1)) continue;
                double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (k > i && k <= c) sacc -= top[i][k] * xcol[k];
                xcol[i] = sacc / rdiag[i];
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) small->Rinv[c * VP_MAX_N + i] = xcol[i];
#pragma unroll
        for (int i = 0; i < N; ++i) small->Rm[c * VP_MAX_N + i] = (i < c) ? top[i][c] : ((i == c) ? rdiag[c] : 0.0);
    }
    if (tid == 32) {
        int t = 0;
#pragma unroll
        for (int e = 0; e < P; ++e)
#pragma unroll
            for (int f = e; f < P; ++f) {
                small->M[f * VP_MAX_P + e] = Mm[t];
                small->M[e * VP_MAX_P + f] = Mm[t];
                ++t;
            }
        small->nonfinite = bad ? 1 : 0;
        small->dropped = dropped | (truncated ? (1 << 30) : 0);
    }
    if (tid >= 64 && tid < 64 + VP_MAX_Q) small->alpha[tid - 64] = alpha_s[tid - 64];
    dbg_mark(dbg, 4);
}

template <typename T, int N, int P, int RPT, int THREADS>
__device__ __forceinline__ void panel_hh_body(const ModelDesc &md, const double (&xi_r)[RPT], const double (&wi_r)[RPT],
                                              const double *alpha_s, const double svd_eps, const int ldp, T *Pq,
                                              PanelSmall *small,
                                              double (*red)[(THREADS / 32) * ((N + P > 8) ? N + P : 8)],
                                              double (*top)[N + P], unsigned long long *dbg, SmallSvd *svd_s)
{
    double a[RPT][N + P];           // working matrix rows
    double d0[RPT][P > 0 ? P : 1];  // the untouched weighted derivative columns
    const int bad = panel_hh_eval<N, P, RPT, THREADS>(md, xi_r, wi_r, alpha_s, a, d0);
    panel_hh_factor<T, N, P, RPT, THREADS>(md, a, d0, bad, alpha_s, svd_eps, ldp, Pq, small, red, top, dbg, svd_s);
}

// pre != nullptr: host-evaluated model -- the unweighted [Phi | D] (m x (n+p), f64, column-major) was uploaded and
// takes the place of the device basis evaluation (vp_model_create_hosteval).
template <typename T, int N, int P, int RPT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
panel_kernel_hh(ModelDesc md, const T *__restrict__ x, const T *__restrict__ w,
                const double *__restrict__ alpha_dev, double svd_eps, int ldp, T *__restrict__ Pq,
                PanelSmall *__restrict__ small, unsigned long long *dbg, const double *__restrict__ pre)
{
    constexpr int NPV = N + P;
    constexpr int NW = THREADS / 32;
    constexpr int KMAX = (NPV > 8) ? NPV : 8;
    __shared__ double red[2][NW * KMAX];
    __shared__ double top[N][NPV];   // rows 0..n-1 of the working matrix (R, Q^T D and v entries)
    __shared__ double alpha_s[VP_MAX_Q];
    __shared__ SmallSvd svd_s;
    const int tid = threadIdx.x;
    dbg_mark(dbg, 0);
    if (tid < VP_MAX_Q) alpha_s[tid] = tid < md.q ? alpha_dev[tid] : 0.0;
    double xi[RPT], wi[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
        const bool in = i < md.m;
        xi[r] = in ? (double)x[i] : 0.0;
        wi[r] = in ? (w ? (double)w[i] : 1.0) : 0.0;
    }
    __syncthreads();
    if (pre) {
        double a[RPT][NPV];
        double d0[RPT][P > 0 ? P : 1];
        int bad = 0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = tid + r * THREADS;
            const bool in = i < md.m;
#pragma unroll
            for (int k = 0; k < NPV; ++k) {
                const double v = in ? wi[r] * pre[(size_t)k * md.m + i] : 0.0;
                if (k < N) bad |= (!isfinite(v) ? 1 : 0) | (fabs(v) > RANK_HUGE_ENTRY ? (2 << k) : 0); // flag word of rank_policy.cuh
                a[r][k] = v;
            }
#pragma unroll
            for (int e = 0; e < P; ++e) d0[r][e] = a[r][N + e];
        }
        panel_hh_factor<T, N, P, RPT, THREADS>(md, a, d0, bad, alpha_s, svd_eps, ldp, Pq, small, red, top, dbg, &svd_s);
        return;
    }
    panel_hh_body<T, N, P, RPT, THREADS>(md, xi, wi, alpha_s, svd_eps, ldp, Pq, small, red, top, dbg, &svd_s);
}

} // namespace vp
