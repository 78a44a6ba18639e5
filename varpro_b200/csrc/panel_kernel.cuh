// panel_kernel.cuh -- K1: the S-independent "panel" of one evaluation.
//
// Replaces, for one parameter vector alpha:
//   model.eval() / eval_partial_deriv(k)          src/model/mod.rs:441-512
//   Phi_w = W*Phi, D_k = W*dPhi/dalpha_k          src/solvers/levmar/mod.rs:47,141
//   thin SVD of Phi_w (U used as projector)       src/solvers/levmar/mod.rs:51,123-124
// by a thin QR  Phi_w = Q R1  (Q spans the same space as the reference's U
// whenever Phi_w has full numerical rank, so Q Q^T = U U^T), plus
//   E_e = (I - Q Q^T) d_e   for every non-zero derivative column d_e,
//   M   = E^T E (p x p),  R1^-1 (n x n).
// One CTA; the whole m x (n+p) panel lives in shared memory in fp64 regardless
// of the problem dtype; Q and E are written to HBM in the problem dtype for the
// streaming kernel.
//
// Orthogonalisation: classical Gram-Schmidt applied twice (CGS2), which gives
// orthogonality at machine precision for numerically non-singular panels and
// needs only one multi-value block reduction per pass.
// Rank policy (rank_policy.cuh): the reference truncates singular values <= eps in the solve
// (src/solvers/levmar/mod.rs:52-54) and keeps the untruncated U in the projector (:123-124); MATLAB's
// relative rule is the alternative. Both are decided on the n x n triangle R1 -- a cheap bound proves the
// full-rank case, otherwise one thread computes the tiny SVD and Q is re-expressed in the singular basis
// (after E has been formed with the full Q). A column that is EXACTLY zero when it is normalised gives q_j = 0.
//
// Implementation note: the kernel runs for microseconds on ONE SM, so its cost
// is dominated by instruction fetch of straight-line code, not by arithmetic
// (measured: a fully unrolled first version spent 58 us on 6.3k SASS
// instructions executed once). The work is therefore expressed as a short
// table of "rounds" -- each round = up to 8 column-pair dot products, one block
// reduction, one post-operation -- interpreted by a single compact loop.
#pragma once

#include "device_common.cuh"
#include "rank_policy.cuh"

namespace vp {

struct PanelSmall {
    double Rinv[VP_MAX_N * VP_MAX_N]; // column-major n x n (upper triangular), ld = VP_MAX_N
    double Rm[VP_MAX_N * VP_MAX_N];   // R1, column-major, ld = VP_MAX_N
    double M[VP_MAX_P * VP_MAX_P];    // E^T E, column-major, ld = VP_MAX_P
    double alpha[VP_MAX_Q];           // parameters this panel was built at
    int nonfinite;                    // 1 if Phi_w or D had a non-finite entry
    int dropped;                      // bit j set: column j dropped (rank policy)
};

constexpr int PANEL_THREADS = 512; // (one CTA; the rounds are latency-bound: 512 threads measured 1.3x faster than 256 at m = 1024)
constexpr int PANEL_MAX_ROUNDS = 3 * VP_MAX_N + 2 * VP_MAX_P + (VP_MAX_P * (VP_MAX_P + 1) / 2 + 7) / 8 + 2;

enum { ROUND_PROJECT = 0, ROUND_NORMALISE = 1, ROUND_GRAM = 2 };

struct PanelRound {
    short type, target, cnt, pad;
    short ia[8], ib[8];
};

// Value and the partial derivatives w.r.t. (up to) two parameters of a built-in
// basis function at x. One call site for exp and one for sincos keeps the code
// small, and the transcendental is shared between value and derivatives.
// Formulas as the reference writes them.
struct BasisVals { double v, d0, d1; };
static __device__ __noinline__ BasisVals basis_eval_all(int kind, double x, double a0, double a1, double scale)
{
    double earg = 0.0, targ = 0.0;
    // exp(-x/tau), d/dtau = exp(-x/tau) x/tau^2 with ONE division per basis function: t = x * (1/tau) (<= 1 ulp from x/tau in the exponent: a relative 1e-15 in the value; every device evaluator uses this form)
    const double inv0 = (kind == VP_BASIS_EXP_DECAY) ? 1.0 / a0 : 0.0;
    const double t0 = x * inv0;
    if (kind == VP_BASIS_EXP_DECAY) earg = -t0;
    else if (kind == VP_BASIS_EXP_RATE_COS) { earg = -a0 * x; targ = a1 * x; }
    else if (kind == VP_BASIS_SIN_PHASE) targ = a0 * x + a1;
    double e = 1.0, sn = 0.0, cs = 1.0;
    if (kind == VP_BASIS_EXP_DECAY || kind == VP_BASIS_EXP_RATE_COS) e = vp_exp(earg);
    if (kind == VP_BASIS_EXP_RATE_COS || kind == VP_BASIS_SIN_PHASE) sincos(targ, &sn, &cs);
    BasisVals r;
    switch (kind) {
    case VP_BASIS_EXP_DECAY: // exp(-x/tau); exp(-x/tau)*x/(tau*tau)  shared_test_code/src/lib.rs:101-114
        r.v = e; r.d0 = e * t0 * inv0; r.d1 = 0.0; break;
    case VP_BASIS_CONSTANT: // lib.rs:123
        r.v = 1.0; r.d0 = 0.0; r.d1 = 0.0; break;
    case VP_BASIS_EXP_RATE_COS: // shared_test_code/src/models.rs:321-322,362-385
        r.v = e * cs; r.d0 = -x * (e * cs); r.d1 = -x * e * sn; break;
    case VP_BASIS_SIN_PHASE: // src/test_helpers/mod.rs:27-51
        r.v = sn; r.d0 = x * cs; r.d1 = cs; break;
    case VP_BASIS_LINEAR_X: // src/model/builder/test.rs:97,101
        r.v = scale * x; r.d0 = 0.0; r.d1 = 0.0; break;
    default:
        r.v = nan(""); r.d0 = r.v; r.d1 = r.v;
    }
    return r;
}

// Sum the first cnt (<= 8) per-thread values over the CTA; all threads get the
// totals (entries >= cnt are unspecified). scratch: (nwarps + 1) * 8 doubles.
__device__ __forceinline__ void block_sum8(double (&v)[8], double *scratch, int cnt = 8)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < cnt) {
            double t = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) scratch[warp * 8 + k] = t;
        }
    }
    __syncthreads();
    if (warp == 0) {
        // lane = 8*g + k handles value k for warps g, g+4, g+8, ...
        const int k = lane & 7, g = lane >> 3;
        double t = 0.0;
        for (int w = g; w < nw; w += 4) t += scratch[w * 8 + k];
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        if (lane < 8) scratch[nw * 8 + lane] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = scratch[nw * 8 + k];
}

template <typename T>
__global__ void __launch_bounds__(PANEL_THREADS)
panel_kernel(ModelDesc md, const T *__restrict__ x, const T *__restrict__ w,
             const double *__restrict__ alpha_dev, double svd_eps, int ldp, T *__restrict__ Pq,
             PanelSmall *__restrict__ small, unsigned long long *dbg, const double *__restrict__ pre)
{
    // pre != nullptr: host-evaluated model -- the unweighted [Phi | D] (m x (n+p), f64) was uploaded
    extern __shared__ __align__(16) double psm[];
    const int m = md.m, n = md.n, p = md.p;
    double *col = psm;                           // (n+p) columns of m doubles
    double *scratch = psm + (size_t)(n + p) * m; // block_sum8 scratch
    __shared__ double Rm_s[VP_MAX_N * VP_MAX_N];
    __shared__ double alpha_s[VP_MAX_Q];
    __shared__ PanelRound rounds[PANEL_MAX_ROUNDS];
    __shared__ SmallSvd svd_s;
    __shared__ double Ri_s[VP_MAX_N * VP_MAX_N];
    __shared__ int nrounds_s, dropped_s;
    const int tid = threadIdx.x, nt = blockDim.x;

    dbg_mark(dbg, 0);
    if (tid < VP_MAX_Q) alpha_s[tid] = tid < md.q ? alpha_dev[tid] : 0.0;
    if (tid < VP_MAX_N * VP_MAX_N) Rm_s[tid] = 0.0;
    if (tid == 0) {
        // build the round table: CGS2 on the n basis columns, double projection
        // of the p derivative columns, then the Gram matrix of E
        int r = 0;
        for (int j = 0; j < n; ++j) {
            if (j > 0)
                for (int pass = 0; pass < 2; ++pass) {
                    PanelRound &R = rounds[r++];
                    R.type = ROUND_PROJECT; R.target = (short)j; R.cnt = (short)j;
                    for (int t = 0; t < 8; ++t) { R.ia[t] = (short)(t < j ? t : 0); R.ib[t] = (short)j; }
                }
            PanelRound &R = rounds[r++];
            R.type = ROUND_NORMALISE; R.target = (short)j; R.cnt = 1;
            for (int t = 0; t < 8; ++t) { R.ia[t] = (short)j; R.ib[t] = (short)j; }
        }
        for (int e = 0; e < p; ++e)
            for (int pass = 0; pass < 2; ++pass) {
                PanelRound &R = rounds[r++];
                R.type = ROUND_PROJECT; R.target = (short)(n + e); R.cnt = (short)n;
                for (int t = 0; t < 8; ++t) { R.ia[t] = (short)(t < n ? t : 0); R.ib[t] = (short)(n + e); }
            }
        int a = 0, b = 0; // upper triangle (a <= b) of M, eight entries per round
        const int npairs = p * (p + 1) / 2;
        for (int base = 0; base < npairs; base += 8) {
            PanelRound &R = rounds[r++];
            R.type = ROUND_GRAM; R.target = (short)base; R.cnt = (short)min(8, npairs - base);
            for (int t = 0; t < 8; ++t) {
                R.ia[t] = (short)(n + a); R.ib[t] = (short)(n + b);
                if (base + t < npairs) { if (++b == p) { ++a; b = a; } }
            }
        }
        nrounds_s = r;
        dropped_s = 0;
    }
    __syncthreads();

    // 1. evaluate the weighted basis functions and derivative columns
    int bad = 0;
    if (pre) {
        for (int idx = tid; idx < m * (n + p); idx += nt) {
            const int i = idx % m;
            const double v = (w ? (double)w[i] : 1.0) * pre[idx];
            if (idx < m * n) bad |= (!isfinite(v) ? 1 : 0) | (fabs(v) > RANK_HUGE_ENTRY ? (2 << (idx / m)) : 0);
            col[idx] = v;
        }
    }
    for (int i = tid; i < m && !pre; i += nt) {
        const double xi = (double)x[i];
        const double wi = w ? (double)w[i] : 1.0;
        int e = 0; // derivative columns are ordered by (basis function, slot)
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const int np = md.npar[j];
            const double a0 = np > 0 ? alpha_s[md.pidx[j][0]] : 0.0;
            const double a1 = np > 1 ? alpha_s[md.pidx[j][1]] : 0.0;
            const BasisVals bv = basis_eval_all(md.kind[j], xi, a0, a1, md.scale[j]);
            const double v = wi * bv.v;
            bad |= (!isfinite(v) ? 1 : 0) | (fabs(v) > RANK_HUGE_ENTRY ? (2 << j) : 0); // flag word of rank_policy.cuh
            col[(size_t)j * m + i] = v;
            if (np > 0) { col[(size_t)(n + e) * m + i] = wi * bv.d0; ++e; }
            if (np > 1) { col[(size_t)(n + e) * m + i] = wi * bv.d1; ++e; }
        }
    }
    bad = block_or_flags(bad, 1 + n);
    if (bad >> 1) { // overflowing basis columns: zero them and their derivative columns
        for (int idx = tid; idx < m * (n + p); idx += nt) {
            const int c = idx / m;
            const int j = c < n ? c : md.e_basis[c - n];
            if ((bad >> (1 + j)) & 1) col[idx] = 0.0;
        }
        __syncthreads();
    }
    bad &= 1;
    dbg_mark(dbg, 1);

    // 2.-4. the rounds
    const int nrounds = nrounds_s;
#pragma unroll 1
    for (int r = 0; r < nrounds; ++r) {
        const PanelRound &R = rounds[r];
        const int cnt = R.cnt, type = R.type, tgt = R.target;
        double d[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) d[t] = 0.0;
        for (int i = tid; i < m; i += nt) {
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (t < cnt) d[t] += col[(size_t)R.ia[t] * m + i] * col[(size_t)R.ib[t] * m + i];
        }
        block_sum8(d, scratch, cnt);
        if (type == ROUND_PROJECT) {
            double *aj = col + (size_t)tgt * m;
            for (int i = tid; i < m; i += nt) {
                double a = aj[i];
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t < cnt) a -= d[t] * col[(size_t)t * m + i];
                aj[i] = a;
            }
            if (tid < cnt && tgt < n) Rm_s[tgt * VP_MAX_N + tid] += scratch[((nt + 31) >> 5) * 8 + tid];
        } else if (type == ROUND_NORMALISE) {
            const double nrm = sqrt(d[0]);
            const bool keep = isfinite(nrm) && nrm > 0.0; // near-dependence is the rank policy's business
            double *aj = col + (size_t)tgt * m;
            for (int i = tid; i < m; i += nt) aj[i] = keep ? aj[i] / nrm : 0.0;
            if (tid == 0) {
                Rm_s[tgt * VP_MAX_N + tgt] = nrm;
                if (!keep) dropped_s |= 1 << tgt;
            }
        } else { // ROUND_GRAM
            if (tid < cnt) {
                const int a = R.ia[tid] - n, b = R.ib[tid] - n;
                const double v = scratch[((nt + 31) >> 5) * 8 + tid];
                small->M[b * VP_MAX_P + a] = v;
                small->M[a * VP_MAX_P + b] = v;
            }
        }
        __syncthreads();
    }

    dbg_mark(dbg, 2);
    // 5. R1^-1 (back substitution; thread c solves R1 x = e_c), then the rank policy on R1
    const int dropped = dropped_s;
    if (tid < VP_MAX_N * VP_MAX_N) Ri_s[tid] = 0.0;
    __syncthreads();
    if (tid < n) {
        const int c = tid;
        double xcol[VP_MAX_N];
#pragma unroll
        for (int i = 0; i < VP_MAX_N; ++i) xcol[i] = 0.0;
#pragma unroll
        for (int i = VP_MAX_N - 1; i >= 0; --i) {
            if (i > c) continue;
            double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < VP_MAX_N; ++k)
                if (k > i && k <= c) s -= Rm_s[k * VP_MAX_N + i] * xcol[k];
            xcol[i] = s / Rm_s[i * VP_MAX_N + i];
        }
#pragma unroll
        for (int i = 0; i < VP_MAX_N; ++i) Ri_s[c * VP_MAX_N + i] = xcol[i];
    }
    __syncthreads();
    if (tid == 0) {
        svd_s.truncated = 0;
        if (!rank_surely_full(n, Rm_s, VP_MAX_N, Ri_s, VP_MAX_N, svd_eps)) rank_policy_svd(n, Rm_s, VP_MAX_N, svd_eps, &svd_s);
    }
    __syncthreads();
    const int truncated = svd_s.truncated;
    if (truncated) { // rare: Q'' = Q Ur diag(keep) (E was formed with the full Q), Rinv = V diag(keep / sigma)
        for (int i = tid; i < m; i += nt) {
            double qo[VP_MAX_N], qn[VP_MAX_N];
            for (int k = 0; k < n; ++k) qo[k] = col[(size_t)k * m + i];
            for (int c = 0; c < n; ++c) {
                double sacc = 0.0;
                for (int k = 0; k < n; ++k) sacc = fma(qo[k], svd_s.Urot[c * n + k], sacc);
                qn[c] = sacc;
            }
            for (int c = 0; c < n; ++c) col[(size_t)c * m + i] = qn[c];
        }
        if (tid < n * n) Ri_s[(tid / n) * VP_MAX_N + (tid % n)] = svd_s.RinvEff[tid];
        __syncthreads();
    }
    if (tid < VP_MAX_N * VP_MAX_N) small->Rinv[tid] = Ri_s[tid];
    if (tid < VP_MAX_N * VP_MAX_N) small->Rm[tid] = Rm_s[tid];
    if (tid < VP_MAX_Q) small->alpha[tid] = alpha_s[tid];
    if (tid == 0) {
        small->nonfinite = bad ? 1 : 0;
        small->dropped = dropped | (truncated ? (1 << 30) : 0);
    }

    dbg_mark(dbg, 3);
    // 6. publish [Q | E | 0] in the problem dtype: n+p+1 columns of ldp rows, rows >= m zero
    for (int i = tid; i < ldp; i += nt) {
#pragma unroll 1
        for (int c = 0; c <= n + p; ++c)
            Pq[(size_t)c * ldp + i] = (i < m && c < n + p) ? (T)col[(size_t)c * m + i] : (T)0;
    }
    dbg_mark(dbg, 4);
}

} // namespace vp
