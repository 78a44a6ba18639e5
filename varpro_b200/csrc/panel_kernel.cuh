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
// Rank policy: a column whose remaining norm is <= svd_eps (the reference's
// absolute singular-value threshold, src/problem/builder.rs:246-251,282) is
// dropped: q_j = 0 and its coefficient is 0. The reference truncates sigma_i <=
// eps in the solve instead (src/solvers/levmar/mod.rs:52-54); the two agree on
// every full-rank panel; behaviour on exactly rank-deficient panels is not
// pinned by any reference test (SURVEY.md 8c "unpinned" (v)).
#pragma once

#include "device_common.cuh"

namespace vp {

struct PanelSmall {
    double Rinv[VP_MAX_N * VP_MAX_N]; // column-major n x n (upper triangular)
    double Rm[VP_MAX_N * VP_MAX_N];   // R1, column-major n x n
    double M[VP_MAX_P * VP_MAX_P];    // E^T E, column-major p x p
    double alpha[VP_MAX_Q];           // parameters this panel was built at
    int nonfinite;                    // 1 if Phi_w or D had a non-finite entry
    int dropped;                      // bit j set: column j dropped (rank policy)
};

constexpr int PANEL_THREADS = 1024;

template <typename T>
__global__ void __launch_bounds__(PANEL_THREADS)
panel_kernel(ModelDesc md, const T *__restrict__ x, const T *__restrict__ w,
             const double *__restrict__ alpha_dev, double svd_eps, int ld, T *__restrict__ Pq,
             T *__restrict__ Pe, PanelSmall *__restrict__ small)
{
    extern __shared__ __align__(16) double psm[];
    const int m = md.m, n = md.n, p = md.p;
    double *col = psm;                            // (n+p) columns of m doubles
    double *scratch = psm + (size_t)(n + p) * m;  // block_sum scratch
    __shared__ double Rm_s[VP_MAX_N * VP_MAX_N];
    __shared__ double alpha_s[VP_MAX_Q];
    const int tid = threadIdx.x, nt = blockDim.x;

    if (tid < VP_MAX_Q) alpha_s[tid] = tid < md.q ? alpha_dev[tid] : 0.0;
    if (tid < VP_MAX_N * VP_MAX_N) Rm_s[tid] = 0.0;
    __syncthreads();

    // 1. evaluate weighted basis functions and derivative columns
    int bad = 0;
    for (int i = tid; i < m; i += nt) {
        const double xi = (double)x[i];
        const double wi = w ? (double)w[i] : 1.0;
        for (int j = 0; j < n; ++j) {
            double a[VP_MAX_BASIS_PARAMS];
#pragma unroll
            for (int s = 0; s < VP_MAX_BASIS_PARAMS; ++s) a[s] = s < md.npar[j] ? alpha_s[md.pidx[j][s]] : 0.0;
            double v = wi * basis_value(md.kind[j], xi, a, md.scale[j]);
            bad |= !isfinite(v);
            col[(size_t)j * m + i] = v;
        }
        for (int e = 0; e < p; ++e) {
            const int j = md.e_basis[e];
            double a[VP_MAX_BASIS_PARAMS];
#pragma unroll
            for (int s = 0; s < VP_MAX_BASIS_PARAMS; ++s) a[s] = s < md.npar[j] ? alpha_s[md.pidx[j][s]] : 0.0;
            double v = wi * basis_deriv(md.kind[j], md.e_slot[e], xi, a);
            bad |= !isfinite(v);
            col[(size_t)(n + e) * m + i] = v;
        }
    }
    bad = __syncthreads_or(bad);

    // 2. CGS2 thin QR of the first n columns
    int dropped = 0;
    for (int j = 0; j < n; ++j) {
        double *aj = col + (size_t)j * m;
        if (j > 0) {
            for (int pass = 0; pass < 2; ++pass) {
                double d[VP_MAX_N];
#pragma unroll
                for (int k = 0; k < VP_MAX_N; ++k) d[k] = 0.0;
                for (int i = tid; i < m; i += nt) {
                    const double a = aj[i];
#pragma unroll
                    for (int k = 0; k < VP_MAX_N; ++k)
                        if (k < j) d[k] += col[(size_t)k * m + i] * a;
                }
                block_sum<VP_MAX_N>(d, scratch);
                for (int i = tid; i < m; i += nt) {
                    double a = aj[i];
#pragma unroll
                    for (int k = 0; k < VP_MAX_N; ++k)
                        if (k < j) a -= d[k] * col[(size_t)k * m + i];
                    aj[i] = a;
                }
                if (tid == 0)
                    for (int k = 0; k < j; ++k) Rm_s[j * VP_MAX_N + k] += d[k];
                __syncthreads();
            }
        }
        double s2[1] = {0.0};
        for (int i = tid; i < m; i += nt) s2[0] += aj[i] * aj[i];
        block_sum<1>(s2, scratch);
        const double nrm = sqrt(s2[0]);
        const bool keep = isfinite(nrm) && nrm > svd_eps;
        if (!keep) dropped |= 1 << j;
        if (tid == 0) Rm_s[j * VP_MAX_N + j] = nrm;
        for (int i = tid; i < m; i += nt) aj[i] = keep ? aj[i] / nrm : 0.0;
        __syncthreads();
    }

    // 3. E_e = (I - Q Q^T) d_e, projected twice
    for (int e = 0; e < p; ++e) {
        double *de = col + (size_t)(n + e) * m;
        for (int pass = 0; pass < 2; ++pass) {
            double d[VP_MAX_N];
#pragma unroll
            for (int k = 0; k < VP_MAX_N; ++k) d[k] = 0.0;
            for (int i = tid; i < m; i += nt) {
                const double a = de[i];
#pragma unroll
                for (int k = 0; k < VP_MAX_N; ++k)
                    if (k < n) d[k] += col[(size_t)k * m + i] * a;
            }
            block_sum<VP_MAX_N>(d, scratch);
            for (int i = tid; i < m; i += nt) {
                double a = de[i];
#pragma unroll
                for (int k = 0; k < VP_MAX_N; ++k)
                    if (k < n) a -= d[k] * col[(size_t)k * m + i];
                de[i] = a;
            }
            __syncthreads();
        }
    }

    // 4. M = E^T E, eight entries of the upper triangle per reduction round
    {
        const int npairs = p * (p + 1) / 2;
        for (int base = 0; base < npairs; base += 8) {
            double d[8];
            int ea[8], eb[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                d[t] = 0.0;
                // unpack linear index -> (a<=b)
                int idx = base + t, a = 0, rem = idx;
                while (a < p && rem >= p - a) { rem -= p - a; ++a; }
                ea[t] = a;
                eb[t] = a + rem;
            }
            for (int i = tid; i < m; i += nt) {
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (base + t < npairs)
                        d[t] += col[(size_t)(n + ea[t]) * m + i] * col[(size_t)(n + eb[t]) * m + i];
            }
            block_sum<8>(d, scratch);
            if (tid == 0) {
                for (int t = 0; t < 8; ++t)
                    if (base + t < npairs) {
                        small->M[eb[t] * VP_MAX_P + ea[t]] = d[t];
                        small->M[ea[t] * VP_MAX_P + eb[t]] = d[t];
                    }
            }
        }
    }

    // 5. R1^-1 restricted to the kept columns (back substitution), bookkeeping
    if (tid == 0) {
        for (int c = 0; c < n; ++c) {
            double xcol[VP_MAX_N];
            for (int i = 0; i < n; ++i) xcol[i] = 0.0;
            if (!((dropped >> c) & 1)) {
                for (int i = c; i >= 0; --i) {
                    if ((dropped >> i) & 1) continue;
                    double s = (i == c) ? 1.0 : 0.0;
                    for (int k = i + 1; k <= c; ++k) s -= Rm_s[k * VP_MAX_N + i] * xcol[k];
                    xcol[i] = s / Rm_s[i * VP_MAX_N + i];
                }
            }
            for (int i = 0; i < n; ++i) small->Rinv[c * VP_MAX_N + i] = xcol[i];
        }
        for (int i = 0; i < VP_MAX_N * VP_MAX_N; ++i) small->Rm[i] = Rm_s[i];
        for (int k = 0; k < VP_MAX_Q; ++k) small->alpha[k] = alpha_s[k];
        small->nonfinite = bad ? 1 : 0;
        small->dropped = dropped;
    }

    // 6. publish Q and E in the problem dtype, zero-padded to ld rows
    for (int i = tid; i < ld; i += nt) {
        for (int j = 0; j < n; ++j) Pq[(size_t)j * ld + i] = i < m ? (T)col[(size_t)j * m + i] : (T)0;
        for (int e = 0; e < p; ++e) Pe[(size_t)e * ld + i] = i < m ? (T)col[(size_t)(n + e) * m + i] : (T)0;
    }
}

} // namespace vp
