// aux_kernels.cuh -- small helper kernels: weighting of Y at problem creation
// and the materialisers behind the trait-mirroring entry points
// (vp_residuals / vp_jacobian / vp_best_fit). The LM loop never uses the
// materialisers: it only needs the reductions produced by stream_kernel.
#pragma once

#include "device_common.cuh"
#include "panel_kernel.cuh"

namespace vp {

// Y_w = W * Y in place (reference: src/problem/builder.rs:307), and zero the
// padding rows i in [m, ld) so that they contribute nothing downstream.
template <typename T>
__global__ void weight_rows_kernel(T *__restrict__ Y, const T *__restrict__ w, int m, int ld, long long S)
{
    const long long total = S * (long long)ld;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % ld);
        if (i >= m)
            Y[idx] = (T)0;
        else if (w)
            Y[idx] = Y[idx] * w[i];
    }
}

// vec(R): r_s = y_s - Q (Q^T y_s)   (src/solvers/levmar/mod.rs:57-59, 91-95)
// one warp per column; out is dense m x S.
template <typename T>
__global__ void residuals_kernel(const T *__restrict__ Y, int ld, int ldp, int m, int S, int n,
                                 const T *__restrict__ Pq, T *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < S; s += gridDim.x * wpb) {
        const T *y = Y + (size_t)s * ld;
        double b[VP_MAX_N];
#pragma unroll
        for (int k = 0; k < VP_MAX_N; ++k) b[k] = 0.0;
        for (int i = lane; i < m; i += 32) {
            const double yi = (double)y[i];
#pragma unroll
            for (int k = 0; k < VP_MAX_N; ++k)
                if (k < n) b[k] += (double)Pq[(size_t)k * ldp + i] * yi;
        }
#pragma unroll
        for (int k = 0; k < VP_MAX_N; ++k) b[k] = warp_sum(b[k]);
        for (int i = lane; i < m; i += 32) {
            double r = (double)y[i];
#pragma unroll
            for (int k = 0; k < VP_MAX_N; ++k)
                if (k < n) r -= (double)Pq[(size_t)k * ldp + i] * b[k];
            out[(size_t)s * m + i] = (T)r;
        }
    }
}

// Kaufman Jacobian, column k, RHS s (src/solvers/levmar/mod.rs:101-201):
//   J[s*m+i, k] = -(P_perp D_k c_s)_i = -sum_{e in k} E_e[i] * C[j(e), s]
template <typename T>
__global__ void jacobian_kernel(int ldp, int m, int S, int n, int p, int q, const T *__restrict__ Pe,
                                const T *__restrict__ C, ModelDesc md, T *__restrict__ out)
{
    const long long total = (long long)m * S;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % m);
        const long long s = idx / m;
        for (int k = 0; k < q; ++k) {
            double acc = 0.0;
            for (int e = 0; e < p; ++e)
                if (md.e_param[e] == k)
                    acc -= (double)Pe[(size_t)e * ldp + i] * (double)C[(size_t)s * n + md.e_basis[e]];
            out[(size_t)k * total + idx] = (T)acc;
        }
    }
}

// unweighted Phi(alpha), m x n dense (model.eval(), src/model/mod.rs:441-471)
template <typename T>
__global__ void phi_kernel(ModelDesc md, const T *__restrict__ x, const double *__restrict__ alpha,
                           double *__restrict__ phi)
{
    const int m = md.m, n = md.n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < m * n; idx += gridDim.x * blockDim.x) {
        const int i = idx % m, j = idx / m;
        const double a0 = md.npar[j] > 0 ? alpha[md.pidx[j][0]] : 0.0;
        const double a1 = md.npar[j] > 1 ? alpha[md.pidx[j][1]] : 0.0;
        phi[idx] = basis_eval_all(md.kind[j], (double)x[i], a0, a1, md.scale[j]).v;
    }
}

// best_fit = Phi * C (src/fit.rs:55-59), dense m x S
template <typename T>
__global__ void best_fit_kernel(int m, int S, int n, const double *__restrict__ phi,
                                const T *__restrict__ C, T *__restrict__ out)
{
    const long long total = (long long)m * S;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % m);
        const long long s = idx / m;
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += phi[(size_t)j * m + i] * (double)C[(size_t)s * n + j];
        out[idx] = (T)acc;
    }
}

} // namespace vp
