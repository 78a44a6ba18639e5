// device_common.cuh -- shared device-side definitions: model descriptor, basis
// evaluation, mbarrier / bulk-copy (TMA) PTX wrappers, block reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/varpro_b200.h"

namespace vp {

// Device-side model description: the table of built-in basis kinds plus the
// parameter-index map (reference: create_index_mapping, src/model/detail.rs:60-78)
// and the list of non-zero derivative columns (the reference stores these as a
// HashMap per basis function, src/model/mod.rs:497-510; MATLAB varpro calls it
// `Ind`, matlab/varpro.m:147-189).
struct ModelDesc {
    int m, n, q, p;
    int kind[VP_MAX_N];
    int npar[VP_MAX_N];
    int pidx[VP_MAX_N][VP_MAX_BASIS_PARAMS];
    double scale[VP_MAX_N];
    int e_basis[VP_MAX_P]; // basis function j(e) of derivative column e
    int e_slot[VP_MAX_P];  // which of that function's parameters
    int e_param[VP_MAX_P]; // model parameter index k(e)
};

// Layout of the per-evaluation reduction vector produced by the streaming pass:
// [0]                 sum_s ||r_s||^2
// [1 .. 1+n(n+1)/2)   G = sum_s c_s c_s^T, upper triangle packed row-wise (i<=j)
// [.. +p)             V_e = sum_s c_{s,j(e)} * (E_e^T y_s)
__host__ __device__ inline int red_count(int n, int p) { return 1 + n * (n + 1) / 2 + p; }
__host__ __device__ inline int g_index(int n, int i, int j)
{ // i <= j
    return 1 + i * n - i * (i - 1) / 2 + (j - i);
}

// --- exp for the basis evaluators ----------------------------------------------
// exp(a) = 2^k * exp(r), k = rint(a / ln 2), r = a - k ln 2 (two-term ln 2), exp(r) = 1 + r + r^2 P9(r) with P9 the
// degree-9 interpolant of (e^r - 1 - r) / r^2 at the Chebyshev nodes of |r| <= ln2 / 2 (truncation 1.6e-17
// relative; coefficients generated with mpmath). The same scheme as the CUDA library's exp, but every constant
// is an operand from the constant bank: the library version re-materialises its 13 constants with ~25 moves
// per call once registers are tight (49 instructions per exp in batch_fit_kernel, 17 of them fp64; this
// one is ~20). Arguments outside |a| < 700 (overflow, underflow to denormals, NaN) take the library path.
#define VP_EXP_CONSTANTS                                                                              \
    1.4426950408889634,          /* 0: 1 / ln 2 */                                                     \
        6755399441055744.0,      /* 1: 1.5 * 2^52 (round-to-nearest-integer shifter) */                \
        -0.6931471805599453,     /* 2: -ln 2 (high) */                                                 \
        -2.3190468138462996e-17, /* 3: -ln 2 (low) */                                                  \
        2.5100424157005067e-08,  /* 4: P9 coefficients, highest degree first */                       \
        2.7620138719733994e-07, 2.7557268378684192e-06, 2.480152119021773e-05, 0.00019841269863105968, \
        0.0013888888917281794, 0.008333333333330051, 0.04166666666662399, 0.16666666666666669,         \
        0.5000000000000001
static __constant__ double VP_EXP_C[14] = {VP_EXP_CONSTANTS};
// The same table as a kernel parameter (constant bank 0): there every coefficient is a direct operand of its
// DFMA, with no load at all (a __constant__ array lives in bank 3 and costs one LDC per coefficient and exp
// once the registers are too few to keep the table resident). Filled by the host with vp_exp_table().
struct ExpTable { double c[14]; };
inline ExpTable vp_exp_table()
{
    const ExpTable t = {{VP_EXP_CONSTANTS}};
    return t;
}
template <typename TAB>
__device__ __forceinline__ double vp_exp_with(const double a, const TAB &C)
{
    const double kd = fma(a, C[0], C[1]);
    const double kf = kd - C[1];
    double r = fma(kf, C[2], a);
    r = fma(kf, C[3], r);
    double p = C[4];
#pragma unroll
    for (int i = 5; i < 14; ++i) p = fma(p, r, C[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double res = __hiloint2double(__double2hiint(p) + (__double2loint(kd) << 20), __double2loint(p));
    if (!(fabs(a) < 700.0)) return exp(a);
    return res;
}
__device__ __forceinline__ double vp_exp(const double a) { return vp_exp_with(a, VP_EXP_C); }
// NV independent arguments, written coefficient by coefficient: every constant is fetched once and feeds NV
// DFMAs of NV independent Horner chains (the row-by-row form reloads the table for every exp when registers
// are tight). Bitwise the same results as vp_exp_with.
template <int NV, typename TAB>
__device__ __forceinline__ void vp_exp_vec(const double (&a)[NV], double (&res)[NV], const TAB &C)
{
    double kd[NV], r[NV], p[NV];
    {
        const double c0 = C[0], c1 = C[1], c2 = C[2], c3 = C[3];
#pragma unroll
        for (int v = 0; v < NV; ++v) kd[v] = fma(a[v], c0, c1);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double kf = kd[v] - c1;
            r[v] = fma(kf, c3, fma(kf, c2, a[v]));
        }
    }
    {
        const double c4 = C[4];
#pragma unroll
        for (int v = 0; v < NV; ++v) p[v] = c4;
    }
#pragma unroll
    for (int i = 5; i < 14; ++i) {
        const double ci = C[i];
#pragma unroll
        for (int v = 0; v < NV; ++v) p[v] = fma(p[v], r[v], ci);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double pp = fma(p[v], r[v], 1.0);
        pp = fma(pp, r[v], 1.0);
        res[v] = __hiloint2double(__double2hiint(pp) + (__double2loint(kd[v]) << 20), __double2loint(pp));
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
        if (!(fabs(a[v]) < 700.0)) res[v] = exp(a[v]);
}

// --- mbarrier + bulk async copy (TMA, SASS: UBLKCP / SYNCS) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// --- optional in-kernel timeline (debug/profiling aid; no-op when dbg == nullptr) ---
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define VP_DBG_SLOTS 16
__device__ __forceinline__ void dbg_mark(unsigned long long *dbg, int slot)
{
    if (dbg && threadIdx.x == 0) dbg[(size_t)blockIdx.x * VP_DBG_SLOTS + slot] = global_timer_ns();
}

// --- reductions ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV per-thread values over the whole CTA; every thread receives the totals.
// scratch: at least (blockDim.x/32)*NV + NV doubles of shared memory.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = warp_sum(v[k]);
        if (lane == 0) scratch[warp * NV + k] = t;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = 0.0;
            for (int w = lane; w < nw; w += 32) t += scratch[w * NV + k];
            t = warp_sum(t);
            if (lane == 0) scratch[nw * NV + k] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = scratch[nw * NV + k];
    __syncthreads();
}

} // namespace vp
