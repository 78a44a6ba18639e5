// dmma_tile.cuh -- the per-tile arithmetic shared by the fused kernels (fit_kernel_dmma,
// fit_queue_kernel): one tile = DMMA_CT = 8 whole columns of the weighted observations, staged
// in shared memory by 1-D bulk copies (see stream_kernel_dmma.cuh for the layout and the two
// fragment access patterns).
//
// Per column y_s (reference: set_params / residuals / jacobian, src/solvers/levmar/mod.rs:42-201):
//   phase 1   [b_s ; u_s] = [Q | E]^T y_s            DMMA, reduced over the CTA's warps
//   solve     c_s = Rinv b_s  -> coefficient buffer;  G += c_s c_s^T,  V_e += c_{s,j(e)} u_{s,e},
//             U_ef += u_{s,e} u_{s,f}  (the extra term of the full Golub-Pereyra Jacobian,
//             matlab/varpro.m:696-731; the reference leaves it as a TODO at levmar/mod.rs:188-190)
//   phase 2   r_s = y_s - Q b_s (explicit residual), ||r_s||^2
//
// Both kernels call the SAME code on the SAME partition of the tiles into "parts" (canonical
// partition below), and fold a part's per-thread accumulators with the same instruction
// sequence, so vp_fit and vp_fit_many produce bitwise identical (||r||^2, g, H) -- and therefore
// identical LM iterates and evaluation counts -- for the same problem.
#pragma once

#include "stream_kernel_dmma.cuh"

namespace vp {

// Canonical partition of a problem's ntiles tiles into nparts contiguous parts: part i covers the
// tiles [part_first_tile(i), part_first_tile(i + 1)). One partial-sum row per part.
struct TilePartition {
    int nparts, base, rem; // the first `rem` parts have base + 1 tiles
};
__host__ __device__ inline TilePartition make_partition(int ntiles, int nparts)
{
    TilePartition tp;
    tp.nparts = nparts < 1 ? 1 : (nparts > ntiles ? ntiles : nparts);
    if (tp.nparts < 1) tp.nparts = 1;
    tp.base = ntiles / tp.nparts;
    tp.rem = ntiles % tp.nparts;
    return tp;
}
__host__ __device__ inline int part_first_tile(const TilePartition &tp, int i)
{
    return i * tp.base + (i < tp.rem ? i : tp.rem);
}

template <int N, int P>
struct TileAcc { // per-thread accumulators of one part (G, V, U live in the threads tid < CT)
    static constexpr int NG = N * (N + 1) / 2, NU = P * (P + 1) / 2;
    double rn2;
    double G[NG];
    double V[P > 0 ? P : 1];
    double U[NU > 0 ? NU : 1];
    __device__ __forceinline__ void clear()
    {
        rn2 = 0.0;
#pragma unroll
        for (int i = 0; i < NG; ++i) G[i] = 0.0;
#pragma unroll
        for (int i = 0; i < (P > 0 ? P : 1); ++i) V[i] = 0.0;
#pragma unroll
        for (int i = 0; i < (NU > 0 ? NU : 1); ++i) U[i] = 0.0;
    }
};

// number of values in a partial row of the fused kernels: [rn2 | G | V | U]
__host__ __device__ inline int fused_red_count(int n, int p) { return red_count(n, p) + p * (p + 1) / 2; }

// One tile. tp: the tile's stage in shared memory (element type TY: double, or float converted on
// load); a1 / a2: the warp's DMMA A fragments of [Q|E]^T and Q; rinv_s: Rinv (n x n, column-major,
// ld N; need not be triangular: the rank policy may replace R1^-1 by V Sigma^+); part / bu: shared
// scratch (NWARPS*64 and 64 doubles); between(): called by every thread between the two barriers
// (the callers refill the TMA ring and flush staged part rows there).
template <typename TY, int N, int P, int KSTEPS, int NWARPS, bool EXACT, typename Between>
__device__ __forceinline__ void dmma_tile(const TY *tp, const int lds, const int nc, const int col0,
                                          const double (&a1)[KSTEPS], const double (&a2)[KSTEPS / 2],
                                          const double *rinv_s, double *part, double *bu, TY *Cout,
                                          const int (&ebasis)[P > 0 ? P : 1], TileAcc<N, P> &acc, Between &&between)
{
    constexpr int RSTEPS = KSTEPS / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    // phase-2 column permutation: mma column n <-> tile column pi(n) = n/2 + 4*(n%2)
    const int pcol_b = (grp >> 1) + 4 * (grp & 1);
    const int pcol_c0 = tig, pcol_c1 = tig + 4;

    // phase 1: C(8 dots x 8 cols) += A1(8 x 4) * Y(4 rows x 8 cols) over the warp's rows
    {
        double c[4][2];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) c[ch][0] = c[ch][1] = 0.0;
        const TY *bp = tp + (size_t)grp * lds + 4 * (warp * KSTEPS) + tig;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            double b;
            if (EXACT) b = (double)bp[4 * ks];
            else b = (4 * (warp * KSTEPS + ks) + tig < lds) ? (double)bp[4 * ks] : 0.0;
            dmma_8x8x4(c[ks & 3][0], c[ks & 3][1], a1[ks], b);
        }
        const double s0 = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
        const double s1 = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
        *reinterpret_cast<double2 *>(&part[warp * 64 + grp * 8 + 2 * tig]) = make_double2(s0, s1);
    }
    __syncthreads(); // (A) every warp is past phase 2 of the previous tile
    between();
    if (tid < 64) {
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NWARPS; ++w2) s += part[w2 * 64 + tid];
        bu[tid] = s;
    }
    __syncthreads(); // (B) b_s, u_s of the 8 columns are complete

    // solve: c_s = Rinv b_s ; accumulate G, V and U (one thread per column)
    if (tid < nc) {
        double coef[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c2 = 0; c2 < N; ++c2) s += rinv_s[c2 * N + r] * bu[c2 * 8 + tid];
            coef[r] = s;
            Cout[(size_t)(col0 + tid) * N + r] = (TY)s;
        }
        int gi = 0;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c2 = r; c2 < N; ++c2) acc.G[gi++] += coef[r] * coef[c2];
        double u[P > 0 ? P : 1];
#pragma unroll
        for (int e2 = 0; e2 < P; ++e2) {
            double cj = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r) cj = (ebasis[e2] == r) ? coef[r] : cj;
            u[e2] = bu[(N + e2) * 8 + tid];
            acc.V[e2] += cj * u[e2];
        }
        int ui = 0;
#pragma unroll
        for (int e2 = 0; e2 < P; ++e2)
#pragma unroll
            for (int f2 = e2; f2 < P; ++f2) acc.U[ui++] += u[e2] * u[f2];
    }

    // phase 2: R(8 rows x 8 cols) = Y + Q(8 x 4) * (-b)(4 x 8); accumulate r^2
    {
        const double b2 = (tig < N) ? -bu[tig * 8 + pcol_b] : 0.0;
        const TY *cp0 = tp + (size_t)pcol_c0 * lds + 8 * (warp * RSTEPS) + grp;
        const TY *cp1 = tp + (size_t)pcol_c1 * lds + 8 * (warp * RSTEPS) + grp;
        double q0 = 0.0, q1 = 0.0;
#pragma unroll
        for (int rs = 0; rs < RSTEPS; ++rs) {
            double d0, d1;
            if (EXACT) { d0 = (double)cp0[8 * rs]; d1 = (double)cp1[8 * rs]; }
            else {
                const bool ok = 8 * (warp * RSTEPS + rs) + grp < lds;
                d0 = ok ? (double)cp0[8 * rs] : 0.0;
                d1 = ok ? (double)cp1[8 * rs] : 0.0;
            }
            dmma_8x8x4(d0, d1, a2[rs], b2);
            q0 = fma(d0, d0, q0);
            q1 = fma(d1, d1, q1);
        }
        acc.rn2 += (pcol_c0 < nc ? q0 : 0.0) + (pcol_c1 < nc ? q1 : 0.0);
    }
}

// Fold of one part's accumulators into its partial row, in two halves so that the second one can be
// deferred behind a barrier the caller has anyway:
//   part_stage : every thread; writes the warp sums of rn2 and the per-column G/V/U into a staging
//                buffer (wsum: NWARPS doubles, gv: CT*(NG+P+NU) doubles)
//   part_flush : after a __syncthreads following part_stage, the threads lane < NVR of ONE warp sum
//                the staging buffer in a fixed order and write the row.
template <int N, int P, int CT, int NWARPS>
__device__ __forceinline__ void part_stage(const TileAcc<N, P> &acc, double *wsum, double *gv)
{
    constexpr int NG = TileAcc<N, P>::NG, NU = TileAcc<N, P>::NU, NGP = NG + P + NU;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double t = warp_sum(acc.rn2);
    if (lane == 0) wsum[warp] = t;
    if (tid < CT) {
#pragma unroll
        for (int u = 0; u < NG; ++u) gv[tid * NGP + u] = acc.G[u];
#pragma unroll
        for (int e = 0; e < P; ++e) gv[tid * NGP + NG + e] = acc.V[e];
#pragma unroll
        for (int e = 0; e < NU; ++e) gv[tid * NGP + NG + P + e] = acc.U[e];
    }
}
template <int N, int P, int CT, int NWARPS>
__device__ __forceinline__ void part_flush(const double *wsum, const double *gv, double *row, const int lane)
{
    constexpr int NG = TileAcc<N, P>::NG, NU = TileAcc<N, P>::NU, NGP = NG + P + NU, NVR = 1 + NGP;
    static_assert(NVR <= 32, "the partial row must be written by one warp");
    if (lane < NVR) {
        double s = 0.0;
        if (lane == 0) {
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) s += wsum[w];
        } else {
#pragma unroll
            for (int c = 0; c < CT; ++c) s += gv[c * NGP + lane - 1];
        }
        row[lane] = s;
    }
}

} // namespace vp
