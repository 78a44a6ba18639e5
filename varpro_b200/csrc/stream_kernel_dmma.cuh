// stream_kernel_dmma.cuh -- K2 (fp64 fast path): the Y-streaming reduce with the
// two contractions on the fp64 tensor pipe (mma.sync m8n8k4 .f64, SASS DMMA).
//
// Same contract, inputs and outputs as stream_kernel (stream_kernel.cuh); see
// there for the mapping to the reference (src/solvers/levmar/mod.rs:42-201).
//
// Why tensor cores here: the SIMT version spends ~40% of its issue slots on the
// warp-shuffle fold of the (n+p) x CT partial dot products and runs at IPC ~1
// (ncu, profiles/r01_*). The per-tile work is exactly two small GEMMs,
//   phase 1:  [Q|E]^T (8 x m, n+p <= 8 rows used)  x  Y_tile (m x 8)   -> b, u
//   phase 2:  R_tile (m x 8) = Y_tile - Q (m x 4, n <= 4) x b (4 x 8)   -> ||r||^2
// and a DMMA reduces over k inside the instruction, so no shuffles are needed:
// ~4x fewer instructions per column. tcgen05 has no fp64 kind; DMMA is the
// fp64 tensor path on sm_100a (measured 37 TFLOP/s, profiles/r01_fp64_peaks.txt).
// The tensor pipe is not the bound -- HBM is; the DMMAs just make the kernel
// cheap enough in issue slots to keep up with the bulk copies.
//
// Layout: a tile is 8 whole columns. Each column is fetched by its own 1-D bulk
// async copy (TMA) into a padded shared-memory slot of `lds` doubles with
// lds = 4 (mod 16): with that stride both fragment load patterns (phase 1:
// 4 consecutive rows x 8 columns; phase 2: 8 consecutive rows x 4 columns, with
// the column permutation pi(n) = n/2 + 4*(n%2)) are bank-conflict free.
// The warp's slices of [Q|E]^T and -Q live in registers as A fragments for the
// whole kernel.
#pragma once

#include "stream_kernel.cuh"

namespace vp {

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}


// KSTEPS: k-steps (4 rows each) per warp; NWARPS warps cover 4*KSTEPS*NWARPS >= ld rows.
// EXACT: 4*KSTEPS*NWARPS <= lds, so no row predicate is needed on the fragment loads.
template <int N, int P, int KSTEPS, int NWARPS, bool EXACT>
__global__ void __launch_bounds__(NWARPS * 32, 1)
stream_kernel_dmma(const StreamArgs<double> a, const int lds)
{
    constexpr int NPV = N + P;
    constexpr int THREADS = NWARPS * 32;
    constexpr int CT = DMMA_CT;
    constexpr int RSTEPS = KSTEPS / 2; // 8-row steps of phase 2
    static_assert(NPV <= 8 && N <= 4, "one DMMA row block / k block only");
    static_assert(KSTEPS % 4 == 0, "four accumulator chains");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STREAM_MAX_STAGES];
    __shared__ __align__(16) double part[NWARPS * 64];
    __shared__ __align__(16) double bu[64]; // bu[dot*8 + col]
    __shared__ double rinv_s[N * N];
    __shared__ double fin_scratch[FIN_SCRATCH];
    __shared__ __align__(8) uint64_t panel_bar;
    __shared__ double wsum_s[NWARPS];
    __shared__ double gv_s[DMMA_CT * (N * (N + 1) / 2 + P)];
    __shared__ double fin_sh[64];
    __shared__ int is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3; // mma "groupID" and "threadID_in_group"
    const int ld = a.ld, S = a.S, nst = a.nstages;
    const size_t stage_elems = (size_t)CT * lds;
    double *tiles = reinterpret_cast<double *>(smem_raw);
    double *Cout = stream_cout(a);

    const int my = a.tiles_base + ((int)blockIdx.x < a.tiles_rem ? 1 : 0);
    dbg_mark(a.dbg, 0);

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        mbar_init(&panel_bar, 1);
        fence_mbar_init();
    }
    // zero the pad rows [ld, lds) of every column slot (never written by the copies)
    if (lds > ld)
        for (int slot = tid; slot < nst * CT; slot += THREADS)
            for (int r = ld; r < lds; ++r) tiles[(size_t)slot * lds + r] = 0.0;
    __syncthreads();
    dbg_mark(a.dbg, 8);

    // Producer: every warp issues the bulk copies of its own columns (column c of a tile is
    // fetched by warp c % NWARPS, lane 0), so the ~70 ns issue cost of a copy is paid in
    // parallel instead of 8x in one thread. Thread 0 arms the barrier with the byte count of
    // the whole tile (a copy may complete before the expect_tx: the tx-count is signed).
    int next_i = 0, next_st = 0; // next tile to fetch and the stage it goes to (uniform)
    const uint32_t col_bytes = (uint32_t)(ld * sizeof(double));
    auto issue = [&]() {
        const int tile = blockIdx.x + next_i * gridDim.x;
        const int col0 = tile * CT;
        const int nc = min(CT, S - col0);
        if (lane == 0) {
            if (warp == 0) mbar_arrive_expect_tx(&full_bar[next_st], col_bytes * nc);
            double *dst = tiles + (size_t)next_st * stage_elems;
            const double *src = a.Y + (size_t)col0 * ld;
#pragma unroll 1
            for (int c = warp; c < nc; c += NWARPS)
                bulk_copy_g2s(dst + (size_t)c * lds, src + (size_t)c * ld, col_bytes, &full_bar[next_st]);
        }
        ++next_i;
        if (++next_st == nst) next_st = 0;
    };
    // The observations do not depend on the panel: start fetching immediately. The last stage
    // first receives the panel columns (one bulk copy each, into padded slots so that the
    // fragment reads are conflict free); its first Y tile is requested once the fragments
    // are in registers. Panel first: the copy engine serves requests in order.
    constexpr int PROWS = 4 * KSTEPS * NWARPS;
    const int prow_copy = min(PROWS, lds); // rows of each panel column staged (ldp >= PROWS, zero padded)
    double *pstage = tiles + (size_t)(nst - 1) * stage_elems;
    if (lane == 0) {
        if (warp == 0) mbar_arrive_expect_tx(&panel_bar, (uint32_t)(NPV * prow_copy * sizeof(double)));
#pragma unroll 1
        for (int c = warp; c < NPV; c += NWARPS)
            bulk_copy_g2s(pstage + (size_t)c * lds, a.Pq + (size_t)c * a.ldp, (uint32_t)(prow_copy * sizeof(double)), &panel_bar);
    }
    for (int i = 0; i < nst - 1 && i < my; ++i) issue();

    // A fragments, resident in registers:
    //   a1[ks] = [Q|E][row = 4*(warp*KSTEPS+ks) + tig][dot = grp]      (phase 1: A is 8 dots x 4 rows)
    //   a2[rs] = Q    [row = 8*(warp*RSTEPS+rs) + grp][k = tig]        (phase 2: A is 8 rows x 4 k)
    // The panel has ldp >= 4*KSTEPS*NWARPS zero-padded rows and an all-zero column NPV, so the
    // loads need no predicates: unused dots / k point at the zero column.
    double a1[KSTEPS], a2[RSTEPS];
    if (tid < N * N) rinv_s[tid] = a.small->Rinv[(tid / N) * VP_MAX_N + (tid % N)];
    dbg_mark(a.dbg, 9);
    mbar_wait(&panel_bar, 0);
    dbg_mark(a.dbg, 10);
    {
        const bool use1 = grp < NPV, use2 = tig < N;
        const double *src1 = pstage + (size_t)(use1 ? grp : 0) * lds + 4 * (warp * KSTEPS) + tig;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const bool ok = use1 && (EXACT || 4 * (warp * KSTEPS + ks) + tig < lds);
            a1[ks] = ok ? src1[4 * ks] : 0.0;
        }
        const double *src2 = pstage + (size_t)(use2 ? tig : 0) * lds + 8 * (warp * RSTEPS) + grp;
#pragma unroll
        for (int rs = 0; rs < RSTEPS; ++rs) {
            const bool ok = use2 && (EXACT || 8 * (warp * RSTEPS + rs) + grp < lds);
            a2[rs] = ok ? src2[8 * rs] : 0.0;
        }
    }
    __syncthreads(); // every warp has its fragments: the last stage may now receive Y
    if (next_i < my) issue();
    dbg_mark(a.dbg, 1);

    double rn2 = 0.0;
    double Gacc[N * (N + 1) / 2];
    double Vacc[P > 0 ? P : 1];
#pragma unroll
    for (int i = 0; i < N * (N + 1) / 2; ++i) Gacc[i] = 0.0;
#pragma unroll
    for (int e = 0; e < (P > 0 ? P : 1); ++e) Vacc[e] = 0.0;

    // phase-2 column permutation: mma column n <-> tile column pi(n) = n/2 + 4*(n%2)
    const int pcol_b = (grp >> 1) + 4 * (grp & 1); // B fragment: mma col = grp
    const int pcol_c0 = tig;                       // C fragment: mma cols 2*tig, 2*tig+1
    const int pcol_c1 = tig + 4;

    int st = 0;
    uint32_t parity = 0;
    for (int i = 0; i < my; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int col0 = tile * CT;
        const int nc = min(CT, S - col0);
        const double *tp = tiles + (size_t)st * stage_elems;
        mbar_wait(&full_bar[st], parity);
        if (i == 0) dbg_mark(a.dbg, 2);
        if (++st == nst) { st = 0; parity ^= 1u; }

        // ---- phase 1: C(8 dots x 8 cols) += A1(8 x 4) * Y(4 rows x 8 cols) over the warp's rows
        {
            double c[4][2];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) c[ch][0] = c[ch][1] = 0.0;
            const double *bp = tp + (size_t)grp * lds + 4 * (warp * KSTEPS) + tig;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                double b;
                if (EXACT) b = bp[4 * ks];
                else b = (4 * (warp * KSTEPS + ks) + tig < lds) ? bp[4 * ks] : 0.0;
                dmma_8x8x4(c[ks & 3][0], c[ks & 3][1], a1[ks], b);
            }
            const double s0 = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
            const double s1 = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
            // C fragment: row (dot) = grp, cols 2*tig, 2*tig+1
            *reinterpret_cast<double2 *>(&part[warp * 64 + grp * 8 + 2 * tig]) = make_double2(s0, s1);
        }
        __syncthreads(); // (A) every warp is past phase 2 of the previous tile
        if (i >= 1 && next_i < my) issue(); // refill the stage tile i-1 used
        if (tid < 64) {
            double s = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < NWARPS; ++w2) s += part[w2 * 64 + tid];
            bu[tid] = s;
        }
        __syncthreads(); // (B) b_s, u_s of the 8 columns are complete

        // ---- solve: c_s = R1^-1 b_s ; accumulate G and V (one thread per column)
        if (tid < nc) {
            double coef[N];
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double s = 0.0;
#pragma unroll
                for (int c2 = 0; c2 < N; ++c2) s += rinv_s[c2 * N + r] * bu[c2 * 8 + tid];
                coef[r] = s;
                Cout[(size_t)(col0 + tid) * N + r] = s;
            }
            int gi = 0;
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c2 = r; c2 < N; ++c2) Gacc[gi++] += coef[r] * coef[c2];
#pragma unroll
            for (int e = 0; e < P; ++e) {
                double cj = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) cj = (a.e_basis[e] == r) ? coef[r] : cj;
                Vacc[e] += cj * bu[(N + e) * 8 + tid];
            }
        }

        // ---- phase 2: R(8 rows x 8 cols) = Y + Q(8 x 4) * (-b)(4 x 8); accumulate r^2
        {
            const double b2 = (tig < N) ? -bu[tig * 8 + pcol_b] : 0.0; // B fragment: row k = tig, mma col = grp
            const double *cp0 = tp + (size_t)pcol_c0 * lds + 8 * (warp * RSTEPS) + grp;
            const double *cp1 = tp + (size_t)pcol_c1 * lds + 8 * (warp * RSTEPS) + grp;
            double q0 = 0.0, q1 = 0.0;
#pragma unroll
            for (int rs = 0; rs < RSTEPS; ++rs) {
                double d0, d1;
                if (EXACT) { d0 = cp0[8 * rs]; d1 = cp1[8 * rs]; }
                else {
                    const bool ok = 8 * (warp * RSTEPS + rs) + grp < lds;
                    d0 = ok ? cp0[8 * rs] : 0.0;
                    d1 = ok ? cp1[8 * rs] : 0.0;
                }
                dmma_8x8x4(d0, d1, a2[rs], b2);
                q0 = fma(d0, d0, q0);
                q1 = fma(d1, d1, q1);
            }
            rn2 += (pcol_c0 < nc ? q0 : 0.0) + (pcol_c1 < nc ? q1 : 0.0);
        }
    }

    // ---- CTA partial -> global, last CTA folds -------------------------------
    dbg_mark(a.dbg, 3);
    __syncthreads();
    cta_publish<double, N, P, CT, NWARPS>(a, rn2, Gacc, Vacc, wsum_s, gv_s, fin_sh, fin_scratch, &is_last);
}

} // namespace vp
