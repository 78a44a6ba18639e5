// fit_kernel_dmma.cuh -- the fused evaluation / whole-fit kernel (fp64).
//
// One launch does, per evaluation, what K1 (panel_kernel_hh) + K2 (stream_kernel_dmma) do in
// two, and -- in fit mode -- runs the WHOLE Levenberg-Marquardt fit as one persistent,
// co-resident grid (cooperative launch, one CTA per SM):
//
//   repeat
//     every CTA : panel at the trial parameters (Phi_w, D, Householder QR, E = P_perp D, M, R1^-1)
//                 redundantly in its own registers / shared memory        [K1 without HBM round trip]
//                 -> DMMA A-fragments -> stream its share of the Y tiles   [K2]
//                 -> publish the CTA partial, take a ticket
//     last CTA  : fold the partials in a fixed order -> (||r||^2, g, H)
//                 -> (multi-GPU: exchange with the peer GPUs over NVLink, see vp_comm)
//                 -> advance the lmder state machine (lm_step.cuh) -> broadcast (x_trial, more)
//     other CTAs: spin on the generation flag (acquire), with the first tiles of the next
//                 evaluation already in flight (Y does not depend on alpha)
//   until the state machine terminates
//
// Why: a C2 evaluation streams 33.5 MB in ~4 us but the two-kernel graph loop costs ~48 us per
// evaluation (K1 12 us on ONE SM while 147 idle, launch gaps, conditional-node relaunch;
// profiles/r01a_timeline_c2.txt). Fusing removes every launch from the loop; the panel costs
// each CTA a few microseconds of latency that overlap the initial tile prefetch.
// Reference mapping: the loop body is impl LeastSquaresProblem for SeparableProblem
// (src/solvers/levmar/mod.rs:42-201) and the loop itself LevenbergMarquardt::minimize (:247).
//
// Single-evaluation mode (a.fit == nullptr): one pass, EvalOut written by the last CTA, no
// grid-wide wait (ordinary launch) -- used by set_params / problem creation / profiling.
#pragma once

#include "panel_kernel_hh.cuh"
#include "stream_kernel_dmma.cuh"

namespace vp {

// what the last CTA broadcasts to the grid after every evaluation of a fit
struct FitBcast {
    double x_trial[VP_MAX_Q];
    int more;         // 1: another evaluation follows
    int cdst;         // coefficient buffer the next evaluation writes
    unsigned int gen; // evaluation e publishes gen = e + 1
    int error;        // 1: a grid-wide wait timed out (should never happen on a co-resident grid)
};

// One-shot all-to-all exchange of the reduced vector between the GPUs of a column-sharded
// global fit (vp_comm): rank r stores its contribution into slot r of every peer's mailbox
// through NVLink peer mappings and then a flag; each rank sums the W slots in rank order, so all
// ranks obtain bitwise identical sums and the replicated LM step stays in lock step.
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_SLOT_DOUBLES = 80; // >= 1 + VP_MAX_Q + VP_MAX_Q^2 (rnorm2, g, H) + nonfinite
struct CommMailbox { // lives in this rank's HBM, mapped into every peer
    double slot[2][COMM_MAX_WORLD][COMM_SLOT_DOUBLES];
    unsigned long long flag[2][COMM_MAX_WORLD];
};
struct CommArgs {
    int world, rank;
    unsigned long long *epoch;              // this rank's exchange counter (device memory)
    CommMailbox *box[COMM_MAX_WORLD];       // box[r]: rank r's mailbox as mapped in THIS process
    int *error;                             // set to 1 on timeout
};

struct FitArgs {
    ModelDesc md;
    const double *x; // m values of the independent variable
    const double *w; // m weights or nullptr
    double svd_eps;
    const double *alpha_dev; // parameters of the first evaluation
    FitBcast *bc;            // fit mode only
    CommArgs comm;           // comm.world == 0: no communicator
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

constexpr unsigned long long SPIN_TIMEOUT_NS = 4000000000ull; // 4 s: a dead peer / lost CTA must not hang the GPU

// Sum `vals[0..nv)` (shared memory, this rank's contribution) over all ranks; executed by the
// whole CTA; on return vals holds the global sums (identical bits on every rank). Returns 0 on
// success, 1 on timeout.
__device__ __forceinline__ int comm_allreduce(const CommArgs &c, double *vals, int nv, int *flag_s)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int W = c.world;
    const unsigned long long ep = *c.epoch + 1; // same on every rank: all ranks make the same sequence of evaluations
    const int par = (int)(ep & 1ull);
    __syncthreads();
    for (int idx = tid; idx < W * nv; idx += nt) {
        const int r = idx / nv, k = idx - r * nv;
        c.box[r]->slot[par][c.rank][k] = vals[k];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < W) st_release_sys_u64(&c.box[tid]->flag[par][c.rank], ep);
    if (tid == 0) *flag_s = 0;
    __syncthreads();
    if (tid < W) {
        const unsigned long long t0 = global_timer_ns();
        const unsigned long long *fl = &c.box[c.rank]->flag[par][tid];
        while (ld_acquire_sys_u64(fl) < ep) {
            if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) { *flag_s = 1; break; }
        }
    }
    __syncthreads();
    const int timed_out = *flag_s;
    if (tid < nv) {
        double s = 0.0;
        const CommMailbox *mine = c.box[c.rank];
        for (int r = 0; r < W; ++r) s += __ldcv(&mine->slot[par][r][tid]);
        vals[tid] = s;
    }
    if (tid == 0) {
        *c.epoch = ep;
        if (timed_out) *c.error = 1;
    }
    __syncthreads();
    return timed_out;
}

// Weighted basis functions and derivative columns of this thread's rows, evaluated through a
// shared-memory staging area (column k at stg + k*lds): the loop over the basis functions is
// ROLLED (one copy of the exp / sincos code, selected by a uniform branch on the kind) while
// the rows of a thread are unrolled inside it, so the long dependent chains of exp() and of
// the fp64 divisions of RPT rows overlap. The values are then read back into statically
// indexed registers. (The straight-line panel_hh_eval calls a non-inlined evaluator once per
// row and basis function, back to back: ~5 us of pure latency at 8 warps per SM.)
template <int N, int P, int RPT, int THREADS>
__device__ __forceinline__ int panel_eval_staged(const ModelDesc &md, const double (&xi)[RPT], const double (&wi)[RPT],
                                                 const double *alpha_s, double *stg, const int lds,
                                                 double (&a)[RPT][N + P], double (&d0)[RPT][P > 0 ? P : 1])
{
    constexpr int NPV = N + P;
    const int tid = threadIdx.x;
    int bad = 0;
    int e = 0;
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        const int kind = md.kind[j], np = md.npar[j];
        const double a0 = np > 0 ? alpha_s[md.pidx[j][0]] : 0.0;
        const double a1 = np > 1 ? alpha_s[md.pidx[j][1]] : 0.0;
        const double scale = md.scale[j];
        double v[RPT], da[RPT], db[RPT];
        if (kind == VP_BASIS_EXP_DECAY) { // exp(-x/tau); exp(-x/tau)*x/(tau*tau)
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const double ex = exp(-xi[r] / a0);
                v[r] = ex; da[r] = ex * xi[r] / (a0 * a0); db[r] = 0.0;
            }
        } else if (kind == VP_BASIS_EXP_RATE_COS) {
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const double ex = exp(-a0 * xi[r]);
                double sn, cs;
                sincos(a1 * xi[r], &sn, &cs);
                v[r] = ex * cs; da[r] = -xi[r] * (ex * cs); db[r] = -xi[r] * ex * sn;
            }
        } else if (kind == VP_BASIS_SIN_PHASE) {
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                double sn, cs;
                sincos(a0 * xi[r] + a1, &sn, &cs);
                v[r] = sn; da[r] = xi[r] * cs; db[r] = cs;
            }
        } else {
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                v[r] = kind == VP_BASIS_CONSTANT ? 1.0 : (kind == VP_BASIS_LINEAR_X ? scale * xi[r] : nan(""));
                da[r] = 0.0; db[r] = 0.0;
            }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = tid + r * THREADS;
            const bool in = i < md.m; // wi = 0 there, but 0 * inf must not poison the column
            const double pv = in ? wi[r] * v[r] : 0.0, pa = in ? wi[r] * da[r] : 0.0, pb = in ? wi[r] * db[r] : 0.0;
            bad |= !isfinite(pv) | ((np > 0) & !isfinite(pa)) | ((np > 1) & !isfinite(pb));
            if (i < lds) {
                stg[(size_t)j * lds + i] = pv;
                if (np > 0) stg[(size_t)(N + e) * lds + i] = pa;
                if (np > 1) stg[(size_t)(N + e + 1) * lds + i] = pb;
            }
        }
        e += np;
    }
    // own rows only: no synchronisation needed between the stores above and these loads
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
#pragma unroll
        for (int k = 0; k < NPV; ++k) a[r][k] = (i < lds) ? stg[(size_t)k * lds + i] : 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) d0[r][k] = a[r][N + k];
    }
    return bad;
}

template <int N, int P, int KSTEPS, int NWARPS, bool EXACT>
__global__ void __launch_bounds__(NWARPS * 32, 1)
fit_kernel_dmma(const StreamArgs<double> a, const int lds, const FitArgs f)
{
    constexpr int NPV = N + P;
    constexpr int THREADS = NWARPS * 32;
    constexpr int CT = DMMA_CT;
    constexpr int RSTEPS = KSTEPS / 2; // 8-row steps of phase 2
    constexpr int PROWS = 4 * KSTEPS * NWARPS;
    constexpr int RPT = PROWS / THREADS; // panel rows per thread
    constexpr int KMAX = (NPV > 8) ? NPV : 8;
    static_assert(NPV + 1 <= CT && N <= 4, "one DMMA row block / k block only; panel + zero column fit one stage");
    static_assert(KSTEPS % 8 == 0, "whole panel rows per thread");
    static_assert(THREADS >= 64 + VP_MAX_Q, "panel_hh_body's small-output writers");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STREAM_MAX_STAGES];
    __shared__ __align__(16) double part[NWARPS * 64];
    __shared__ __align__(16) double bu[64]; // bu[dot*8 + col]
    __shared__ double rinv_s[N * N];
    __shared__ double fin_scratch[FIN_SCRATCH];
    __shared__ double wsum_s[NWARPS];
    __shared__ double gv_s[DMMA_CT * (N * (N + 1) / 2 + P)];
    __shared__ double fin_sh[COMM_SLOT_DOUBLES];
    __shared__ double red[2][NWARPS * KMAX];
    __shared__ double top[N][NPV];
    __shared__ double alpha_s[VP_MAX_Q];
    __shared__ PanelSmall small_s;
    __shared__ int is_last, ctrl_more, ctrl_cdst, flag_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3; // mma "groupID" and "threadID_in_group"
    const int ld = a.ld, S = a.S, nst = a.nstages, q = a.q;
    const size_t stage_elems = (size_t)CT * lds;
    double *tiles = reinterpret_cast<double *>(smem_raw);
    const bool fit_mode = a.fit != nullptr;
    StreamArgs<double> al = a; // local copy whose `small` points at this CTA's shared-memory panel outputs
    al.small = &small_s;

    const int my = a.tiles_base + ((int)blockIdx.x < a.tiles_rem ? 1 : 0);
    dbg_mark(a.dbg, 0);

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
        ctrl_cdst = fit_mode ? (__ldcg(&a.fit->cur) ^ 1) : a.cdst;
        ctrl_more = 0;
    }
    // zero the pad rows [ld, lds) of every column slot (never written by the copies)
    if (lds > ld)
        for (int slot = tid; slot < nst * CT; slot += THREADS)
            for (int r = ld; r < lds; ++r) tiles[(size_t)slot * lds + r] = 0.0;
    // this thread's rows of x and w stay in registers for the whole fit
    double xi[RPT], wi[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
        const bool in = i < f.md.m;
        xi[r] = in ? f.x[i] : 0.0;
        wi[r] = in ? (f.w ? f.w[i] : 1.0) : 0.0;
    }
    __syncthreads();

    // Producer (see stream_kernel_dmma): column c of a tile is fetched by warp c % NWARPS, lane 0.
    int next_i = 0, next_st = 0;
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
    // the observations do not depend on the parameters: start fetching immediately. The last
    // stage is where the panel is staged for the fragment loads, so it is filled afterwards.
    for (int i = 0; i < nst - 1 && i < my; ++i) issue();

    double *pstage = tiles + (size_t)(nst - 1) * stage_elems;
    uint32_t phase_bits = 0; // bit st: parity of the next completion to wait for on stage st
    // phase-2 column permutation: mma column n <-> tile column pi(n) = n/2 + 4*(n%2)
    const int pcol_b = (grp >> 1) + 4 * (grp & 1);
    const int pcol_c0 = tig, pcol_c1 = tig + 4;

    for (unsigned int e = 0;; ++e) {
        // ---- parameters of this evaluation ------------------------------------------------
        if (tid < VP_MAX_Q)
            alpha_s[tid] = tid < q ? (e == 0 ? __ldcg(&f.alpha_dev[tid]) : __ldcg(&f.bc->x_trial[tid])) : 0.0;
        __syncthreads();
        double *Cout = ctrl_cdst ? a.C1 : a.C0;

        // ---- K1: the panel, in this CTA ---------------------------------------------------
        dbg_mark(a.dbg, 8);
        {
            double pa[RPT][NPV], pd0[RPT][P > 0 ? P : 1];
            const int bad = panel_eval_staged<N, P, RPT, THREADS>(f.md, xi, wi, alpha_s, pstage, lds, pa, pd0);
            dbg_mark(a.dbg, 9);
            panel_hh_factor<double, N, P, RPT, THREADS>(f.md, pa, pd0, bad, alpha_s, f.svd_eps, lds, pstage, &small_s, red, top, nullptr);
        }
        __syncthreads();
        dbg_mark(a.dbg, 10);
        double a1[KSTEPS], a2[RSTEPS];
        if (tid < N * N) rinv_s[tid] = small_s.Rinv[(tid / N) * VP_MAX_N + (tid % N)];
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
        fence_proxy_async_smem(); // generic accesses to the panel stage before the bulk copy overwrites it
        __syncthreads();
        if (next_i < my) issue();
        dbg_mark(a.dbg, 1);

        // ---- K2: stream this CTA's tiles ----------------------------------------------------
        double rn2 = 0.0;
        double Gacc[N * (N + 1) / 2];
        double Vacc[P > 0 ? P : 1];
#pragma unroll
        for (int i = 0; i < N * (N + 1) / 2; ++i) Gacc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < (P > 0 ? P : 1); ++i) Vacc[i] = 0.0;

        int st = 0;
        for (int i = 0; i < my; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int col0 = tile * CT;
            const int nc = min(CT, S - col0);
            const double *tp = tiles + (size_t)st * stage_elems;
            mbar_wait(&full_bar[st], (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
            if (i == 0) dbg_mark(a.dbg, 2);
            if (++st == nst) st = 0;

            // phase 1: C(8 dots x 8 cols) += A1(8 x 4) * Y(4 rows x 8 cols) over the warp's rows
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

            // solve: c_s = R1^-1 b_s ; accumulate G and V (one thread per column)
            if (tid < nc) {
                double coef[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double s = 0.0;
#pragma unroll
                    for (int c2 = r; c2 < N; ++c2) s += rinv_s[c2 * N + r] * bu[c2 * 8 + tid];
                    coef[r] = s;
                    Cout[(size_t)(col0 + tid) * N + r] = s;
                }
                int gi = 0;
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c2 = r; c2 < N; ++c2) Gacc[gi++] += coef[r] * coef[c2];
#pragma unroll
                for (int e2 = 0; e2 < P; ++e2) {
                    double cj = 0.0;
#pragma unroll
                    for (int r = 0; r < N; ++r) cj = (a.e_basis[e2] == r) ? coef[r] : cj;
                    Vacc[e2] += cj * bu[(N + e2) * 8 + tid];
                }
            }

            // phase 2: R(8 rows x 8 cols) = Y + Q(8 x 4) * (-b)(4 x 8); accumulate r^2
            {
                const double b2 = (tig < N) ? -bu[tig * 8 + pcol_b] : 0.0;
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
        dbg_mark(a.dbg, 3);
        __syncthreads(); // every warp is done with every stage

        // ---- the next evaluation reads the same tiles: put the first ones in flight now ----------
        next_i = 0;
        next_st = 0;
        if (fit_mode)
            for (int i = 0; i < nst - 1 && i < my; ++i) issue();

        // ---- CTA partial -> global; the last CTA folds, steps the LM state machine, broadcasts -----
        cta_publish_partial<double, N, P, CT, NWARPS>(al, rn2, Gacc, Vacc, wsum_s, gv_s, &is_last);
        if (is_last) {
            stream_finalize<double, true, false>(al, N, P, gridDim.x, fin_sh, fin_scratch);
            __syncthreads();
            if (f.comm.world >= 1) { // world == 1: exchange with itself (exercises the protocol on one GPU)
                // fold this GPU's (||r||^2, g, H) with the peers' (all linear in the column sums)
                const int nv = 2 + q + q * q;
                if (tid == 0) {
                    fin_sh[0] = a.out->rnorm2;
                    for (int k = 0; k < q; ++k) fin_sh[1 + k] = a.out->g[k];
                    for (int k = 0; k < q * q; ++k) fin_sh[1 + q + k] = a.out->H[k];
                    fin_sh[1 + q + q * q] = a.out->finite ? 0.0 : 1.0;
                }
                comm_allreduce(f.comm, fin_sh, nv, &flag_s);
                if (tid == 0) {
                    EvalOut *o = a.out;
                    o->rnorm2 = fin_sh[0];
                    for (int k = 0; k < q; ++k) o->g[k] = fin_sh[1 + k];
                    for (int k = 0; k < q * q; ++k) o->H[k] = fin_sh[1 + q + k];
                    o->finite = (fin_sh[1 + q + q * q] == 0.0) && isfinite(fin_sh[0]);
                }
                __syncthreads();
            }
            dbg_mark(a.dbg, 7);
            if (fit_mode) {
                // advance the lmder state machine on a shared-memory copy of the state
                unsigned long long *fw = reinterpret_cast<unsigned long long *>(a.fit);
                unsigned long long *lw = reinterpret_cast<unsigned long long *>(fin_scratch);
                for (int i = tid; i < FIT_WORDS; i += THREADS) lw[i] = __ldcg(fw + i);
                __syncthreads();
                if (tid == 0) {
                    FitDevice *fd = reinterpret_cast<FitDevice *>(lw);
                    LmEval ev;
                    ev.rnorm2 = a.out->rnorm2;
                    ev.finite = a.out->finite;
                    for (int kk = 0; kk < VP_LM_MAXQ; ++kk) ev.g[kk] = kk < q ? a.out->g[kk] : 0.0;
                    for (int kk = 0; kk < VP_LM_MAXQ * VP_LM_MAXQ; ++kk) ev.H[kk] = kk < q * q ? a.out->H[kk] : 0.0;
                    const bool more = lm_advance(fd->st, fd->cfg, ev);
                    if (fd->st.last_accepted) {
                        fd->cur ^= 1;
                        fd->accepted = ev;
                    }
                    if (fd->evals < 48) {
                        double *tr = fd->trace + 4 * fd->evals;
                        tr[0] = sqrt(ev.rnorm2); tr[1] = fd->st.par; tr[2] = fd->st.delta; tr[3] = fd->st.last_accepted;
                    }
                    fd->evals += 1;
                    for (int kk = 0; kk < VP_MAX_Q; ++kk) f.bc->x_trial[kk] = fd->st.x_trial[kk];
                    f.bc->cdst = fd->cur ^ 1;
                    f.bc->more = more ? 1 : 0;
                    ctrl_more = more ? 1 : 0;
                    ctrl_cdst = fd->cur ^ 1;
                    dbg_mark(a.dbg, 15);
                }
                __syncthreads();
                for (int i = tid; i < FIT_WORDS; i += THREADS) fw[i] = lw[i];
                __threadfence();
                __syncthreads();
                if (tid == 0) st_release_gpu_u32(&f.bc->gen, e + 1u);
            }
            dbg_mark(a.dbg, 6);
        } else if (fit_mode) {
            if (tid == 0) {
                const unsigned long long t0 = global_timer_ns();
                int ok = 1;
                while (ld_acquire_gpu_u32(&f.bc->gen) != e + 1u) {
                    if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) { ok = 0; break; }
                }
                dbg_mark(a.dbg, 6);
                if (ok) {
                    ctrl_more = __ldcg(&f.bc->more);
                    ctrl_cdst = __ldcg(&f.bc->cdst);
                } else {
                    ctrl_more = 0;
                    f.bc->error = 1;
                }
            }
        }
        __syncthreads();
        if (!fit_mode || !ctrl_more) break;
    }
    // drain the tiles that were put in flight for an evaluation that is not going to happen
    if (fit_mode)
        for (int s = 0; s < next_i; ++s) mbar_wait(&full_bar[s], (phase_bits >> s) & 1u);
}

} // namespace vp
