// fit_kernel_dmma.cuh -- the fused evaluation / whole-fit kernel (fp64).
//
// One launch does, per evaluation, what K1 (panel_kernel_hh) + K2 (stream_kernel_dmma) do in
// two, and -- in fit mode -- runs the WHOLE Levenberg-Marquardt fit as one persistent,
// co-resident grid (cooperative launch, one CTA per SM):
//
//   repeat
//     every CTA : panel at the trial parameters (Phi_w, D, Householder QR, E = P_perp D, M, R1^-1)
//                 redundantly in its own registers / shared memory        [K1 without HBM round trip]
//                 -> DMMA A-fragments -> stream its part of the Y tiles    [K2, dmma_tile.cuh]
//                 -> publish the part's partial row, count itself in (one atomic)
//                 -> wait until every CTA is in, fold ALL partial rows in a fixed order
//                    -> (||r||^2, g, H)      [multi-GPU: the last CTA in sends this GPU's sums to
//                    the peers over NVLink; every CTA sums the mailbox slots in rank order]
//                 -> advance its own shared-memory copy of the lmder state machine (lm_step.cuh)
//   until the state machine terminates; CTA 0 writes the final state back.
//
// Every CTA computes the same sums in the same order, so all copies of the LM state take the same
// decisions: no broadcast hop, no state round trip through HBM, and the only grid-wide
// communication per evaluation is the partial rows + one counter. (Round 1 had the last CTA fold,
// step and broadcast: ~6 us more per evaluation.) While a CTA waits, the first tiles of the next
// evaluation are already in flight (Y does not depend on alpha).
//
// Why: a C2 evaluation streams 33.5 MB in ~4 us but the two-kernel graph loop costs ~48 us per
// evaluation (K1 12 us on ONE SM while 147 idle, launch gaps, conditional-node relaunch;
// profiles/r01a_timeline_c2.txt). Fusing removes every launch from the loop.
// Reference mapping: the loop body is impl LeastSquaresProblem for SeparableProblem
// (src/solvers/levmar/mod.rs:42-201) and the loop itself LevenbergMarquardt::minimize (:247).
//
// Single-evaluation mode (a.fit == nullptr): one pass, EvalOut written by the last CTA, no
// grid-wide wait (ordinary launch) -- used by set_params / problem creation / profiling.
#pragma once

#include "dmma_tile.cuh"
#include "panel_kernel_hh.cuh"

namespace vp {

// grid-wide control words of one fit launch (device memory, zeroed by the host before the launch)
struct FitCtl {
    unsigned int arrived; // partial rows published so far, counted over all evaluations of the launch
    int error;            // 1: a grid-wide wait timed out (should never happen on a co-resident grid)
};

// One-shot all-to-all exchange of the reduced vector between the GPUs of a column-sharded
// global fit (vp_comm): rank r stores its contribution into slot r of every peer's mailbox
// through NVLink peer mappings and then a flag; every CTA of every rank sums the W slots in rank
// order, so all ranks obtain bitwise identical sums and the replicated LM step stays in lock step.
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_SLOT_DOUBLES = 80; // >= fused_red_count(n, p) + 1 for every instantiated shape
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
    const void *x; // m values of the independent variable, in the problem dtype
    const void *w; // m weights or nullptr, in the problem dtype
    double svd_eps;
    const double *alpha_dev; // parameters of the first evaluation
    FitCtl *ctl;             // fit mode only
    CommArgs comm;           // comm.world == 0: no communicator
    int jac_full;            // 1: add the second Golub-Pereyra term to J^T J (vp_problem_set_jacobian)
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
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu_u32(unsigned int *p, unsigned int v)
{
    unsigned int prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(prev) : "l"(p), "r"(v) : "memory");
    return prev;
}
// the same without a return value: nothing waits for the L2 round trip
__device__ __forceinline__ void red_add_release_gpu_u32(unsigned int *p, unsigned int v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

// Send `vals[0..nv)` (shared memory, this rank's contribution to exchange number `ep`) to every rank's
// mailbox. Executed by ONE whole CTA per rank.
__device__ __forceinline__ void comm_send(const CommArgs &c, const double *vals, const int nv, const unsigned long long ep)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int W = c.world, par = (int)(ep & 1ull);
    for (int idx = tid; idx < W * nv; idx += nt) {
        const int r = idx / nv, k = idx - r * nv;
        c.box[r]->slot[par][c.rank][k] = vals[k];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < W) st_release_sys_u64(&c.box[tid]->flag[par][c.rank], ep);
}
// Wait for the W contributions to exchange `ep` in this rank's own mailbox and sum them in rank order
// into vals[0..nv) (shared memory). Executed by whole CTAs (any number per rank). Returns 1 on timeout.
__device__ __forceinline__ int comm_recv(const CommArgs &c, double *vals, const int nv, const unsigned long long ep, int *flag_s)
{
    const int tid = threadIdx.x;
    const int W = c.world, par = (int)(ep & 1ull);
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
    __syncthreads();
    return timed_out;
}

// Sum nv values over the nrows partial rows (written by other CTAs: ld.global.cg) into sh[0..nv).
// 16 threads per value: thread (k = tid%16, c0 = tid/16) sums value k over the rows c0, c0 + nt/16, ...
// Summation order (fixed => bitwise reproducible for a given partition): four accumulators take the rows of every
// full group of four, the rows after the last full group go to the first accumulator in order, (s0+s1)+(s2+s3),
// then the 16-column table is folded in order. Up to 12 rows per thread (192 rows: the 148 part rows of a
// whole-GPU fit) are loaded in ONE pass -- one L2 round trip instead of three -- and then summed in exactly that
// order. scratch: >= 16 * (nt/16) doubles.
__device__ __forceinline__ void fold_rows(const double *rows, const int rs, const int nrows, const int nv, double *sh, double *scratch)
{
    constexpr int NL = 12;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int k = tid & 15, c0 = tid >> 4, nc0 = nt >> 4;
#pragma unroll 1
    for (int base = 0; base < nv; base += 16) {
        double s = 0.0;
        if (base + k < nv) {
            const double *src = rows + base + k;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (nrows <= NL * nc0) {
                double l[NL];
#pragma unroll
                for (int j = 0; j < NL; ++j) l[j] = (c0 + j * nc0 < nrows) ? __ldcg(src + (size_t)(c0 + j * nc0) * rs) : 0.0;
#pragma unroll
                for (int g = 0; g < NL / 4; ++g) {
                    if (c0 + (4 * g + 3) * nc0 < nrows) { // a full group of four rows
                        s0 += l[4 * g]; s1 += l[4 * g + 1]; s2 += l[4 * g + 2]; s3 += l[4 * g + 3];
                    } else { // the tail (entries past nrows are zero)
                        s0 += l[4 * g]; s0 += l[4 * g + 1]; s0 += l[4 * g + 2]; s0 += l[4 * g + 3];
                    }
                }
            } else {
                int c = c0;
                for (; c + 3 * nc0 < nrows; c += 4 * nc0) {
                    const double l0 = __ldcg(src + (size_t)c * rs), l1 = __ldcg(src + (size_t)(c + nc0) * rs);
                    const double l2 = __ldcg(src + (size_t)(c + 2 * nc0) * rs), l3 = __ldcg(src + (size_t)(c + 3 * nc0) * rs);
                    s0 += l0; s1 += l1; s2 += l2; s3 += l3;
                }
                for (; c < nrows; c += nc0) s0 += __ldcg(src + (size_t)c * rs);
            }
            s = (s0 + s1) + (s2 + s3);
        }
        scratch[c0 * 16 + k] = s;
        __syncthreads();
        if (tid < 16 && base + tid < nv) {
            double t = 0.0;
            for (int c = 0; c < nc0; ++c) t += scratch[c * 16 + tid];
            sh[base + tid] = t;
        }
        __syncthreads();
    }
}

// (||r||^2, G, V, U) sums in sh -> the evaluation (rnorm2, g = J^T r, H = J^T J) in *ev (shared memory):
//   g_k = -(sum_{e in k} V_e),   H_kl = sum_{e in k, f in l} M_ef G_{j(e) j(f)}   [Kaufman]
//   jac_full:  H_kl += sum_{e in k, f in l} (R1^-1 R1^-T)_{j(e) j(f)} U_ef        [Golub-Pereyra]
// One thread per entry of g and of the lower triangle of H. Msh: E^T E (ld ldm); rinv: Rinv (n x n, ld N).
template <int N, int P>
__device__ __forceinline__ void fused_assemble(const double *sh, const double *Msh, const int ldm, const double *rinv,
                                               const int *e_basis, const int *e_param, const int q, const int jac_full,
                                               const int nonfinite, LmEval *ev)
{
    constexpr int NG = N * (N + 1) / 2;
    const int tid = threadIdx.x;
    int ok = 1; // derivatives finite (every thread votes on its entry)
    if (tid == 0) ev->rnorm2 = sh[0];
    if (tid >= 32 && tid < 32 + q) {
        const int kk = tid - 32;
        double gk = 0.0;
        for (int e = 0; e < P; ++e)
            if (e_param[e] == kk) gk -= sh[1 + NG + e];
        ev->g[kk] = gk;
        ok = isfinite(gk);
    }
    if (tid >= 64 && tid < 64 + q * q) {
        const int kk = (tid - 64) / q, l = (tid - 64) % q;
        if (l <= kk) {
            double h = 0.0;
            for (int e = 0; e < P; ++e) {
                if (e_param[e] != kk) continue;
                for (int f = 0; f < P; ++f) {
                    if (e_param[f] != l) continue;
                    int i = e_basis[e], j = e_basis[f];
                    if (i > j) { const int t = i; i = j; j = t; }
                    h += Msh[f * ldm + e] * sh[g_index(N, i, j)];
                    if (jac_full) {
                        double wij = 0.0;
                        for (int c = 0; c < N; ++c) wij += rinv[c * N + i] * rinv[c * N + j];
                        const int e1 = e < f ? e : f, f1 = e < f ? f : e;
                        h += wij * sh[1 + NG + P + e1 * P - e1 * (e1 - 1) / 2 + (f1 - e1)];
                    }
                }
            }
            ev->H[l * q + kk] = h;
            ev->H[kk * q + l] = h;
            ok = isfinite(h);
        }
    }
    ok = __syncthreads_and(ok);
    if (tid == 0) ev->finite = ((isfinite(sh[0]) && !nonfinite) ? VP_EVAL_RESIDUAL_OK : 0) | (ok ? VP_EVAL_DERIVS_OK : 0);
    __syncthreads();
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
            const double inv0 = 1.0 / a0;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const double t = xi[r] * inv0, ex = vp_exp(-t); // one division per basis function, see basis_eval_all
                v[r] = ex; da[r] = ex * t * inv0; db[r] = 0.0;
            }
        } else if (kind == VP_BASIS_EXP_RATE_COS) {
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const double ex = vp_exp(-a0 * xi[r]);
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
            bad |= (!isfinite(pv) ? 1 : 0) | (fabs(pv) > RANK_HUGE_ENTRY ? (2 << j) : 0); // flag word of rank_policy.cuh
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

// TY: element type of the observations / coefficients in HBM (double or float); all arithmetic is fp64 (fp32
// problems stream half the bytes and are converted when the tile fragments are loaded, see fit_queue_kernel.cuh).
template <typename TY, int N, int P, int KSTEPS, int NWARPS, bool EXACT>
__global__ void __launch_bounds__(NWARPS * 32, 1)
fit_kernel_dmma(const StreamArgs<TY> a, const int lds, const FitArgs f)
{
    constexpr int NPV = N + P;
    constexpr int THREADS = NWARPS * 32;
    constexpr int CT = DMMA_CT;
    constexpr int RSTEPS = KSTEPS / 2; // 8-row steps of phase 2
    constexpr int PROWS = 4 * KSTEPS * NWARPS;
    constexpr int RPT = PROWS / THREADS; // panel rows per thread
    constexpr int KMAX = (NPV > 8) ? NPV : 8;
    constexpr int NGPU = TileAcc<N, P>::NG + P + TileAcc<N, P>::NU;
    static_assert(NPV + 1 <= CT && N <= 4, "one DMMA row block / k block only; panel + zero column fit one stage");
    static_assert(KSTEPS % 8 == 0, "whole panel rows per thread");
    static_assert(THREADS >= 64 + VP_MAX_Q * VP_MAX_Q, "fused_assemble / panel_hh_body's small-output writers");
    static_assert(1 + NGPU + 1 <= COMM_SLOT_DOUBLES, "mailbox slot too small");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STREAM_MAX_STAGES];
    __shared__ __align__(16) double part[NWARPS * 64];
    __shared__ __align__(16) double bu[64]; // bu[dot*8 + col]
    __shared__ double rinv_s[N * N];
    __shared__ double fold_scratch[16 * (THREADS / 16)];
    __shared__ double wsum_s[NWARPS];
    __shared__ double gv_s[CT * NGPU];
    __shared__ double sums_s[COMM_SLOT_DOUBLES];
    __shared__ double red[2][NWARPS * KMAX];
    __shared__ double top[N][NPV];
    __shared__ double alpha_s[VP_MAX_Q];
    __shared__ PanelSmall small_s;
    __shared__ SmallSvd svd_s;
    __shared__ LmEval ev_s;
    __shared__ __align__(8) FitDevice fd_s; // this CTA's copy of the LM state (fit mode)
    __shared__ int is_last, ctrl_more, flag_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3; // mma "groupID" and "threadID_in_group"
    const int ld = a.ld, S = a.S, nst = a.nstages, q = a.q;
    const size_t stage_elems = (size_t)CT * lds;
    TY *tiles = reinterpret_cast<TY *>(smem_raw);
    // the panel is staged (fp64, column stride lds) in the LAST pst stages of the ring: one for double tiles, two for
    // float tiles (n + p columns of lds doubles must fit)
    const int pst = (int)(((size_t)(NPV + 1) * lds * sizeof(double) + stage_elems * sizeof(TY) - 1) / (stage_elems * sizeof(TY)));
    const bool fit_mode = a.fit != nullptr;
    const int grid = (int)gridDim.x;

    // canonical partition: CTA b owns part b = the contiguous tiles [tile0, tile0 + my)
    const int my = a.tiles_base + ((int)blockIdx.x < a.tiles_rem ? 1 : 0);
    const int tile0 = (int)blockIdx.x * a.tiles_base + min((int)blockIdx.x, a.tiles_rem);
    dbg_mark(a.dbg, 0);

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
        ctrl_more = 0;
    }
    // zero the pad rows [ld, lds) of every column slot (never written by the copies)
    if (lds > ld)
        for (int slot = tid; slot < nst * CT; slot += THREADS)
            for (int r = ld; r < lds; ++r) tiles[(size_t)slot * lds + r] = (TY)0;
    // this thread's rows of x and w stay in registers for the whole fit
    double xi[RPT], wi[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
        const bool in = i < f.md.m;
        xi[r] = in ? (double)static_cast<const TY *>(f.x)[i] : 0.0;
        wi[r] = in ? (f.w ? (double)static_cast<const TY *>(f.w)[i] : 1.0) : 0.0;
    }
    // every CTA keeps its own copy of the LM state
    if (fit_mode) {
        const unsigned long long *fw = reinterpret_cast<const unsigned long long *>(a.fit);
        unsigned long long *lw = reinterpret_cast<unsigned long long *>(&fd_s);
        for (int i = tid; i < FIT_WORDS; i += THREADS) lw[i] = __ldcg(fw + i);
    }
    for (int i = tid; i < (int)(sizeof(LmEval) / 8); i += THREADS) reinterpret_cast<unsigned long long *>(&ev_s)[i] = 0ull;
    const unsigned long long epoch0 = f.comm.world >= 1 ? *f.comm.epoch : 0ull;
    int ebasis[P > 0 ? P : 1];
#pragma unroll
    for (int e2 = 0; e2 < P; ++e2) ebasis[e2] = a.e_basis[e2];
    __syncthreads();

    // Producer (see stream_kernel_dmma): column c of a tile is fetched by warp c % NWARPS, lane 0.
    int next_i = 0, next_st = 0;
    const uint32_t col_bytes = (uint32_t)(ld * sizeof(TY));
    auto issue = [&]() {
        const int col0 = (tile0 + next_i) * CT;
        const int nc = min(CT, S - col0);
        if (lane == 0) {
            if (warp == 0) mbar_arrive_expect_tx(&full_bar[next_st], col_bytes * nc);
            TY *dst = tiles + (size_t)next_st * stage_elems;
            const TY *src = a.Y + (size_t)col0 * ld;
#pragma unroll 1
            for (int c = warp; c < nc; c += NWARPS)
                bulk_copy_g2s(dst + (size_t)c * lds, src + (size_t)c * ld, col_bytes, &full_bar[next_st]);
        }
        ++next_i;
        if (++next_st == nst) next_st = 0;
    };
    // the observations do not depend on the parameters: start fetching immediately. The last pst
    // stages are where the panel is staged for the fragment loads, so they are filled afterwards.
    for (int i = 0; i < nst - pst && i < my; ++i) issue();

    double *pstage = reinterpret_cast<double *>(tiles + (size_t)(nst - pst) * stage_elems);
    uint32_t phase_bits = 0; // bit st: parity of the next completion to wait for on stage st

    for (unsigned int e = 0;; ++e) {
        // ---- parameters of this evaluation ------------------------------------------------
        if (tid < VP_MAX_Q)
            alpha_s[tid] = tid < q ? ((e == 0 || !fit_mode) ? __ldcg(&f.alpha_dev[tid]) : fd_s.st.x_trial[tid]) : 0.0;
        __syncthreads();
        const int cdst = fit_mode ? (fd_s.cur ^ 1) : a.cdst;
        TY *Cout = cdst ? a.C1 : a.C0;

        // ---- K1: the panel, in this CTA ---------------------------------------------------
        dbg_mark(a.dbg, 8);
        {
            double pa[RPT][NPV], pd0[RPT][P > 0 ? P : 1];
            const int bad = panel_eval_staged<N, P, RPT, THREADS>(f.md, xi, wi, alpha_s, pstage, lds, pa, pd0);
            dbg_mark(a.dbg, 9);
            panel_hh_factor<double, N, P, RPT, THREADS>(f.md, pa, pd0, bad, alpha_s, f.svd_eps, lds, pstage, &small_s, red, top, nullptr, &svd_s);
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
        if (sizeof(TY) != sizeof(double)) {
            // the f64 staging overlaid float slots: restore the zero pad rows [ld, lds) of the staging stages
            __syncthreads();
            if (lds > ld)
                for (int slot = (nst - pst) * CT + tid; slot < nst * CT; slot += THREADS)
                    for (int r = ld; r < lds; ++r) tiles[(size_t)slot * lds + r] = (TY)0;
        }
        fence_proxy_async_smem(); // generic accesses to the panel stages before the bulk copies overwrite them
        __syncthreads();
        for (int k = 0; k < pst && next_i < my; ++k) issue();
        dbg_mark(a.dbg, 1);

        // ---- K2: stream this CTA's part -----------------------------------------------------
        TileAcc<N, P> acc;
        acc.clear();
        int st = 0;
        for (int i = 0; i < my; ++i) {
            const int col0 = (tile0 + i) * CT;
            const int nc = min(CT, S - col0);
            const TY *tp = tiles + (size_t)st * stage_elems;
            mbar_wait(&full_bar[st], (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
            if (i == 0) dbg_mark(a.dbg, 2);
            if (++st == nst) st = 0;
            dmma_tile<TY, N, P, KSTEPS, NWARPS, EXACT>(tp, lds, nc, col0, a1, a2, rinv_s, part, bu, Cout, ebasis, acc, [&]() {
                if (i >= 1 && next_i < my) issue(); // refill the stage tile i-1 used
            });
        }
        dbg_mark(a.dbg, 3);
        __syncthreads(); // every warp is done with every stage

        // ---- the next evaluation reads the same tiles: put the first ones in flight now ----------
        next_i = 0;
        next_st = 0;
        if (fit_mode)
            for (int i = 0; i < nst - pst && i < my; ++i) issue();

        // ---- part row -> global (double-buffered by evaluation parity), count this CTA in ------------
        double *rows = a.partials + (size_t)(fit_mode ? (e & 1u) : 0u) * grid * a.red_stride;
        const bool comm_on = f.comm.world >= 1; // world == 1: exchange with itself (exercises the protocol on one GPU)
        dbg_mark(a.dbg, 4);
        part_stage<N, P, CT, NWARPS>(acc, wsum_s, gv_s);
        __syncthreads();
        if (warp == 0) {
            part_flush<N, P, CT, NWARPS>(wsum_s, gv_s, rows + (size_t)blockIdx.x * a.red_stride, lane);
            __syncwarp();
            if (lane == 0) {
                // release/acquire at gpu scope: orders the row before the count and, in whoever
                // observes the full count, the count before the reads of everybody's rows
                if (fit_mode && !comm_on) {
                    // every CTA waits for the full count below and nobody needs to know who was last: a reduction
                    // (no return value, no L2 round trip before the polling starts)
                    red_add_release_gpu_u32(&f.ctl->arrived, 1u);
                    is_last = 0;
                } else {
                    const unsigned int prev = atom_add_acq_rel_gpu_u32(fit_mode ? &f.ctl->arrived : a.ticket, 1u);
                    is_last = fit_mode ? (prev == (e + 1u) * (unsigned int)grid - 1u) : (prev == (unsigned int)grid - 1u);
                }
            }
        }
        dbg_mark(a.dbg, 14);
        constexpr int NVF = 1 + NGPU; // values of a partial row
        int timed_out = 0;
        if (!fit_mode) {
            __syncthreads();
            if (!is_last) break;
            __threadfence();
        } else if (!comm_on) {
            if (tid == 0) {
                const unsigned int want = (e + 1u) * (unsigned int)grid;
                const unsigned long long t0 = global_timer_ns();
                int ok = 1;
                while (ld_acquire_gpu_u32(&f.ctl->arrived) < want) {
                    if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) { ok = 0; break; }
                }
                flag_s = ok ? 0 : 1;
            }
            __syncthreads();
            timed_out = flag_s;
        } else {
            __syncthreads();
        }
        dbg_mark(a.dbg, 5);
        // ---- fold: every CTA (fit mode) / the last CTA (single evaluation, or the sender of a sharded fit) ----
        const bool folds_local = !comm_on || is_last;
        if (folds_local && !timed_out) {
            fold_rows(rows, a.red_stride, grid, NVF, sums_s, fold_scratch);
            if (tid == 0) sums_s[NVF] = small_s.nonfinite ? 1.0 : 0.0;
            __syncthreads();
        }
        if (comm_on) {
            const unsigned long long ep = epoch0 + e + 1ull;
            if (is_last) comm_send(f.comm, sums_s, NVF + 1, ep);
            timed_out = comm_recv(f.comm, sums_s, NVF + 1, ep, &flag_s);
            if (timed_out && tid == 0) *f.comm.error = 1;
        }
        dbg_mark(a.dbg, 12);
        fused_assemble<N, P>(sums_s, small_s.M, VP_MAX_P, rinv_s, a.e_basis, a.e_param, q, f.jac_full, sums_s[NVF] != 0.0, &ev_s);
        dbg_mark(a.dbg, 13);
        if (!fit_mode) {
            if (tid == 0) {
                EvalOut *o = a.out;
                o->rnorm2 = ev_s.rnorm2;
                for (int k = 0; k < q; ++k) o->g[k] = ev_s.g[k];
                for (int k = 0; k < q * q; ++k) o->H[k] = ev_s.H[k];
                o->finite = ev_s.finite;
                *a.ticket = 0; // re-arm for the next launch
                if (comm_on) *f.comm.epoch = epoch0 + 1ull;
            }
            dbg_mark(a.dbg, 6);
            break;
        }
        // ---- the LM step, redundantly in every CTA (identical inputs => identical decisions) ----------
        if (tid == 0) {
            bool more = false;
            if (timed_out) {
                f.ctl->error = 1;
            } else {
                more = lm_advance(fd_s.st, fd_s.cfg, ev_s);
                if (fd_s.st.last_accepted) fd_s.cur ^= 1;
                if (fd_s.evals < 48) {
                    double *tr = fd_s.trace + 4 * fd_s.evals;
                    tr[0] = sqrt(ev_s.rnorm2); tr[1] = fd_s.st.par; tr[2] = fd_s.st.delta; tr[3] = fd_s.st.last_accepted;
                }
                fd_s.evals += 1;
            }
            ctrl_more = more ? 1 : 0;
            dbg_mark(a.dbg, 15);
        }
        __syncthreads();
        dbg_mark(a.dbg, 6);
        if (!timed_out && fd_s.st.last_accepted) { // the evaluation becomes the accepted one (cooperative copy)
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(&fd_s.accepted);
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&ev_s);
            for (int i = tid; i < (int)(sizeof(LmEval) / 8); i += THREADS) dst[i] = src[i];
        }
        if (!ctrl_more) {
            __syncthreads();
            if (blockIdx.x == 0) { // final state -> global (the host reads it after the launch)
                unsigned long long *fw = reinterpret_cast<unsigned long long *>(a.fit);
                const unsigned long long *lw = reinterpret_cast<const unsigned long long *>(&fd_s);
                for (int i = tid; i < FIT_WORDS; i += THREADS) fw[i] = lw[i];
                if (comm_on && tid == 0) *f.comm.epoch = epoch0 + e + 1ull;
            }
            break;
        }
    }
    // drain the tiles that were put in flight for an evaluation that is not going to happen
    if (fit_mode)
        for (int s = 0; s < next_i; ++s) mbar_wait(&full_bar[s], (phase_bits >> s) & 1u);
}

} // namespace vp
