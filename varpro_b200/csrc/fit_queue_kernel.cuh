// fit_queue_kernel.cuh -- many independent fits on ONE persistent grid with a device-side work queue
// (the throughput path of vp_fit_many).
//
// fit_kernel_dmma gives each fit a fixed slice of SMs for its whole life; since fits need
// different numbers of evaluations (17..24 on the noise-free benchmark), slices go idle as fits
// finish, the panel is recomputed by every CTA of a slice, and each slice stalls during its own
// serial LM step (profiles/: 53 % of the HBM peak at K = 20). Here all SMs serve all fits:
//
//   work item  = (fit k, chunk j): a contiguous range of tiles of the observations of fit k at the
//                fit's current trial parameters
//   any CTA    : claims the next item (atomic head counter, per-slot sequence flag), loads the
//                DMMA A-fragments of fit k's panel [Q|E] from L2, streams the chunk's tiles through
//                the TMA ring (same tile math as fit_kernel_dmma), publishes the chunk's partial sums
//                and takes a ticket of fit k
//   the CTA that completes the LAST chunk of an evaluation becomes that fit's finisher: it folds the
//                chunk partials in fixed order, advances the lmder state machine, and -- if the fit
//                goes on -- computes the panel at the new trial parameters ONCE (Householder QR in
//                its registers, written to HBM/L2), then pushes the next evaluation's items.
//   Meanwhile every other CTA keeps streaming other fits' chunks: the serial phases of one fit are
//   hidden behind the streaming of the others, the panel is computed by one CTA per evaluation
//   instead of by every CTA, and load balance is dynamic at chunk granularity.
//
// Results are those of fit_kernel_dmma up to the summation order of partial sums (chunks instead
// of CTA-strided tiles). All fits of one launch share the kernel instantiation (model shape, row
// tiling) and the padded row count; the host groups problems accordingly.
// Reference mapping as in fit_kernel_dmma.cuh (src/solvers/levmar/mod.rs:42-201, :247).
#pragma once

#include "fit_kernel_dmma.cuh"

namespace vp {

struct QueueFit {
    ModelDesc md;
    const void *Y;    // ld x S weighted observations in the problem dtype TY
    void *C0, *C1;    // coefficient buffers (TY)
    const void *x, *w; // TY
    double *Pq;       // panel [Q | E | 0] in HBM, ALWAYS f64: (n+p+1) columns of ldp rows (ldp >= the kernel's row tiling)
    PanelSmall *small;
    double *partials; // nchunks rows of red_stride doubles
    unsigned int *ticket;
    EvalOut *out;
    FitDevice *fit;
    double svd_eps;
    int ld, S, ldp, red_stride;
    int ntiles, min_chunk_tiles, max_chunks, adaptive, items_per_cta;
    // set by the finisher for every evaluation (read by the consumers with ld.global.cg):
    int chunk_tiles, nchunks;
    int cdst;         // coefficient buffer the current evaluation writes
};

struct QueueItem {
    int fit, chunk;
    unsigned long long seq; // item index + 1 once the item is valid
};

struct QueueCtl {
    unsigned long long head; // next item index to claim
    unsigned long long tail; // items reserved so far
    int fits_left;
    int error;
    QueueItem *items;
    unsigned int cap;
    // finish queue (optional, VP_QUEUE_FINISHERS=n): the CTA that completes the last chunk of an evaluation hands
    // the fit to one of n dedicated CTAs (blockIdx.x < n) that only finalize / step the LM state machine / build
    // panels. Built to test the hypothesis that the rarely executed LM + panel code was slow because of a cold
    // instruction cache; the measurement (VP_QUEUE_DBG) rejected it -- the LM step itself averages ~31 us on one
    // thread -- so the default stays 0: the last chunk's CTA finishes the evaluation itself.
    unsigned long long fhead, ftail;
    QueueItem *fitems;
    unsigned int fcap;
    int nfinishers;          // 0: the last chunk's CTA finishes the evaluation itself
    unsigned long long *dbg; // optional (VP_QUEUE_DBG): per CTA 8 accumulators in ns / counts, see QDBG_*
};
enum { QDBG_ITEMS = 0, QDBG_CLAIM = 1, QDBG_FRAG = 2, QDBG_STREAM = 3, QDBG_PUBLISH = 4, QDBG_FINISH = 5, QDBG_NFINISH = 6, QDBG_TOTAL = 7 };

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu_s32(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// TY: element type of the observations / coefficients in HBM (double or float). All arithmetic is
// fp64 either way: fp32 problems (BASELINE config 4) stream half the bytes and are converted when
// the tile fragments are loaded, so they are MORE accurate than an fp32 reference and share every
// other piece of the machinery. lds (column stride of a tile slot, in elements) = 4 (mod 16) for
// double, 8 (mod 32) for float: both fragment access patterns are then bank-conflict free.
template <typename TY, int N, int P, int KSTEPS, int NWARPS, bool EXACT>
__global__ void __launch_bounds__(NWARPS * 32, 1)
fit_queue_kernel(QueueCtl *ctl, QueueFit *fits, const int nfits, const int lds, const int nst)
{
    constexpr int NPV = N + P;
    constexpr int THREADS = NWARPS * 32;
    constexpr int CT = DMMA_CT;
    constexpr int RSTEPS = KSTEPS / 2;
    constexpr int PROWS = 4 * KSTEPS * NWARPS;
    constexpr int RPT = PROWS / THREADS;
    constexpr int KMAX = (NPV > 8) ? NPV : 8;
    static_assert(NPV + 1 <= CT && N <= 4, "one DMMA row block / k block only");
    static_assert(KSTEPS % 8 == 0, "whole panel rows per thread");
    static_assert(THREADS >= 64 + VP_MAX_Q, "panel_hh_factor's small-output writers");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STREAM_MAX_STAGES];
    __shared__ __align__(16) double part[NWARPS * 64];
    __shared__ __align__(16) double bu[64];
    __shared__ double rinv_s[N * N];
    __shared__ double fin_scratch[FIN_SCRATCH];
    __shared__ double wsum_s[NWARPS];
    __shared__ double gv_s[DMMA_CT * (N * (N + 1) / 2 + P)];
    __shared__ double fin_sh[64];
    __shared__ double red[2][NWARPS * KMAX];
    __shared__ double top[N][NPV];
    __shared__ double alpha_s[VP_MAX_Q];
    __shared__ int is_last, item_fit, item_chunk, more_s, push_n;
    __shared__ unsigned long long fin_acc[8]; // VP_QUEUE_DBG: finisher sub-phases (finalize, state load, LM, state store, x/w + basis, factor + panel store, push, count)
    __shared__ unsigned long long push_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    const size_t stage_elems = (size_t)CT * lds;
    TY *tiles = reinterpret_cast<TY *>(smem_raw);
    double *staging = reinterpret_cast<double *>(smem_raw); // panel staging (f64, stride lds): fits in the ring for nst >= 2

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
        for (int i = 0; i < 8; ++i) fin_acc[i] = 0;
    }
    __syncthreads();
    uint32_t phase_bits = 0;
    const int pcol_b = (grp >> 1) + 4 * (grp & 1);
    const int pcol_c0 = tig, pcol_c1 = tig + 4;

    // Panel of fit k at its trial parameters (fit->st.x_trial) -> HBM, then publish the items of the
    // evaluation. Executed by the whole CTA; all TMA stages must be idle (stage 0 is the staging area).
    auto start_evaluation = [&](const int k) {
        QueueFit *qf = &fits[k];
        const ModelDesc &md = qf->md;
        const unsigned long long ts0 = global_timer_ns();
        if (tid < VP_MAX_Q) alpha_s[tid] = tid < md.q ? __ldcg(&qf->fit->st.x_trial[tid]) : 0.0;
        double xi[RPT], wi[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = tid + r * THREADS;
            const bool in = i < md.m;
            xi[r] = in ? (double)static_cast<const TY *>(qf->x)[i] : 0.0;
            wi[r] = in ? (qf->w ? (double)static_cast<const TY *>(qf->w)[i] : 1.0) : 0.0;
        }
        // the staging rows [ld, lds) of stage 0 may hold zeros only by construction of the evaluator
        __syncthreads();
        {
            double pa[RPT][NPV], pd0[RPT][P > 0 ? P : 1];
            const int bad = panel_eval_staged<N, P, RPT, THREADS>(md, xi, wi, alpha_s, staging, lds, pa, pd0);
            if (tid == 0) fin_acc[4] += global_timer_ns() - ts0;
            panel_hh_factor<double, N, P, RPT, THREADS>(md, pa, pd0, bad, alpha_s, qf->svd_eps, qf->ldp, qf->Pq, qf->small, red, top,
                                                        nullptr);
        }
        if (sizeof(TY) != sizeof(double)) {
            // the f64 staging overlaid the float slots: restore the zero pad rows [ld, lds) of every slot
            __syncthreads();
            const int ld0 = qf->ld;
            for (int slot = tid; slot < nst * CT; slot += THREADS)
                for (int r = ld0; r < lds; ++r) tiles[(size_t)slot * lds + r] = (TY)0;
        }
        fence_proxy_async_smem(); // generic writes to the staging area before later bulk copies into it
        __threadfence();
        __syncthreads();
        const unsigned long long ts1 = global_timer_ns();
        if (tid == 0) fin_acc[5] += ts1 - ts0;
        if (tid == 0) {
            const int cdst = __ldcg(&qf->fit->cur) ^ 1;
            qf->cdst = cdst;
            // chunk size of this evaluation: about four items per CTA over the fits still running, so
            // that the last few fits are spread over the whole grid too
            // qf->adaptive == 0: size the chunks for the initial number of fits, i.e. the SAME partition (and
            // therefore bitwise the same rounding of the partial sums) in every evaluation of a fit
            int active = qf->adaptive ? ld_acquire_gpu_s32(&ctl->fits_left) : nfits;
            if (active < 1) active = 1;
            // many fits: items_per_cta (default 2) items per CTA per round -- measured: per-item overhead (claim,
            // fragment load, pipeline fill, publish ~ 7 us) against balance; few fits: larger items
            const int per_cta = active >= 8 ? qf->items_per_cta : (active >= 3 ? min(2, qf->items_per_cta) : 1);
            const int target = (per_cta * (int)gridDim.x + active - 1) / active;
            int ct = (qf->ntiles + target - 1) / target;
            if (ct < qf->min_chunk_tiles) ct = qf->min_chunk_tiles;
            while ((qf->ntiles + ct - 1) / ct > qf->max_chunks) ct *= 2;
            const int nch = (qf->ntiles + ct - 1) / ct;
            qf->chunk_tiles = ct;
            qf->nchunks = nch;
            push_n = nch;
            push_base = atomicAdd(&ctl->tail, (unsigned long long)nch);
        }
        __syncthreads();
        // publish the items: all threads fill slots, one fence, then the sequence flags (a slot is
        // valid once its flag holds item index + 1; consumers read it with ld.acquire)
        {
            const int nch = push_n;
            const unsigned long long base = push_base;
            for (int j = tid; j < nch; j += THREADS) {
                QueueItem *it = &ctl->items[(base + j) % ctl->cap];
                it->fit = k;
                it->chunk = j;
            }
            __threadfence();
            __syncthreads();
            for (int j = tid; j < nch; j += THREADS)
                st_release_gpu_u64(&ctl->items[(base + j) % ctl->cap].seq, base + j + 1ull);
        }
        __syncthreads();
        if (tid == 0) { fin_acc[6] += global_timer_ns() - ts1; fin_acc[7] += 1; }
    };

    // Fold the chunk partials of fit k, advance its lmder state machine, and either start the next
    // evaluation or retire the fit. Executed by a whole CTA with an idle TMA ring.
    auto finish_evaluation = [&](const int k) {
        QueueFit *qf = &fits[k];
        StreamArgs<double> al{};
        al.small = qf->small;
        al.partials = qf->partials; al.red_stride = qf->red_stride; al.ticket = qf->ticket; al.out = qf->out;
        al.q = qf->md.q; al.dbg = nullptr; al.fit = nullptr;
#pragma unroll
        for (int e2 = 0; e2 < VP_MAX_P; ++e2) { al.e_basis[e2] = qf->md.e_basis[e2]; al.e_param[e2] = qf->md.e_param[e2]; }
        const int nchunks = __ldcg(&qf->nchunks);
        const unsigned long long tf0 = global_timer_ns();
        stream_finalize<double, false, false>(al, N, P, nchunks, fin_sh, fin_scratch);
        __syncthreads();
        const unsigned long long tf1 = global_timer_ns();
        // advance the lmder state machine of fit k on a shared-memory copy of its state
        const int q = qf->md.q;
        unsigned long long *fw = reinterpret_cast<unsigned long long *>(qf->fit);
        unsigned long long *lw = reinterpret_cast<unsigned long long *>(fin_scratch);
        for (int i = tid; i < FIT_WORDS; i += THREADS) lw[i] = __ldcg(fw + i);
        __syncthreads();
        const unsigned long long tf2 = global_timer_ns();
        if (tid == 0) {
            FitDevice *fd = reinterpret_cast<FitDevice *>(lw);
            LmEval ev;
            ev.rnorm2 = __ldcg(&qf->out->rnorm2);
            ev.finite = __ldcg(&qf->out->finite);
            for (int kk = 0; kk < VP_LM_MAXQ; ++kk) ev.g[kk] = kk < q ? __ldcg(&qf->out->g[kk]) : 0.0;
            for (int kk = 0; kk < VP_LM_MAXQ * VP_LM_MAXQ; ++kk) ev.H[kk] = kk < q * q ? __ldcg(&qf->out->H[kk]) : 0.0;
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
            more_s = more ? 1 : 0;
        }
        __syncthreads();
        const unsigned long long tf3 = global_timer_ns();
        for (int i = tid; i < FIT_WORDS; i += THREADS) fw[i] = lw[i];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            fin_acc[0] += tf1 - tf0; fin_acc[1] += tf2 - tf1; fin_acc[2] += tf3 - tf2; fin_acc[3] += global_timer_ns() - tf3;
        }
        if (more_s) {
            start_evaluation(k);
        } else if (tid == 0) {
            __threadfence();
            atomicSub(&ctl->fits_left, 1);
        }
        __syncthreads();
    };

    // zero the pad rows [ld, lds) of every column slot once (the bulk copies never write them; the
    // panel evaluator writes zeros there)
    {
        const int ld0 = fits[0].ld;
        if (lds > ld0)
            for (int slot = tid; slot < nst * CT; slot += THREADS)
                for (int r = ld0; r < lds; ++r) tiles[(size_t)slot * lds + r] = (TY)0;
        __syncthreads();
    }

    // prologue: first panels (the host has advanced every fit to its first trial point)
    for (int k = blockIdx.x; k < nfits; k += gridDim.x) start_evaluation(k);

    // ---- dedicated finisher CTAs -------------------------------------------------------------------
    const int nfin = ctl->nfinishers;
    if (nfin > 0 && (int)blockIdx.x < nfin) {
        for (;;) {
            if (tid == 0) {
                const unsigned long long idx = atomicAdd(&ctl->fhead, 1ull);
                QueueItem *it = &ctl->fitems[idx % ctl->fcap];
                const unsigned long long t0 = global_timer_ns();
                int got = 0;
                for (;;) {
                    if (ld_acquire_gpu_u64(&it->seq) == idx + 1ull) { got = 1; break; }
                    if (ld_acquire_gpu_s32(&ctl->fits_left) <= 0) break;
                    if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) { ctl->error = 1; break; }
                }
                item_fit = got ? it->fit : -1;
            }
            __syncthreads();
            const int k = item_fit;
            if (k < 0) break;
            finish_evaluation(k);
        }
        if (ctl->dbg != nullptr && tid == 0)
            for (int i = 0; i < 8; ++i) ctl->dbg[(size_t)blockIdx.x * 8 + i] = fin_acc[i];
        return;
    }

    unsigned long long dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // tid 0, only when ctl->dbg is set
    const bool dbg_on = ctl->dbg != nullptr;
    const unsigned long long t_kernel0 = dbg_on ? global_timer_ns() : 0ull;
    for (;;) {
        // ---- claim an item ----------------------------------------------------------------------
        const unsigned long long t_a = dbg_on ? global_timer_ns() : 0ull;
        if (tid == 0) {
            const unsigned long long idx = atomicAdd(&ctl->head, 1ull);
            QueueItem *it = &ctl->items[idx % ctl->cap];
            const unsigned long long t0 = global_timer_ns();
            int got = 0;
            for (;;) {
                if (ld_acquire_gpu_u64(&it->seq) == idx + 1ull) { got = 1; break; }
                if (ld_acquire_gpu_s32(&ctl->fits_left) <= 0 && idx >= ld_acquire_gpu_u64(&ctl->tail)) break;
                if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) { ctl->error = 1; break; }
            }
            item_fit = got ? it->fit : -1;
            item_chunk = got ? it->chunk : 0;
        }
        __syncthreads();
        const int k = item_fit, chunk = item_chunk;
        if (k < 0) break;
        const unsigned long long t_b = dbg_on ? global_timer_ns() : 0ull;
        QueueFit *qf = &fits[k];
        const int ld = qf->ld, S = qf->S, ldp = qf->ldp;
        const TY *Yk = static_cast<const TY *>(qf->Y);
        const int chunk_tiles = __ldcg(&qf->chunk_tiles), nchunks = __ldcg(&qf->nchunks);
        const int t_begin = chunk * chunk_tiles;
        const int t_end = min(qf->ntiles, t_begin + chunk_tiles);
        const int my = t_end - t_begin;
        TY *Cout = static_cast<TY *>(__ldcg(&qf->cdst) ? qf->C1 : qf->C0);
        int ebasis[P > 0 ? P : 1];
#pragma unroll
        for (int e2 = 0; e2 < P; ++e2) ebasis[e2] = qf->md.e_basis[e2];

        // ---- TMA producer for this chunk ------------------------------------------------------------
        int next_i = 0, next_st = 0;
        const uint32_t col_bytes = (uint32_t)(ld * sizeof(TY));
        auto issue = [&]() {
            const int col0 = (t_begin + next_i) * CT;
            const int nc = min(CT, S - col0);
            if (lane == 0) {
                if (warp == 0) mbar_arrive_expect_tx(&full_bar[next_st], col_bytes * nc);
                TY *dst = tiles + (size_t)next_st * stage_elems;
                const TY *src = Yk + (size_t)col0 * ld;
#pragma unroll 1
                for (int c = warp; c < nc; c += NWARPS)
                    bulk_copy_g2s(dst + (size_t)c * lds, src + (size_t)c * ld, col_bytes, &full_bar[next_st]);
            }
            ++next_i;
            if (++next_st == nst) next_st = 0;
        };
        for (int i = 0; i < nst && i < my; ++i) issue();

        // ---- A fragments of fit k's panel straight from L2 (the panel changes every evaluation:
        //      ld.global.cg, never L1) ------------------------------------------------------------------
        double a1[KSTEPS], a2[RSTEPS];
        {
            const double *Pq = qf->Pq;
            const bool use1 = grp < NPV, use2 = tig < N;
            const double *src1 = Pq + (size_t)(use1 ? grp : 0) * ldp + 4 * (warp * KSTEPS) + tig;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const double v = __ldcg(src1 + 4 * ks);
                a1[ks] = use1 ? v : 0.0;
            }
            const double *src2 = Pq + (size_t)(use2 ? tig : 0) * ldp + 8 * (warp * RSTEPS) + grp;
#pragma unroll
            for (int rs = 0; rs < RSTEPS; ++rs) {
                const double v = __ldcg(src2 + 8 * rs);
                a2[rs] = use2 ? v : 0.0;
            }
            if (tid < N * N) rinv_s[tid] = __ldcg(&qf->small->Rinv[(tid / N) * VP_MAX_N + (tid % N)]);
        }
        __syncthreads();
        const unsigned long long t_c = dbg_on ? global_timer_ns() : 0ull;

        // ---- stream the chunk (tile math of fit_kernel_dmma) -------------------------------------------
        double rn2 = 0.0;
        double Gacc[N * (N + 1) / 2];
        double Vacc[P > 0 ? P : 1];
#pragma unroll
        for (int i = 0; i < N * (N + 1) / 2; ++i) Gacc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < (P > 0 ? P : 1); ++i) Vacc[i] = 0.0;
        int st = 0;
        for (int i = 0; i < my; ++i) {
            const int col0 = (t_begin + i) * CT;
            const int nc = min(CT, S - col0);
            const TY *tp = tiles + (size_t)st * stage_elems;
            mbar_wait(&full_bar[st], (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
            if (++st == nst) st = 0;
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
            __syncthreads(); // (A)
            if (i >= 1 && next_i < my) issue(); // refill the stage tile i-1 used
            if (tid < 64) {
                double s = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < NWARPS; ++w2) s += part[w2 * 64 + tid];
                bu[tid] = s;
            }
            __syncthreads(); // (B)
            if (tid < nc) {
                double coef[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double s = 0.0;
#pragma unroll
                    for (int c2 = r; c2 < N; ++c2) s += rinv_s[c2 * N + r] * bu[c2 * 8 + tid];
                    coef[r] = s;
                    Cout[(size_t)(col0 + tid) * N + r] = (TY)s;
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
                    for (int r = 0; r < N; ++r) cj = (ebasis[e2] == r) ? coef[r] : cj;
                    Vacc[e2] += cj * bu[(N + e2) * 8 + tid];
                }
            }
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
                rn2 += (pcol_c0 < nc ? q0 : 0.0) + (pcol_c1 < nc ? q1 : 0.0);
            }
        }
        __syncthreads(); // every warp is done with every stage: the ring is idle
        const unsigned long long t_d = dbg_on ? global_timer_ns() : 0ull;

        // ---- chunk partial -> global; the last chunk's CTA finishes the evaluation ---------------------
        StreamArgs<double> al{};
        al.small = qf->small;
        al.partials = qf->partials; al.red_stride = qf->red_stride; al.ticket = qf->ticket; al.out = qf->out;
        al.q = qf->md.q; al.dbg = nullptr; al.fit = nullptr;
#pragma unroll
        for (int e2 = 0; e2 < VP_MAX_P; ++e2) { al.e_basis[e2] = qf->md.e_basis[e2]; al.e_param[e2] = qf->md.e_param[e2]; }
        cta_publish_partial<double, N, P, CT, NWARPS>(al, rn2, Gacc, Vacc, wsum_s, gv_s, &is_last, chunk, nchunks);
        const unsigned long long t_e = dbg_on ? global_timer_ns() : 0ull;
        if (is_last) {
            if (nfin > 0) {
                // hand the fit to a finisher CTA (the partial rows were released by the ticket atomics)
                if (tid == 0) {
                    const unsigned long long base = atomicAdd(&ctl->ftail, 1ull);
                    QueueItem *it = &ctl->fitems[base % ctl->fcap];
                    it->fit = k;
                    it->chunk = -1;
                    __threadfence();
                    st_release_gpu_u64(&it->seq, base + 1ull);
                }
            } else {
                finish_evaluation(k);
            }
        }
        __syncthreads();
        if (dbg_on && tid == 0) {
            const unsigned long long t_f = global_timer_ns();
            dacc[QDBG_ITEMS] += 1; dacc[QDBG_CLAIM] += t_b - t_a; dacc[QDBG_FRAG] += t_c - t_b; dacc[QDBG_STREAM] += t_d - t_c;
            dacc[QDBG_PUBLISH] += t_e - t_d;
            if (is_last) { dacc[QDBG_FINISH] += t_f - t_e; dacc[QDBG_NFINISH] += 1; }
        }
    }
    if (dbg_on && tid == 0) {
        dacc[QDBG_TOTAL] = global_timer_ns() - t_kernel0;
        for (int i = 0; i < 8; ++i) ctl->dbg[(size_t)blockIdx.x * 8 + i] = dacc[i];
    }
}

} // namespace vp
