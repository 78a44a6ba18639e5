// fit_queue_kernel.cuh -- many independent fits on ONE persistent grid with a device-side work queue
// (the throughput path of vp_fit_many).
//
// fit_kernel_dmma gives one fit the whole GPU; its evaluations are separated by serial phases (panel,
// grid-wide fold, LM step) during which HBM idles. Independent fits hide each other's serial phases:
//
//   part       = one entry of the canonical partition of a problem's tiles (dmma_tile.cuh): the SAME
//                partition, and therefore bitwise the same partial sums, as fit_kernel_dmma uses
//   work item  = (fit k, item j): `parts_per_item` consecutive parts of fit k at its current trial
//                parameters. The finisher sizes the items of every evaluation from the number of fits still
//                running (about items_per_cta items per CTA and round): many fits -> few large items per
//                fit (little per-item overhead), last fits -> one part per item (the whole grid serves the
//                tail). The partial sums are per PART, so the item size never changes a result.
//   any CTA    : claims the next item (atomic head counter, per-slot sequence flag), loads the
//                DMMA A-fragments of fit k's panel [Q|E] from L2, streams the item's tiles through
//                the TMA ring (tile math of dmma_tile.cuh), writes one partial row per part and takes
//                a ticket of fit k
//   the CTA that completes the LAST item of an evaluation becomes that fit's finisher: it folds the
//                part rows in fixed order, advances the lmder state machine, and -- if the fit
//                goes on -- computes the panel at the new trial parameters ONCE (Householder QR in
//                its registers, written to HBM/L2), then pushes the next evaluation's items.
//   Meanwhile every other CTA keeps streaming other fits' items.
//
// Results are bitwise those of fit_kernel_dmma (same partition, same fold, same LM code): vp_fit_many
// and vp_fit walk the same iterates and need the same number of evaluations for the same problem.
// All fits of one launch share the kernel instantiation (model shape, row tiling) and the padded row
// count; the host groups problems accordingly. The grid is launched cooperatively (co-residency: the
// consumers spin on queue slots that only other CTAs fill).
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
    double *partials; // part.nparts rows of red_stride doubles
    unsigned int *ticket;
    FitDevice *fit;
    double svd_eps;
    int ld, S, ldp, red_stride;
    int ntiles;
    TilePartition part; // canonical partition (one partial row per part)
    int items_per_cta;  // work-item sizing target (ctx option queue_items_per_cta)
    int jac_full;
    // set by the finisher for every evaluation (the consumers copy the descriptor with ld.global.cg):
    int parts_per_item, nitems; // work items of the current evaluation: item j = parts [j*ppi, (j+1)*ppi)
    int cdst;           // coefficient buffer the current evaluation writes
    int pad_;
};

// A queue slot is ONE 64-bit word, written with a single st.release and read with a single ld.acquire:
// [63:32] sequence = queue index + 1 once the item is valid, [31:12] fit, [11:0] item of the fit's evaluation.
typedef unsigned long long QueueItem;
constexpr int QUEUE_ITEM_BITS = 12, QUEUE_FIT_BITS = 20;
__host__ __device__ inline QueueItem queue_pack(unsigned long long idx, int fit, int item)
{
    return ((idx + 1ull) << 32) | ((unsigned long long)(unsigned)fit << QUEUE_ITEM_BITS) | (unsigned long long)(unsigned)item;
}

struct QueueCtl {
    unsigned long long head; // next item index to claim
    unsigned long long tail; // items reserved so far
    int fits_left;
    int error;
    QueueItem *items;
    unsigned int cap;
    unsigned long long *dbg; // optional (queue_dbg option): per CTA 16 accumulators in ns / counts, see QDBG_*
};
enum { QDBG_ITEMS = 0, QDBG_CLAIM = 1, QDBG_FRAG = 2, QDBG_STREAM = 3, QDBG_PUBLISH = 4, QDBG_FINISH = 5, QDBG_NFINISH = 6, QDBG_TOTAL = 7,
       QDBG_F_FOLD = 8, QDBG_F_LM = 9, QDBG_F_BASIS = 10, QDBG_F_FACTOR = 11, QDBG_F_PUSH = 12, QDBG_SLOTS = 16 };

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
    constexpr int NGPU = TileAcc<N, P>::NG + P + TileAcc<N, P>::NU;
    constexpr int NVF = 1 + NGPU; // values of a partial row
    static_assert(NPV + 1 <= CT && N <= 4, "one DMMA row block / k block only");
    static_assert(KSTEPS % 8 == 0, "whole panel rows per thread");
    static_assert(THREADS >= 64 + VP_MAX_Q * VP_MAX_Q, "fused_assemble / panel_hh_factor's small-output writers");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STREAM_MAX_STAGES];
    __shared__ __align__(16) double part[NWARPS * 64];
    __shared__ __align__(16) double bu[64];
    __shared__ double rinv_s[N * N];
    __shared__ double fold_scratch[16 * (THREADS / 16)];
    __shared__ double wsum_s[3][NWARPS];   // [0], [1]: parts that end inside an item (double-buffered); [2]: its last part
    __shared__ double gv_s[3][CT * NGPU];
    __shared__ double sums_s[64];
    __shared__ double Msh[P > 0 ? P * P : 1];
    __shared__ double red[2][NWARPS * KMAX];
    __shared__ double top[N][NPV];
    __shared__ double alpha_s[VP_MAX_Q];
    __shared__ SmallSvd svd_s;
    __shared__ LmEval ev_s;
    __shared__ __align__(8) FitDevice fd_s;
    __shared__ __align__(8) QueueFit qf_s; // the claimed fit's descriptor: ONE cooperative L2 read per item instead of chains of dependent loads
    __shared__ int is_last, item_fit, item_idx, more_s, nonfinite_s, push_n;
    __shared__ unsigned long long fin_acc[8]; // queue_dbg: finisher sub-phases
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
    for (int i = tid; i < (int)(sizeof(LmEval) / 8); i += THREADS) reinterpret_cast<unsigned long long *>(&ev_s)[i] = 0ull;
    __syncthreads();
    uint32_t phase_bits = 0;
    const bool dbg_on = ctl->dbg != nullptr;

    // Panel of fit k at the parameters in alpha_s -> HBM, then publish the items of the evaluation.
    // Executed by the whole CTA; all TMA stages must be idle (stage 0 is the staging area). xi / wi:
    // this thread's rows of the fit's x and w.
    auto start_evaluation = [&](const int k, const QueueFit *qf, const double (&xi)[RPT], const double (&wi)[RPT], const int cdst) {
        const ModelDesc &md = qf->md;
        const unsigned long long ts0 = dbg_on ? global_timer_ns() : 0ull;
        // the number of fits still running only sizes the items: requested now with a relaxed load, looked at after
        // the panel (an L2 round trip off the serial path of the finisher)
        int active = 1;
        if (tid == 0) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(active) : "l"(&ctl->fits_left));
        {
            double pa[RPT][NPV], pd0[RPT][P > 0 ? P : 1];
            const int bad = panel_eval_staged<N, P, RPT, THREADS>(md, xi, wi, alpha_s, staging, lds, pa, pd0);
            if (dbg_on && tid == 0) fin_acc[2] += global_timer_ns() - ts0;
            panel_hh_factor<double, N, P, RPT, THREADS>(md, pa, pd0, bad, alpha_s, qf->svd_eps, qf->ldp, qf->Pq, qf->small, red, top,
                                                        nullptr, &svd_s);
        }
        if (sizeof(TY) != sizeof(double)) {
            // the f64 staging overlaid the float slots: restore the zero pad rows [ld, lds) of every slot
            __syncthreads();
            const int ld0 = qf->ld;
            for (int slot = tid; slot < nst * CT; slot += THREADS)
                for (int r = ld0; r < lds; ++r) tiles[(size_t)slot * lds + r] = (TY)0;
        }
        fence_proxy_async_smem(); // generic writes to the staging area before later bulk copies into it
        if (tid == 0) {
            // items of this evaluation: about items_per_cta items per CTA over the fits still running
            if (active < 1) active = 1;
            const int nparts = qf->part.nparts;
            int want = (qf->items_per_cta * (int)gridDim.x + active - 1) / active;
            want = want < 1 ? 1 : (want > nparts ? nparts : want);
            int ppi = (nparts + want - 1) / want;
            if (qf->items_per_cta < 0) ppi = -qf->items_per_cta > nparts ? nparts : -qf->items_per_cta; // fixed size (diagnostics)
            push_n = (nparts + ppi - 1) / ppi;
            fits[k].parts_per_item = ppi;
            fits[k].nitems = push_n;
            fits[k].cdst = cdst;
            *qf->ticket = 0u;
            // reserve the queue slots now: the round trip of the atomic overlaps the fence below (the consumers look
            // at the slot words, which are written after the fence, not at the tail)
            push_base = atomicAdd(&ctl->tail, (unsigned long long)push_n);
        }
        __threadfence(); // panel, small outputs, cdst, ticket and (finisher) the stored LM state before the items
        __syncthreads();
        const unsigned long long ts1 = dbg_on ? global_timer_ns() : 0ull;
        if (dbg_on && tid == 0) fin_acc[3] += ts1 - ts0;
        const int nitems = push_n;
        // publish the items: one st.release per slot (the fence + barrier above ordered everything the consumers
        // will read before these stores)
        {
            const unsigned long long base = push_base;
            for (int j = tid; j < nitems; j += THREADS)
                st_release_gpu_u64(&ctl->items[(base + j) % ctl->cap], queue_pack(base + j, k, j));
        }
        __syncthreads();
        if (dbg_on && tid == 0) { fin_acc[4] += global_timer_ns() - ts1; fin_acc[5] += 1; }
    };

    auto load_xw = [&](const QueueFit *qf, double (&xi)[RPT], double (&wi)[RPT]) {
        const int m = qf->md.m;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = tid + r * THREADS;
            const bool in = i < m;
            xi[r] = in ? (double)static_cast<const TY *>(qf->x)[i] : 0.0;
            wi[r] = in ? (qf->w ? (double)static_cast<const TY *>(qf->w)[i] : 1.0) : 0.0;
        }
    };

    // Fold the part rows of fit k, advance its lmder state machine, and either start the next
    // evaluation or retire the fit. Executed by a whole CTA with an idle TMA ring.
    auto finish_evaluation = [&](const int k) {
        const QueueFit *qf = &qf_s; // the finisher just processed an item of fit k
        const int q = qf->md.q;
        const unsigned long long tf0 = dbg_on ? global_timer_ns() : 0ull;
        // everything the LM step and the next panel need that does not depend on the fold is requested
        // first, so that its L2 latency overlaps the fold: the LM state, M, x and w
        // (into REGISTERS: a store to shared memory would stall the warp until the load has returned, before the
        // fold's own loads are even issued)
        unsigned long long *fw = reinterpret_cast<unsigned long long *>(qf->fit);
        unsigned long long *lw = reinterpret_cast<unsigned long long *>(&fd_s);
        constexpr int FWPT = (FIT_WORDS + THREADS - 1) / THREADS;
        unsigned long long fwreg[FWPT];
#pragma unroll
        for (int t = 0; t < FWPT; ++t) fwreg[t] = (tid + t * THREADS < FIT_WORDS) ? __ldcg(fw + tid + t * THREADS) : 0ull;
        const double m_reg = (tid < P * P) ? __ldcg(&qf->small->M[(tid / (P > 0 ? P : 1)) * VP_MAX_P + (tid % (P > 0 ? P : 1))]) : 0.0;
        const int nonfinite_reg = (tid == 0) ? __ldcg(&qf->small->nonfinite) : 0;
        double xi[RPT], wi[RPT];
        load_xw(qf, xi, wi);
        // rinv_s still holds Rinv of this fit's current panel (loaded with the item's fragments)
        fold_rows(qf->partials, qf->red_stride, qf->part.nparts, NVF, sums_s, fold_scratch);
#pragma unroll
        for (int t = 0; t < FWPT; ++t)
            if (tid + t * THREADS < FIT_WORDS) lw[tid + t * THREADS] = fwreg[t];
        if (tid < P * P) Msh[tid] = m_reg;
        if (tid == 0) {
            nonfinite_s = nonfinite_reg;
            *qf->ticket = 0u; // re-arm: the next evaluation of this problem may be a single-evaluation launch (vp_set_params)
        }
        __syncthreads();
        fused_assemble<N, P>(sums_s, Msh, P, rinv_s, qf->md.e_basis, qf->md.e_param, q, qf->jac_full, nonfinite_s, &ev_s);
        const unsigned long long tf1 = dbg_on ? global_timer_ns() : 0ull;
        if (tid == 0) {
            const bool more = lm_advance(fd_s.st, fd_s.cfg, ev_s);
            if (fd_s.st.last_accepted) fd_s.cur ^= 1;
            if (fd_s.evals < 48) {
                double *tr = fd_s.trace + 4 * fd_s.evals;
                tr[0] = sqrt(ev_s.rnorm2); tr[1] = fd_s.st.par; tr[2] = fd_s.st.delta; tr[3] = fd_s.st.last_accepted;
            }
            fd_s.evals += 1;
            more_s = more ? 1 : 0;
        }
        __syncthreads();
        const unsigned long long tf2 = dbg_on ? global_timer_ns() : 0ull;
        if (fd_s.st.last_accepted) { // the evaluation becomes the accepted one
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(&fd_s.accepted);
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&ev_s);
            for (int i = tid; i < (int)(sizeof(LmEval) / 8); i += THREADS) dst[i] = src[i];
        }
        if (tid < VP_MAX_Q) alpha_s[tid] = tid < q ? fd_s.st.x_trial[tid] : 0.0;
        __syncthreads();
        for (int i = tid; i < FIT_WORDS; i += THREADS) fw[i] = lw[i];
        if (dbg_on && tid == 0) { fin_acc[0] += tf1 - tf0; fin_acc[1] += tf2 - tf1; }
        if (more_s) {
            start_evaluation(k, qf, xi, wi, fd_s.cur ^ 1); // fences the state stores before the items
        } else {
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicSub(&ctl->fits_left, 1);
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
    for (int k = blockIdx.x; k < nfits; k += gridDim.x) {
        const QueueFit *qf = &fits[k];
        if (tid < VP_MAX_Q) alpha_s[tid] = tid < qf->md.q ? __ldcg(&qf->fit->st.x_trial[tid]) : 0.0;
        double xi[RPT], wi[RPT];
        load_xw(qf, xi, wi);
        const int cdst = __ldcg(&qf->fit->cur) ^ 1;
        __syncthreads();
        start_evaluation(k, qf, xi, wi, cdst);
    }

    unsigned long long dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // tid 0, only when ctl->dbg is set
    const unsigned long long t_kernel0 = dbg_on ? global_timer_ns() : 0ull;
    for (;;) {
        // ---- claim an item ----------------------------------------------------------------------
        const unsigned long long t_a = dbg_on ? global_timer_ns() : 0ull;
        if (tid == 0) {
            const unsigned long long idx = atomicAdd(&ctl->head, 1ull);
            const QueueItem *it = &ctl->items[idx % ctl->cap];
            const unsigned long long t0 = global_timer_ns();
            unsigned long long word = 0ull;
            int got = 0;
            for (;;) {
                word = ld_acquire_gpu_u64(it);
                if ((word >> 32) == ((idx + 1ull) & 0xffffffffull)) { got = 1; break; }
                if (ld_acquire_gpu_s32(&ctl->fits_left) <= 0 && idx >= ld_acquire_gpu_u64(&ctl->tail)) break;
                if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) { ctl->error = 1; break; }
            }
            item_fit = got ? (int)((word >> QUEUE_ITEM_BITS) & ((1u << QUEUE_FIT_BITS) - 1u)) : -1;
            item_idx = got ? (int)(word & ((1u << QUEUE_ITEM_BITS) - 1u)) : 0;
        }
        __syncthreads();
        const int k = item_fit, item = item_idx;
        if (k < 0) break;
        {
            static_assert(sizeof(QueueFit) % 8 == 0, "QueueFit is copied in 8-byte words");
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&fits[k]);
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(&qf_s);
            for (int i = tid; i < (int)(sizeof(QueueFit) / 8); i += THREADS) dst[i] = __ldcg(src + i);
        }
        __syncthreads();
        const unsigned long long t_b = dbg_on ? global_timer_ns() : 0ull;
        const QueueFit *qf = &qf_s;
        const int ld = qf->ld, S = qf->S, ldp = qf->ldp;
        const TY *Yk = static_cast<const TY *>(qf->Y);
        const TilePartition tpn = qf->part;
        const int p_begin = item * qf->parts_per_item;
        const int p_end = min(tpn.nparts, p_begin + qf->parts_per_item);
        const int t_begin = part_first_tile(tpn, p_begin);
        const int my = part_first_tile(tpn, p_end) - t_begin;
        const int nitems = qf->nitems;
        TY *Cout = static_cast<TY *>(qf->cdst ? qf->C1 : qf->C0);
        int ebasis[P > 0 ? P : 1];
#pragma unroll
        for (int e2 = 0; e2 < P; ++e2) ebasis[e2] = qf->md.e_basis[e2];

        // ---- TMA producer for this item ------------------------------------------------------------
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

        // ---- stream the item part by part (tile math and part fold of dmma_tile.cuh) -------------------
        TileAcc<N, P> acc;
        acc.clear();
        int st = 0;
        int cur_part = p_begin;
        int part_last = part_first_tile(tpn, p_begin + 1) - t_begin - 1; // item-relative index of the current part's last tile
        int pending_row = -1, pending_buf = 0, stage_buf = 0;
        for (int i = 0; i < my; ++i) {
            const int col0 = (t_begin + i) * CT;
            const int nc = min(CT, S - col0);
            const TY *tp = tiles + (size_t)st * stage_elems;
            mbar_wait(&full_bar[st], (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
            if (++st == nst) st = 0;
            dmma_tile<TY, N, P, KSTEPS, NWARPS, EXACT>(tp, lds, nc, col0, a1, a2, rinv_s, part, bu, Cout, ebasis, acc, [&]() {
                if (i >= 1 && next_i < my) issue(); // refill the stage tile i-1 used
                if (pending_row >= 0) {             // the row of the part that ended with the previous tile
                    if (warp == NWARPS - 1) {
                        // no fence here: the row is ordered before the item's ticket by the barriers in between
                        // and the release of the ticket atomic (cumulative over the CTA's earlier writes)
                        part_flush<N, P, CT, NWARPS>(wsum_s[pending_buf], gv_s[pending_buf],
                                                     qf->partials + (size_t)pending_row * qf->red_stride, lane);
                    }
                    pending_row = -1;
                }
            });
            if (i == part_last && i + 1 < my) { // a part ends inside the item: stage its fold, start the next part
                part_stage<N, P, CT, NWARPS>(acc, wsum_s[stage_buf], gv_s[stage_buf]);
                acc.clear();
                pending_row = cur_part;
                pending_buf = stage_buf;
                stage_buf ^= 1;
                ++cur_part;
                part_last = part_first_tile(tpn, cur_part + 1) - t_begin - 1;
            }
        }
        __syncthreads(); // every warp is done with every stage: the ring is idle
        const unsigned long long t_d = dbg_on ? global_timer_ns() : 0ull;

        // ---- the item's last part -> its row; ticket; the last item's CTA finishes the evaluation -------
        part_stage<N, P, CT, NWARPS>(acc, wsum_s[2], gv_s[2]);
        __syncthreads();
        if (warp == 0) {
            part_flush<N, P, CT, NWARPS>(wsum_s[2], gv_s[2], qf->partials + (size_t)(p_end - 1) * qf->red_stride, lane);
            __syncwarp();
            if (lane == 0) {
                const unsigned int prev = atom_add_acq_rel_gpu_u32(qf->ticket, 1u);
                is_last = (prev == (unsigned int)nitems - 1u);
            }
        }
        __syncthreads();
        const unsigned long long t_e = dbg_on ? global_timer_ns() : 0ull;
        if (is_last) {
            __threadfence();
            finish_evaluation(k);
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
        unsigned long long *o = ctl->dbg + (size_t)blockIdx.x * QDBG_SLOTS;
        for (int i = 0; i < 8; ++i) o[i] = dacc[i];
        o[QDBG_F_FOLD] = fin_acc[0]; o[QDBG_F_LM] = fin_acc[1]; o[QDBG_F_BASIS] = fin_acc[2]; o[QDBG_F_FACTOR] = fin_acc[3];
        o[QDBG_F_PUSH] = fin_acc[4]; o[13] = fin_acc[5]; o[14] = 0; o[15] = 0;
    }
}

} // namespace vp
