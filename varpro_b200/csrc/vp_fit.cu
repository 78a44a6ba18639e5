// vp_fit.cu -- the LM drivers of the C ABI (include/varpro_b200.h): vp_fit (persistent whole-fit kernel, work-queue
// kernel, CUDA-graph loop, host loop), vp_fit_many (work-queue kernel) and the vp_comm communicator of
// column-sharded global fits.
#include "vp_internal.h"
#include "fit_queue_kernel.cuh"

using namespace vp;

// ----------------------------------------------------------------------------
// vp_comm: column-sharded global fit over the GPUs of one box
// ----------------------------------------------------------------------------
extern "C" int vp_comm_create(vp_ctx *ctx, int rank, int world, vp_comm **out, void *local_handle_out)
{
    if (!ctx || !out || !local_handle_out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world)
        return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_comm_create: need 0 <= rank < world <= 8");
    static_assert(sizeof(cudaIpcMemHandle_t) == VP_COMM_HANDLE_BYTES, "handle size");
    cudaSetDevice(ctx->device);
    vp_comm *c = new (std::nothrow) vp_comm();
    if (!c) return VP_ERR_OUT_OF_MEMORY;
    c->ctx = ctx; c->world = world; c->rank = rank;
    // plain cudaMalloc (not the pool): the allocation is exported through CUDA IPC
    cudaError_t e = cudaMalloc(&c->local_box, sizeof(CommMailbox));
    if (e == cudaSuccess) e = cudaMemset(c->local_box, 0, sizeof(CommMailbox));
    if (e == cudaSuccess) e = cudaMalloc(&c->epoch, 256);
    if (e == cudaSuccess) e = cudaMemset(c->epoch, 0, 256);
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&h, c->local_box);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(c->local_box); cudaFree(c->epoch);
        delete c;
        return vp_fail(ctx, VP_ERR_COMM, std::string("vp_comm_create: ") + cudaGetErrorString(e));
    }
    c->error = reinterpret_cast<int *>(c->epoch + 8);
    memcpy(local_handle_out, &h, sizeof(h));
    c->peer[rank] = c->local_box;
    if (world == 1) { // nothing to map
        c->args.world = 1; c->args.rank = 0; c->args.epoch = c->epoch; c->args.error = c->error;
        c->args.box[0] = c->local_box;
        c->connected = true;
    }
    *out = c;
    return VP_OK;
}

extern "C" int vp_comm_connect(vp_comm *c, const void *all_handles)
{
    if (!c || !all_handles) return VP_ERR_INVALID_ARGUMENT;
    if (c->connected) return VP_OK;
    vp_ctx *ctx = c->ctx;
    cudaSetDevice(ctx->device);
    const unsigned char *hs = static_cast<const unsigned char *>(all_handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)r * VP_COMM_HANDLE_BYTES, sizeof(h));
        cudaError_t e = cudaIpcOpenMemHandle(&c->peer[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return vp_fail(ctx, VP_ERR_COMM, "vp_comm_connect: cannot map the mailbox of rank " + std::to_string(r) + ": " +
                                              cudaGetErrorString(e));
    }
    c->ipc_peers = true;
    c->args.world = c->world; c->args.rank = c->rank; c->args.epoch = c->epoch; c->args.error = c->error;
    for (int r = 0; r < c->world; ++r) c->args.box[r] = static_cast<CommMailbox *>(c->peer[r]);
    c->connected = true;
    return VP_OK;
}

// One process driving several GPUs (or several contexts of one GPU, one host thread each): the mailboxes are ordinary
// device pointers of this process, no IPC handles needed. comms: the `world` communicators in rank order.
extern "C" int vp_comm_connect_local(vp_comm **comms, int world)
{
    if (!comms || world < 1 || world > COMM_MAX_WORLD) return VP_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < world; ++r)
        if (!comms[r] || comms[r]->world != world || comms[r]->rank != r) return VP_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < world; ++r) {
        vp_comm *c = comms[r];
        if (c->connected) continue;
        vp_ctx *ctx = c->ctx;
        cudaSetDevice(ctx->device);
        for (int s = 0; s < world; ++s) {
            const int peer_dev = comms[s]->ctx->device;
            if (peer_dev != ctx->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, ctx->device, peer_dev);
                if (!can) return vp_fail(ctx, VP_ERR_COMM, "vp_comm_connect_local: no peer access between the devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return vp_fail(ctx, VP_ERR_COMM, std::string("vp_comm_connect_local: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            c->peer[s] = comms[s]->local_box;
        }
        c->ipc_peers = false;
        c->args.world = world; c->args.rank = r; c->args.epoch = c->epoch; c->args.error = c->error;
        for (int s = 0; s < world; ++s) c->args.box[s] = static_cast<CommMailbox *>(c->peer[s]);
        c->connected = true;
    }
    return VP_OK;
}

extern "C" int vp_comm_destroy(vp_comm *c)
{
    if (!c) return VP_OK;
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->ipc_peers && r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local_box);
    cudaFree(c->epoch);
    delete c;
    return VP_OK;
}

extern "C" int vp_problem_set_comm(vp_problem *pr, vp_comm *c)
{
    if (!pr) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = pr->ctx;
    if (c && (c->ctx != ctx || !c->connected)) return vp_fail(ctx, VP_ERR_COMM, "vp_problem_set_comm: communicator not connected / wrong context");
    if (c && pr->plan_fit < 0)
        return vp_fail(ctx, VP_ERR_COMM, "column-sharded fits need the fused fp64 kernel (no instantiation for this model shape)");
    pr->comm = c;
    // the evaluation made at creation covered this rank's columns only: redo it collectively
    int rc = vp_refresh_cached_evaluation(pr);
    if (rc == VP_OK && !pr->cached) rc = vp_fail(ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
    return rc;
}

// ----------------------------------------------------------------------------
// LM configuration and reports
// ----------------------------------------------------------------------------
// vp_lm_options: a NEGATIVE (or NaN) tolerance / stepbound / patience selects the crate default; 0 is a legal value
// that passes through (ftol = xtol = gtol = 0 disable the respective criterion in levenberg-marquardt 0.14).
void vp_lm_config_from_options(int dtype, int q, const vp_lm_options *opt, LmConfig &cfg)
{
    const double eps = dtype == VP_F32 ? (double)FLT_EPSILON : DBL_EPSILON;
    auto pick = [](bool have, double v, double dflt) { return (have && v >= 0.0) ? v : dflt; }; // NaN >= 0 is false
    cfg.epsmch = eps;
    cfg.ftol = pick(opt != nullptr, opt ? opt->ftol : 0.0, 30.0 * eps);
    cfg.xtol = pick(opt != nullptr, opt ? opt->xtol : 0.0, 30.0 * eps);
    cfg.gtol = pick(opt != nullptr, opt ? opt->gtol : 0.0, 30.0 * eps);
    cfg.stepbound = (opt && opt->stepbound > 0.0) ? opt->stepbound : 100.0; // must be positive
    const int patience = (opt && opt->patience > 0) ? opt->patience : 100;
    cfg.maxfev = patience * (q + 1);
    cfg.scale_diag = (opt && opt->scale_diag >= 0) ? (opt->scale_diag != 0) : 1;
}

static void fill_report(const LmState &st, vp_fit_report *rep)
{
    rep->termination = st.termination;
    rep->number_of_evaluations = st.nfev;
    rep->objective_function = 0.5 * st.fnorm * st.fnorm;
    rep->successful = lm_successful(st.termination) ? 1 : 0;
}

static void trace_fit(const char *who, const FitDevice *fh)
{
    for (int i = 0; i < fh->evals && i < 48; ++i)
        fprintf(stderr, "[vp_fit %s] eval %d fnorm_trial=%.6e par=%.3e delta=%.3e acc=%d\n", who, i + 2, fh->trace[4 * i],
                fh->trace[4 * i + 1], fh->trace[4 * i + 2], (int)fh->trace[4 * i + 3]);
}

static int ensure_timeline_buffer(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    if (ctx->opt.dbg_fit && !pr->dbg) { // in-kernel timeline of the last evaluation (vp_debug_timeline reads it)
        const size_t nd = ((size_t)pr->max_grid + 1) * VP_DBG_SLOTS;
        VP_CUDA(ctx, cudaMalloc(&pr->dbg, nd * sizeof(unsigned long long)));
        VP_CUDA(ctx, cudaMemset(pr->dbg, 0, nd * sizeof(unsigned long long)));
    }
    return VP_OK;
}

// ----------------------------------------------------------------------------
// driver 1: the persistent whole-fit kernel (fit_kernel_dmma in fit mode): ONE cooperative launch per fit
// ----------------------------------------------------------------------------
static int fit_persistent(vp_problem *pr, LmState &st, const LmConfig &cfg)
{
    vp_ctx *ctx = pr->ctx;
    const int q = pr->model->md.q;
    cudaSetDevice(ctx->device);
    int rc = ensure_timeline_buffer(pr);
    if (rc != VP_OK) return rc;
    FitDevice *fh = pr->fit_host;
    memset(fh, 0, VP_FIT_BLOCK_BYTES); // the state and the (zeroed) control word behind it: one copy in, one copy out
    fh->st = st; fh->cfg = cfg; fh->accepted = pr->eval; fh->cur = pr->cur; fh->evals = 0;
    cudaStream_t stream = ctx->stream;
    VP_CUDA(ctx, cudaMemcpyAsync(pr->fit_dev, fh, VP_FIT_BLOCK_BYTES, cudaMemcpyHostToDevice, stream));
    rc = vp_launch_fused(pr, 0, /*fit_mode=*/true);
    if (rc != VP_OK) return rc;
    VP_CUDA(ctx, cudaMemcpyAsync(fh, pr->fit_dev, VP_FIT_BLOCK_BYTES, cudaMemcpyDeviceToHost, stream));
    const FitCtl *ch = reinterpret_cast<const FitCtl *>(reinterpret_cast<const unsigned char *>(fh) + VP_FIT_CTL_OFFSET);
    VP_CUDA(ctx, cudaStreamSynchronize(stream));
    if (ch->error) {
        pr->cached = false; // the coefficient buffers no longer belong to pr->alpha
        return vp_fail(ctx, VP_ERR_CUDA, "vp_fit: a grid-wide wait timed out inside the persistent fit kernel");
    }
    rc = vp_comm_check(pr);
    if (rc != VP_OK) { pr->cached = false; return rc; }
    st = fh->st;
    pr->cur = fh->cur;
    pr->eval = fh->accepted;
    for (int k = 0; k < q; ++k) pr->alpha[k] = st.x[k];
    if (ctx->opt.trace) trace_fit("persistent", fh);
    return VP_OK;
}

// ----------------------------------------------------------------------------
// driver 2: the work-queue kernel (fit_queue_kernel): all fits of a group on one persistent grid
// ----------------------------------------------------------------------------
// The work-queue kernel instantiation a problem can use (index into the table, -1 = none) and the
// padded tile-column stride (elements) that goes with it.
static int queue_kernel_for(const vp_problem *pr, int *lds_out)
{
    const vp_model *mo = pr->model;
    if (mo->hosteval || pr->comm || !pr->cached) return -1;
    const ModelDesc &md = mo->md;
    int lds = 0; // column stride of a tile slot (vp_tile_lds)
    const std::vector<QueueKernelEntry> &tab = vp_kernel_tables().queue;
    int pick = -1;
    for (size_t i = 0; i < tab.size(); ++i) {
        const QueueKernelEntry &k = tab[i];
        if (k.dtype != mo->dtype || k.n != md.n || k.p != md.p) continue;
        const int rows = 4 * k.ksteps * k.nwarps;
        if (rows < mo->ld) continue;
        const int lds_k = vp_tile_lds(mo->dtype, mo->ld, rows, k.exact);
        if (lds_k < 0) continue;
        if (pick < 0 || vp_better_tiling(k.ksteps, k.nwarps, k.exact, tab[(size_t)pick].ksteps, tab[(size_t)pick].nwarps,
                                         tab[(size_t)pick].exact, pr->ctx->opt.fit_warps)) {
            pick = (int)i;
            lds = lds_k;
        }
    }
    if (lds_out) *lds_out = lds;
    return pick;
}

// One launch of fit_queue_kernel for a group of problems that share the kernel instantiation and the
// padded row count. states[i] has been advanced past the cached evaluation at the starting point.
// Returns VP_OK after the fits completed and their final states were adopted; VP_ERR_UNSUPPORTED_BASIS if
// the group has no work-queue kernel (the caller falls back to the per-fit drivers).
static int fit_queue_group(vp_ctx *ctx, const std::vector<vp_problem *> &prs, std::vector<LmState> &states,
                           const std::vector<LmConfig> &cfgs)
{
    const int K = (int)prs.size();
    if (K >= (1 << QUEUE_FIT_BITS)) return VP_ERR_UNSUPPORTED_BASIS; // does not fit a queue slot: per-fit drivers
    vp_problem *p0 = prs[0];
    int lds = 0;
    const int qidx = queue_kernel_for(p0, &lds);
    if (qidx < 0) return VP_ERR_UNSUPPORTED_BASIS;
    const QueueKernelEntry *qk = &vp_kernel_tables().queue[(size_t)qidx];
    const size_t es = vp_esize(p0->model->dtype);
    const int prow = 4 * qk->ksteps * qk->nwarps;
    const size_t stage_bytes = (size_t)DMMA_CT * lds * es;
    // launch configuration of (kernel, stage size), cached: the attribute / occupancy queries cost tens of
    // microseconds each and this function sits inside callers' timed regions
    struct QueuePlan { const void *fn; size_t stage_bytes; int device; int nst; size_t smem; int occ; };
    static thread_local std::vector<QueuePlan> plans;
    const QueuePlan *plan = nullptr;
    for (const QueuePlan &qp : plans)
        if (qp.fn == qk->fn && qp.stage_bytes == stage_bytes && qp.device == ctx->device) plan = &qp;
    if (!plan) {
        cudaFuncAttributes fa{};
        VP_CUDA(ctx, cudaFuncGetAttributes(&fa, qk->fn));
        QueuePlan qp{qk->fn, stage_bytes, ctx->device, 0, 0, 0};
        if (fa.sharedSizeBytes + 1024 + 2 * stage_bytes <= 227 * 1024) {
            qp.nst = (int)((227 * 1024 - fa.sharedSizeBytes - 1024) / stage_bytes);
            if (qp.nst > STREAM_MAX_STAGES) qp.nst = STREAM_MAX_STAGES;
            if (ctx->opt.stream_stages >= 2 && qp.nst > ctx->opt.stream_stages) qp.nst = ctx->opt.stream_stages;
            qp.smem = (size_t)qp.nst * stage_bytes;
            VP_CUDA(ctx, vp_ensure_dynamic_smem(ctx->device, qk->fn, qp.smem));
            VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&qp.occ, qk->fn, qk->nwarps * 32, qp.smem));
        }
        plans.push_back(qp);
        plan = &plans.back();
    }
    if (plan->occ < 1) return VP_ERR_UNSUPPORTED_BASIS;
    const int nst = plan->nst, occ = plan->occ;
    const size_t smem = plan->smem;
    const long long grid = (long long)ctx->sm_count * occ;

    std::vector<QueueFit> hq((size_t)K);
    long long total_items = 0;
    for (int i = 0; i < K; ++i) {
        vp_problem *pr = prs[(size_t)i];
        const ModelDesc &md = pr->model->md;
        QueueFit &f = hq[(size_t)i];
        memset(&f, 0, sizeof(f));
        f.md = md;
        f.Y = pr->Yw; f.C0 = pr->C[0]; f.C1 = pr->C[1];
        f.x = pr->model->x_dev; f.w = pr->w_dev;
        // the kernel's panel buffer is f64 with at least `prow` zero-padded rows per column
        if (pr->model->dtype == VP_F64 && pr->ldp >= prow) {
            f.Pq = (double *)pr->Pq; f.ldp = pr->ldp;
        } else {
            const int ldp64 = ((prow > lds ? prow : lds) + 3) / 4 * 4;
            if (!pr->Pq64 || pr->ldp64 < ldp64) {
                DEV_FREE(ctx, pr->Pq64);
                pr->Pq64 = nullptr;
                VP_CUDA(ctx, DEV_ALLOC(ctx, &pr->Pq64, sizeof(double) * (size_t)ldp64 * (md.n + md.p + 1)));
                pr->ldp64 = ldp64;
            }
            f.Pq = pr->Pq64; f.ldp = pr->ldp64;
        }
        f.small = pr->small; f.partials = pr->partials; f.ticket = pr->ticket;
        f.fit = pr->fit_dev; f.svd_eps = pr->rank_tol;
        f.ld = pr->model->ld; f.S = (int)pr->S; f.red_stride = pr->red_stride;
        f.ntiles = (int)((pr->S + DMMA_CT - 1) / DMMA_CT);
        // the partition vp_fit's persistent kernel uses for this problem (one part per CTA of ITS grid)
        int nparts = pr->plan_fit >= 0 ? pr->fit_grid : (f.ntiles < ctx->sm_count ? f.ntiles : ctx->sm_count);
        if (nparts > pr->max_grid) nparts = pr->max_grid;
        if (nparts >= (1 << QUEUE_ITEM_BITS)) nparts = (1 << QUEUE_ITEM_BITS) - 1;
        f.part = make_partition(f.ntiles, nparts);
        // the kernel sizes the work items of every evaluation itself (negative: fixed number of parts per item)
        f.items_per_cta = ctx->opt.queue_parts_per_item > 0 ? -ctx->opt.queue_parts_per_item : ctx->opt.queue_items_per_cta;
        f.parts_per_item = f.part.nparts;
        f.nitems = 1;
        f.jac_full = pr->jac_full;
        f.cdst = pr->cur ^ 1;
        total_items += f.part.nparts; // upper bound of the items one evaluation of this fit can have
    }
    // One pinned host block and one device block: [FitDevice x K | QueueCtl | QueueFit x K]. One copy
    // in; one copy out (states + control word).
    const size_t off_ctl = sizeof(FitDevice) * (size_t)K;
    const size_t off_q = (off_ctl + sizeof(QueueCtl) + 255) / 256 * 256;
    const size_t total_bytes = off_q + sizeof(QueueFit) * (size_t)K;
    const unsigned int cap = (unsigned int)(total_items + 2048);
    unsigned char *hb = nullptr, *db = nullptr;
    QueueItem *ditems = nullptr;
    cudaError_t e = HOST_ALLOC(ctx, &hb, total_bytes);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &db, total_bytes);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &ditems, sizeof(QueueItem) * (size_t)cap);
    if (e != cudaSuccess) {
        HOST_FREE(ctx, hb); DEV_FREE(ctx, db); DEV_FREE(ctx, ditems);
        return vp_fail(ctx, VP_ERR_OUT_OF_MEMORY, std::string("vp_fit_many (queue): ") + cudaGetErrorString(e));
    }
    FitDevice *hstate = reinterpret_cast<FitDevice *>(hb), *dstate = reinterpret_cast<FitDevice *>(db);
    QueueCtl *hctl = reinterpret_cast<QueueCtl *>(hb + off_ctl), *dctl = reinterpret_cast<QueueCtl *>(db + off_ctl);
    QueueFit *hqp = reinterpret_cast<QueueFit *>(hb + off_q), *dq = reinterpret_cast<QueueFit *>(db + off_q);
    memset(hb, 0, off_q);
    for (int i = 0; i < K; ++i) {
        vp_problem *pr = prs[(size_t)i];
        FitDevice *fh = &hstate[i];
        fh->st = states[(size_t)i]; fh->cfg = cfgs[(size_t)i]; fh->accepted = pr->eval; fh->cur = pr->cur; fh->evals = 0;
        hq[(size_t)i].fit = dstate + i;
        hqp[i] = hq[(size_t)i];
    }
    hctl->head = 0; hctl->tail = 0; hctl->fits_left = K; hctl->error = 0; hctl->items = ditems; hctl->cap = cap;
    unsigned long long *ddbg = nullptr;
    if (ctx->opt.queue_dbg) { // per-CTA phase accumulators (diagnostics; printed to stderr after the launch)
        if (DEV_ALLOC(ctx, &ddbg, sizeof(unsigned long long) * QDBG_SLOTS * (size_t)grid) == cudaSuccess)
            cudaMemsetAsync(ddbg, 0, sizeof(unsigned long long) * QDBG_SLOTS * (size_t)grid, ctx->stream);
        else
            ddbg = nullptr;
    }
    hctl->dbg = ddbg;
    e = cudaMemcpyAsync(db, hb, total_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ditems, 0, sizeof(QueueItem) * (size_t)cap, ctx->stream);
    if (e == cudaSuccess) {
        int nf = K, lds_arg = lds, nst_arg = nst;
        void *args[] = {(void *)&dctl, (void *)&dq, (void *)&nf, (void *)&lds_arg, (void *)&nst_arg};
        // cooperative: the consumers spin on queue slots that only other CTAs fill, so the whole grid must be
        // co-resident (an ordinary launch next to other work of the process could leave it partly resident)
        e = cudaLaunchCooperativeKernel(qk->fn, dim3((unsigned)grid), dim3(qk->nwarps * 32), args, smem, ctx->stream);
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hb, db, off_ctl + sizeof(QueueCtl), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    const int qerr = hctl->error;
    if (ddbg && e == cudaSuccess) {
        std::vector<unsigned long long> hd(QDBG_SLOTS * (size_t)grid);
        if (cudaMemcpy(hd.data(), ddbg, sizeof(unsigned long long) * hd.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
            double acc[QDBG_SLOTS] = {0};
            for (long long b = 0; b < grid; ++b)
                for (int i = 0; i < QDBG_SLOTS; ++i) acc[i] += (double)hd[(size_t)b * QDBG_SLOTS + i];
            const double it = acc[QDBG_ITEMS] > 0 ? acc[QDBG_ITEMS] : 1, nfin = acc[QDBG_NFINISH] > 0 ? acc[QDBG_NFINISH] : 1;
            const double nstart = acc[13] > 0 ? acc[13] : 1;
            fprintf(stderr, "[vp queue dbg] fits %d items %.0f (%.1f per CTA) | per item: claim %.2f us, fragments %.2f us, "
                            "stream %.2f us, publish %.2f us | finisher %.2f us x %.0f = fold+assemble %.2f, LM step %.2f, basis %.2f, "
                            "panel (basis+factor+store) %.2f, push %.2f | CTA lifetime %.1f us, busy %.1f %%\n",
                    K, acc[QDBG_ITEMS], acc[QDBG_ITEMS] / grid, 1e-3 * acc[QDBG_CLAIM] / it,
                    1e-3 * acc[QDBG_FRAG] / it, 1e-3 * acc[QDBG_STREAM] / it, 1e-3 * acc[QDBG_PUBLISH] / it,
                    1e-3 * acc[QDBG_FINISH] / nfin, acc[QDBG_NFINISH], 1e-3 * acc[QDBG_F_FOLD] / nfin, 1e-3 * acc[QDBG_F_LM] / nfin,
                    1e-3 * acc[QDBG_F_BASIS] / nstart, 1e-3 * acc[QDBG_F_FACTOR] / nstart, 1e-3 * acc[QDBG_F_PUSH] / nstart,
                    1e-3 * acc[QDBG_TOTAL] / grid,
                    100.0 * (acc[QDBG_FRAG] + acc[QDBG_STREAM] + acc[QDBG_PUBLISH] + acc[QDBG_FINISH]) / (acc[QDBG_TOTAL] > 0 ? acc[QDBG_TOTAL] : 1));
        }
    }
    DEV_FREE(ctx, ddbg);
    DEV_FREE(ctx, db); DEV_FREE(ctx, ditems);
    if (e != cudaSuccess || qerr) {
        // the device has overwritten both coefficient buffers of the fits it started: nothing cached is trustworthy
        for (vp_problem *pr : prs) pr->cached = false;
        HOST_FREE(ctx, hb);
        if (e != cudaSuccess) return vp_fail(ctx, VP_ERR_CUDA, std::string("vp_fit_many (queue): ") + cudaGetErrorString(e));
        return vp_fail(ctx, VP_ERR_CUDA, "vp_fit_many: a wait inside the work-queue kernel timed out");
    }
    for (int i = 0; i < K; ++i) {
        vp_problem *pr = prs[(size_t)i];
        FitDevice *fh = &hstate[i];
        states[(size_t)i] = fh->st;
        pr->cur = fh->cur;
        pr->eval = fh->accepted;
        for (int k = 0; k < pr->model->md.q; ++k) pr->alpha[k] = fh->st.x[k];
        if (ctx->opt.trace) trace_fit("queue", fh);
    }
    HOST_FREE(ctx, hb);
    return VP_OK;
}

// ----------------------------------------------------------------------------
// driver 3: CUDA-graph loop for model shapes without a fused kernel: a conditional WHILE node whose body is one
// evaluation, K1 (panel at the trial parameters) -> K2 (streaming reduce; its last CTA advances the lmder state
// machine on the device and sets the loop condition). One graph launch per fit.
// ----------------------------------------------------------------------------
static int ensure_fit_graph(vp_problem *pr)
{
    if (pr->fit_exec) return VP_OK;
    vp_ctx *ctx = pr->ctx;
    int rc = ensure_timeline_buffer(pr);
    if (rc != VP_OK) return rc;
    VP_CUDA(ctx, cudaGraphCreate(&pr->fit_graph, 0));
    VP_CUDA(ctx, cudaGraphConditionalHandleCreate(&pr->fit_cond, pr->fit_graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = pr->fit_cond;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t node;
    VP_CUDA(ctx, cudaGraphAddNode(&node, pr->fit_graph, nullptr, 0, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    VP_CUDA(ctx, cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    const int64_t launches_before = ctx->launches;
    rc = vp_launch_panel(pr);
    if (rc == VP_OK) rc = vp_launch_stream(pr, 0, /*graph_mode=*/true);
    ctx->launches = launches_before; // captured, not launched
    cudaGraph_t captured = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &captured);
    if (rc != VP_OK) return rc;
    if (e != cudaSuccess) return vp_fail(ctx, VP_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    VP_CUDA(ctx, cudaGraphInstantiate(&pr->fit_exec, pr->fit_graph, 0));
    return VP_OK;
}

static int fit_graph_loop(vp_problem *pr, LmState &st, const LmConfig &cfg)
{
    vp_ctx *ctx = pr->ctx;
    const int q = pr->model->md.q;
    cudaSetDevice(ctx->device);
    int rc = ensure_fit_graph(pr);
    if (rc != VP_OK) return rc;
    FitDevice *fh = pr->fit_host;
    memset(fh, 0, sizeof(FitDevice));
    fh->st = st; fh->cfg = cfg; fh->accepted = pr->eval; fh->cur = pr->cur; fh->evals = 0;
    VP_CUDA(ctx, cudaMemcpyAsync(pr->fit_dev, fh, sizeof(FitDevice), cudaMemcpyHostToDevice, ctx->stream));
    VP_CUDA(ctx, cudaGraphLaunch(pr->fit_exec, ctx->stream));
    VP_CUDA(ctx, cudaMemcpyAsync(fh, pr->fit_dev, sizeof(FitDevice), cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    st = fh->st;
    pr->cur = fh->cur;
    pr->eval = fh->accepted;
    for (int k = 0; k < q; ++k) pr->alpha[k] = st.x[k];
    ctx->launches += 2 * (int64_t)fh->evals;
    if (ctx->opt.trace) trace_fit("graph", fh);
    return VP_OK;
}

// ----------------------------------------------------------------------------
// driver 4: host-driven loop (host-evaluated models; the full Jacobian without a fused kernel; fit_mode = host):
// one synchronisation per evaluation
// ----------------------------------------------------------------------------
static int fit_host_loop(vp_problem *pr, LmState &st, const LmConfig &cfg)
{
    const int q = pr->model->md.q;
    bool more = true;
    while (more) {
        const int dst = pr->cur ^ 1;
        int rc = vp_evaluate_sync(pr, st.x_trial, dst);
        if (rc == VP_ERR_NO_CACHED_CALCULATION) { // residuals() is None -> the LM loop stops with a User termination
            st.termination = TERM_USER;
            pr->cached = false;
            break;
        }
        if (rc != VP_OK) return rc;
        LmEval ev;
        vp_evalout_to_lm(*pr->out_host, q, ev);
        more = lm_advance(st, cfg, ev);
        if (pr->ctx->opt.trace)
            fprintf(stderr, "[vp_fit host] nfev=%d fnorm_trial=%.6e fnorm=%.6e par=%.3e delta=%.3e acc=%d\n", st.nfev, sqrt(ev.rnorm2),
                    st.fnorm, st.par, st.delta, st.last_accepted);
        if (st.last_accepted) {
            pr->cur = dst;
            pr->eval = ev;
            for (int k = 0; k < q; ++k) pr->alpha[k] = st.x[k];
        }
    }
    return VP_OK;
}

// ----------------------------------------------------------------------------
// LevMarSolver::fit  (src/solvers/levmar/mod.rs:238-254)
// ----------------------------------------------------------------------------
extern "C" int vp_fit(vp_problem *pr, const vp_lm_options *opt, vp_fit_report *rep)
{
    VP_NVTX("vp_fit");
    if (!pr || !rep) return vp_fail(pr ? pr->ctx : nullptr, VP_ERR_INVALID_ARGUMENT, "vp_fit: problem and report must not be NULL");
    vp_ctx *ctx = pr->ctx;
    const int q = pr->model->md.q;
    LmConfig cfg;
    vp_lm_config_from_options(pr->model->dtype, q, opt, cfg);
    memset(rep, 0, sizeof(*rep));

    LmState st;
    lm_init(st, q, pr->alpha);
    if (!pr->cached) {
        // residuals() is None -> the LM crate stops with a User termination
        rep->termination = VP_TERM_USER;
        rep->number_of_evaluations = 0;
        rep->objective_function = NAN;
        rep->successful = 0;
        return VP_OK;
    }
    // the evaluation at the current parameters is cached (builder / set_params)
    const bool more = lm_advance(st, cfg, pr->eval);
    int rc = VP_OK;
    if (more) {
        const int mode = ctx->opt.fit_mode;
        const bool host_only = pr->model->hosteval || (pr->jac_full && pr->plan_fit < 0);
        if (pr->comm) {
            if (pr->plan_fit < 0 || mode != VP_FITMODE_AUTO)
                return vp_fail(ctx, VP_ERR_COMM, "column-sharded fits run on the persistent fit kernel only");
            rc = fit_persistent(pr, st, cfg);
        } else if (mode == VP_FITMODE_HOST || host_only) {
            rc = fit_host_loop(pr, st, cfg);
        } else if (mode == VP_FITMODE_GRAPH && !pr->jac_full) {
            rc = fit_graph_loop(pr, st, cfg);
        } else if (pr->plan_fit >= 0) {
            rc = fit_persistent(pr, st, cfg);
        } else {
            // no fused kernel for this problem: fp32 problems have a work-queue kernel (panel once per evaluation,
            // fp64 arithmetic); everything else runs the CUDA-graph loop over K1 + K2
            rc = VP_ERR_UNSUPPORTED_BASIS;
            if (queue_kernel_for(pr, nullptr) >= 0) {
                cudaSetDevice(ctx->device);
                std::vector<vp_problem *> prs{pr};
                std::vector<LmState> sts{st};
                std::vector<LmConfig> cfs{cfg};
                rc = fit_queue_group(ctx, prs, sts, cfs);
                if (rc == VP_OK) st = sts[0];
            }
            if (rc == VP_ERR_UNSUPPORTED_BASIS) rc = fit_graph_loop(pr, st, cfg);
        }
    }
    if (rc != VP_OK) return rc;
    fill_report(st, rep);
    return VP_OK;
}

// Many independent fits at once (throughput mode). A single fit on the whole GPU is latency bound: per evaluation
// the panel, the grid-wide fold and the serial LM step cost more than streaming 33.5 MB does. Independent problems
// are therefore fitted TOGETHER: like-shaped problems share ONE persistent grid through a device-side work queue
// (fit_queue_kernel.cuh): every CTA streams items of whichever fit has work, the panel of an evaluation is computed
// once, and the serial phases of one fit hide behind the streaming of the others. Problems without a work-queue
// kernel are fitted one after the other (vp_fit). Results are bitwise those of vp_fit.
extern "C" int vp_fit_many(vp_problem **problems, int64_t n, const vp_lm_options *opt, vp_fit_report *reports,
                           int32_t /*reserved*/)
{
    VP_NVTX("vp_fit_many");
    if (n < 0 || (n > 0 && (!problems || !reports))) return VP_ERR_INVALID_ARGUMENT;
    if (n == 0) return VP_OK;
    for (int64_t i = 0; i < n; ++i)
        if (!problems[i] || problems[i]->ctx != problems[0]->ctx) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = problems[0]->ctx;
    cudaSetDevice(ctx->device);
    // done: 0 = to be fitted by vp_fit, 1 = report filled, 2 = waiting for a queue launch, 3 = failed
    std::vector<char> done((size_t)n, 0);
    std::vector<LmState> all_states((size_t)n);
    std::vector<LmConfig> all_cfgs((size_t)n);
    std::vector<int> qidx((size_t)n, -1);
    const bool queue_ok = ctx->opt.fit_mode == VP_FITMODE_AUTO && n > 1;
    for (int64_t i = 0; i < n && queue_ok; ++i) {
        vp_problem *pr = problems[i];
        memset(&reports[i], 0, sizeof(vp_fit_report));
        vp_lm_config_from_options(pr->model->dtype, pr->model->md.q, opt, all_cfgs[(size_t)i]);
        lm_init(all_states[(size_t)i], pr->model->md.q, pr->alpha);
        qidx[(size_t)i] = queue_kernel_for(pr, nullptr);
        if (qidx[(size_t)i] < 0) continue;
        if (!lm_advance(all_states[(size_t)i], all_cfgs[(size_t)i], pr->eval)) { // terminated at the starting point
            fill_report(all_states[(size_t)i], &reports[i]);
            done[(size_t)i] = 1;
            continue;
        }
        done[(size_t)i] = 2;
    }
    int first_error = VP_OK;
    for (int64_t i = 0; i < n; ++i) {
        if (done[(size_t)i] != 2) continue;
        // group: same kernel instantiation, same padded rows
        std::vector<int64_t> idx;
        for (int64_t j = i; j < n; ++j)
            if (done[(size_t)j] == 2 && qidx[(size_t)j] == qidx[(size_t)i] && problems[j]->model->ld == problems[i]->model->ld)
                idx.push_back(j);
        std::vector<vp_problem *> prs;
        std::vector<LmState> sts;
        std::vector<LmConfig> cfs;
        for (int64_t j : idx) { prs.push_back(problems[j]); sts.push_back(all_states[(size_t)j]); cfs.push_back(all_cfgs[(size_t)j]); }
        // a group of one f64 fit is better served by the whole-GPU persistent kernel (vp_fit)
        const bool use_queue = idx.size() > 1 || problems[i]->plan_fit < 0;
        const int rc = use_queue ? fit_queue_group(ctx, prs, sts, cfs) : VP_ERR_UNSUPPORTED_BASIS;
        for (size_t t = 0; t < idx.size(); ++t) {
            const int64_t j = idx[t];
            if (rc == VP_OK) {
                fill_report(sts[t], &reports[j]);
                done[(size_t)j] = 1;
            } else if (rc == VP_ERR_UNSUPPORTED_BASIS) {
                done[(size_t)j] = 0; // per-fit driver below
            } else {
                done[(size_t)j] = 3;
                if (first_error == VP_OK) first_error = rc;
            }
        }
    }
    for (int64_t i = 0; i < n; ++i)
        if (done[(size_t)i] == 0) {
            const int rc = vp_fit(problems[i], opt, &reports[i]);
            if (rc != VP_OK && first_error == VP_OK) first_error = rc;
        }
    return first_error;
}
