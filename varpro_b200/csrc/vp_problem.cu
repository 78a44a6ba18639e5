// vp_problem.cu -- launch planning, kernel launches, problems, the trait-mirroring entry points
// (set_params / residuals / jacobian / ...), the materialisers and the batched statistics of the C ABI
// (include/varpro_b200.h).
#include "vp_internal.h"
#include "aux_kernels.cuh"

using namespace vp;

constexpr int GENERIC_THREADS = 512; // stream_kernel_generic

// the fused evaluation / persistent fit kernel (fit_kernel_dmma<TY>) for this problem, if instantiated for its shape
static int plan_fit_kernel(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    const vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    pr->plan_fit = -1;
    if (ctx->opt.eval_split) return VP_OK;  // K1 + K2 requested
    if (mo->hosteval) return VP_OK; // the fused kernels evaluate the built-in device basis functions
    if (ctx->opt.stream_kernel != VP_STREAMK_AUTO) return VP_OK;
    const size_t es = vp_esize(mo->dtype);
    int lds = 0; // column stride of a tile slot (vp_tile_lds)
    const KernelTables &KT = vp_kernel_tables();
    // the best row tiling among the fused instantiations of this shape
    int pick = -1;
    for (size_t i = 0; i < KT.fit.size(); ++i) {
        const FitKernelEntry &k = KT.fit[i];
        if (k.dtype != mo->dtype || k.n != md.n || k.p != md.p) continue;
        const int rows = 4 * k.ksteps * k.nwarps;
        const int lds_k = vp_tile_lds(mo->dtype, mo->ld, rows, k.exact);
        if (rows < mo->ld || lds_k < 0) continue;
        if (pick < 0 || vp_better_tiling(k.ksteps, k.nwarps, k.exact, KT.fit[(size_t)pick].ksteps, KT.fit[(size_t)pick].nwarps,
                                         KT.fit[(size_t)pick].exact, ctx->opt.fit_warps)) {
            pick = (int)i;
            lds = lds_k;
        }
    }
    if (pick < 0) return VP_OK;
    const FitKernelEntry &k = KT.fit[(size_t)pick];
    const size_t stage_bytes = (size_t)DMMA_CT * lds * es;
    const int pst = (int)(((size_t)(md.n + md.p + 1) * lds * sizeof(double) + stage_bytes - 1) / stage_bytes); // panel staging stages
    cudaFuncAttributes fa{};
    VP_CUDA(ctx, cudaFuncGetAttributes(&fa, k.fn));
    if (fa.sharedSizeBytes + 1024 + (size_t)(pst + 1) * stage_bytes > 227 * 1024) return VP_OK;
    const size_t budget = 227 * 1024 - fa.sharedSizeBytes - 1024;
    int nst = (int)(budget / stage_bytes);
    if (nst > STREAM_MAX_STAGES) nst = STREAM_MAX_STAGES;
    const int max_st = ctx->opt.stream_stages;
    if (max_st >= 2 && nst > max_st) nst = max_st;
    if (nst < pst + 1) return VP_OK;
    const size_t smem = (size_t)nst * stage_bytes;
    VP_CUDA(ctx, vp_ensure_dynamic_smem(ctx->device, k.fn, smem));
    int occ = 0;
    VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, k.nwarps * 32, smem));
    if (occ < 1) return VP_OK;
    const long long ntiles = (pr->S + DMMA_CT - 1) / DMMA_CT;
    long long grid = (long long)ctx->sm_count * occ; // co-resident by construction (cooperative launch checks it)
    if (ctx->opt.max_ctas > 0 && grid > ctx->opt.max_ctas) grid = ctx->opt.max_ctas;
    if (grid > ntiles) grid = ntiles;
    if (grid > pr->max_grid / 2) grid = pr->max_grid / 2; // two row buffers (evaluation parity)
    pr->plan_fit = pick;
    if (4 * k.ksteps * k.nwarps > pr->plan_rows) pr->plan_rows = 4 * k.ksteps * k.nwarps;
    pr->fit_lds = lds;
    pr->fit_grid = (int)grid;
    pr->fit_nst = nst;
    pr->fit_smem = smem;
    return VP_OK;
}

// choose the kernel instantiation, stage count and grid for a problem
static int plan_stream(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    const vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const int want_ct = ctx->opt.stream_ct;
    const int want_occ = ctx->opt.stream_occ;
    const bool force_generic = ctx->opt.stream_kernel == VP_STREAMK_GENERIC;
    const size_t es = vp_esize(mo->dtype);
    const KernelTables &KT = vp_kernel_tables();
    pr->plan_kind = -1;
    pr->plan_dmma = -1;
    const bool allow_dmma = mo->dtype == VP_F64 && ctx->opt.stream_kernel == VP_STREAMK_AUTO;
    if (allow_dmma) {
        int lds = 0; // column stride of a tile slot (vp_tile_lds)
        int pick = -1;
        for (int i = 0; i < (int)KT.dmma.size(); ++i) {
            const DmmaKernelEntry &k = KT.dmma[i];
            if (k.n != md.n || k.p != md.p) continue;
            const int rows = 4 * k.ksteps * k.nwarps;
            if (rows < mo->ld) continue;
            const int lds_k = vp_tile_lds(VP_F64, mo->ld, rows, k.exact); // the unpredicated variant needs rows <= lds
            if (lds_k < 0) continue;
            const int prow = pick < 0 ? 0 : 4 * KT.dmma[pick].ksteps * KT.dmma[pick].nwarps;
            if (pick < 0 || rows < prow || (rows == prow && k.exact && !KT.dmma[pick].exact)) { pick = i; lds = lds_k; }
        }
        if (pick >= 0) {
            const DmmaKernelEntry &k = KT.dmma[pick];
            const size_t stage_bytes = (size_t)DMMA_CT * lds * sizeof(double);
            cudaFuncAttributes fa{};
            VP_CUDA(ctx, cudaFuncGetAttributes(&fa, k.fn));
            const size_t budget = 227 * 1024 - fa.sharedSizeBytes - 1024;
            int nst = (int)(budget / stage_bytes);
            if (nst > STREAM_MAX_STAGES) nst = STREAM_MAX_STAGES;
            const int max_st = ctx->opt.stream_stages;
            if (max_st >= 2 && nst > max_st) nst = max_st;
            if (nst >= 2) {
                const size_t smem = (size_t)nst * stage_bytes;
                VP_CUDA(ctx, vp_ensure_dynamic_smem(ctx->device, k.fn, smem));
                int occ_real = 0;
                VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, k.fn, k.nwarps * 32, smem));
                if (occ_real >= 1) {
                    const long long ntiles = (pr->S + DMMA_CT - 1) / DMMA_CT;
                    long long grid = (long long)ctx->sm_count * occ_real;
                    if (ctx->opt.max_ctas > 0 && grid > ctx->opt.max_ctas) grid = ctx->opt.max_ctas;
                    if (grid > ntiles) grid = ntiles;
                    if (grid > pr->max_grid) grid = pr->max_grid;
                    pr->plan_dmma = pick;
                    pr->plan_lds = lds;
                    pr->plan_rows = 4 * k.ksteps * k.nwarps;
                    pr->plan_ct = DMMA_CT;
                    pr->plan_grid = (int)grid;
                    pr->plan_nst = nst;
                    pr->plan_smem = smem;
                    return plan_fit_kernel(pr);
                }
            }
        }
    }
    int best = -1;
    if (!force_generic) {
        for (int i = 0; i < (int)KT.stream.size(); ++i) {
            const StreamKernelEntry &k = KT.stream[i];
            if (k.dtype != mo->dtype || k.n != md.n || k.p != md.p) continue;
            if ((long long)k.threads * k.chunks * vp_vec_of(mo->dtype) < mo->ld) continue;
            // smallest covering (threads*chunks) first, then the requested tile width
            if (best < 0) { best = i; continue; }
            const StreamKernelEntry &b = KT.stream[best];
            const long long ck = (long long)k.threads * k.chunks, cb = (long long)b.threads * b.chunks;
            if (ck < cb || (ck == cb && k.threads < b.threads) ||
                (ck == cb && k.threads == b.threads && std::abs(k.ct - want_ct) < std::abs(b.ct - want_ct)))
                best = i;
        }
    }
    if (best >= 0) {
        const StreamKernelEntry &k = KT.stream[best];
        const size_t stage_bytes = (size_t)k.ct * mo->ld * es;
        // static shared memory of the kernel + a safety margin
        cudaFuncAttributes fa{};
        VP_CUDA(ctx, cudaFuncGetAttributes(&fa, k.fn));
        const size_t per_sm = 227 * 1024; // usable shared memory per SM
        int occ = want_occ < 1 ? 1 : want_occ;
        int nst = 0;
        for (; occ >= 1; --occ) {
            const size_t budget = per_sm / occ - fa.sharedSizeBytes - 1024;
            nst = (int)(budget / stage_bytes);
            if (nst >= 2) break;
        }
        if (nst >= 2) {
            if (nst > STREAM_MAX_STAGES) nst = STREAM_MAX_STAGES;
            const int max_st = ctx->opt.stream_stages;
            if (max_st >= 2 && nst > max_st) nst = max_st;
            const size_t smem = (size_t)nst * stage_bytes;
            if (smem + fa.sharedSizeBytes <= ctx->smem_optin) {
                VP_CUDA(ctx, vp_ensure_dynamic_smem(ctx->device, k.fn, smem));
                int occ_real = 0;
                VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, k.fn, k.threads, smem));
                if (occ_real >= 1) {
                    const long long ntiles = (pr->S + k.ct - 1) / k.ct;
                    long long grid = (long long)ctx->sm_count * occ_real;
                    if (ctx->opt.max_ctas > 0 && grid > ctx->opt.max_ctas) grid = ctx->opt.max_ctas;
                    if (grid > ntiles) grid = ntiles;
                    if (grid > pr->max_grid) grid = pr->max_grid;
                    pr->plan_kind = best;
                    pr->plan_rows = k.threads * k.chunks * vp_vec_of(mo->dtype);
                    pr->plan_ct = k.ct;
                    pr->plan_grid = (int)grid;
                    pr->plan_nst = nst;
                    pr->plan_smem = smem;
                    return plan_fit_kernel(pr);
                }
            }
        }
    }
    // generic fallback (stream_kernel_generic): one persistent CTA per SM, every warp two columns at a time, the panel
    // staged in shared memory when (n + p) * m doubles fit next to the kernel's static shared memory
    pr->plan_kind = -1;
    {
        long long grid = (pr->S + 2 * (GENERIC_THREADS / 32) - 1) / (2 * (GENERIC_THREADS / 32));
        if (grid > ctx->sm_count) grid = ctx->sm_count;
        if (ctx->opt.max_ctas > 0 && grid > ctx->opt.max_ctas) grid = ctx->opt.max_ctas;
        if (grid > pr->max_grid) grid = pr->max_grid;
        pr->plan_grid = (int)(grid < 1 ? 1 : grid);
        const int mp = (md.m + 1) / 2 * 2;
        const size_t panel_bytes = sizeof(double) * (size_t)(md.n + md.p) * mp;
        pr->plan_nst = 0;
        pr->plan_smem = panel_bytes + 24 * 1024 <= ctx->smem_optin ? panel_bytes : 0; // (24 KB: static shared memory of the kernel)
        pr->plan_rows = mo->ld;
        pr->plan_ct = 1;
    }
    return VP_OK;
}

template <typename T>
static int launch_panel_t(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    unsigned long long *dbg = pr->dbg ? pr->dbg + (size_t)pr->max_grid * VP_DBG_SLOTS : nullptr;
    // fast path: register-resident Householder panel
    if (!ctx->opt.panel_generic) { // (host-evaluated models too: their uploaded [Phi | D] replaces the basis evaluation)
        for (const PanelHHEntry &k : vp_kernel_tables().panel) {
            if (k.dtype != mo->dtype || k.n != md.n || k.p != md.p || (long long)k.rpt * k.threads < md.m) continue;
            ModelDesc mdc = md;
            const void *xp = mo->x_dev, *wp = pr->w_dev;
            const double *ap = pr->alpha_dev;
            double eps = pr->rank_tol;
            int ldp = pr->ldp;
            void *pq = pr->Pq;
            PanelSmall *sm = pr->small;
            const double *pre = mo->hosteval ? mo->pre_dev : nullptr;
            void *args[] = {&mdc, &xp, &wp, &ap, &eps, &ldp, &pq, &sm, &dbg, &pre};
            VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3(1), dim3(k.threads), args, 0, ctx->stream));
            ctx->launches++;
            return VP_OK;
        }
    }
    // generic path: CGS2 round interpreter with the panel in shared memory
    int threads = PANEL_THREADS;
    if (md.m < threads) threads = ((md.m + 31) / 32) * 32;
    const size_t smem = sizeof(double) * ((size_t)(md.n + md.p) * md.m + (size_t)(threads / 32 + 1) * 8 + 64);
    if (smem > ctx->smem_optin)
        return vp_fail(ctx, VP_ERR_MODEL_TOO_LARGE, "m*(n+p) panel does not fit in shared memory");
    if (smem > 48 * 1024) VP_CUDA(ctx, vp_ensure_dynamic_smem(ctx->device, (const void *)panel_kernel<T>, smem));
    panel_kernel<T><<<1, threads, smem, ctx->stream>>>(md, (const T *)mo->x_dev, (const T *)pr->w_dev, pr->alpha_dev,
                                                       pr->rank_tol, pr->ldp, (T *)pr->Pq, pr->small, dbg,
                                                       mo->hosteval ? mo->pre_dev : nullptr);
    ctx->launches++;
    VP_CUDA(ctx, cudaGetLastError());
    return VP_OK;
}

template <typename T>
static int launch_stream_t(vp_problem *pr, int cdst, bool graph_mode)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    StreamArgs<T> a{};
    a.Y = (const T *)pr->Yw; a.ld = mo->ld; a.S = (int)pr->S;
    a.Pq = (const T *)pr->Pq; a.Pe = (const T *)pr->Pq + (size_t)md.n * pr->ldp; a.ldp = pr->ldp; a.small = pr->small;
    {
        const long long ntiles = (pr->S + pr->plan_ct - 1) / pr->plan_ct;
        a.tiles_base = (int)(ntiles / pr->plan_grid);
        a.tiles_rem = (int)(ntiles % pr->plan_grid);
    }
    a.C0 = (T *)pr->C[0]; a.C1 = (T *)pr->C[1]; a.cdst = cdst;
    a.fit = graph_mode ? pr->fit_dev : nullptr;
    a.cond = graph_mode ? (unsigned long long)pr->fit_cond : 0ull;
    a.partials = pr->partials; a.red_stride = pr->red_stride; a.ticket = pr->ticket; a.out = pr->out_dev;
    a.nstages = pr->plan_nst; a.q = md.q; a.dbg = pr->dbg;
    for (int e = 0; e < VP_MAX_P; ++e) { a.e_basis[e] = md.e_basis[e]; a.e_param[e] = md.e_param[e]; }
    if (pr->plan_dmma >= 0) {
        if constexpr (std::is_same<T, double>::value) {
            const DmmaKernelEntry &k = vp_kernel_tables().dmma[pr->plan_dmma];
            int lds = pr->plan_lds;
            void *args[] = {(void *)&a, (void *)&lds};
            VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3(pr->plan_grid), dim3(k.nwarps * 32), args, pr->plan_smem, ctx->stream));
        }
    } else if (pr->plan_kind >= 0) {
        const StreamKernelEntry &k = vp_kernel_tables().stream[pr->plan_kind];
        void *args[] = {(void *)&a};
        VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3(pr->plan_grid), dim3(k.threads), args, pr->plan_smem, ctx->stream));
    } else {
        // generic shape: panel staged in shared memory when it fits (plan_smem from the launch plan)
        const int npv = md.n + md.p;
        const void *fn = npv <= 8 ? (const void *)stream_kernel_generic<T, GENERIC_THREADS, 8>
                                  : (npv <= 12 ? (const void *)stream_kernel_generic<T, GENERIC_THREADS, 12>
                                               : (const void *)stream_kernel_generic<T, GENERIC_THREADS, 20>);
        int n_arg = md.n, p_arg = md.p, m_arg = md.m, mp_arg = pr->plan_smem > 0 ? (md.m + 1) / 2 * 2 : 0;
        if (pr->plan_smem > 0) VP_CUDA(ctx, vp_ensure_dynamic_smem(ctx->device, fn, pr->plan_smem)); // (static + dynamic may exceed 48 KB)
        void *args[] = {(void *)&a, (void *)&n_arg, (void *)&p_arg, (void *)&m_arg, (void *)&mp_arg};
        VP_CUDA(ctx, cudaLaunchKernel(fn, dim3(pr->plan_grid), dim3(GENERIC_THREADS), args, pr->plan_smem, ctx->stream));
    }
    ctx->launches++;
    VP_CUDA(ctx, cudaGetLastError());
    return VP_OK;
}

// One launch of fit_kernel_dmma: a single fused evaluation (fit_mode = false; result in out_dev,
// coefficients into buffer cdst) or a whole fit (fit_mode = true; cooperative launch, state in fit_dev).
// Either way CTA b works on part b of the canonical partition (dmma_tile.cuh).
template <typename T>
static int launch_fused_t(vp_problem *pr, int cdst, bool fit_mode)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const FitKernelEntry &k = vp_kernel_tables().fit[pr->plan_fit];
    cudaStream_t stream = ctx->stream;
    const int grid = pr->fit_grid;
    StreamArgs<T> a{};
    a.Y = (const T *)pr->Yw; a.ld = mo->ld; a.S = (int)pr->S;
    a.Pq = nullptr; a.Pe = nullptr; a.ldp = pr->ldp; a.small = nullptr;
    {
        const TilePartition tp = make_partition((int)((pr->S + DMMA_CT - 1) / DMMA_CT), grid);
        a.tiles_base = tp.base;
        a.tiles_rem = tp.rem;
    }
    a.C0 = (T *)pr->C[0]; a.C1 = (T *)pr->C[1]; a.cdst = cdst;
    a.fit = fit_mode ? pr->fit_dev : nullptr;
    a.cond = 0ull;
    a.partials = pr->partials; a.red_stride = pr->red_stride; a.ticket = pr->ticket; a.out = pr->out_dev;
    a.nstages = pr->fit_nst; a.q = md.q; a.dbg = pr->dbg;
    for (int e = 0; e < VP_MAX_P; ++e) { a.e_basis[e] = md.e_basis[e]; a.e_param[e] = md.e_param[e]; }
    FitArgs f{};
    f.md = md;
    f.x = mo->x_dev; f.w = pr->w_dev;
    f.svd_eps = pr->rank_tol;
    f.alpha_dev = pr->alpha_dev;
    f.ctl = pr->fit_ctl;
    f.jac_full = pr->jac_full;
    if (pr->comm) f.comm = pr->comm->args;
    int lds = pr->fit_lds;
    void *args[] = {(void *)&a, (void *)&lds, (void *)&f};
    if (fit_mode) { // (the caller's copy of the LM state has zeroed the control word behind it)
        VP_CUDA(ctx, cudaLaunchCooperativeKernel(k.fn, dim3(grid), dim3(k.nwarps * 32), args, pr->fit_smem, stream));
    } else {
        VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3(grid), dim3(k.nwarps * 32), args, pr->fit_smem, stream));
    }
    ctx->launches++;
    return VP_OK;
}

int vp_launch_fused(vp_problem *pr, int cdst, bool fit_mode)
{
    return pr->model->dtype == VP_F32 ? launch_fused_t<float>(pr, cdst, fit_mode) : launch_fused_t<double>(pr, cdst, fit_mode);
}

int vp_launch_panel(vp_problem *pr)
{
    return pr->model->dtype == VP_F32 ? launch_panel_t<float>(pr) : launch_panel_t<double>(pr);
}
int vp_launch_stream(vp_problem *pr, int cdst, bool graph_mode)
{
    return pr->model->dtype == VP_F32 ? launch_stream_t<float>(pr, cdst, graph_mode) : launch_stream_t<double>(pr, cdst, graph_mode);
}
int vp_launch_eval(vp_problem *pr, int cdst)
{
    if (pr->plan_fit >= 0) return vp_launch_fused(pr, cdst, false);
    int rc = vp_launch_panel(pr);
    return rc != VP_OK ? rc : vp_launch_stream(pr, cdst);
}

// after a collective evaluation: did the NVLink exchange time out?
int vp_comm_check(vp_problem *pr)
{
    if (!pr->comm) return VP_OK;
    int err = 0;
    vp_ctx *ctx = pr->ctx;
    VP_CUDA(ctx, cudaMemcpyAsync(&err, pr->comm->error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (err) return vp_fail(ctx, VP_ERR_COMM, "timed out waiting for a peer GPU's contribution (ranks must make the same sequence of calls)");
    return VP_OK;
}

// host-evaluated model: run the user's callback at `alpha` and upload the unweighted [Phi | D]
static int host_eval_push(vp_problem *pr, const double *alpha)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const size_t nphi = (size_t)md.m * md.n, nd = (size_t)md.m * md.p;
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the previous upload has been consumed
    const int rc = mo->eval_fn(mo->eval_user, alpha, mo->pre_host, mo->pre_host + nphi);
    if (rc != 0) return vp_fail(ctx, VP_ERR_NO_CACHED_CALCULATION, "the model's evaluation callback reported an error");
    VP_CUDA(ctx, cudaMemcpyAsync(mo->pre_dev, mo->pre_host, sizeof(double) * (nphi + nd), cudaMemcpyHostToDevice, ctx->stream));
    return VP_OK;
}

// Full Golub-Pereyra mode for problems WITHOUT a fused kernel (fp32, host-evaluated models, shapes outside
// kernel_tables.h): H += sum (R^-1 R^-T)_{j(e) j(f)} U_ef on top of the Kaufman H in pr->out_host, from one extra pass
// over Y (ugram_kernel, aux_kernels.cuh; K1 has left the panel [Q | E] in HBM). The fused kernels accumulate U in
// their single pass (dmma_tile.cuh) and need none of this.
template <typename T>
static int add_full_jacobian_term_t(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const int p = md.p, n = md.n, q = md.q, nu = p * (p + 1) / 2;
    if (p == 0) return VP_OK;
    long long blocks = (pr->S + 7) / 8;
    if (blocks > (long long)ctx->sm_count * 4) blocks = (long long)ctx->sm_count * 4;
    const size_t nrows = (size_t)blocks * 8;
    double *rows = nullptr;
    VP_CUDA(ctx, DEV_ALLOC(ctx, &rows, sizeof(double) * nrows * nu));
    ugram_kernel<T><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const T *)pr->Yw, mo->ld, pr->ldp, md.m, (int)pr->S, p,
                                                               (const T *)pr->Pq + (size_t)n * pr->ldp, rows);
    ctx->launches++;
    std::vector<double> h(nrows * nu);
    PanelSmall sm;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), rows, sizeof(double) * nrows * nu, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&sm, pr->small, sizeof(sm), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    DEV_FREE(ctx, rows);
    if (e != cudaSuccess) return vp_fail(ctx, VP_ERR_CUDA, std::string("full Jacobian term: ") + cudaGetErrorString(e));
    double U[VP_MAX_P][VP_MAX_P];
    {
        int t = 0;
        for (int a = 0; a < p; ++a)
            for (int b = a; b < p; ++b, ++t) {
                double acc = 0.0;
                for (size_t r = 0; r < nrows; ++r) acc += h[r * nu + t];
                U[a][b] = U[b][a] = acc;
            }
    }
    double Wm[VP_MAX_N][VP_MAX_N]; // R^-1 R^-T
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int c = 0; c < n; ++c) acc += sm.Rinv[c * VP_MAX_N + i] * sm.Rinv[c * VP_MAX_N + j];
            Wm[i][j] = acc;
        }
    EvalOut *o = pr->out_host;
    for (int a = 0; a < p; ++a)
        for (int b = 0; b < p; ++b)
            o->H[md.e_param[b] * q + md.e_param[a]] += Wm[md.e_basis[a]][md.e_basis[b]] * U[a][b];
    for (int i = 0; i < q * q; ++i)
        if (!std::isfinite(o->H[i])) o->finite &= ~VP_EVAL_DERIVS_OK;
    return VP_OK;
}
static int add_full_jacobian_term(vp_problem *pr)
{
    return pr->model->dtype == VP_F32 ? add_full_jacobian_term_t<float>(pr) : add_full_jacobian_term_t<double>(pr);
}

// Evaluate at `alpha` into coefficient buffer `cdst`; result in pr->out_host.
int vp_evaluate_sync(vp_problem *pr, const double *alpha, int cdst)
{
    vp_ctx *ctx = pr->ctx;
    const int q = pr->model->md.q;
    cudaSetDevice(ctx->device);
    if (pr->model->hosteval) {
        int rce = host_eval_push(pr, alpha);
        if (rce != VP_OK) return rce;
    }
    for (int k = 0; k < q; ++k) pr->alpha_stage[k] = alpha[k];
    if (q > 0)
        VP_CUDA(ctx, cudaMemcpyAsync(pr->alpha_dev, pr->alpha_stage, sizeof(double) * q, cudaMemcpyHostToDevice, ctx->stream));
    int rc = vp_launch_eval(pr, cdst);
    if (rc != VP_OK) return rc;
    VP_CUDA(ctx, cudaMemcpyAsync(pr->out_host, pr->out_dev, sizeof(EvalOut), cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rc = vp_comm_check(pr);
    if (rc == VP_OK && pr->jac_full && pr->plan_fit < 0) rc = add_full_jacobian_term(pr);
    return rc;
}

// Re-evaluate at pr->alpha into the other coefficient buffer and adopt the result (after a change of the
// Jacobian mode, the communicator, ...).
int vp_refresh_cached_evaluation(vp_problem *pr)
{
    const int dst = pr->cur ^ 1;
    int rc = vp_evaluate_sync(pr, pr->alpha, dst);
    if (rc == VP_ERR_NO_CACHED_CALCULATION) { pr->cached = false; return VP_OK; } // model error -> cache = None
    if (rc != VP_OK) { pr->cached = false; return rc; }
    pr->cur = dst;
    vp_evalout_to_lm(*pr->out_host, pr->model->md.q, pr->eval);
    pr->cached = (pr->eval.finite & VP_EVAL_RESIDUAL_OK) != 0;
    return VP_OK;
}

void vp_evalout_to_lm(const EvalOut &o, int q, LmEval &ev)
{
    ev.rnorm2 = o.rnorm2;
    ev.finite = o.finite;
    for (int k = 0; k < q; ++k) ev.g[k] = o.g[k];
    for (int i = 0; i < q * q; ++i) ev.H[i] = o.H[i];
}

// ----------------------------------------------------------------------------
// problem
// ----------------------------------------------------------------------------
static int problem_create_common(vp_ctx *ctx, vp_model *model, int64_t S, const void *Y, int64_t ldY, bool y_on_device,
                                 const void *w_host, double svd_eps, const double *alpha0, vp_problem **out)
{
    if (!ctx || !model || !out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (model->ctx != ctx) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "model belongs to a different context");
    if (!Y) return vp_fail(ctx, VP_ERR_Y_DATA_MISSING, vp_status_string(VP_ERR_Y_DATA_MISSING));
    const ModelDesc &md = model->md;
    if (S <= 0 || md.m <= 0) return vp_fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (ldY < md.m)
        return vp_fail(ctx, VP_ERR_INVALID_LENGTH_OF_DATA, "Vectors x and y must have same lengths. Given x length = " +
                                                            std::to_string(md.m) + " and y length = " + std::to_string(ldY));
    if (S > INT32_MAX / 2) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "S too large for one problem handle");
    if (md.q > 0 && !alpha0) return vp_fail(ctx, VP_ERR_INVALID_PARAMETER_COUNT, vp_status_string(VP_ERR_INVALID_PARAMETER_COUNT));
    cudaSetDevice(ctx->device);
    vp_problem *pr = new (std::nothrow) vp_problem();
    if (!pr) return VP_ERR_OUT_OF_MEMORY;
    pr->ctx = ctx; pr->model = model; pr->S = S;
    const int dtype = model->dtype;
    const size_t es = vp_esize(dtype);
    const int ld = model->ld, m = md.m;
    // default epsilon = machine epsilon of the scalar (src/problem/builder.rs:282), |eps| otherwise (:248)
    pr->svd_eps = svd_eps < 0 ? (dtype == VP_F32 ? (double)FLT_EPSILON : DBL_EPSILON) : fabs(svd_eps);
    pr->rank_policy = VP_RANK_ABSOLUTE;
    pr->rank_tol = vp_rank_tol(pr->rank_policy, pr->svd_eps, md.m, dtype);
    pr->red_stride = 64;
    pr->max_grid = ctx->sm_count * 8;

    auto cleanup = [&](int code, const std::string &msg) {
        vp_problem_destroy(pr);
        return vp_fail(ctx, code, msg);
    };
#define VP_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return cleanup(_e == cudaErrorMemoryAllocation ? VP_ERR_OUT_OF_MEMORY : VP_ERR_CUDA,       \
                           std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
    } while (0)

    VP_TRY(DEV_ALLOC(ctx, &pr->Yw, es * (size_t)ld * S));
    VP_TRY(DEV_ALLOC(ctx, &pr->small, sizeof(PanelSmall)));
    VP_TRY(DEV_ALLOC(ctx, &pr->C[0], es * (size_t)md.n * S));
    VP_TRY(DEV_ALLOC(ctx, &pr->C[1], es * (size_t)md.n * S));
    VP_TRY(DEV_ALLOC(ctx, &pr->partials, sizeof(double) * (size_t)pr->red_stride * pr->max_grid));
    VP_TRY(DEV_ALLOC(ctx, &pr->ticket, sizeof(unsigned int)));
    VP_TRY(cudaMemsetAsync(pr->ticket, 0, sizeof(unsigned int), ctx->stream));
    VP_TRY(DEV_ALLOC(ctx, &pr->out_dev, sizeof(EvalOut)));
    // the LM state and, right behind it, the control word of the persistent fit kernel: ONE copy in (which also
    // zeroes the control word) and ONE copy out per fit
    VP_TRY(DEV_ALLOC(ctx, &pr->fit_dev, VP_FIT_BLOCK_BYTES));
    VP_TRY(cudaMemsetAsync(pr->fit_dev, 0, VP_FIT_BLOCK_BYTES, ctx->stream));
    VP_TRY(HOST_ALLOC(ctx, &pr->fit_host, VP_FIT_BLOCK_BYTES));
    pr->fit_ctl = reinterpret_cast<FitCtl *>(reinterpret_cast<unsigned char *>(pr->fit_dev) + VP_FIT_CTL_OFFSET);
    pr->alpha_dev = &pr->fit_dev->st.x_trial[0];
    VP_TRY(DEV_ALLOC(ctx, &pr->phi_scratch, sizeof(double) * (size_t)m * md.n));
    VP_TRY(HOST_ALLOC(ctx, &pr->out_host, sizeof(EvalOut)));
    VP_TRY(HOST_ALLOC(ctx, &pr->alpha_stage, sizeof(double) * VP_MAX_Q + sizeof(FitCtl)));
    if (w_host) {
        VP_TRY(DEV_ALLOC(ctx, &pr->w_dev, es * (size_t)m));
        VP_TRY(cudaMemcpyAsync(pr->w_dev, w_host, es * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    }
    // observations -> device buffer with leading dimension ld
    const cudaMemcpyKind kind = y_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (ldY == ld)
        VP_TRY(cudaMemcpyAsync(pr->Yw, Y, es * (size_t)ld * S, kind, ctx->stream));
    else
        VP_TRY(cudaMemcpy2DAsync(pr->Yw, es * ld, Y, es * (size_t)ldY, es * (size_t)m, (size_t)S, kind, ctx->stream));
    if (w_host || ld != m) {
        const long long total = (long long)ld * S;
        int blocks = (int)((total + 255) / 256 < (long long)ctx->sm_count * 16 ? (total + 255) / 256 : (long long)ctx->sm_count * 16);
        if (dtype == VP_F32)
            weight_rows_kernel<float><<<blocks, 256, 0, ctx->stream>>>((float *)pr->Yw, (const float *)pr->w_dev, m, ld, S);
        else
            weight_rows_kernel<double><<<blocks, 256, 0, ctx->stream>>>((double *)pr->Yw, (const double *)pr->w_dev, m, ld, S);
        ctx->launches++;
        VP_TRY(cudaGetLastError());
    }
#undef VP_TRY
    int rc = plan_stream(pr);
    if (rc != VP_OK) { vp_problem_destroy(pr); return rc; }
    {
        int ldp = pr->plan_rows > ld ? pr->plan_rows : ld;
        if (pr->plan_lds > ldp) ldp = pr->plan_lds;
        if (pr->fit_lds > ldp) ldp = pr->fit_lds;
        pr->ldp = (ldp + 3) / 4 * 4;
        cudaError_t e = DEV_ALLOC(ctx, &pr->Pq, es * (size_t)pr->ldp * (md.n + md.p + 1));
        if (e != cudaSuccess) { vp_problem_destroy(pr); return vp_fail(ctx, VP_ERR_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    }
    // first evaluation at the initial guess (src/problem/builder.rs:321)
    for (int k = 0; k < md.q; ++k) pr->alpha[k] = alpha0[k];
    rc = vp_evaluate_sync(pr, pr->alpha, pr->cur);
    if (rc == VP_ERR_NO_CACHED_CALCULATION) { // model error: the problem is built with cache = None (levmar/mod.rs:43-45)
        pr->cached = false;
        *out = pr;
        return VP_OK;
    }
    if (rc != VP_OK) { vp_problem_destroy(pr); return rc; }
    vp_evalout_to_lm(*pr->out_host, md.q, pr->eval);
    pr->cached = (pr->eval.finite & VP_EVAL_RESIDUAL_OK) != 0;
    *out = pr;
    return VP_OK;
}

extern "C" int vp_problem_create(vp_ctx *ctx, vp_model *model, int64_t S, const void *Y_host, int64_t ldY,
                                 const void *w_host, double svd_eps, const double *alpha0, vp_problem **out)
{
    return problem_create_common(ctx, model, S, Y_host, ldY, false, w_host, svd_eps, alpha0, out);
}

extern "C" int vp_problem_create_device(vp_ctx *ctx, vp_model *model, int64_t S, const void *Y_device, int64_t ldY,
                                        const void *w_host, double svd_eps, const double *alpha0, vp_problem **out)
{
    return problem_create_common(ctx, model, S, Y_device, ldY, true, w_host, svd_eps, alpha0, out);
}

extern "C" int vp_problem_destroy(vp_problem *pr)
{
    if (!pr) return VP_OK;
    cudaSetDevice(pr->ctx->device);
    cudaStreamSynchronize(pr->ctx->stream);
    vp_ctx *ctx = pr->ctx;
    DEV_FREE(ctx, pr->Yw); DEV_FREE(ctx, pr->w_dev); DEV_FREE(ctx, pr->Pq); DEV_FREE(ctx, pr->small);
    DEV_FREE(ctx, pr->C[0]); DEV_FREE(ctx, pr->C[1]); DEV_FREE(ctx, pr->partials); DEV_FREE(ctx, pr->ticket);
    DEV_FREE(ctx, pr->out_dev); DEV_FREE(ctx, pr->fit_dev); DEV_FREE(ctx, pr->phi_scratch); // (fit_ctl lives inside fit_dev)
    DEV_FREE(ctx, pr->Pq64);
    cudaFree(pr->dbg);
    if (pr->fit_exec) cudaGraphExecDestroy(pr->fit_exec);
    if (pr->fit_graph) cudaGraphDestroy(pr->fit_graph);
    HOST_FREE(ctx, pr->fit_host); HOST_FREE(ctx, pr->out_host); HOST_FREE(ctx, pr->alpha_stage);
    delete pr;
    return VP_OK;
}

// ----------------------------------------------------------------------------
// trait-mirroring entry points
// ----------------------------------------------------------------------------
extern "C" int vp_set_params(vp_problem *pr, const double *alpha)
{
    if (!pr || (!alpha && pr->model->md.q > 0)) return VP_ERR_INVALID_ARGUMENT;
    VP_NVTX("vp_set_params");
    const int q = pr->model->md.q;
    for (int k = 0; k < q; ++k) pr->alpha[k] = alpha[k];
    return vp_refresh_cached_evaluation(pr);
}

extern "C" int vp_params(const vp_problem *pr, double *alpha_out)
{
    if (!pr || !alpha_out) return VP_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < pr->model->md.q; ++k) alpha_out[k] = pr->alpha[k];
    return VP_OK;
}

extern "C" int vp_problem_set_jacobian(vp_problem *pr, int mode)
{
    if (!pr || (mode != VP_JACOBIAN_KAUFMAN && mode != VP_JACOBIAN_FULL)) return VP_ERR_INVALID_ARGUMENT;
    pr->jac_full = mode == VP_JACOBIAN_FULL ? 1 : 0;
    if (!pr->cached) return VP_OK;
    return vp_refresh_cached_evaluation(pr); // its J^T J depends on the mode
}

extern "C" int vp_problem_set_rank_policy(vp_problem *pr, int policy)
{
    if (!pr || (policy != VP_RANK_ABSOLUTE && policy != VP_RANK_RELATIVE)) return VP_ERR_INVALID_ARGUMENT;
    pr->rank_policy = policy;
    pr->rank_tol = vp_rank_tol(policy, pr->svd_eps, pr->model->md.m, pr->model->dtype);
    return vp_refresh_cached_evaluation(pr);
}

extern "C" int vp_reduce(vp_problem *pr, vp_reduced *out)
{
    if (!pr || !out) return VP_ERR_INVALID_ARGUMENT;
    const int q = pr->model->md.q;
    memset(out, 0, sizeof(*out));
    out->rnorm2 = pr->eval.rnorm2;
    out->finite = pr->eval.finite == VP_EVAL_ALL_OK ? 1 : 0;
    out->q = q;
    for (int k = 0; k < q; ++k) out->g[k] = pr->eval.g[k];
    for (int i = 0; i < q * q; ++i) out->H[i] = pr->eval.H[i];
    return pr->cached ? VP_OK : vp_fail(pr->ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
}

// The materialisers read the panel [Q | E] from HBM: build it with K1 at the accepted parameters
// (the fused kernel keeps the panel on chip, and a rejected LM trial leaves K1's buffers at the
// trial point). The coefficients C[cur] already belong to pr->alpha.
static int ensure_panel_current(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    const int q = pr->model->md.q;
    if (pr->model->hosteval) {
        int rce = host_eval_push(pr, pr->alpha);
        if (rce != VP_OK) return rce;
    }
    for (int k = 0; k < q; ++k) pr->alpha_stage[k] = pr->alpha[k];
    if (q > 0)
        VP_CUDA(ctx, cudaMemcpyAsync(pr->alpha_dev, pr->alpha_stage, sizeof(double) * q, cudaMemcpyHostToDevice, ctx->stream));
    return vp_launch_panel(pr);
}

template <typename T>
static int materialise_t(vp_problem *pr, int what, void *out_host, void *out_device = nullptr)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const size_t mS = (size_t)md.m * pr->S;
    const size_t count = what == 1 ? mS * md.q : mS;
    if (count == 0) return VP_OK;
    T *buf = static_cast<T *>(out_device); // device-resident output requested: write it in place, no copy
    if (!buf) VP_CUDA(ctx, DEV_ALLOC(ctx, &buf, sizeof(T) * count));
    const int blocks = ctx->sm_count * 8;
    if (what == 0) {
        residuals_kernel<T><<<blocks, 256, 0, ctx->stream>>>((const T *)pr->Yw, mo->ld, pr->ldp, md.m, (int)pr->S, md.n,
                                                             (const T *)pr->Pq, buf);
    } else if (what == 1) {
        jacobian_kernel<T><<<blocks, 256, 0, ctx->stream>>>(pr->ldp, md.m, (int)pr->S, md.n, md.p, md.q,
                                                            (const T *)pr->Pq + (size_t)md.n * pr->ldp, (const T *)pr->C[pr->cur], md, buf);
        if (pr->jac_full) {
            jacobian_full_term_kernel<T><<<blocks, 256, 0, ctx->stream>>>(md, (const T *)pr->Yw, mo->ld, pr->ldp, (int)pr->S,
                                                                          (const T *)pr->Pq, pr->small, buf);
            ctx->launches++;
        }
    } else {
        if (mo->hosteval) {
            cudaMemcpyAsync(pr->phi_scratch, mo->pre_dev, sizeof(double) * (size_t)md.m * md.n, cudaMemcpyDeviceToDevice, ctx->stream);
        } else {
            phi_kernel<T><<<32, 256, 0, ctx->stream>>>(md, (const T *)mo->x_dev, pr->alpha_dev, pr->phi_scratch);
            ctx->launches++;
        }
        best_fit_kernel<T><<<blocks, 256, 0, ctx->stream>>>(md.m, (int)pr->S, md.n, pr->phi_scratch,
                                                            (const T *)pr->C[pr->cur], buf);
    }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !out_device) e = cudaMemcpyAsync(out_host, buf, sizeof(T) * count, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (!out_device) DEV_FREE(ctx, buf);
    if (e != cudaSuccess) return vp_fail(ctx, VP_ERR_CUDA, std::string("materialise: ") + cudaGetErrorString(e));
    return VP_OK;
}

static int materialise(vp_problem *pr, int what, void *out_host, void *out_device = nullptr)
{
    if (!pr || (!out_host && !out_device)) return VP_ERR_INVALID_ARGUMENT;
    if (!pr->cached) return vp_fail(pr->ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
    cudaSetDevice(pr->ctx->device);
    int rc = ensure_panel_current(pr);
    if (rc != VP_OK) return rc;
    return pr->model->dtype == VP_F32 ? materialise_t<float>(pr, what, out_host, out_device)
                                      : materialise_t<double>(pr, what, out_host, out_device);
}

extern "C" int vp_residuals(vp_problem *pr, void *out_host) { return materialise(pr, 0, out_host); }
extern "C" int vp_jacobian(vp_problem *pr, void *out_host) { return materialise(pr, 1, out_host); }
extern "C" int vp_best_fit(vp_problem *pr, void *out_host) { return materialise(pr, 2, out_host); }
// the same three, written straight into a caller-owned DEVICE buffer (no host copy)
extern "C" int vp_residuals_device(vp_problem *pr, void *out_device) { return materialise(pr, 0, nullptr, out_device); }
extern "C" int vp_jacobian_device(vp_problem *pr, void *out_device) { return materialise(pr, 1, nullptr, out_device); }
extern "C" int vp_best_fit_device(vp_problem *pr, void *out_device) { return materialise(pr, 2, nullptr, out_device); }

extern "C" int vp_linear_coefficients(vp_problem *pr, void *out_host)
{
    if (!pr || !out_host) return VP_ERR_INVALID_ARGUMENT;
    if (!pr->cached) return vp_fail(pr->ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
    vp_ctx *ctx = pr->ctx;
    cudaSetDevice(ctx->device);
    const size_t bytes = vp_esize(pr->model->dtype) * (size_t)pr->model->md.n * pr->S;
    VP_CUDA(ctx, cudaMemcpyAsync(out_host, pr->C[pr->cur], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VP_OK;
}

// ----------------------------------------------------------------------------
// vp_statistics: FitStatistics::try_calculate per right-hand side (src/statistics/mod.rs:352-441)
// ----------------------------------------------------------------------------
template <typename T>
static int statistics_t(vp_problem *pr, double *cov_host, double *chi2_host, double *conf_host)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const int m = md.m, t = md.n + md.q, t0 = md.n + md.p;
    const size_t S = (size_t)pr->S;
    double *B = nullptr, *Gm = nullptr, *cov = nullptr, *chi2 = nullptr, *conf = nullptr;
    int *flag = nullptr;
    cudaError_t e = DEV_ALLOC(ctx, &B, sizeof(double) * (size_t)m * t0);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &Gm, sizeof(double) * (size_t)t0 * t0);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &cov, sizeof(double) * (size_t)t * t * S);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &chi2, sizeof(double) * S);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &flag, sizeof(int));
    if (e == cudaSuccess && conf_host) e = DEV_ALLOC(ctx, &conf, sizeof(double) * (size_t)m * S);
    int host_flag = 0;
    if (e == cudaSuccess) e = cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream);
    if (e == cudaSuccess) {
        if (mo->hosteval)
            cudaMemcpyAsync(B, mo->pre_dev, sizeof(double) * (size_t)m * t0, cudaMemcpyDeviceToDevice, ctx->stream);
        else
            basis_kernel<T><<<(m + 255) / 256, 256, 0, ctx->stream>>>(md, (const T *)mo->x_dev, pr->alpha_dev, B);
        gram_kernel<T><<<(t0 * t0 + 127) / 128, 128, 0, ctx->stream>>>(m, t0, B, (const T *)pr->w_dev, Gm);
        long long blocks = ((long long)S + 7) / 8;
        if (blocks > (long long)ctx->sm_count * 8) blocks = (long long)ctx->sm_count * 8;
        statistics_kernel<T><<<(unsigned)blocks, 256, 0, ctx->stream>>>(md, (const T *)pr->Yw, mo->ld, pr->ldp, (int)pr->S,
                                                                        (const T *)pr->Pq, (const T *)pr->C[pr->cur], Gm, B, cov,
                                                                        chi2, conf, flag);
        ctx->launches += 3;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(cov_host, cov, sizeof(double) * (size_t)t * t * S, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && chi2_host) e = cudaMemcpyAsync(chi2_host, chi2, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && conf_host) e = cudaMemcpyAsync(conf_host, conf, sizeof(double) * (size_t)m * S, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    DEV_FREE(ctx, B); DEV_FREE(ctx, Gm); DEV_FREE(ctx, cov); DEV_FREE(ctx, chi2); DEV_FREE(ctx, flag); DEV_FREE(ctx, conf);
    if (e != cudaSuccess)
        return vp_fail(ctx, e == cudaErrorMemoryAllocation ? VP_ERR_OUT_OF_MEMORY : VP_ERR_CUDA, std::string("vp_statistics: ") + cudaGetErrorString(e));
    if (host_flag) return vp_fail(ctx, VP_ERR_MATRIX_INVERSION, vp_status_string(VP_ERR_MATRIX_INVERSION));
    return VP_OK;
}

extern "C" int vp_statistics(vp_problem *pr, double *cov_out, double *reduced_chi2_out, double *conf_sigma_out)
{
    if (!pr || !cov_out) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = pr->ctx;
    if (pr->comm) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_statistics: call it on each rank's own columns after detaching the communicator");
    if (!pr->cached) return vp_fail(ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
    const ModelDesc &md = pr->model->md;
    if (md.m <= md.n + md.q) return vp_fail(ctx, VP_ERR_UNDERDETERMINED, vp_status_string(VP_ERR_UNDERDETERMINED)); // :377-379
    cudaSetDevice(ctx->device);
    int rc = ensure_panel_current(pr); // Q at the accepted parameters, in HBM
    if (rc != VP_OK) return rc;
    return pr->model->dtype == VP_F32 ? statistics_t<float>(pr, cov_out, reduced_chi2_out, conf_sigma_out)
                                      : statistics_t<double>(pr, cov_out, reduced_chi2_out, conf_sigma_out);
}
