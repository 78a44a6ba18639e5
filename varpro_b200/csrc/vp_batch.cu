// vp_batch.cu -- vp_batch_*: P independent problems (BASELINE config 3), one CTA per problem (batch_fit_kernel).
#include "vp_internal.h"
#include "batch_fit_kernel.cuh"

using namespace vp;

// ----------------------------------------------------------------------------
// vp_batch: P independent problems (BASELINE config 3), one CTA per problem (batch_fit_kernel)
// ----------------------------------------------------------------------------
struct vp_batch {
    vp_ctx *ctx = nullptr;
    vp_model *model = nullptr;
    int64_t P = 0;
    int64_t ld = 0;
    double *Y = nullptr, *w_dev = nullptr;
    double *alpha0 = nullptr, *alpha = nullptr, *C = nullptr, *obj = nullptr;
    int *term = nullptr, *nfev = nullptr;
    unsigned long long *next = nullptr;
    double svd_eps = 0.0;
    int rank_policy = 0;
    int kernel = -1;
    int mpad = 0;
    size_t smem = 0;
    bool fitted = false;
};

static int batch_create_common(vp_ctx *ctx, vp_model *model, int64_t P, const void *Y, int64_t ldY, bool on_device,
                               const void *w_host, double svd_eps, const double *alpha0, vp_batch **out)
{
    if (!ctx || !model || !out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (model->ctx != ctx) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "model belongs to a different context");
    if (model->dtype != VP_F64) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_batch: fp64 models only");
    if (model->hosteval) return vp_fail(ctx, VP_ERR_UNSUPPORTED_BASIS, "vp_batch: needs the built-in device basis functions");
    if (!Y) return vp_fail(ctx, VP_ERR_Y_DATA_MISSING, vp_status_string(VP_ERR_Y_DATA_MISSING));
    const ModelDesc &md = model->md;
    if (P <= 0) return vp_fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (ldY < md.m)
        return vp_fail(ctx, VP_ERR_INVALID_LENGTH_OF_DATA, "Vectors x and y must have same lengths. Given x length = " +
                                                            std::to_string(md.m) + " and y length = " + std::to_string(ldY));
    if (md.q > 0 && !alpha0) return vp_fail(ctx, VP_ERR_INVALID_PARAMETER_COUNT, vp_status_string(VP_ERR_INVALID_PARAMETER_COUNT));
    const KernelTables &KT = vp_kernel_tables();
    cudaSetDevice(ctx->device);
    int pick = -1;
    bool shape_known = false;
    for (size_t i = 0; i < KT.batch.size(); ++i) {
        const BatchKernelEntry &k = KT.batch[i];
        if (k.n != md.n || k.p != md.p) continue;
        shape_known = true;
        if ((long long)k.rpt * k.threads < md.m) continue;
        // the working matrix (rpt * threads rows, rows >= m are zero) + the kernel's static shared memory must fit
        cudaFuncAttributes fa{};
        VP_CUDA(ctx, cudaFuncGetAttributes(&fa, k.fn));
        if (sizeof(double) * (size_t)(md.n + md.p) * k.rpt * k.threads + fa.sharedSizeBytes > ctx->smem_optin) continue;
        // fewest rows per thread among the tilings that cover m; among those the requested number of problem slots
        const int want = ctx->opt.batch_slots;
        if (pick < 0 || k.rpt < KT.batch[pick].rpt ||
            (k.rpt == KT.batch[pick].rpt && std::abs(k.slots - want) < std::abs(KT.batch[pick].slots - want)))
            pick = (int)i;
    }
    if (pick < 0)
        return vp_fail(ctx, VP_ERR_MODEL_TOO_LARGE, shape_known ? "vp_batch: m*(n+p) working matrix does not fit in shared memory (m > 4096?)"
                                                                  : "vp_batch: no independent-batch kernel instantiated for this model shape");
    const int mpad = KT.batch[pick].rpt * KT.batch[pick].threads; // rows of the shared-memory working matrix
    const size_t smem = sizeof(double) * (size_t)(md.n + md.p) * mpad;
    VP_CUDA(ctx, vp_ensure_dynamic_smem(ctx->device, KT.batch[pick].fn, smem));
    vp_batch *b = new (std::nothrow) vp_batch();
    if (!b) return VP_ERR_OUT_OF_MEMORY;
    b->ctx = ctx; b->model = model; b->P = P; b->ld = md.m; b->kernel = pick; b->mpad = mpad; b->smem = smem;
    b->svd_eps = svd_eps < 0 ? DBL_EPSILON : fabs(svd_eps);
    const int q = md.q, n = md.n, m = md.m;
    cudaError_t e = DEV_ALLOC(ctx, &b->Y, sizeof(double) * (size_t)m * P);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->alpha0, sizeof(double) * (size_t)(q > 0 ? q : 1) * P);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->alpha, sizeof(double) * (size_t)(q > 0 ? q : 1) * P);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->C, sizeof(double) * (size_t)n * P);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->obj, sizeof(double) * (size_t)P);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->term, sizeof(int) * (size_t)P);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->nfev, sizeof(int) * (size_t)P);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->next, 2 * sizeof(unsigned long long)); // [0] work counter, [1] error word
    if (e == cudaSuccess && w_host) {
        e = DEV_ALLOC(ctx, &b->w_dev, sizeof(double) * (size_t)m);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->w_dev, w_host, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream);
    }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (e == cudaSuccess) {
        if (ldY == m) e = cudaMemcpyAsync(b->Y, Y, sizeof(double) * (size_t)m * P, kind, ctx->stream);
        else e = cudaMemcpy2DAsync(b->Y, sizeof(double) * m, Y, sizeof(double) * (size_t)ldY, sizeof(double) * (size_t)m, (size_t)P, kind, ctx->stream);
    }
    if (e == cudaSuccess && q > 0)
        e = cudaMemcpyAsync(b->alpha0, alpha0, sizeof(double) * (size_t)q * P, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && q > 0)
        e = cudaMemcpyAsync(b->alpha, b->alpha0, sizeof(double) * (size_t)q * P, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        vp_batch_destroy(b);
        return vp_fail(ctx, e == cudaErrorMemoryAllocation ? VP_ERR_OUT_OF_MEMORY : VP_ERR_CUDA, std::string("vp_batch_create: ") + cudaGetErrorString(e));
    }
    *out = b;
    return VP_OK;
}

extern "C" int vp_batch_create(vp_ctx *ctx, vp_model *model, int64_t P, const void *Y_host, int64_t ldY, const void *w_host,
                               double svd_eps, const double *alpha0, vp_batch **out)
{
    return batch_create_common(ctx, model, P, Y_host, ldY, false, w_host, svd_eps, alpha0, out);
}
extern "C" int vp_batch_create_device(vp_ctx *ctx, vp_model *model, int64_t P, const void *Y_device, int64_t ldY,
                                      const void *w_host, double svd_eps, const double *alpha0, vp_batch **out)
{
    return batch_create_common(ctx, model, P, Y_device, ldY, true, w_host, svd_eps, alpha0, out);
}

extern "C" int vp_batch_destroy(vp_batch *b)
{
    if (!b) return VP_OK;
    vp_ctx *ctx = b->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DEV_FREE(ctx, b->Y); DEV_FREE(ctx, b->w_dev); DEV_FREE(ctx, b->alpha0); DEV_FREE(ctx, b->alpha); DEV_FREE(ctx, b->C);
    DEV_FREE(ctx, b->obj); DEV_FREE(ctx, b->term); DEV_FREE(ctx, b->nfev); DEV_FREE(ctx, b->next);
    delete b;
    return VP_OK;
}

// Enqueue the batch fit on the context's stream (no synchronisation): starts every problem from the
// CURRENT parameters (the initial guess, or the result of the previous fit).
static int batch_fit_launch(vp_batch *b, const vp_lm_options *opt)
{
    vp_ctx *ctx = b->ctx;
    const ModelDesc &md = b->model->md;
    const BatchKernelEntry &k = vp_kernel_tables().batch[b->kernel];
    BatchArgs a{};
    a.md = md;
    a.x = (const double *)b->model->x_dev; a.w = b->w_dev; a.Y = b->Y; a.ld = b->ld; a.P = b->P;
    a.svd_eps = vp_rank_tol(b->rank_policy, b->svd_eps, md.m, VP_F64);
    vp_lm_config_from_options(VP_F64, md.q, opt, a.cfg);
    // start from the current parameters: alpha -> alpha0 (device copy), results go to alpha
    if (md.q > 0)
        VP_CUDA(ctx, cudaMemcpyAsync(b->alpha0, b->alpha, sizeof(double) * (size_t)md.q * b->P, cudaMemcpyDeviceToDevice, ctx->stream));
    a.alpha0 = b->alpha0; a.alpha_out = b->alpha; a.C_out = b->C; a.obj_out = b->obj; a.term_out = b->term; a.nfev_out = b->nfev;
    a.next = b->next;
    a.error = reinterpret_cast<unsigned int *>(b->next + 1);
    a.expc = vp_exp_table();
    VP_CUDA(ctx, cudaMemsetAsync(b->next, 0, 2 * sizeof(unsigned long long), ctx->stream));
    int occ = 0;
    VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, k.threads + 32 * k.lm_warps, b->smem));
    if (occ < 1) return vp_fail(ctx, VP_ERR_MODEL_TOO_LARGE, "vp_batch: kernel does not fit on an SM");
    long long grid = (long long)ctx->sm_count * occ;
    if (grid > b->P) grid = b->P;
    unsigned long long *ddbg = nullptr;
    if (ctx->opt.batch_dbg && DEV_ALLOC(ctx, &ddbg, sizeof(unsigned long long) * 8 * (size_t)grid) == cudaSuccess)
        cudaMemsetAsync(ddbg, 0, sizeof(unsigned long long) * 8 * (size_t)grid, ctx->stream);
    a.dbg = ddbg;
    void *args[] = {(void *)&a};
    VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3((unsigned)grid), dim3(k.threads + 32 * k.lm_warps), args, b->smem, ctx->stream));
    ctx->launches++;
    if (ddbg) { // diagnostics: per-CTA phase accumulators on stderr
        std::vector<unsigned long long> hd((size_t)grid * 8);
        if (cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
            cudaMemcpy(hd.data(), ddbg, sizeof(unsigned long long) * hd.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
            double ev = 0, lm = 0, wait = 0, evals = 0, basis = 0, sw0 = 0, nlm = 0;
            for (long long c = 0; c < grid; ++c) {
                ev += hd[c * 8]; lm += hd[c * 8 + 1]; wait += hd[c * 8 + 2]; evals += hd[c * 8 + 3]; basis += hd[c * 8 + 5]; sw0 += hd[c * 8 + 6]; nlm += hd[c * 8 + 7];
            }
            fprintf(stderr, "[vp batch dbg] CTAs %lld evaluations %.0f | per evaluation %.2f us (y + basis %.2f, sweep 0 + reduction %.2f) + %.2f us waiting for the LM warps | LM step %.2f us (on the LM warps, overlapped) | busy per CTA %.2f ms\n",
                    grid, evals, 1e-3 * ev / evals, 1e-3 * basis / evals, 1e-3 * sw0 / evals, 1e-3 * wait / evals, 1e-3 * lm / (nlm > 0 ? nlm : 1), 1e-6 * (ev + wait) / grid);
        }
        DEV_FREE(ctx, ddbg);
    }
    b->fitted = true;
    return VP_OK;
}

extern "C" int vp_batch_fit(vp_batch *b, const vp_lm_options *opt, vp_fit_report *reports)
{
    if (!b) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = b->ctx;
    cudaSetDevice(ctx->device);
    int rc = batch_fit_launch(b, opt);
    if (rc != VP_OK) return rc;
    unsigned int kerr = 0;
    VP_CUDA(ctx, cudaMemcpyAsync(&kerr, b->next + 1, sizeof(kerr), cudaMemcpyDeviceToHost, ctx->stream));
    if (!reports) {
        VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (kerr) return vp_fail(ctx, VP_ERR_CUDA, "vp_batch_fit: a hand-off inside the batch kernel timed out (GPU shared with another long-running kernel?)");
        return VP_OK;
    }
    std::vector<double> obj((size_t)b->P);
    std::vector<int> term((size_t)b->P), nfev((size_t)b->P);
    VP_CUDA(ctx, cudaMemcpyAsync(obj.data(), b->obj, sizeof(double) * (size_t)b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaMemcpyAsync(term.data(), b->term, sizeof(int) * (size_t)b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaMemcpyAsync(nfev.data(), b->nfev, sizeof(int) * (size_t)b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (kerr) return vp_fail(ctx, VP_ERR_CUDA, "vp_batch_fit: a hand-off inside the batch kernel timed out (GPU shared with another long-running kernel?)");
    for (int64_t p = 0; p < b->P; ++p) {
        reports[p].termination = term[(size_t)p];
        reports[p].number_of_evaluations = nfev[(size_t)p];
        reports[p].objective_function = obj[(size_t)p];
        reports[p].successful = lm_successful(term[(size_t)p]) ? 1 : 0;
        reports[p].reserved = 0;
    }
    return VP_OK;
}

extern "C" int vp_batch_params(vp_batch *b, double *alpha_out)
{
    if (!b || !alpha_out) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = b->ctx;
    cudaSetDevice(ctx->device);
    const int q = b->model->md.q;
    if (q == 0) return VP_OK;
    VP_CUDA(ctx, cudaMemcpyAsync(alpha_out, b->alpha, sizeof(double) * (size_t)q * b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VP_OK;
}

extern "C" int vp_batch_set_params(vp_batch *b, const double *alpha)
{
    if (!b || !alpha) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = b->ctx;
    cudaSetDevice(ctx->device);
    const int q = b->model->md.q;
    if (q == 0) return VP_OK;
    VP_CUDA(ctx, cudaMemcpyAsync(b->alpha, alpha, sizeof(double) * (size_t)q * b->P, cudaMemcpyHostToDevice, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    b->fitted = false;
    return VP_OK;
}

extern "C" int vp_batch_set_rank_policy(vp_batch *b, int policy)
{
    if (!b || (policy != VP_RANK_ABSOLUTE && policy != VP_RANK_RELATIVE)) return VP_ERR_INVALID_ARGUMENT;
    b->rank_policy = policy;
    return VP_OK;
}

extern "C" int vp_batch_linear_coefficients(vp_batch *b, double *C_out)
{
    if (!b || !C_out) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = b->ctx;
    if (!b->fitted) return vp_fail(ctx, VP_ERR_NO_CACHED_CALCULATION, "vp_batch: coefficients are available after vp_batch_fit");
    cudaSetDevice(ctx->device);
    VP_CUDA(ctx, cudaMemcpyAsync(C_out, b->C, sizeof(double) * (size_t)b->model->md.n * b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VP_OK;
}
