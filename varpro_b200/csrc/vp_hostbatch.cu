// vp_hostbatch.cu -- vp_fit_host_batch: many problems whose observations live in HOST memory, pipelined.
//
// What a caller of the reference does for a set of measured data sets -- for each: SeparableProblemBuilder::
// observations(Y_i).build(), LevMarSolver::fit, read the parameters and coefficients -- is bound by the
// host-to-device copy of Y_i (33.6 MB for BASELINE config 2) once the fit itself takes a fraction of a millisecond.
// This entry point keeps the copy engine busy: a few worker threads of the library, each with its own context
// (= CUDA stream + buffer pool) on the caller's device, take problems from a shared counter and run the ordinary
// entry points (vp_problem_create -> vp_fit -> vp_params / vp_linear_coefficients -> vp_problem_destroy), so the
// copy of one problem overlaps the fit and the read-back of the others. No Python, no GIL, no per-call allocation
// (the worker contexts and their pools live as long as the caller's context).
#include <atomic>
#include <thread>

#include "vp_internal.h"

using namespace vp;

static std::mutex g_workers_mu; // guards the lazily created worker contexts of a caller context

extern "C" int vp_fit_host_batch(vp_ctx *ctx, int dtype, int64_t m, const void *x_host, int32_t q, int32_t n,
                                 const vp_basis_desc *basis, int64_t n_problems, int64_t S, const void *const *Y_hosts,
                                 int64_t ldY, const void *w_host, double svd_eps, const double *alpha0,
                                 const vp_lm_options *opt, int32_t workers, vp_fit_report *reports, double *alpha_out,
                                 void *const *C_outs)
{
    VP_NVTX("vp_fit_host_batch");
    if (!ctx || n_problems < 0 || (n_problems > 0 && (!Y_hosts || !reports))) return VP_ERR_INVALID_ARGUMENT;
    if (n_problems == 0) return VP_OK;
    if (workers <= 0) workers = 3;
    if (workers > 16) workers = 16;
    if (workers > n_problems) workers = (int32_t)n_problems;
    std::vector<vp_ctx *> wctx;
    {
        std::lock_guard<std::mutex> lock(g_workers_mu);
        while ((int)ctx->workers.size() < workers) { // live as long as the caller's context (vp_ctx_destroy)
            vp_ctx *c = nullptr;
            const int rc = vp_ctx_create(ctx->device, &c);
            if (rc != VP_OK) return vp_fail(ctx, rc, std::string("vp_fit_host_batch: cannot create a worker context: ") + vp_last_error(nullptr));
            ctx->workers.push_back(c);
        }
        wctx.assign(ctx->workers.begin(), ctx->workers.begin() + workers);
        for (vp_ctx *c : wctx) c->opt = ctx->opt; // same tunables as the caller's context
    }
    std::atomic<int64_t> next{0};
    std::vector<int> status((size_t)workers, VP_OK);
    std::vector<std::string> message((size_t)workers);
    auto work = [&](int t) {
        vp_ctx *c = wctx[(size_t)t];
        vp_model *model = nullptr;
        int rc = vp_model_create(c, dtype, m, x_host, q, n, basis, &model);
        while (rc == VP_OK) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_problems) break;
            vp_problem *pr = nullptr;
            rc = vp_problem_create(c, model, S, Y_hosts[i], ldY, w_host, svd_eps, alpha0, &pr); // the H2D copy of this step
            if (rc != VP_OK) break;
            rc = vp_fit(pr, opt, &reports[i]);
            if (rc == VP_OK && alpha_out && q > 0) rc = vp_params(pr, alpha_out + (size_t)i * q);
            if (rc == VP_OK && C_outs && C_outs[i]) {
                rc = vp_linear_coefficients(pr, C_outs[i]); // the D2H read of this step
                if (rc == VP_ERR_NO_CACHED_CALCULATION) rc = VP_OK; // unsuccessful fit: the report says so
            }
            vp_problem_destroy(pr);
        }
        if (rc != VP_OK) {
            status[(size_t)t] = rc;
            message[(size_t)t] = vp_last_error(c);
            next.store(n_problems); // stop the others
        }
        vp_model_destroy(model);
    };
    std::vector<std::thread> threads;
    for (int t = 1; t < workers; ++t) threads.emplace_back(work, t);
    work(0);
    for (std::thread &th : threads) th.join();
    for (int t = 0; t < workers; ++t) {
        ctx->launches += wctx[(size_t)t]->launches;
        wctx[(size_t)t]->launches = 0;
    }
    for (int t = 0; t < workers; ++t)
        if (status[(size_t)t] != VP_OK) return vp_fail(ctx, status[(size_t)t], "vp_fit_host_batch: " + message[(size_t)t]);
    return VP_OK;
}
