// vp_diag.cu -- diagnostics entry points of the C ABI: per-kernel device times of one evaluation (CUDA events on the
// context's stream; bench.py's single-evaluation roofline line) and the in-kernel timeline.
#include "vp_internal.h"

using namespace vp;

// ----------------------------------------------------------------------------
// diagnostics: per-kernel device times of one evaluation, CUDA events on the
// context's stream (used by bench.py for the roofline line)
// ----------------------------------------------------------------------------
extern "C" int vp_profile_evaluation(vp_problem *pr, int iters, int64_t flush_bytes, double *panel_us,
                                     double *stream_us, int64_t *stream_grid, int64_t *stream_smem)
{
    if (!pr || iters <= 0 || iters > 4096) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = pr->ctx;
    cudaSetDevice(ctx->device);
    std::vector<cudaEvent_t> ev(3 * (size_t)iters);
    for (auto &e : ev) VP_CUDA(ctx, cudaEventCreate(&e));
    void *flush = nullptr;
    if (flush_bytes > 0) VP_CUDA(ctx, cudaMalloc(&flush, (size_t)flush_bytes));
    const int q = pr->model->md.q;
    for (int k = 0; k < q; ++k) pr->alpha_stage[k] = pr->alpha[k];
    if (q > 0)
        VP_CUDA(ctx, cudaMemcpyAsync(pr->alpha_dev, pr->alpha_stage, sizeof(double) * q, cudaMemcpyHostToDevice, ctx->stream));
    int rc = VP_OK;
    // warm-up: keep the device busy long enough for the clocks to ramp up
    const bool fused = pr->plan_fit >= 0;
    if (pr->comm) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_profile_evaluation: not for column-sharded problems");
    if (pr->model->hosteval) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_profile_evaluation: not for host-evaluated models");
    for (int it = 0; it < 64 && rc == VP_OK; ++it) {
        if (fused) { rc = vp_launch_fused(pr, pr->cur ^ 1, false); continue; }
        rc = vp_launch_panel(pr);
        if (rc == VP_OK) rc = vp_launch_stream(pr, pr->cur ^ 1);
    }
    // timed launches are enqueued back to back; no host synchronisation in between
    for (int it = 0; it < iters && rc == VP_OK; ++it) {
        if (flush) cudaMemsetAsync(flush, it & 0xff, (size_t)flush_bytes, ctx->stream);
        cudaEventRecord(ev[3 * it + 0], ctx->stream);
        if (!fused) rc = vp_launch_panel(pr);
        cudaEventRecord(ev[3 * it + 1], ctx->stream);
        if (rc == VP_OK) rc = fused ? vp_launch_fused(pr, pr->cur ^ 1, false) : vp_launch_stream(pr, pr->cur ^ 1);
        cudaEventRecord(ev[3 * it + 2], ctx->stream);
    }
    cudaStreamSynchronize(ctx->stream);
    double tp = 0, ts = 0;
    if (rc == VP_OK)
        for (int it = 0; it < iters; ++it) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ev[3 * it + 0], ev[3 * it + 1]);
            cudaEventElapsedTime(&b, ev[3 * it + 1], ev[3 * it + 2]);
            tp += a; ts += b;
        }
    if (flush) cudaFree(flush);
    for (auto &e : ev) cudaEventDestroy(e);
    if (rc != VP_OK) return rc;
    if (panel_us) *panel_us = 1e3 * tp / iters;
    if (stream_us) *stream_us = 1e3 * ts / iters;
    if (stream_grid) *stream_grid = fused ? pr->fit_grid : pr->plan_grid;
    if (stream_smem) *stream_smem = (int64_t)(fused ? pr->fit_smem : pr->plan_smem);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return vp_fail(ctx, VP_ERR_CUDA, cudaGetErrorString(e));
    return VP_OK;
}

// One evaluation with the in-kernel timeline enabled: out receives
// (grid + 1) * VP_DBG_SLOTS %globaltimer stamps in ns (the last row is the panel
// kernel's), relative to the smallest stamp. Diagnostics only.
extern "C" int vp_debug_timeline(vp_problem *pr, long long *out, int64_t capacity, int64_t *grid_out)
{
    if (!pr || !out) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = pr->ctx;
    cudaSetDevice(ctx->device);
    const size_t n = ((size_t)pr->max_grid + 1) * VP_DBG_SLOTS;
    const bool read_only = pr->dbg != nullptr; // VP_DBG_FIT: stamps of the last evaluation of the last fit
    if (!pr->dbg) VP_CUDA(ctx, cudaMalloc(&pr->dbg, n * sizeof(unsigned long long)));
    int rc = VP_OK;
    for (int it = 0; it < 3 && rc == VP_OK && !read_only; ++it) { // warm, then the recorded one
        VP_CUDA(ctx, cudaMemsetAsync(pr->dbg, 0, n * sizeof(unsigned long long), ctx->stream));
        rc = vp_launch_eval(pr, pr->cur ^ 1);
    }
    std::vector<unsigned long long> h(n);
    VP_CUDA(ctx, cudaMemcpyAsync(h.data(), pr->dbg, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!read_only) {
        cudaFree(pr->dbg);
        pr->dbg = nullptr;
    }
    if (rc != VP_OK) return rc;
    unsigned long long t0 = ~0ull;
    for (auto v : h) if (v && v < t0) t0 = v;
    const size_t rows = (size_t)(pr->plan_fit >= 0 ? pr->fit_grid : pr->plan_grid);
    size_t k = 0;
    for (size_t b = 0; b <= rows && k + VP_DBG_SLOTS <= (size_t)capacity; ++b) {
        const size_t src = (b < rows ? b : (size_t)pr->max_grid) * VP_DBG_SLOTS;
        for (int s2 = 0; s2 < VP_DBG_SLOTS; ++s2) out[k++] = h[src + s2] ? (long long)(h[src + s2] - t0) : -1;
    }
    if (grid_out) *grid_out = (int64_t)rows;
    return VP_OK;
}

// ----------------------------------------------------------------------------
// fp64 ALU peaks of this GPU, measured on the context's stream (bench.py: denominator of the fp64-ALU
// roofline of the independent-batch kernel; MEASURED_PEAKS.json has no fp64 entry)
// ----------------------------------------------------------------------------
__global__ void vp_peak_dfma_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void vp_peak_dexp_kernel(double *out, int iters)
{
    double x = -1.0 - threadIdx.x * 1e-3, s = 0;
    for (int i = 0; i < iters; ++i) { s += exp(x); x -= 1e-6; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int vp_measure_fp64_peaks(vp_ctx *ctx, double *dfma_tflops, double *dexp_gexps)
{
    if (!ctx) return VP_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 20000;
    double *out = nullptr;
    VP_CUDA(ctx, DEV_ALLOC(ctx, &out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t ev[4];
    for (auto &e : ev) cudaEventCreate(&e);
    for (int rep = 0; rep < 2; ++rep) { // first round warms up
        cudaEventRecord(ev[0], ctx->stream);
        vp_peak_dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters);
        cudaEventRecord(ev[1], ctx->stream);
        cudaEventRecord(ev[2], ctx->stream);
        vp_peak_dexp_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters / 10);
        cudaEventRecord(ev[3], ctx->stream);
        ctx->launches += 2;
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    float ms_fma = 0, ms_exp = 0;
    cudaEventElapsedTime(&ms_fma, ev[0], ev[1]);
    cudaEventElapsedTime(&ms_exp, ev[2], ev[3]);
    for (auto &x : ev) cudaEventDestroy(x);
    DEV_FREE(ctx, out);
    if (e != cudaSuccess) return vp_fail(ctx, VP_ERR_CUDA, cudaGetErrorString(e));
    if (dfma_tflops) *dfma_tflops = 2.0 * 8 * iters * (double)blocks * threads / ms_fma / 1e9;
    if (dexp_gexps) *dexp_gexps = (double)(iters / 10) * blocks * threads / ms_exp / 1e6;
    return VP_OK;
}
