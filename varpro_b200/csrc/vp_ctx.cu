// vp_ctx.cu -- contexts, options, buffer pools, status strings, kernel tables and models of the C ABI
// (include/varpro_b200.h).
#include "vp_internal.h"

using namespace vp;

// ----------------------------------------------------------------------------
// buffer pools
// ----------------------------------------------------------------------------
cudaError_t pool_alloc(BufferPool &pool, bool host, void **out, size_t bytes)
{
    bytes = (bytes + 255) / 256 * 256;
    auto hit = pool.free_list.find(bytes);
    if (hit != pool.free_list.end()) {
        *out = hit->second;
        pool.free_list.erase(hit);
        pool.free_bytes -= bytes;
        pool.live[*out] = bytes;
        return cudaSuccess;
    }
    cudaError_t e = host ? cudaMallocHost(out, bytes) : cudaMalloc(out, bytes);
    if (e != cudaSuccess && !pool.free_list.empty()) {
        // out of memory: drop the cache and retry once
        pool_release(pool, host);
        cudaGetLastError();
        e = host ? cudaMallocHost(out, bytes) : cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) pool.live[*out] = bytes;
    return e;
}

void pool_free(BufferPool &pool, bool host, void *p)
{
    if (!p) return;
    auto it = pool.live.find(p);
    if (it == pool.live.end()) { host ? cudaFreeHost(p) : cudaFree(p); return; }
    const size_t bytes = it->second;
    pool.live.erase(it);
    if (pool.free_bytes + bytes <= pool.cap_bytes) {
        pool.free_list.emplace(bytes, p);
        pool.free_bytes += bytes;
    } else {
        host ? cudaFreeHost(p) : cudaFree(p);
    }
}

void pool_release(BufferPool &pool, bool host)
{
    for (auto &b : pool.free_list) host ? cudaFreeHost(b.second) : cudaFree(b.second);
    pool.free_list.clear();
    pool.free_bytes = 0;
}

// ----------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------
static thread_local std::string g_last_error_noctx;

int vp_fail(vp_ctx *ctx, int code, const std::string &msg)
{
    if (ctx)
        ctx->last_error = msg;
    else
        g_last_error_noctx = msg;
    return code;
}

cudaError_t vp_ensure_dynamic_smem(int device, const void *fn, size_t bytes)
{
    struct Key {
        int device;
        const void *fn;
        bool operator==(const Key &o) const { return device == o.device && fn == o.fn; }
    };
    struct KeyHash {
        size_t operator()(const Key &k) const { return std::hash<const void *>()(k.fn) ^ (std::hash<int>()(k.device) * 0x9e3779b97f4a7c15ull); }
    };
    static std::mutex mu;
    static std::unordered_map<Key, size_t, KeyHash> current;
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = current[Key{device, fn}];
    if (bytes <= cur) return cudaSuccess;
    int now = -1;
    cudaGetDevice(&now);
    if (now != device) cudaSetDevice(device);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

extern "C" int vp_abi_version(void) { return VP_ABI_VERSION; }

extern "C" const char *vp_status_string(int s)
{
    switch (s) {
    case VP_OK: return "ok";
    case VP_ERR_Y_DATA_MISSING: return "Right hand side(s) not provided";
    case VP_ERR_INVALID_LENGTH_OF_DATA: return "Vectors x and y must have same lengths";
    case VP_ERR_ZERO_LENGTH_VECTOR: return "x or y must have nonzero number of elements";
    case VP_ERR_INVALID_PARAMETER_COUNT: return "Initial guess vector must have same length as parameters";
    case VP_ERR_INVALID_LENGTH_OF_WEIGHTS: return "The weights must have the same length as the data y";
    case VP_ERR_PARAMETER_NOT_IN_MODEL: return "Parameter is not in model";
    case VP_ERR_DERIVATIVE_INDEX_OUT_OF_BOUNDS: return "Index for derivative is out of bounds";
    case VP_ERR_INCORRECT_PARAMETER_COUNT: return "Model expects a different number of parameters";
    case VP_ERR_EMPTY_MODEL: return "Model contains no basis functions";
    case VP_ERR_UNUSED_PARAMETER: return "A model parameter is not used by any basis function";
    case VP_ERR_UNSUPPORTED_BASIS: return "Unsupported basis-function kind";
    case VP_ERR_MODEL_TOO_LARGE: return "Model exceeds the compiled-in size limits";
    case VP_ERR_NO_CACHED_CALCULATION: return "No cached calculation (evaluation failed)";
    case VP_ERR_UNDERDETERMINED: return "Problem is underdetermined";
    case VP_ERR_MATRIX_INVERSION: return "Matrix inversion failed";
    case VP_ERR_INVALID_ARGUMENT: return "Invalid argument";
    case VP_ERR_CUDA: return "CUDA error";
    case VP_ERR_OUT_OF_MEMORY: return "Out of device memory";
    case VP_ERR_COMM: return "Communicator error";
    default: return "unknown status";
    }
}

// ----------------------------------------------------------------------------
// options
// ----------------------------------------------------------------------------
static int parse_option(CtxOptions &o, const char *key, const char *value)
{
    if (!key || !value) return VP_ERR_INVALID_ARGUMENT;
    const std::string k(key), v(value);
    auto as_int = [&](int &dst) { dst = atoi(value); return VP_OK; };
    if (k == "fit_mode") {
        if (v == "auto" || v == "persistent" || v == "queue") o.fit_mode = VP_FITMODE_AUTO;
        else if (v == "host") o.fit_mode = VP_FITMODE_HOST;
        else if (v == "graph") o.fit_mode = VP_FITMODE_GRAPH;
        else return VP_ERR_INVALID_ARGUMENT;
        return VP_OK;
    }
    if (k == "eval_kernel") {
        if (v == "fused") o.eval_split = 0;
        else if (v == "split") o.eval_split = 1;
        else return VP_ERR_INVALID_ARGUMENT;
        return VP_OK;
    }
    if (k == "stream_kernel") {
        if (v == "auto" || v == "dmma") o.stream_kernel = VP_STREAMK_AUTO;
        else if (v == "simt") o.stream_kernel = VP_STREAMK_SIMT;
        else if (v == "generic") o.stream_kernel = VP_STREAMK_GENERIC;
        else return VP_ERR_INVALID_ARGUMENT;
        return VP_OK;
    }
    if (k == "panel_generic") return as_int(o.panel_generic);
    if (k == "stream_stages") return as_int(o.stream_stages);
    if (k == "stream_ct") return as_int(o.stream_ct);
    if (k == "stream_occ") return as_int(o.stream_occ);
    if (k == "queue_items_per_cta") { as_int(o.queue_items_per_cta); if (o.queue_items_per_cta < 1) o.queue_items_per_cta = 1; return VP_OK; }
    if (k == "queue_dbg") return as_int(o.queue_dbg);
    if (k == "batch_dbg") return as_int(o.batch_dbg);
    if (k == "queue_parts_per_item") return as_int(o.queue_parts_per_item);
    if (k == "trace") return as_int(o.trace);
    if (k == "dbg_fit") return as_int(o.dbg_fit);
    if (k == "batch_slots") return as_int(o.batch_slots);
    if (k == "pool_mb") return as_int(o.pool_mb);
    if (k == "max_ctas") return as_int(o.max_ctas);
    if (k == "fit_warps") return as_int(o.fit_warps);
    return VP_ERR_INVALID_ARGUMENT;
}

// environment defaults, read once per context
static void options_from_env(CtxOptions &o)
{
    static const char *const map[][2] = {
        {"VP_FIT_MODE", "fit_mode"}, {"VP_EVAL_KERNEL", "eval_kernel"}, {"VP_STREAM_KERNEL", "stream_kernel"},
        {"VP_PANEL_GENERIC", "panel_generic"}, {"VP_STREAM_STAGES", "stream_stages"}, {"VP_STREAM_CT", "stream_ct"},
        {"VP_STREAM_OCC", "stream_occ"}, {"VP_QUEUE_ITEMS_PER_CTA", "queue_items_per_cta"}, {"VP_QUEUE_DBG", "queue_dbg"}, {"VP_BATCH_DBG", "batch_dbg"}, {"VP_QUEUE_PARTS_PER_ITEM", "queue_parts_per_item"},
        {"VP_TRACE", "trace"}, {"VP_DBG_FIT", "dbg_fit"}, {"VP_BATCH_SLOTS", "batch_slots"}, {"VP_POOL_MB", "pool_mb"}, {"VP_MAX_CTAS", "max_ctas"},
        {"VP_FIT_WARPS", "fit_warps"}};
    for (const auto &m : map) {
        const char *s = getenv(m[0]);
        if (s && *s) parse_option(o, m[1], s);
    }
}

// ----------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------
extern "C" int vp_ctx_create(int device, vp_ctx **out)
{
    VP_NVTX("vp_ctx_create");
    if (!out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return vp_fail(nullptr, VP_ERR_CUDA,
                       std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return vp_fail(nullptr, VP_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    vp_ctx *ctx = new (std::nothrow) vp_ctx();
    if (!ctx) return VP_ERR_OUT_OF_MEMORY;
    ctx->device = device;
    cudaDeviceProp prop{};
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return vp_fail(nullptr, VP_ERR_CUDA, "cannot initialise device");
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    options_from_env(ctx->opt);
    ctx->dev_pool.cap_bytes = (size_t)(ctx->opt.pool_mb > 0 ? ctx->opt.pool_mb : 0) << 20;
    ctx->host_pool.cap_bytes = (size_t)64 << 20;
    *out = ctx;
    return VP_OK;
}

extern "C" int vp_ctx_set_option(vp_ctx *ctx, const char *key, const char *value)
{
    if (!ctx) return VP_ERR_INVALID_ARGUMENT;
    const int rc = parse_option(ctx->opt, key, value);
    if (rc != VP_OK)
        return vp_fail(ctx, rc, std::string("vp_ctx_set_option: unknown option or value: ") + (key ? key : "(null)") + " = " + (value ? value : "(null)"));
    ctx->dev_pool.cap_bytes = (size_t)(ctx->opt.pool_mb > 0 ? ctx->opt.pool_mb : 0) << 20;
    return VP_OK;
}

extern "C" int vp_ctx_trim(vp_ctx *ctx)
{
    if (!ctx) return VP_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pool_release(ctx->dev_pool, false);
    pool_release(ctx->host_pool, true);
    return VP_OK;
}

extern "C" int vp_ctx_destroy(vp_ctx *ctx)
{
    if (!ctx) return VP_OK;
    for (vp_ctx *w : ctx->workers) vp_ctx_destroy(w);
    ctx->workers.clear();
    cudaSetDevice(ctx->device);
    pool_release(ctx->dev_pool, false);
    pool_release(ctx->host_pool, true);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VP_OK;
}

extern "C" const char *vp_last_error(const vp_ctx *ctx)
{
    return ctx ? ctx->last_error.c_str() : g_last_error_noctx.c_str();
}
extern "C" int64_t vp_ctx_kernel_launches(const vp_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *vp_ctx_stream(const vp_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// ----------------------------------------------------------------------------
// kernel tables
// ----------------------------------------------------------------------------
// The templated fast-path kernels are instantiated in separate translation units (inst.cu, one per
// model shape and kernel family; see kernel_tables.h); gathered once (function-local static: thread safe).
const KernelTables &vp_kernel_tables()
{
    static const KernelTables tables = [] {
        KernelTables t;
#define VP_GATHER(tag, T, DT, N, P, PART)                                          \
    {                                                                              \
        const KernelGroup *g = vp_kernel_group_##tag();                            \
        t.stream.insert(t.stream.end(), g->simt, g->simt + g->nsimt);              \
        t.dmma.insert(t.dmma.end(), g->dmma, g->dmma + g->ndmma);                  \
        t.panel.insert(t.panel.end(), g->panel, g->panel + g->npanel);             \
        t.fit.insert(t.fit.end(), g->fit, g->fit + g->nfit);                       \
        t.batch.insert(t.batch.end(), g->batch, g->batch + g->nbatch);             \
        t.queue.insert(t.queue.end(), g->queue, g->queue + g->nqueue);             \
    }
        VP_KERNEL_GROUPS(VP_GATHER)
#undef VP_GATHER
        return t;
    }();
    return tables;
}

// ----------------------------------------------------------------------------
// model
// ----------------------------------------------------------------------------
static int basis_arity(int kind)
{
    switch (kind) {
    case VP_BASIS_EXP_DECAY: return 1;
    case VP_BASIS_CONSTANT: return 0;
    case VP_BASIS_EXP_RATE_COS: return 2;
    case VP_BASIS_SIN_PHASE: return 2;
    case VP_BASIS_LINEAR_X: return 0;
    default: return -1;
    }
}

extern "C" int vp_model_create(vp_ctx *ctx, int dtype, int64_t m, const void *x_host, int32_t q, int32_t n,
                               const vp_basis_desc *basis, vp_model **out)
{
    VP_NVTX("vp_model_create");
    if (!ctx || !out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (dtype != VP_F64 && dtype != VP_F32) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "dtype must be VP_F64 or VP_F32");
    if (n <= 0 || !basis) return vp_fail(ctx, VP_ERR_EMPTY_MODEL, vp_status_string(VP_ERR_EMPTY_MODEL));
    if (m <= 0 || !x_host) return vp_fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (n > VP_MAX_N || q > VP_MAX_Q || q < 0 || m > (1 << 24))
        return vp_fail(ctx, VP_ERR_MODEL_TOO_LARGE, vp_status_string(VP_ERR_MODEL_TOO_LARGE));
    ModelDesc md{};
    md.m = (int)m; md.n = n; md.q = q; md.p = 0;
    std::vector<int> used(q > 0 ? q : 1, 0);
    for (int j = 0; j < n; ++j) {
        const vp_basis_desc &b = basis[j];
        const int ar = basis_arity(b.kind);
        if (ar < 0) return vp_fail(ctx, VP_ERR_UNSUPPORTED_BASIS, vp_status_string(VP_ERR_UNSUPPORTED_BASIS));
        if (b.n_params != ar)
            return vp_fail(ctx, VP_ERR_INCORRECT_PARAMETER_COUNT, "basis function " + std::to_string(j) + " expects " +
                                                                      std::to_string(ar) + " parameters, but got " +
                                                                      std::to_string(b.n_params));
        md.kind[j] = b.kind;
        md.npar[j] = ar;
        md.scale[j] = b.scale;
        for (int s = 0; s < ar; ++s) {
            const int k = b.param_idx[s];
            if (k < 0 || k >= q) return vp_fail(ctx, VP_ERR_PARAMETER_NOT_IN_MODEL, vp_status_string(VP_ERR_PARAMETER_NOT_IN_MODEL));
            md.pidx[j][s] = k;
            used[k] = 1;
            if (md.p >= VP_MAX_P) return vp_fail(ctx, VP_ERR_MODEL_TOO_LARGE, vp_status_string(VP_ERR_MODEL_TOO_LARGE));
            md.e_basis[md.p] = j;
            md.e_slot[md.p] = s;
            md.e_param[md.p] = k;
            md.p++;
        }
    }
    for (int k = 0; k < q; ++k)
        if (!used[k]) return vp_fail(ctx, VP_ERR_UNUSED_PARAMETER, vp_status_string(VP_ERR_UNUSED_PARAMETER));

    vp_model *mo = new (std::nothrow) vp_model();
    if (!mo) return VP_ERR_OUT_OF_MEMORY;
    mo->ctx = ctx; mo->dtype = dtype; mo->md = md;
    const int v = vp_vec_of(dtype);
    mo->ld = (int)((m + v - 1) / v * v);
    cudaSetDevice(ctx->device);
    cudaError_t e = DEV_ALLOC(ctx, &mo->x_dev, vp_esize(dtype) * (size_t)m);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mo->x_dev, x_host, vp_esize(dtype) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        DEV_FREE(ctx, mo->x_dev);
        delete mo;
        return vp_fail(ctx, VP_ERR_CUDA, std::string("vp_model_create: ") + cudaGetErrorString(e));
    }
    *out = mo;
    return VP_OK;
}

extern "C" int vp_model_create_hosteval(vp_ctx *ctx, int dtype, int64_t m, int32_t q, int32_t n, int32_t p, const int32_t *ind,
                                        vp_host_eval_fn eval, void *user, vp_model **out)
{
    VP_NVTX("vp_model_create_hosteval");
    if (!ctx || !out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (dtype != VP_F64 && dtype != VP_F32) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "dtype must be VP_F64 or VP_F32");
    if (!eval) return vp_fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_model_create_hosteval: eval callback is NULL");
    if (n <= 0) return vp_fail(ctx, VP_ERR_EMPTY_MODEL, vp_status_string(VP_ERR_EMPTY_MODEL));
    if (m <= 0) return vp_fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (n > VP_MAX_N || q > VP_MAX_Q || q < 0 || p < 0 || p > VP_MAX_P || m > (1 << 24) || (p > 0 && !ind))
        return vp_fail(ctx, VP_ERR_MODEL_TOO_LARGE, vp_status_string(VP_ERR_MODEL_TOO_LARGE));
    ModelDesc md{};
    md.m = (int)m; md.n = n; md.q = q; md.p = p;
    for (int j = 0; j < n; ++j) { md.kind[j] = VP_BASIS_HOST; md.npar[j] = 0; }
    std::vector<int> used(q > 0 ? q : 1, 0);
    for (int e = 0; e < p; ++e) {
        const int j = ind[2 * e], k = ind[2 * e + 1];
        if (j < 0 || j >= n) return vp_fail(ctx, VP_ERR_DERIVATIVE_INDEX_OUT_OF_BOUNDS, vp_status_string(VP_ERR_DERIVATIVE_INDEX_OUT_OF_BOUNDS));
        if (k < 0 || k >= q) return vp_fail(ctx, VP_ERR_PARAMETER_NOT_IN_MODEL, vp_status_string(VP_ERR_PARAMETER_NOT_IN_MODEL));
        md.e_basis[e] = j; md.e_param[e] = k; md.e_slot[e] = 0;
        used[k] = 1;
    }
    for (int k = 0; k < q; ++k)
        if (!used[k]) return vp_fail(ctx, VP_ERR_UNUSED_PARAMETER, vp_status_string(VP_ERR_UNUSED_PARAMETER));
    vp_model *mo = new (std::nothrow) vp_model();
    if (!mo) return VP_ERR_OUT_OF_MEMORY;
    mo->ctx = ctx; mo->dtype = dtype; mo->md = md;
    mo->hosteval = true; mo->eval_fn = eval; mo->eval_user = user;
    const int v = vp_vec_of(dtype);
    mo->ld = (int)((m + v - 1) / v * v);
    cudaSetDevice(ctx->device);
    const size_t bytes = sizeof(double) * (size_t)m * (n + p);
    cudaError_t e = DEV_ALLOC(ctx, &mo->pre_dev, bytes);
    if (e == cudaSuccess) e = HOST_ALLOC(ctx, &mo->pre_host, bytes);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &mo->x_dev, vp_esize(dtype) * (size_t)m); // unused by the kernels; keeps the layout uniform
    if (e == cudaSuccess) e = cudaMemsetAsync(mo->x_dev, 0, vp_esize(dtype) * (size_t)m, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        DEV_FREE(ctx, mo->pre_dev); HOST_FREE(ctx, mo->pre_host); DEV_FREE(ctx, mo->x_dev);
        delete mo;
        return vp_fail(ctx, VP_ERR_CUDA, std::string("vp_model_create_hosteval: ") + cudaGetErrorString(e));
    }
    *out = mo;
    return VP_OK;
}

extern "C" int vp_model_destroy(vp_model *model)
{
    if (!model) return VP_OK;
    cudaSetDevice(model->ctx->device);
    DEV_FREE(model->ctx, model->pre_dev);
    HOST_FREE(model->ctx, model->pre_host);
    DEV_FREE(model->ctx, model->x_dev);
    delete model;
    return VP_OK;
}
