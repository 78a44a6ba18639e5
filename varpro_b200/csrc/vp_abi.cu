// vp_abi.cu -- implementation of the C ABI declared in include/varpro_b200.h.
// Host-side handle management, kernel dispatch and the LM driver. There is no
// CPU compute path in this file: every evaluation launches the sm_100a kernels.
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../../include/varpro_b200.h"
#include "aux_kernels.cuh"
#include "device_common.cuh"
#include "kernel_tables.h"
#include "lm_step.cuh"
#include "panel_kernel.cuh"
#include "stream_kernel.cuh"
#include "fit_kernel_dmma.cuh"
#include "batch_fit_kernel.cuh"
#include "fit_queue_kernel.cuh"

using namespace vp;

// ----------------------------------------------------------------------------
// handles
// ----------------------------------------------------------------------------
// Size-keyed free lists for device and pinned-host buffers: problems of the same shape are
// created and destroyed per fit by callers that mirror the reference API (the builder produces
// a new SeparableProblem every time), and cudaMalloc / cudaMallocHost cost more than a fit.
struct BufferPool {
    std::unordered_map<void *, size_t> live;
    std::vector<std::pair<size_t, void *>> free_list;
    size_t free_bytes = 0;
    size_t cap_bytes = 0;
};

struct vp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t launches = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    BufferPool dev_pool, host_pool;
    std::vector<cudaStream_t> side_streams; // vp_fit_many: one per concurrent fit
    cudaEvent_t fork_event = nullptr, join_event = nullptr;
};

static cudaError_t pool_alloc(BufferPool &pool, bool host, void **out, size_t bytes)
{
    bytes = (bytes + 255) / 256 * 256;
    for (size_t i = 0; i < pool.free_list.size(); ++i)
        if (pool.free_list[i].first == bytes) {
            *out = pool.free_list[i].second;
            pool.free_list[i] = pool.free_list.back();
            pool.free_list.pop_back();
            pool.free_bytes -= bytes;
            pool.live[*out] = bytes;
            return cudaSuccess;
        }
    cudaError_t e = host ? cudaMallocHost(out, bytes) : cudaMalloc(out, bytes);
    if (e != cudaSuccess && !pool.free_list.empty()) {
        // out of memory: drop the cache and retry once
        for (auto &b : pool.free_list) host ? cudaFreeHost(b.second) : cudaFree(b.second);
        pool.free_list.clear();
        pool.free_bytes = 0;
        cudaGetLastError();
        e = host ? cudaMallocHost(out, bytes) : cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) pool.live[*out] = bytes;
    return e;
}

static void pool_free(BufferPool &pool, bool host, void *p)
{
    if (!p) return;
    auto it = pool.live.find(p);
    if (it == pool.live.end()) { host ? cudaFreeHost(p) : cudaFree(p); return; }
    const size_t bytes = it->second;
    pool.live.erase(it);
    if (pool.free_bytes + bytes <= pool.cap_bytes) {
        pool.free_list.emplace_back(bytes, p);
        pool.free_bytes += bytes;
    } else {
        host ? cudaFreeHost(p) : cudaFree(p);
    }
}

static void pool_release(BufferPool &pool, bool host)
{
    for (auto &b : pool.free_list) host ? cudaFreeHost(b.second) : cudaFree(b.second);
    pool.free_list.clear();
    pool.free_bytes = 0;
}
#define DEV_ALLOC(ctx, pp, bytes) pool_alloc((ctx)->dev_pool, false, (void **)(pp), (bytes))
#define HOST_ALLOC(ctx, pp, bytes) pool_alloc((ctx)->host_pool, true, (void **)(pp), (bytes))
#define DEV_FREE(ctx, p) pool_free((ctx)->dev_pool, false, (void *)(p))
#define HOST_FREE(ctx, p) pool_free((ctx)->host_pool, true, (void *)(p))

struct vp_model {
    vp_ctx *ctx = nullptr;
    int dtype = VP_F64;
    ModelDesc md{};
    void *x_dev = nullptr;
    int ld = 0; // padded row count (multiple of 16/sizeof(T))
    // host-evaluated model (vp_model_create_hosteval)
    bool hosteval = false;
    vp_host_eval_fn eval_fn = nullptr;
    void *eval_user = nullptr;
    double *pre_host = nullptr; // pinned m x (n+p)
    double *pre_dev = nullptr;  // the unweighted [Phi | D] of the last callback
};

// Column-sharded global fit across the GPUs of one box (one process per GPU): this rank's
// mailbox (device memory exported through CUDA IPC) plus the peers' mailboxes mapped over NVLink.
struct vp_comm {
    vp_ctx *ctx = nullptr;
    int world = 1, rank = 0;
    bool connected = false;
    CommMailbox *local_box = nullptr;
    unsigned long long *epoch = nullptr; // device
    int *error = nullptr;                // device
    void *peer[COMM_MAX_WORLD] = {nullptr};
    CommArgs args{};
};

struct vp_problem {
    vp_ctx *ctx = nullptr;
    vp_comm *comm = nullptr;
    vp_model *model = nullptr;
    int64_t S = 0;
    void *Yw = nullptr;    // ld x S
    void *w_dev = nullptr; // m or null
    double svd_eps = 0.0;
    double alpha[VP_MAX_Q] = {0};
    double *alpha_dev = nullptr; // = &fit_dev->st.x_trial[0]: the parameters the next evaluation is made at
    FitDevice *fit_dev = nullptr; // device-resident LM state (graph path)
    FitDevice *fit_host = nullptr; // pinned staging copy
    cudaGraph_t fit_graph = nullptr;
    cudaGraphExec_t fit_exec = nullptr;
    cudaGraphConditionalHandle fit_cond = 0;
    void *Pq = nullptr; // panel [Q | E | 0], (n+p+1) columns of ldp rows
    int ldp = 0;
    PanelSmall *small = nullptr;
    void *C[2] = {nullptr, nullptr}; // coefficient buffers (n x S); C[cur] belongs to `alpha`
    int cur = 0;
    double *partials = nullptr;
    int red_stride = 0;
    int max_grid = 0;
    unsigned int *ticket = nullptr;
    EvalOut *out_dev = nullptr;
    EvalOut *out_host = nullptr; // pinned
    double *alpha_stage = nullptr; // pinned
    double *phi_scratch = nullptr; // m x n (best_fit)
    unsigned long long *dbg = nullptr; // optional in-kernel timeline (vp_debug_timeline)
    LmEval eval{};                 // reduction at `alpha`
    bool cached = false;
    // streaming-kernel launch plan
    int plan_kind = -1; // index into the dispatch table, -1 = generic
    int plan_dmma = -1; // index into the DMMA dispatch table (fp64 fast path), -1 = not used
    int plan_lds = 0;   // padded shared-memory column stride of the DMMA path
    int plan_rows = 0;  // rows the chosen kernel touches (panel must be zero-padded that far)
    int plan_ct = 1;    // columns per tile
    int plan_grid = 0, plan_nst = 0;
    size_t plan_smem = 0;
    // fused evaluation / persistent fit kernel (fit_kernel_dmma), -1 = not available
    int plan_fit = -1;
    int fit_grid = 0, fit_nst = 0;
    size_t fit_smem = 0;
    FitBcast *bcast = nullptr;
    int jac_full = 0;       // 1: full Golub-Pereyra Jacobian (vp_problem_set_jacobian); 0: Kaufman (the reference)
    double *Pq64 = nullptr; // f64 panel buffer used by the work-queue kernel for fp32 problems (lazily allocated)
    int ldp64 = 0;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PROCESS-WIDE property of a kernel: problems of different
// sizes (and host threads) share one instantiation, so the limit may only ever be raised -- a later,
// smaller request must not lower it under a launch that was planned with the larger value.
static cudaError_t ensure_dynamic_smem(const void *fn, size_t bytes)
{
    static std::mutex mu;
    static std::unordered_map<const void *, size_t> current;
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = current[fn];
    if (bytes <= cur) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

static thread_local std::string g_last_error_noctx;

static int fail(vp_ctx *ctx, int code, const std::string &msg)
{
    if (ctx)
        ctx->last_error = msg;
    else
        g_last_error_noctx = msg;
    return code;
}

#define VP_CUDA(ctx, expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail((ctx), (_e == cudaErrorMemoryAllocation) ? VP_ERR_OUT_OF_MEMORY : VP_ERR_CUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
    } while (0)

static size_t esize(int dtype) { return dtype == VP_F32 ? 4 : 8; }
static int vec_of(int dtype) { return dtype == VP_F32 ? 4 : 2; }

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

extern "C" int vp_ctx_create(int device, vp_ctx **out)
{
    if (!out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VP_ERR_CUDA,
                    std::string("no CUDA device available (this library has no CPU fallback): ") +
                        cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, VP_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    vp_ctx *ctx = new (std::nothrow) vp_ctx();
    if (!ctx) return VP_ERR_OUT_OF_MEMORY;
    ctx->device = device;
    cudaDeviceProp prop{};
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, VP_ERR_CUDA, "cannot initialise device");
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->dev_pool.cap_bytes = (size_t)env_int("VP_POOL_MB", 8192) << 20;
    ctx->host_pool.cap_bytes = (size_t)64 << 20;
    *out = ctx;
    return VP_OK;
}

extern "C" int vp_ctx_destroy(vp_ctx *ctx)
{
    if (!ctx) return VP_OK;
    cudaSetDevice(ctx->device);
    pool_release(ctx->dev_pool, false);
    pool_release(ctx->host_pool, true);
    for (cudaStream_t s2 : ctx->side_streams) cudaStreamDestroy(s2);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    if (ctx->join_event) cudaEventDestroy(ctx->join_event);
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
    if (!ctx || !out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (dtype != VP_F64 && dtype != VP_F32) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "dtype must be VP_F64 or VP_F32");
    if (n <= 0 || !basis) return fail(ctx, VP_ERR_EMPTY_MODEL, vp_status_string(VP_ERR_EMPTY_MODEL));
    if (m <= 0 || !x_host) return fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (n > VP_MAX_N || q > VP_MAX_Q || q < 0 || m > (1 << 24))
        return fail(ctx, VP_ERR_MODEL_TOO_LARGE, vp_status_string(VP_ERR_MODEL_TOO_LARGE));
    ModelDesc md{};
    md.m = (int)m; md.n = n; md.q = q; md.p = 0;
    std::vector<int> used(q > 0 ? q : 1, 0);
    for (int j = 0; j < n; ++j) {
        const vp_basis_desc &b = basis[j];
        const int ar = basis_arity(b.kind);
        if (ar < 0) return fail(ctx, VP_ERR_UNSUPPORTED_BASIS, vp_status_string(VP_ERR_UNSUPPORTED_BASIS));
        if (b.n_params != ar)
            return fail(ctx, VP_ERR_INCORRECT_PARAMETER_COUNT, "basis function " + std::to_string(j) + " expects " +
                                                                   std::to_string(ar) + " parameters, but got " +
                                                                   std::to_string(b.n_params));
        md.kind[j] = b.kind;
        md.npar[j] = ar;
        md.scale[j] = b.scale;
        for (int s = 0; s < ar; ++s) {
            const int k = b.param_idx[s];
            if (k < 0 || k >= q) return fail(ctx, VP_ERR_PARAMETER_NOT_IN_MODEL, vp_status_string(VP_ERR_PARAMETER_NOT_IN_MODEL));
            md.pidx[j][s] = k;
            used[k] = 1;
            if (md.p >= VP_MAX_P) return fail(ctx, VP_ERR_MODEL_TOO_LARGE, vp_status_string(VP_ERR_MODEL_TOO_LARGE));
            md.e_basis[md.p] = j;
            md.e_slot[md.p] = s;
            md.e_param[md.p] = k;
            md.p++;
        }
    }
    for (int k = 0; k < q; ++k)
        if (!used[k]) return fail(ctx, VP_ERR_UNUSED_PARAMETER, vp_status_string(VP_ERR_UNUSED_PARAMETER));

    vp_model *mo = new (std::nothrow) vp_model();
    if (!mo) return VP_ERR_OUT_OF_MEMORY;
    mo->ctx = ctx; mo->dtype = dtype; mo->md = md;
    const int v = vec_of(dtype);
    mo->ld = (int)((m + v - 1) / v * v);
    cudaSetDevice(ctx->device);
    cudaError_t e = DEV_ALLOC(ctx, &mo->x_dev, esize(dtype) * (size_t)m);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mo->x_dev, x_host, esize(dtype) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        DEV_FREE(ctx, mo->x_dev);
        delete mo;
        return fail(ctx, VP_ERR_CUDA, std::string("vp_model_create: ") + cudaGetErrorString(e));
    }
    *out = mo;
    return VP_OK;
}

extern "C" int vp_model_create_hosteval(vp_ctx *ctx, int dtype, int64_t m, int32_t q, int32_t n, int32_t p, const int32_t *ind,
                                        vp_host_eval_fn eval, void *user, vp_model **out)
{
    if (!ctx || !out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (dtype != VP_F64 && dtype != VP_F32) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "dtype must be VP_F64 or VP_F32");
    if (!eval) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_model_create_hosteval: eval callback is NULL");
    if (n <= 0) return fail(ctx, VP_ERR_EMPTY_MODEL, vp_status_string(VP_ERR_EMPTY_MODEL));
    if (m <= 0) return fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (n > VP_MAX_N || q > VP_MAX_Q || q < 0 || p < 0 || p > VP_MAX_P || m > (1 << 24) || (p > 0 && !ind))
        return fail(ctx, VP_ERR_MODEL_TOO_LARGE, vp_status_string(VP_ERR_MODEL_TOO_LARGE));
    ModelDesc md{};
    md.m = (int)m; md.n = n; md.q = q; md.p = p;
    for (int j = 0; j < n; ++j) { md.kind[j] = VP_BASIS_HOST; md.npar[j] = 0; }
    std::vector<int> used(q > 0 ? q : 1, 0);
    for (int e = 0; e < p; ++e) {
        const int j = ind[2 * e], k = ind[2 * e + 1];
        if (j < 0 || j >= n) return fail(ctx, VP_ERR_DERIVATIVE_INDEX_OUT_OF_BOUNDS, vp_status_string(VP_ERR_DERIVATIVE_INDEX_OUT_OF_BOUNDS));
        if (k < 0 || k >= q) return fail(ctx, VP_ERR_PARAMETER_NOT_IN_MODEL, vp_status_string(VP_ERR_PARAMETER_NOT_IN_MODEL));
        md.e_basis[e] = j; md.e_param[e] = k; md.e_slot[e] = 0;
        used[k] = 1;
    }
    for (int k = 0; k < q; ++k)
        if (!used[k]) return fail(ctx, VP_ERR_UNUSED_PARAMETER, vp_status_string(VP_ERR_UNUSED_PARAMETER));
    vp_model *mo = new (std::nothrow) vp_model();
    if (!mo) return VP_ERR_OUT_OF_MEMORY;
    mo->ctx = ctx; mo->dtype = dtype; mo->md = md;
    mo->hosteval = true; mo->eval_fn = eval; mo->eval_user = user;
    const int v = vec_of(dtype);
    mo->ld = (int)((m + v - 1) / v * v);
    cudaSetDevice(ctx->device);
    const size_t bytes = sizeof(double) * (size_t)m * (n + p);
    cudaError_t e = DEV_ALLOC(ctx, &mo->pre_dev, bytes);
    if (e == cudaSuccess) e = HOST_ALLOC(ctx, &mo->pre_host, bytes);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &mo->x_dev, esize(dtype) * (size_t)m); // unused by the kernels; keeps the layout uniform
    if (e == cudaSuccess) e = cudaMemsetAsync(mo->x_dev, 0, esize(dtype) * (size_t)m, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        DEV_FREE(ctx, mo->pre_dev); HOST_FREE(ctx, mo->pre_host); DEV_FREE(ctx, mo->x_dev);
        delete mo;
        return fail(ctx, VP_ERR_CUDA, std::string("vp_model_create_hosteval: ") + cudaGetErrorString(e));
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

// ----------------------------------------------------------------------------
// streaming-kernel dispatch table
// ----------------------------------------------------------------------------
// The templated fast-path kernels are instantiated in separate translation units (inst.cu, one per
// model shape and kernel family; see kernel_tables.h); gather their entries once.
static std::vector<StreamKernelEntry> g_stream_kernels;
static std::vector<DmmaKernelEntry> g_dmma_kernels;
static std::vector<PanelHHEntry> g_panel_kernels;
static std::vector<FitKernelEntry> g_fit_kernels;
static std::vector<BatchKernelEntry> g_batch_kernels;
static std::vector<QueueKernelEntry> g_queue_kernels;
static int g_num_stream_kernels = 0, g_num_dmma_kernels = 0;
static void gather_kernel_tables()
{
    static bool done = false;
    if (done) return;
#define VP_GATHER(tag, T, DT, N, P, PART)                                                        \
    {                                                                                            \
        const KernelGroup *g = vp_kernel_group_##tag();                                          \
        g_stream_kernels.insert(g_stream_kernels.end(), g->simt, g->simt + g->nsimt);            \
        g_dmma_kernels.insert(g_dmma_kernels.end(), g->dmma, g->dmma + g->ndmma);                \
        g_panel_kernels.insert(g_panel_kernels.end(), g->panel, g->panel + g->npanel);           \
        g_fit_kernels.insert(g_fit_kernels.end(), g->fit, g->fit + g->nfit);                     \
        g_batch_kernels.insert(g_batch_kernels.end(), g->batch, g->batch + g->nbatch);           \
        g_queue_kernels.insert(g_queue_kernels.end(), g->queue, g->queue + g->nqueue);           \
    }
    VP_KERNEL_GROUPS(VP_GATHER)
#undef VP_GATHER
    g_num_stream_kernels = (int)g_stream_kernels.size();
    g_num_dmma_kernels = (int)g_dmma_kernels.size();
    done = true;
}

// the fused evaluation / persistent fit kernel matching a DMMA plan (same tiling), if instantiated
static int plan_fit_kernel(vp_problem *pr, const DmmaKernelEntry &dk, int lds)
{
    vp_ctx *ctx = pr->ctx;
    pr->plan_fit = -1;
    const char *which = getenv("VP_EVAL_KERNEL"); // "fused" (default) or "split" (K1 + K2)
    if (which && !strcmp(which, "split")) return VP_OK;
    if (pr->model->hosteval) return VP_OK; // the fused kernels evaluate the built-in device basis functions
    for (size_t i = 0; i < g_fit_kernels.size(); ++i) {
        const FitKernelEntry &k = g_fit_kernels[i];
        if (k.n != dk.n || k.p != dk.p || k.ksteps != dk.ksteps || k.nwarps != dk.nwarps || k.exact != dk.exact) continue;
        const size_t stage_bytes = (size_t)DMMA_CT * lds * sizeof(double);
        cudaFuncAttributes fa{};
        VP_CUDA(ctx, cudaFuncGetAttributes(&fa, k.fn));
        if (fa.sharedSizeBytes + 1024 + 2 * stage_bytes > 227 * 1024) return VP_OK;
        const size_t budget = 227 * 1024 - fa.sharedSizeBytes - 1024;
        int nst = (int)(budget / stage_bytes);
        if (nst > STREAM_MAX_STAGES) nst = STREAM_MAX_STAGES;
        const int max_st = env_int("VP_STREAM_STAGES", 0);
        if (max_st >= 2 && nst > max_st) nst = max_st;
        if (nst < 2) return VP_OK;
        const size_t smem = (size_t)nst * stage_bytes;
        VP_CUDA(ctx, ensure_dynamic_smem(k.fn, smem));
        int occ = 0;
        VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, k.nwarps * 32, smem));
        if (occ < 1) return VP_OK;
        const long long ntiles = (pr->S + DMMA_CT - 1) / DMMA_CT;
        long long grid = (long long)ctx->sm_count * occ; // co-resident by construction (cooperative launch checks it)
        if (grid > ntiles) grid = ntiles;
        if (grid > pr->max_grid) grid = pr->max_grid;
        pr->plan_fit = (int)i;
        pr->fit_grid = (int)grid;
        pr->fit_nst = nst;
        pr->fit_smem = smem;
        return VP_OK;
    }
    return VP_OK;
}

// choose the kernel instantiation, stage count and grid for a problem
static int plan_stream(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    const vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const int want_ct = env_int("VP_STREAM_CT", 4);
    const int want_occ = env_int("VP_STREAM_OCC", 2);
    const int force_generic = env_int("VP_STREAM_GENERIC", 0);
    const size_t es = esize(mo->dtype);
    gather_kernel_tables();
    pr->plan_kind = -1;
    pr->plan_dmma = -1;
    const char *which = getenv("VP_STREAM_KERNEL"); // "dmma" (default for fp64), "simt", "generic"
    const bool allow_dmma = !force_generic && mo->dtype == VP_F64 && !(which && (!strcmp(which, "simt") || !strcmp(which, "generic")));
    if (which && !strcmp(which, "generic")) { /* handled below */ }
    if (allow_dmma) {
        int lds = mo->ld;
        while (lds % 16 != 4) lds += 2; // conflict-free fragment loads (see stream_kernel_dmma.cuh)
        int pick = -1;
        for (int i = 0; i < g_num_dmma_kernels; ++i) {
            const DmmaKernelEntry &k = g_dmma_kernels[i];
            if (k.n != md.n || k.p != md.p) continue;
            const int rows = 4 * k.ksteps * k.nwarps;
            if (rows < mo->ld) continue;
            if (k.exact && rows > lds) continue; // the unpredicated variant needs rows <= lds
            const int prow = pick < 0 ? 0 : 4 * g_dmma_kernels[pick].ksteps * g_dmma_kernels[pick].nwarps;
            if (pick < 0 || rows < prow || (rows == prow && k.exact && !g_dmma_kernels[pick].exact)) pick = i;
        }
        if (pick >= 0) {
            const DmmaKernelEntry &k = g_dmma_kernels[pick];
            const size_t stage_bytes = (size_t)DMMA_CT * lds * sizeof(double);
            cudaFuncAttributes fa{};
            VP_CUDA(ctx, cudaFuncGetAttributes(&fa, k.fn));
            const size_t budget = 227 * 1024 - fa.sharedSizeBytes - 1024;
            int nst = (int)(budget / stage_bytes);
            if (nst > STREAM_MAX_STAGES) nst = STREAM_MAX_STAGES;
            const int max_st = env_int("VP_STREAM_STAGES", 0);
            if (max_st >= 2 && nst > max_st) nst = max_st;
            if (nst >= 2) {
                const size_t smem = (size_t)nst * stage_bytes;
                VP_CUDA(ctx, ensure_dynamic_smem(k.fn, smem));
                int occ_real = 0;
                VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, k.fn, k.nwarps * 32, smem));
                if (occ_real >= 1) {
                    const long long ntiles = (pr->S + DMMA_CT - 1) / DMMA_CT;
                    long long grid = (long long)ctx->sm_count * occ_real;
                    if (grid > ntiles) grid = ntiles;
                    if (grid > pr->max_grid) grid = pr->max_grid;
                    pr->plan_dmma = pick;
                    pr->plan_lds = lds;
                    pr->plan_rows = 4 * k.ksteps * k.nwarps;
                    pr->plan_ct = DMMA_CT;
                    pr->plan_grid = (int)grid;
                    pr->plan_nst = nst;
                    pr->plan_smem = smem;
                    return plan_fit_kernel(pr, k, lds);
                }
            }
        }
    }
    int best = -1;
    if (!force_generic && !(which && !strcmp(which, "generic"))) {
        for (int i = 0; i < g_num_stream_kernels; ++i) {
            const StreamKernelEntry &k = g_stream_kernels[i];
            if (k.dtype != mo->dtype || k.n != md.n || k.p != md.p) continue;
            if ((long long)k.threads * k.chunks * vec_of(mo->dtype) < mo->ld) continue;
            // smallest covering (threads*chunks) first, then the requested tile width
            if (best < 0) { best = i; continue; }
            const StreamKernelEntry &b = g_stream_kernels[best];
            const long long ck = (long long)k.threads * k.chunks, cb = (long long)b.threads * b.chunks;
            if (ck < cb || (ck == cb && k.threads < b.threads) ||
                (ck == cb && k.threads == b.threads && std::abs(k.ct - want_ct) < std::abs(b.ct - want_ct)))
                best = i;
        }
    }
    if (best >= 0) {
        const StreamKernelEntry &k = g_stream_kernels[best];
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
            const int max_st = env_int("VP_STREAM_STAGES", 0);
            if (max_st >= 2 && nst > max_st) nst = max_st;
            const size_t smem = (size_t)nst * stage_bytes;
            if (smem + fa.sharedSizeBytes <= ctx->smem_optin) {
                VP_CUDA(ctx, ensure_dynamic_smem(k.fn, smem));
                int occ_real = 0;
                VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, k.fn, k.threads, smem));
                if (occ_real >= 1) {
                    const long long ntiles = (pr->S + k.ct - 1) / k.ct;
                    long long grid = (long long)ctx->sm_count * occ_real;
                    if (grid > ntiles) grid = ntiles;
                    if (grid > pr->max_grid) grid = pr->max_grid;
                    pr->plan_kind = best;
                    pr->plan_rows = k.threads * k.chunks * vec_of(mo->dtype);
                    pr->plan_ct = k.ct;
                    pr->plan_grid = (int)grid;
                    pr->plan_nst = nst;
                    pr->plan_smem = smem;
                    return VP_OK;
                }
            }
        }
    }
    // generic fallback: one warp per column
    pr->plan_kind = -1;
    long long grid = (pr->S + 7) / 8;
    const long long cap = (long long)ctx->sm_count * 8;
    pr->plan_grid = (int)(grid < cap ? grid : cap);
    pr->plan_nst = 0;
    pr->plan_smem = 0;
    pr->plan_rows = mo->ld;
    pr->plan_ct = 1;
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
    if (!env_int("VP_PANEL_GENERIC", 0) && !mo->hosteval) {
        for (const PanelHHEntry &k : g_panel_kernels) {
            if (k.dtype != mo->dtype || k.n != md.n || k.p != md.p || (long long)k.rpt * k.threads < md.m) continue;
            ModelDesc mdc = md;
            const void *xp = mo->x_dev, *wp = pr->w_dev;
            const double *ap = pr->alpha_dev;
            double eps = pr->svd_eps;
            int ldp = pr->ldp;
            void *pq = pr->Pq;
            PanelSmall *sm = pr->small;
            void *args[] = {&mdc, &xp, &wp, &ap, &eps, &ldp, &pq, &sm, &dbg};
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
        return fail(ctx, VP_ERR_MODEL_TOO_LARGE, "m*(n+p) panel does not fit in shared memory");
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        VP_CUDA(ctx, ensure_dynamic_smem((const void *)panel_kernel<T>, smem));
        configured = smem;
    }
    panel_kernel<T><<<1, threads, smem, ctx->stream>>>(md, (const T *)mo->x_dev, (const T *)pr->w_dev, pr->alpha_dev,
                                                       pr->svd_eps, pr->ldp, (T *)pr->Pq, pr->small, dbg,
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
            const DmmaKernelEntry &k = g_dmma_kernels[pr->plan_dmma];
            int lds = pr->plan_lds;
            void *args[] = {(void *)&a, (void *)&lds};
            VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3(pr->plan_grid), dim3(k.nwarps * 32), args, pr->plan_smem, ctx->stream));
        }
    } else if (pr->plan_kind >= 0) {
        const StreamKernelEntry &k = g_stream_kernels[pr->plan_kind];
        void *args[] = {(void *)&a};
        VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3(pr->plan_grid), dim3(k.threads), args, pr->plan_smem, ctx->stream));
    } else {
        stream_kernel_generic<T, 256><<<pr->plan_grid, 256, 0, ctx->stream>>>(a, md.n, md.p, md.m);
    }
    ctx->launches++;
    VP_CUDA(ctx, cudaGetLastError());
    return VP_OK;
}

// One launch of fit_kernel_dmma: a single fused evaluation (fit_mode = false; result in out_dev,
// coefficients into buffer cdst) or a whole fit (fit_mode = true; cooperative launch, state in fit_dev).
static int launch_fused(vp_problem *pr, int cdst, bool fit_mode, cudaStream_t stream = nullptr, int grid = 0)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const FitKernelEntry &k = g_fit_kernels[pr->plan_fit];
    if (!stream) stream = ctx->stream;
    if (grid <= 0 || grid > pr->fit_grid) grid = pr->fit_grid;
    StreamArgs<double> a{};
    a.Y = (const double *)pr->Yw; a.ld = mo->ld; a.S = (int)pr->S;
    a.Pq = nullptr; a.Pe = nullptr; a.ldp = pr->ldp; a.small = nullptr;
    {
        const long long ntiles = (pr->S + DMMA_CT - 1) / DMMA_CT;
        a.tiles_base = (int)(ntiles / grid);
        a.tiles_rem = (int)(ntiles % grid);
    }
    a.C0 = (double *)pr->C[0]; a.C1 = (double *)pr->C[1]; a.cdst = cdst;
    a.fit = fit_mode ? pr->fit_dev : nullptr;
    a.cond = 0ull;
    a.partials = pr->partials; a.red_stride = pr->red_stride; a.ticket = pr->ticket; a.out = pr->out_dev;
    a.nstages = pr->fit_nst; a.q = md.q; a.dbg = pr->dbg;
    for (int e = 0; e < VP_MAX_P; ++e) { a.e_basis[e] = md.e_basis[e]; a.e_param[e] = md.e_param[e]; }
    FitArgs f{};
    f.md = md;
    f.x = (const double *)mo->x_dev; f.w = (const double *)pr->w_dev;
    f.svd_eps = pr->svd_eps;
    f.alpha_dev = pr->alpha_dev;
    f.bc = pr->bcast;
    if (pr->comm) f.comm = pr->comm->args;
    int lds = pr->plan_lds;
    void *args[] = {(void *)&a, (void *)&lds, (void *)&f};
    if (fit_mode) {
        VP_CUDA(ctx, cudaMemsetAsync(pr->bcast, 0, sizeof(FitBcast), stream));
        VP_CUDA(ctx, cudaLaunchCooperativeKernel(k.fn, dim3(grid), dim3(k.nwarps * 32), args, pr->fit_smem, stream));
    } else {
        VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3(grid), dim3(k.nwarps * 32), args, pr->fit_smem, stream));
    }
    ctx->launches++;
    return VP_OK;
}

static int launch_panel(vp_problem *pr)
{
    return pr->model->dtype == VP_F32 ? launch_panel_t<float>(pr) : launch_panel_t<double>(pr);
}
static int launch_stream(vp_problem *pr, int cdst, bool graph_mode = false)
{
    return pr->model->dtype == VP_F32 ? launch_stream_t<float>(pr, cdst, graph_mode) : launch_stream_t<double>(pr, cdst, graph_mode);
}
static int launch_eval(vp_problem *pr, int cdst)
{
    if (pr->plan_fit >= 0) return launch_fused(pr, cdst, false);
    int rc = launch_panel(pr);
    return rc != VP_OK ? rc : launch_stream(pr, cdst);
}

// after a collective evaluation: did the NVLink exchange time out?
static int comm_check(vp_problem *pr)
{
    if (!pr->comm) return VP_OK;
    int err = 0;
    vp_ctx *ctx = pr->ctx;
    VP_CUDA(ctx, cudaMemcpyAsync(&err, pr->comm->error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (err) return fail(ctx, VP_ERR_COMM, "timed out waiting for a peer GPU's contribution (ranks must make the same sequence of calls)");
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
    if (rc != 0) return fail(ctx, VP_ERR_NO_CACHED_CALCULATION, "the model's evaluation callback reported an error");
    VP_CUDA(ctx, cudaMemcpyAsync(mo->pre_dev, mo->pre_host, sizeof(double) * (nphi + nd), cudaMemcpyHostToDevice, ctx->stream));
    return VP_OK;
}

// Full Golub-Pereyra mode: H += sum (R^-1 R^-T)_{j(e) j(f)} U_ef on top of the Kaufman H in pr->out_host
// (one extra pass over Y; see aux_kernels.cuh). The panel [Q | E] must be in HBM at the evaluation point.
template <typename T>
static int add_full_jacobian_term_t(vp_problem *pr)
{
    vp_ctx *ctx = pr->ctx;
    vp_model *mo = pr->model;
    const ModelDesc &md = mo->md;
    const int p = md.p, n = md.n, q = md.q, nu = p * (p + 1) / 2;
    if (p == 0) return VP_OK;
    if (pr->plan_fit >= 0) { // the fused kernel keeps the panel on chip: build it in HBM with K1
        int rc = launch_panel(pr);
        if (rc != VP_OK) return rc;
    }
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
    if (e != cudaSuccess) return fail(ctx, VP_ERR_CUDA, std::string("full Jacobian term: ") + cudaGetErrorString(e));
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
    for (int i = 0; i < q * q; ++i) o->finite = o->finite && std::isfinite(o->H[i]);
    return VP_OK;
}
static int add_full_jacobian_term(vp_problem *pr)
{
    return pr->model->dtype == VP_F32 ? add_full_jacobian_term_t<float>(pr) : add_full_jacobian_term_t<double>(pr);
}

// Evaluate at `alpha` into coefficient buffer `cdst`; result in pr->out_host.
static int evaluate_sync(vp_problem *pr, const double *alpha, int cdst)
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
    int rc = launch_eval(pr, cdst);
    if (rc != VP_OK) return rc;
    VP_CUDA(ctx, cudaMemcpyAsync(pr->out_host, pr->out_dev, sizeof(EvalOut), cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rc = comm_check(pr);
    if (rc == VP_OK && pr->jac_full) rc = add_full_jacobian_term(pr);
    return rc;
}

static void evalout_to_lm(const EvalOut &o, int q, LmEval &ev)
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
    if (model->ctx != ctx) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "model belongs to a different context");
    if (!Y) return fail(ctx, VP_ERR_Y_DATA_MISSING, vp_status_string(VP_ERR_Y_DATA_MISSING));
    const ModelDesc &md = model->md;
    if (S <= 0 || md.m <= 0) return fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (ldY < md.m)
        return fail(ctx, VP_ERR_INVALID_LENGTH_OF_DATA, "Vectors x and y must have same lengths. Given x length = " +
                                                            std::to_string(md.m) + " and y length = " + std::to_string(ldY));
    if (S > INT32_MAX / 2) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "S too large for one problem handle");
    if (md.q > 0 && !alpha0) return fail(ctx, VP_ERR_INVALID_PARAMETER_COUNT, vp_status_string(VP_ERR_INVALID_PARAMETER_COUNT));
    cudaSetDevice(ctx->device);
    vp_problem *pr = new (std::nothrow) vp_problem();
    if (!pr) return VP_ERR_OUT_OF_MEMORY;
    pr->ctx = ctx; pr->model = model; pr->S = S;
    const int dtype = model->dtype;
    const size_t es = esize(dtype);
    const int ld = model->ld, m = md.m;
    // default epsilon = machine epsilon of the scalar (src/problem/builder.rs:282), |eps| otherwise (:248)
    pr->svd_eps = svd_eps < 0 ? (dtype == VP_F32 ? (double)FLT_EPSILON : DBL_EPSILON) : fabs(svd_eps);
    pr->red_stride = 64;
    pr->max_grid = ctx->sm_count * 8;

    auto cleanup = [&](int code, const std::string &msg) {
        vp_problem_destroy(pr);
        return fail(ctx, code, msg);
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
    VP_TRY(DEV_ALLOC(ctx, &pr->fit_dev, sizeof(FitDevice)));
    VP_TRY(cudaMemsetAsync(pr->fit_dev, 0, sizeof(FitDevice), ctx->stream));
    VP_TRY(HOST_ALLOC(ctx, &pr->fit_host, sizeof(FitDevice)));
    VP_TRY(DEV_ALLOC(ctx, &pr->bcast, sizeof(FitBcast)));
    VP_TRY(cudaMemsetAsync(pr->bcast, 0, sizeof(FitBcast), ctx->stream));
    pr->alpha_dev = &pr->fit_dev->st.x_trial[0];
    VP_TRY(DEV_ALLOC(ctx, &pr->phi_scratch, sizeof(double) * (size_t)m * md.n));
    VP_TRY(HOST_ALLOC(ctx, &pr->out_host, sizeof(EvalOut)));
    VP_TRY(HOST_ALLOC(ctx, &pr->alpha_stage, sizeof(double) * VP_MAX_Q + sizeof(FitBcast)));
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
        pr->ldp = (ldp + 3) / 4 * 4;
        cudaError_t e = DEV_ALLOC(ctx, &pr->Pq, es * (size_t)pr->ldp * (md.n + md.p + 1));
        if (e != cudaSuccess) { vp_problem_destroy(pr); return fail(ctx, VP_ERR_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    }
    // first evaluation at the initial guess (src/problem/builder.rs:321)
    for (int k = 0; k < md.q; ++k) pr->alpha[k] = alpha0[k];
    rc = evaluate_sync(pr, pr->alpha, pr->cur);
    if (rc == VP_ERR_NO_CACHED_CALCULATION) { // model error: the problem is built with cache = None (levmar/mod.rs:43-45)
        pr->cached = false;
        *out = pr;
        return VP_OK;
    }
    if (rc != VP_OK) { vp_problem_destroy(pr); return rc; }
    evalout_to_lm(*pr->out_host, md.q, pr->eval);
    pr->cached = pr->eval.finite != 0;
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
    DEV_FREE(ctx, pr->out_dev); DEV_FREE(ctx, pr->fit_dev); DEV_FREE(ctx, pr->phi_scratch); DEV_FREE(ctx, pr->bcast);
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
    const int q = pr->model->md.q;
    for (int k = 0; k < q; ++k) pr->alpha[k] = alpha[k];
    const int dst = pr->cur ^ 1;
    int rc = evaluate_sync(pr, pr->alpha, dst);
    if (rc == VP_ERR_NO_CACHED_CALCULATION) { pr->cached = false; return VP_OK; } // model error -> cache = None
    if (rc != VP_OK) { pr->cached = false; return rc; }
    pr->cur = dst;
    evalout_to_lm(*pr->out_host, q, pr->eval);
    pr->cached = pr->eval.finite != 0;
    return VP_OK;
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
    if (pr->comm && mode == VP_JACOBIAN_FULL)
        return fail(pr->ctx, VP_ERR_COMM, "the full Jacobian is not available for column-sharded problems");
    pr->jac_full = mode == VP_JACOBIAN_FULL ? 1 : 0;
    if (!pr->cached) return VP_OK;
    // refresh the cached evaluation (its J^T J depends on the mode)
    const int dst = pr->cur ^ 1;
    int rc = evaluate_sync(pr, pr->alpha, dst);
    if (rc == VP_ERR_NO_CACHED_CALCULATION) { pr->cached = false; return VP_OK; }
    if (rc != VP_OK) { pr->cached = false; return rc; }
    pr->cur = dst;
    evalout_to_lm(*pr->out_host, pr->model->md.q, pr->eval);
    pr->cached = pr->eval.finite != 0;
    return VP_OK;
}

extern "C" int vp_reduce(vp_problem *pr, vp_reduced *out)
{
    if (!pr || !out) return VP_ERR_INVALID_ARGUMENT;
    const int q = pr->model->md.q;
    memset(out, 0, sizeof(*out));
    out->rnorm2 = pr->eval.rnorm2;
    out->finite = pr->eval.finite;
    out->q = q;
    for (int k = 0; k < q; ++k) out->g[k] = pr->eval.g[k];
    for (int i = 0; i < q * q; ++i) out->H[i] = pr->eval.H[i];
    return pr->cached ? VP_OK : fail(pr->ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
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
    return launch_panel(pr);
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
    if (e != cudaSuccess) return fail(ctx, VP_ERR_CUDA, std::string("materialise: ") + cudaGetErrorString(e));
    return VP_OK;
}

static int materialise(vp_problem *pr, int what, void *out_host, void *out_device = nullptr)
{
    if (!pr || (!out_host && !out_device)) return VP_ERR_INVALID_ARGUMENT;
    if (!pr->cached) return fail(pr->ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
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
    if (!pr->cached) return fail(pr->ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
    vp_ctx *ctx = pr->ctx;
    cudaSetDevice(ctx->device);
    const size_t bytes = esize(pr->model->dtype) * (size_t)pr->model->md.n * pr->S;
    VP_CUDA(ctx, cudaMemcpyAsync(out_host, pr->C[pr->cur], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VP_OK;
}

// ----------------------------------------------------------------------------
// vp_comm: column-sharded global fit over the GPUs of one box
// ----------------------------------------------------------------------------
extern "C" int vp_comm_create(vp_ctx *ctx, int rank, int world, vp_comm **out, void *local_handle_out)
{
    if (!ctx || !out || !local_handle_out) return VP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world)
        return fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_comm_create: need 0 <= rank < world <= 8");
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
        return fail(ctx, VP_ERR_COMM, std::string("vp_comm_create: ") + cudaGetErrorString(e));
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
            return fail(ctx, VP_ERR_COMM, "vp_comm_connect: cannot map the mailbox of rank " + std::to_string(r) + ": " +
                                              cudaGetErrorString(e));
    }
    c->args.world = c->world; c->args.rank = c->rank; c->args.epoch = c->epoch; c->args.error = c->error;
    for (int r = 0; r < c->world; ++r) c->args.box[r] = static_cast<CommMailbox *>(c->peer[r]);
    c->connected = true;
    return VP_OK;
}

extern "C" int vp_comm_destroy(vp_comm *c)
{
    if (!c) return VP_OK;
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local_box);
    cudaFree(c->epoch);
    delete c;
    return VP_OK;
}

extern "C" int vp_problem_set_comm(vp_problem *pr, vp_comm *c)
{
    if (!pr) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = pr->ctx;
    if (c && (c->ctx != ctx || !c->connected)) return fail(ctx, VP_ERR_COMM, "vp_problem_set_comm: communicator not connected / wrong context");
    if (c && pr->plan_fit < 0)
        return fail(ctx, VP_ERR_COMM, "column-sharded fits need the fused fp64 kernel (no instantiation for this model shape)");
    pr->comm = c;
    // the evaluation made at creation covered this rank's columns only: redo it collectively
    const int dst = pr->cur ^ 1;
    int rc = evaluate_sync(pr, pr->alpha, dst);
    if (rc != VP_OK) { pr->cached = false; return rc; }
    pr->cur = dst;
    evalout_to_lm(*pr->out_host, pr->model->md.q, pr->eval);
    pr->cached = pr->eval.finite != 0;
    return VP_OK;
}

// Build (once per problem) the CUDA graph of a whole fit: a conditional WHILE node whose body
// is one evaluation, K1 (panel at the trial parameters) -> K2 (streaming reduce; its last CTA
// advances the lmder state machine on the device and sets the loop condition).
static int ensure_fit_graph(vp_problem *pr)
{
    if (pr->fit_exec) return VP_OK;
    vp_ctx *ctx = pr->ctx;
    if (env_int("VP_DBG_FIT", 0) && !pr->dbg) {
        const size_t n = ((size_t)pr->max_grid + 1) * VP_DBG_SLOTS;
        VP_CUDA(ctx, cudaMalloc(&pr->dbg, n * sizeof(unsigned long long)));
        VP_CUDA(ctx, cudaMemset(pr->dbg, 0, n * sizeof(unsigned long long)));
    }
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
    int rc = launch_panel(pr);
    if (rc == VP_OK) rc = launch_stream(pr, 0, /*graph_mode=*/true);
    ctx->launches = launches_before; // captured, not launched
    cudaGraph_t captured = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &captured);
    if (rc != VP_OK) return rc;
    if (e != cudaSuccess) return fail(ctx, VP_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    VP_CUDA(ctx, cudaGraphInstantiate(&pr->fit_exec, pr->fit_graph, 0));
    return VP_OK;
}

// ----------------------------------------------------------------------------
// LevMarSolver::fit  (src/solvers/levmar/mod.rs:238-254)
// ----------------------------------------------------------------------------
static int queue_kernel_for(const vp_problem *pr, int *lds_out);
static int fit_queue_group(vp_ctx *ctx, const std::vector<vp_problem *> &prs, std::vector<LmState> &states,
                           const std::vector<LmConfig> &cfgs);

static void lm_config_from_options(const vp_problem *pr, const vp_lm_options *opt, LmConfig &cfg)
{
    const int q = pr->model->md.q;
    const double eps = pr->model->dtype == VP_F32 ? (double)FLT_EPSILON : DBL_EPSILON;
    cfg.epsmch = eps;
    cfg.ftol = (opt && opt->ftol > 0) ? opt->ftol : 30.0 * eps;
    cfg.xtol = (opt && opt->xtol > 0) ? opt->xtol : 30.0 * eps;
    cfg.gtol = (opt && opt->gtol > 0) ? opt->gtol : 30.0 * eps;
    cfg.stepbound = (opt && opt->stepbound > 0) ? opt->stepbound : 100.0;
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

// Persistent-kernel fit, asynchronous halves. begin: upload the LM state (already advanced past the
// cached evaluation at the starting point) and enqueue the ONE cooperative launch plus the
// read-back on `stream`; end (after the stream has been synchronised): adopt the final state.
static int fit_persistent_begin(vp_problem *pr, const LmState &st, const LmConfig &cfg, cudaStream_t stream, int grid)
{
    vp_ctx *ctx = pr->ctx;
    if (env_int("VP_DBG_FIT", 0) && !pr->dbg) { // in-kernel timeline of the last evaluation (vp_debug_timeline reads it)
        const size_t nd = ((size_t)pr->max_grid + 1) * VP_DBG_SLOTS;
        VP_CUDA(ctx, cudaMalloc(&pr->dbg, nd * sizeof(unsigned long long)));
        VP_CUDA(ctx, cudaMemset(pr->dbg, 0, nd * sizeof(unsigned long long)));
    }
    FitDevice *fh = pr->fit_host;
    memset(fh, 0, sizeof(FitDevice));
    fh->st = st; fh->cfg = cfg; fh->accepted = pr->eval; fh->cur = pr->cur; fh->evals = 0;
    VP_CUDA(ctx, cudaMemcpyAsync(pr->fit_dev, fh, sizeof(FitDevice), cudaMemcpyHostToDevice, stream));
    int rc = launch_fused(pr, 0, /*fit_mode=*/true, stream, grid);
    if (rc != VP_OK) return rc;
    VP_CUDA(ctx, cudaMemcpyAsync(fh, pr->fit_dev, sizeof(FitDevice), cudaMemcpyDeviceToHost, stream));
    FitBcast *bh = reinterpret_cast<FitBcast *>(pr->alpha_stage + VP_MAX_Q); // pinned staging (allocated with alpha_stage)
    VP_CUDA(ctx, cudaMemcpyAsync(bh, pr->bcast, sizeof(FitBcast), cudaMemcpyDeviceToHost, stream));
    return VP_OK;
}

static int fit_persistent_end(vp_problem *pr, LmState &st)
{
    vp_ctx *ctx = pr->ctx;
    const int q = pr->model->md.q;
    FitDevice *fh = pr->fit_host;
    const FitBcast *bh = reinterpret_cast<const FitBcast *>(pr->alpha_stage + VP_MAX_Q);
    if (bh->error) return fail(ctx, VP_ERR_CUDA, "vp_fit: grid-wide wait timed out inside the persistent fit kernel");
    int rc = comm_check(pr);
    if (rc != VP_OK) return rc;
    st = fh->st;
    pr->cur = fh->cur;
    pr->eval = fh->accepted;
    for (int k = 0; k < q; ++k) pr->alpha[k] = st.x[k];
    if (env_int("VP_TRACE", 0))
        for (int i = 0; i < fh->evals && i < 48; ++i)
            fprintf(stderr, "[vp_fit persistent] eval %d fnorm_trial=%.6e par=%.3e delta=%.3e acc=%d\n", i + 2, fh->trace[4 * i],
                    fh->trace[4 * i + 1], fh->trace[4 * i + 2], (int)fh->trace[4 * i + 3]);
    return VP_OK;
}

extern "C" int vp_fit(vp_problem *pr, const vp_lm_options *opt, vp_fit_report *rep)
{
    if (!pr || !rep) return VP_ERR_INVALID_ARGUMENT;
    const int q = pr->model->md.q;
    LmConfig cfg;
    lm_config_from_options(pr, opt, cfg);
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
    bool more = lm_advance(st, cfg, pr->eval);
    const char *mode = getenv("VP_FIT_MODE"); // "persistent" (default when the fused kernel exists), "graph" or "host"
    if (more && pr->plan_fit >= 0 && !pr->jac_full && !(mode && (!strcmp(mode, "host") || !strcmp(mode, "graph")))) {
        // ---- persistent grid: the whole fit is ONE cooperative kernel launch ------------------
        vp_ctx *ctx = pr->ctx;
        cudaSetDevice(ctx->device);
        int rc = fit_persistent_begin(pr, st, cfg, ctx->stream, 0);
        if (rc != VP_OK) return rc;
        VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        rc = fit_persistent_end(pr, st);
        if (rc != VP_OK) return rc;
        more = false;
    }
    if (more && pr->comm && (pr->jac_full || pr->plan_fit < 0))
        return fail(pr->ctx, VP_ERR_COMM, "column-sharded fits run on the persistent fit kernel only (Kaufman Jacobian)");
    if (more && pr->comm) return fail(pr->ctx, VP_ERR_COMM, "column-sharded fits run on the persistent fit kernel only");
    const char *single = getenv("VP_FIT_SINGLE"); // "queue": run single f64 fits on the work-queue kernel too
    if (more && (pr->plan_fit < 0 || (single && !strcmp(single, "queue"))) &&
        !(mode && (!strcmp(mode, "host") || !strcmp(mode, "graph"))) && queue_kernel_for(pr, nullptr) >= 0) {
        // ---- fp32 problems: the work-queue kernel with a single fit (panel once per evaluation, fp64 math) ---
        cudaSetDevice(pr->ctx->device);
        std::vector<vp_problem *> prs{pr};
        std::vector<LmState> sts{st};
        std::vector<LmConfig> cfs{cfg};
        int rc = fit_queue_group(pr->ctx, prs, sts, cfs);
        if (rc == VP_OK) {
            st = sts[0];
            more = false;
        } else if (rc != VP_ERR_UNSUPPORTED_BASIS) {
            return rc;
        }
    }
    if (more && !(mode && !strcmp(mode, "host")) && !pr->model->hosteval && !pr->jac_full) {
        // ---- device-driven loop: one CUDA graph launch per fit ------------------------
        vp_ctx *ctx = pr->ctx;
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
        if (env_int("VP_TRACE", 0))
            for (int i = 0; i < fh->evals && i < 48; ++i)
                fprintf(stderr, "[vp_fit graph] eval %d fnorm_trial=%.6e par=%.3e delta=%.3e acc=%d\n", i + 2, fh->trace[4 * i],
                        fh->trace[4 * i + 1], fh->trace[4 * i + 2], (int)fh->trace[4 * i + 3]);
        more = false;
    }
    while (more) {
        // ---- host-driven loop (VP_FIT_MODE=host): one synchronisation per evaluation ---
        const int dst = pr->cur ^ 1;
        int rc = evaluate_sync(pr, st.x_trial, dst);
        if (rc == VP_ERR_NO_CACHED_CALCULATION) { // residuals() is None -> the LM loop stops with a User termination
            st.termination = TERM_USER;
            pr->cached = false;
            break;
        }
        if (rc != VP_OK) return rc;
        LmEval ev;
        evalout_to_lm(*pr->out_host, q, ev);
        more = lm_advance(st, cfg, ev);
        if (env_int("VP_TRACE", 0))
            fprintf(stderr, "[vp_fit] nfev=%d fnorm_trial=%.6e fnorm=%.6e par=%.3e delta=%.3e acc=%d x_trial=(%.15g, %.15g) g=(%.3e,%.3e)\n",
                    st.nfev, sqrt(ev.rnorm2), st.fnorm, st.par, st.delta, st.last_accepted, st.x_trial[0], st.x_trial[1], ev.g[0], ev.g[1]);
        if (st.last_accepted) {
            pr->cur = dst;
            pr->eval = ev;
            for (int k = 0; k < q; ++k) pr->alpha[k] = st.x[k];
        }
    }
    fill_report(st, rep);
    return VP_OK;
}

// The work-queue kernel instantiation a problem can use (index into g_queue_kernels, -1 = none) and the
// padded tile-column stride (elements) that goes with it.
static int queue_kernel_for(const vp_problem *pr, int *lds_out)
{
    const vp_model *mo = pr->model;
    if (mo->hosteval || pr->comm || !pr->cached || pr->jac_full) return -1;
    const ModelDesc &md = mo->md;
    int lds = mo->ld;
    if (mo->dtype == VP_F64) { while (lds % 16 != 4) lds += 2; }
    else { while (lds % 32 != 8) lds += 4; }
    int pick = -1;
    for (size_t i = 0; i < g_queue_kernels.size(); ++i) {
        const QueueKernelEntry &k = g_queue_kernels[i];
        if (k.dtype != mo->dtype || k.n != md.n || k.p != md.p) continue;
        const int rows = 4 * k.ksteps * k.nwarps;
        if (rows < mo->ld) continue;
        if (k.exact && rows > lds) continue;
        const int prow = pick < 0 ? 0 : 4 * g_queue_kernels[pick].ksteps * g_queue_kernels[pick].nwarps;
        if (pick < 0 || rows < prow || (rows == prow && k.exact && !g_queue_kernels[pick].exact)) pick = (int)i;
    }
    if (lds_out) *lds_out = lds;
    return pick;
}

// One launch of fit_queue_kernel for a group of problems that share the kernel instantiation and the
// padded row count. states[i] has been advanced past the cached evaluation at the starting point.
// Returns VP_OK after the fits completed and their final states were adopted.
static int fit_queue_group(vp_ctx *ctx, const std::vector<vp_problem *> &prs, std::vector<LmState> &states,
                           const std::vector<LmConfig> &cfgs)
{
    const int K = (int)prs.size();
    vp_problem *p0 = prs[0];
    int lds = 0;
    const int qidx = queue_kernel_for(p0, &lds);
    if (qidx < 0) return VP_ERR_UNSUPPORTED_BASIS; // caller falls back to the per-fit kernels
    const QueueKernelEntry *qk = &g_queue_kernels[(size_t)qidx];
    const size_t es = esize(p0->model->dtype);
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
            qp.smem = (size_t)qp.nst * stage_bytes;
            VP_CUDA(ctx, ensure_dynamic_smem(qk->fn, qp.smem));
            VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&qp.occ, qk->fn, qk->nwarps * 32, qp.smem));
        }
        plans.push_back(qp);
        plan = &plans.back();
    }
    if (plan->occ < 1) return VP_ERR_UNSUPPORTED_BASIS;
    const int nst = plan->nst, occ = plan->occ;
    const size_t smem = plan->smem;

    const int min_chunk = env_int("VP_QUEUE_MIN_CHUNK", 4); // tiles; the kernel picks the chunk size per evaluation
    std::vector<QueueFit> hq((size_t)K);
    long long total_chunks = 0;
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
        f.out = pr->out_dev; f.fit = pr->fit_dev; f.svd_eps = pr->svd_eps;
        f.ld = pr->model->ld; f.S = (int)pr->S; f.red_stride = pr->red_stride;
        f.ntiles = (int)((pr->S + DMMA_CT - 1) / DMMA_CT);
        f.min_chunk_tiles = min_chunk < 1 ? 1 : min_chunk;
        f.max_chunks = pr->max_grid; // one partial row per chunk
        f.adaptive = env_int("VP_QUEUE_ADAPTIVE", 1);
        f.items_per_cta = env_int("VP_QUEUE_ITEMS_PER_CTA", 2);
        if (f.items_per_cta < 1) f.items_per_cta = 1;
        f.chunk_tiles = f.ntiles;
        f.nchunks = 1;
        f.cdst = pr->cur ^ 1;
        total_chunks += (f.ntiles + f.min_chunk_tiles - 1) / f.min_chunk_tiles < f.max_chunks ? (f.ntiles + f.min_chunk_tiles - 1) / f.min_chunk_tiles : f.max_chunks;
    }
    // One pinned host block and one device block: [FitDevice x K | QueueCtl | QueueFit x K]. One copy
    // in; one copy out (states + control word).
    const size_t off_ctl = sizeof(FitDevice) * (size_t)K;
    const size_t off_q = (off_ctl + sizeof(QueueCtl) + 255) / 256 * 256;
    const size_t total_bytes = off_q + sizeof(QueueFit) * (size_t)K;
    const unsigned int cap = (unsigned int)(total_chunks + 2048);
    unsigned char *hb = nullptr, *db = nullptr;
    QueueItem *ditems = nullptr;
    cudaError_t e = HOST_ALLOC(ctx, &hb, total_bytes);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &db, total_bytes);
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &ditems, sizeof(QueueItem) * (size_t)cap);
    if (e != cudaSuccess) {
        HOST_FREE(ctx, hb); DEV_FREE(ctx, db); DEV_FREE(ctx, ditems);
        return fail(ctx, VP_ERR_OUT_OF_MEMORY, std::string("vp_fit_many (queue): ") + cudaGetErrorString(e));
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
    // finish queue + dedicated finisher CTAs: OFF by default. Measured (profiles/r02v_queue_phase_breakdown.txt):
    // 8 finishers cannot keep up with 20-60 fits because one evaluation costs them ~54 us, of which the
    // single-thread LM step is ~31 us -- the work is slow by itself, not because of a cold instruction cache.
    QueueItem *dfitems = nullptr;
    const unsigned int fcap = (unsigned int)K + 64u;
    int nfin = env_int("VP_QUEUE_FINISHERS", 0);
    if (nfin > ctx->sm_count / 4) nfin = ctx->sm_count / 4;
    if (nfin > 0) {
        if (DEV_ALLOC(ctx, &dfitems, sizeof(QueueItem) * (size_t)fcap) == cudaSuccess)
            cudaMemsetAsync(dfitems, 0, sizeof(QueueItem) * (size_t)fcap, ctx->stream);
        else
            nfin = 0;
    }
    hctl->fhead = 0; hctl->ftail = 0; hctl->fitems = dfitems; hctl->fcap = fcap; hctl->nfinishers = nfin;
    unsigned long long *ddbg = nullptr;
    const long long dbg_grid = (long long)ctx->sm_count * occ;
    if (env_int("VP_QUEUE_DBG", 0)) { // per-CTA phase accumulators (diagnostics; printed to stderr after the launch)
        if (DEV_ALLOC(ctx, &ddbg, sizeof(unsigned long long) * 8 * (size_t)dbg_grid) == cudaSuccess)
            cudaMemsetAsync(ddbg, 0, sizeof(unsigned long long) * 8 * (size_t)dbg_grid, ctx->stream);
        else
            ddbg = nullptr;
    }
    hctl->dbg = ddbg;
    e = cudaMemcpyAsync(db, hb, total_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ditems, 0, sizeof(QueueItem) * (size_t)cap, ctx->stream);
    int rc = VP_OK;
    if (e == cudaSuccess) {
        long long grid = (long long)ctx->sm_count * occ;
        int nf = K, lds_arg = lds, nst_arg = nst;
        void *args[] = {(void *)&dctl, (void *)&dq, (void *)&nf, (void *)&lds_arg, (void *)&nst_arg};
        e = cudaLaunchKernel(qk->fn, dim3((unsigned)grid), dim3(qk->nwarps * 32), args, smem, ctx->stream);
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hb, db, off_ctl + sizeof(QueueCtl), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    const int qerr = hctl->error;
    if (ddbg && e == cudaSuccess) {
        std::vector<unsigned long long> hd(8 * (size_t)dbg_grid);
        if (cudaMemcpy(hd.data(), ddbg, sizeof(unsigned long long) * hd.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
            double acc[8] = {0}, fin[8] = {0};
            for (long long b = 0; b < dbg_grid; ++b)
                for (int i = 0; i < 8; ++i) (b < nfin ? fin[i] : acc[i]) += (double)hd[(size_t)b * 8 + i];
            if (nfin > 0 && fin[7] > 0)
                fprintf(stderr, "[vp queue dbg] %d finisher CTAs, %.0f evaluations started | per evaluation: finalize %.2f us, state load %.2f us, "
                                "LM step %.2f us, state store %.2f us, x/w + basis %.2f us, basis + factor + panel store %.2f us, push %.2f us\n",
                        nfin, fin[7], 1e-3 * fin[0] / fin[7], 1e-3 * fin[1] / fin[7], 1e-3 * fin[2] / fin[7], 1e-3 * fin[3] / fin[7],
                        1e-3 * fin[4] / fin[7], 1e-3 * fin[5] / fin[7], 1e-3 * fin[6] / fin[7]);
            const double it = acc[0] > 0 ? acc[0] : 1, nf = acc[6] > 0 ? acc[6] : 1;
            fprintf(stderr, "[vp queue dbg] fits %d items %.0f (%.1f per CTA) | per item: claim %.2f us, fragments %.2f us, stream %.2f us, "
                            "publish %.2f us | finisher (finalize + LM + panel + push) %.2f us x %.0f | CTA lifetime %.1f us, busy %.1f %%\n",
                    K, acc[0], acc[0] / dbg_grid, 1e-3 * acc[1] / it, 1e-3 * acc[2] / it, 1e-3 * acc[3] / it, 1e-3 * acc[4] / it,
                    1e-3 * acc[5] / nf, acc[6], 1e-3 * acc[7] / dbg_grid,
                    100.0 * (acc[2] + acc[3] + acc[4] + acc[5]) / (acc[7] > 0 ? acc[7] : 1));
        }
    }
    DEV_FREE(ctx, ddbg);
    DEV_FREE(ctx, dfitems);
    DEV_FREE(ctx, db); DEV_FREE(ctx, ditems);
    if (e != cudaSuccess) { HOST_FREE(ctx, hb); return fail(ctx, VP_ERR_CUDA, std::string("vp_fit_many (queue): ") + cudaGetErrorString(e)); }
    if (qerr) { HOST_FREE(ctx, hb); return fail(ctx, VP_ERR_CUDA, "vp_fit_many: a wait inside the work-queue kernel timed out"); }
    for (int i = 0; i < K; ++i) {
        vp_problem *pr = prs[(size_t)i];
        FitDevice *fh = &hstate[i];
        states[(size_t)i] = fh->st;
        pr->cur = fh->cur;
        pr->eval = fh->accepted;
        for (int k = 0; k < pr->model->md.q; ++k) pr->alpha[k] = fh->st.x[k];
    }
    HOST_FREE(ctx, hb);
    return rc;
}

// Many independent fits at once (throughput mode). A single fit on the whole GPU is latency
// bound: per evaluation the panel, the grid-wide reduction and the serial LM step cost more than
// streaming 33.5 MB does. Independent problems are therefore fitted TOGETHER. Default
// (VP_FIT_MANY=queue): like-shaped problems share ONE persistent grid through a device-side work
// queue (fit_queue_kernel.cuh): every CTA streams chunks of whichever fit has work, the panel of
// an evaluation is computed once, and the serial phases of one fit hide behind the streaming of
// the others. VP_FIT_MANY=streams: each fit is its own persistent kernel (fit_kernel_dmma) on
// #SMs / n SMs on its own stream. Problems that can use neither are fitted one after the other.
extern "C" int vp_fit_many(vp_problem **problems, int64_t n, const vp_lm_options *opt, vp_fit_report *reports,
                           int32_t max_concurrent)
{
    if (n < 0 || (n > 0 && (!problems || !reports))) return VP_ERR_INVALID_ARGUMENT;
    if (n == 0) return VP_OK;
    for (int64_t i = 0; i < n; ++i)
        if (!problems[i] || problems[i]->ctx != problems[0]->ctx) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = problems[0]->ctx;
    cudaSetDevice(ctx->device);
    const char *mode = getenv("VP_FIT_MODE");
    const bool persistent_ok = !(mode && (!strcmp(mode, "host") || !strcmp(mode, "graph")));
    // ---- default: ONE persistent grid with a device-side work queue per group of like-shaped problems
    const char *many = getenv("VP_FIT_MANY"); // "queue" (default) or "streams"
    if (persistent_ok && n > 1 && !(many && !strcmp(many, "streams"))) {
        std::vector<char> done((size_t)n, 0);
        std::vector<LmState> all_states((size_t)n);
        std::vector<LmConfig> all_cfgs((size_t)n);
        bool any_left = false;
        for (int64_t i = 0; i < n; ++i) {
            vp_problem *pr = problems[i];
            memset(&reports[i], 0, sizeof(vp_fit_report));
            lm_config_from_options(pr, opt, all_cfgs[(size_t)i]);
            lm_init(all_states[(size_t)i], pr->model->md.q, pr->alpha);
            if (queue_kernel_for(pr, nullptr) < 0) { any_left = true; continue; }
            if (!lm_advance(all_states[(size_t)i], all_cfgs[(size_t)i], pr->eval)) { // terminated at the starting point
                fill_report(all_states[(size_t)i], &reports[i]);
                done[(size_t)i] = 1;
                continue;
            }
            done[(size_t)i] = 2; // to be fitted by a queue launch
        }
        int first_error = VP_OK;
        for (int64_t i = 0; i < n; ++i) {
            if (done[(size_t)i] != 2) continue;
            // group: same kernel instantiation, same padded rows
            std::vector<int64_t> idx;
            for (int64_t j = i; j < n; ++j)
                if (done[(size_t)j] == 2 && queue_kernel_for(problems[j], nullptr) == queue_kernel_for(problems[i], nullptr) &&
                    problems[j]->model->ld == problems[i]->model->ld)
                    idx.push_back(j);
            std::vector<vp_problem *> prs;
            std::vector<LmState> sts;
            std::vector<LmConfig> cfs;
            for (int64_t j : idx) { prs.push_back(problems[j]); sts.push_back(all_states[(size_t)j]); cfs.push_back(all_cfgs[(size_t)j]); }
            // a group of one f64 fit is better served by the whole-GPU persistent kernel (vp_fit)
            const bool use_queue = idx.size() > 1 || problems[i]->plan_fit < 0;
            int rc = use_queue ? fit_queue_group(ctx, prs, sts, cfs) : VP_ERR_UNSUPPORTED_BASIS;
            for (size_t t = 0; t < idx.size(); ++t) {
                const int64_t j = idx[t];
                if (rc == VP_OK) {
                    fill_report(sts[t], &reports[j]);
                    done[(size_t)j] = 1;
                } else if (rc == VP_ERR_UNSUPPORTED_BASIS) {
                    done[(size_t)j] = 0; // no queue kernel for this shape (or a group of one): per-fit kernel below
                    any_left = true;
                } else {
                    done[(size_t)j] = 3;
                    if (first_error == VP_OK) first_error = rc;
                }
            }
        }
        if (any_left)
            for (int64_t i = 0; i < n; ++i)
                if (done[(size_t)i] == 0) {
                    int rc = vp_fit(problems[i], opt, &reports[i]);
                    if (rc != VP_OK && first_error == VP_OK) first_error = rc;
                }
        return first_error;
    }
    // ---- VP_FIT_MANY=streams: one persistent kernel per fit on a slice of the SMs, on separate streams
    int64_t width = max_concurrent > 0 ? max_concurrent : ctx->sm_count;
    if (width > n) width = n;
    if (width > ctx->sm_count) width = ctx->sm_count;
    // SMs per fit: the first (sm_count % width) fits get one more, so that all SMs are used
    const int grid_base = (int)(ctx->sm_count / width) > 0 ? (int)(ctx->sm_count / width) : 1;
    const int grid_rem = (int)(ctx->sm_count / width) > 0 ? (int)(ctx->sm_count % width) : 0;
    while ((int64_t)ctx->side_streams.size() < width) {
        cudaStream_t s2 = nullptr;
        VP_CUDA(ctx, cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        ctx->side_streams.push_back(s2);
    }
    if (!ctx->fork_event) VP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    if (!ctx->join_event) VP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->join_event, cudaEventDisableTiming));
    // the side streams start after everything already enqueued on the context's stream ...
    VP_CUDA(ctx, cudaEventRecord(ctx->fork_event, ctx->stream));
    for (int64_t w = 0; w < width; ++w) VP_CUDA(ctx, cudaStreamWaitEvent(ctx->side_streams[w], ctx->fork_event, 0));
    std::vector<LmState> states((size_t)n);
    std::vector<char> launched((size_t)n, 0);
    int first_error = VP_OK;
    for (int64_t i = 0; i < n; ++i) {
        vp_problem *pr = problems[i];
        memset(&reports[i], 0, sizeof(vp_fit_report));
        LmConfig cfg;
        lm_config_from_options(pr, opt, cfg);
        LmState &st = states[(size_t)i];
        lm_init(st, pr->model->md.q, pr->alpha);
        if (!pr->cached || pr->plan_fit < 0 || pr->comm || !persistent_ok) continue; // handled by vp_fit below
        if (!lm_advance(st, cfg, pr->eval)) { launched[(size_t)i] = 2; continue; }   // terminated at the starting point
        const int grid = grid_base + ((i % width) < grid_rem ? 1 : 0);
        int rc = fit_persistent_begin(pr, st, cfg, ctx->side_streams[(size_t)(i % width)], grid);
        if (rc != VP_OK) { if (first_error == VP_OK) first_error = rc; continue; }
        launched[(size_t)i] = 1;
    }
    // ... and the context's stream continues after all of them (events recorded by the caller on
    // the context's stream therefore bracket the whole batch)
    for (int64_t w = 0; w < width; ++w) {
        VP_CUDA(ctx, cudaEventRecord(ctx->join_event, ctx->side_streams[w]));
        VP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->join_event, 0));
    }
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i < n; ++i) {
        vp_problem *pr = problems[i];
        if (launched[(size_t)i] == 1) {
            int rc = fit_persistent_end(pr, states[(size_t)i]);
            if (rc != VP_OK) { if (first_error == VP_OK) first_error = rc; continue; }
            fill_report(states[(size_t)i], &reports[i]);
        } else if (launched[(size_t)i] == 2) {
            fill_report(states[(size_t)i], &reports[i]);
        } else if (first_error == VP_OK || pr->plan_fit < 0) {
            int rc = vp_fit(pr, opt, &reports[i]);
            if (rc != VP_OK && first_error == VP_OK) first_error = rc;
        }
    }
    return first_error;
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
        return fail(ctx, e == cudaErrorMemoryAllocation ? VP_ERR_OUT_OF_MEMORY : VP_ERR_CUDA, std::string("vp_statistics: ") + cudaGetErrorString(e));
    if (host_flag) return fail(ctx, VP_ERR_MATRIX_INVERSION, vp_status_string(VP_ERR_MATRIX_INVERSION));
    return VP_OK;
}

extern "C" int vp_statistics(vp_problem *pr, double *cov_out, double *reduced_chi2_out, double *conf_sigma_out)
{
    if (!pr || !cov_out) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = pr->ctx;
    if (pr->comm) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_statistics: call it on each rank's own columns after detaching the communicator");
    if (!pr->cached) return fail(ctx, VP_ERR_NO_CACHED_CALCULATION, vp_status_string(VP_ERR_NO_CACHED_CALCULATION));
    const ModelDesc &md = pr->model->md;
    if (md.m <= md.n + md.q) return fail(ctx, VP_ERR_UNDERDETERMINED, vp_status_string(VP_ERR_UNDERDETERMINED)); // :377-379
    cudaSetDevice(ctx->device);
    int rc = ensure_panel_current(pr); // Q at the accepted parameters, in HBM
    if (rc != VP_OK) return rc;
    return pr->model->dtype == VP_F32 ? statistics_t<float>(pr, cov_out, reduced_chi2_out, conf_sigma_out)
                                      : statistics_t<double>(pr, cov_out, reduced_chi2_out, conf_sigma_out);
}

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
    if (model->ctx != ctx) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "model belongs to a different context");
    if (model->dtype != VP_F64) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_batch: fp64 models only");
    if (model->hosteval) return fail(ctx, VP_ERR_UNSUPPORTED_BASIS, "vp_batch: needs the built-in device basis functions");
    if (!Y) return fail(ctx, VP_ERR_Y_DATA_MISSING, vp_status_string(VP_ERR_Y_DATA_MISSING));
    const ModelDesc &md = model->md;
    if (P <= 0) return fail(ctx, VP_ERR_ZERO_LENGTH_VECTOR, vp_status_string(VP_ERR_ZERO_LENGTH_VECTOR));
    if (ldY < md.m)
        return fail(ctx, VP_ERR_INVALID_LENGTH_OF_DATA, "Vectors x and y must have same lengths. Given x length = " +
                                                            std::to_string(md.m) + " and y length = " + std::to_string(ldY));
    if (md.q > 0 && !alpha0) return fail(ctx, VP_ERR_INVALID_PARAMETER_COUNT, vp_status_string(VP_ERR_INVALID_PARAMETER_COUNT));
    gather_kernel_tables();
    cudaSetDevice(ctx->device);
    int pick = -1;
    for (size_t i = 0; i < g_batch_kernels.size(); ++i) {
        const BatchKernelEntry &k = g_batch_kernels[i];
        if (k.n != md.n || k.p != md.p || (long long)k.rpt * k.threads < md.m) continue;
        // smallest row tiling that covers m; among those the requested number of problem slots (default 8)
        const int want = env_int("VP_BATCH_SLOTS", 8);
        if (pick < 0 || k.rpt < g_batch_kernels[pick].rpt ||
            (k.rpt == g_batch_kernels[pick].rpt && std::abs(k.slots - want) < std::abs(g_batch_kernels[pick].slots - want)))
            pick = (int)i;
    }
    if (pick < 0)
        return fail(ctx, VP_ERR_MODEL_TOO_LARGE, "vp_batch: no independent-batch kernel instantiated for this model shape / m > 4096");
    const int mpad = (md.m + 1) / 2 * 2;
    const size_t smem = sizeof(double) * (size_t)(md.n + md.p) * mpad;
    cudaFuncAttributes fa{};
    VP_CUDA(ctx, cudaFuncGetAttributes(&fa, g_batch_kernels[pick].fn));
    if (smem + fa.sharedSizeBytes > ctx->smem_optin)
        return fail(ctx, VP_ERR_MODEL_TOO_LARGE, "vp_batch: m*(n+p) working matrix does not fit in shared memory");
    VP_CUDA(ctx, ensure_dynamic_smem(g_batch_kernels[pick].fn, smem));
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
    if (e == cudaSuccess) e = DEV_ALLOC(ctx, &b->next, sizeof(unsigned long long));
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
        return fail(ctx, e == cudaErrorMemoryAllocation ? VP_ERR_OUT_OF_MEMORY : VP_ERR_CUDA, std::string("vp_batch_create: ") + cudaGetErrorString(e));
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
    const BatchKernelEntry &k = g_batch_kernels[b->kernel];
    BatchArgs a{};
    a.md = md;
    a.x = (const double *)b->model->x_dev; a.w = b->w_dev; a.Y = b->Y; a.ld = b->ld; a.P = b->P;
    a.svd_eps = b->svd_eps;
    {
        const double eps = DBL_EPSILON;
        a.cfg.epsmch = eps;
        a.cfg.ftol = (opt && opt->ftol > 0) ? opt->ftol : 30.0 * eps;
        a.cfg.xtol = (opt && opt->xtol > 0) ? opt->xtol : 30.0 * eps;
        a.cfg.gtol = (opt && opt->gtol > 0) ? opt->gtol : 30.0 * eps;
        a.cfg.stepbound = (opt && opt->stepbound > 0) ? opt->stepbound : 100.0;
        const int patience = (opt && opt->patience > 0) ? opt->patience : 100;
        a.cfg.maxfev = patience * (md.q + 1);
        a.cfg.scale_diag = (opt && opt->scale_diag >= 0) ? (opt->scale_diag != 0) : 1;
    }
    // start from the current parameters: alpha -> alpha0 (device copy), results go to alpha
    if (md.q > 0)
        VP_CUDA(ctx, cudaMemcpyAsync(b->alpha0, b->alpha, sizeof(double) * (size_t)md.q * b->P, cudaMemcpyDeviceToDevice, ctx->stream));
    a.alpha0 = b->alpha0; a.alpha_out = b->alpha; a.C_out = b->C; a.obj_out = b->obj; a.term_out = b->term; a.nfev_out = b->nfev;
    a.next = b->next; a.mpad = b->mpad;
    VP_CUDA(ctx, cudaMemsetAsync(b->next, 0, sizeof(unsigned long long), ctx->stream));
    int occ = 0;
    VP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, k.threads, b->smem));
    if (occ < 1) return fail(ctx, VP_ERR_MODEL_TOO_LARGE, "vp_batch: kernel does not fit on an SM");
    long long grid = (long long)ctx->sm_count * occ;
    if (grid > b->P) grid = b->P;
    void *args[] = {(void *)&a};
    VP_CUDA(ctx, cudaLaunchKernel(k.fn, dim3((unsigned)grid), dim3(k.threads), args, b->smem, ctx->stream));
    ctx->launches++;
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
    if (!reports) {
        VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return VP_OK;
    }
    std::vector<double> obj((size_t)b->P);
    std::vector<int> term((size_t)b->P), nfev((size_t)b->P);
    VP_CUDA(ctx, cudaMemcpyAsync(obj.data(), b->obj, sizeof(double) * (size_t)b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaMemcpyAsync(term.data(), b->term, sizeof(int) * (size_t)b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaMemcpyAsync(nfev.data(), b->nfev, sizeof(int) * (size_t)b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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

extern "C" int vp_batch_linear_coefficients(vp_batch *b, double *C_out)
{
    if (!b || !C_out) return VP_ERR_INVALID_ARGUMENT;
    vp_ctx *ctx = b->ctx;
    if (!b->fitted) return fail(ctx, VP_ERR_NO_CACHED_CALCULATION, "vp_batch: coefficients are available after vp_batch_fit");
    cudaSetDevice(ctx->device);
    VP_CUDA(ctx, cudaMemcpyAsync(C_out, b->C, sizeof(double) * (size_t)b->model->md.n * b->P, cudaMemcpyDeviceToHost, ctx->stream));
    VP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VP_OK;
}

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
    if (pr->comm) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_profile_evaluation: not for column-sharded problems");
    if (pr->model->hosteval) return fail(ctx, VP_ERR_INVALID_ARGUMENT, "vp_profile_evaluation: not for host-evaluated models");
    for (int it = 0; it < 64 && rc == VP_OK; ++it) {
        if (fused) { rc = launch_fused(pr, pr->cur ^ 1, false); continue; }
        rc = launch_panel(pr);
        if (rc == VP_OK) rc = launch_stream(pr, pr->cur ^ 1);
    }
    // timed launches are enqueued back to back; no host synchronisation in between
    for (int it = 0; it < iters && rc == VP_OK; ++it) {
        if (flush) cudaMemsetAsync(flush, it & 0xff, (size_t)flush_bytes, ctx->stream);
        cudaEventRecord(ev[3 * it + 0], ctx->stream);
        if (!fused) rc = launch_panel(pr);
        cudaEventRecord(ev[3 * it + 1], ctx->stream);
        if (rc == VP_OK) rc = fused ? launch_fused(pr, pr->cur ^ 1, false) : launch_stream(pr, pr->cur ^ 1);
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
    if (e != cudaSuccess) return fail(ctx, VP_ERR_CUDA, cudaGetErrorString(e));
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
        rc = launch_eval(pr, pr->cur ^ 1);
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
