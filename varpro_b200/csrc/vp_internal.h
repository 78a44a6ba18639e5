// vp_internal.h -- host-side handle types and helpers shared by the translation units that implement the C ABI
// (include/varpro_b200.h): vp_ctx.cu (contexts, options, buffer pools, models), vp_problem.cu (planning, launches,
// problems, trait-mirroring entry points, statistics), vp_fit.cu (LM drivers, vp_fit / vp_fit_many, vp_comm),
// vp_batch.cu (independent batch), vp_diag.cu (profiling entry points). There is no CPU compute path anywhere:
// every evaluation launches the sm_100a kernels.
#pragma once

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
#include "device_common.cuh"
#include "kernel_tables.h"
#include "lm_step.cuh"
#include "panel_kernel.cuh"
#include "stream_kernel.cuh"
#include "fit_kernel_dmma.cuh"

// NVTX ranges around the ABI entry points (tracing row): visible in Nsight Systems / ncu.
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
struct VpNvtxRange {
    explicit VpNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~VpNvtxRange() { nvtxRangePop(); }
};
#define VP_NVTX(name) VpNvtxRange vp_nvtx_range_(name)
#else
#define VP_NVTX(name) do { } while (0)
#endif

// ----------------------------------------------------------------------------
// buffer pools
// ----------------------------------------------------------------------------
// Size-keyed free lists for device and pinned-host buffers: problems of the same shape are
// created and destroyed per fit by callers that mirror the reference API (the builder produces
// a new SeparableProblem every time), and cudaMalloc / cudaMallocHost cost more than a fit.
struct BufferPool {
    std::unordered_map<void *, size_t> live;
    std::unordered_multimap<size_t, void *> free_list; // exact-size buckets
    size_t free_bytes = 0;
    size_t cap_bytes = 0;
};
cudaError_t pool_alloc(BufferPool &pool, bool host, void **out, size_t bytes);
void pool_free(BufferPool &pool, bool host, void *p);
void pool_release(BufferPool &pool, bool host);
#define DEV_ALLOC(ctx, pp, bytes) pool_alloc((ctx)->dev_pool, false, (void **)(pp), (bytes))
#define HOST_ALLOC(ctx, pp, bytes) pool_alloc((ctx)->host_pool, true, (void **)(pp), (bytes))
#define DEV_FREE(ctx, p) pool_free((ctx)->dev_pool, false, (void *)(p))
#define HOST_FREE(ctx, p) pool_free((ctx)->host_pool, true, (void *)(p))

// ----------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------
// Tunables of a context. Defaults come from the environment ONCE, in vp_ctx_create (VP_* variables, for
// experiments on a box); vp_ctx_set_option changes them afterwards (what the tests use). Nothing on the
// fit path calls getenv.
enum { VP_FITMODE_AUTO = 0, VP_FITMODE_HOST = 1, VP_FITMODE_GRAPH = 2 };
enum { VP_STREAMK_AUTO = 0, VP_STREAMK_SIMT = 1, VP_STREAMK_GENERIC = 2 };
struct CtxOptions {
    int fit_mode = VP_FITMODE_AUTO;   // auto: persistent kernel / work queue / CUDA-graph loop / host loop by availability
    int eval_split = 0;               // 1: K1 + K2 instead of the fused evaluation kernel
    int stream_kernel = VP_STREAMK_AUTO;
    int panel_generic = 0;            // 1: always the CGS2 interpreter panel kernel
    int stream_stages = 0;            // > 1: cap on the TMA ring depth
    int stream_ct = 4, stream_occ = 2; // SIMT streaming kernel: preferred columns per tile, CTAs per SM
    int queue_items_per_cta = 2;      // work-queue kernel: items per CTA and evaluation round the item size aims at
    int queue_parts_per_item = 0;     // > 0: fixed work-item size in parts (diagnostics); 0: sized per evaluation by the kernel
    int batch_dbg = 0;                // 1: per-phase accumulators of the independent-batch kernel on stderr
    int queue_dbg = 0;                // 1: per-phase accumulators of the work-queue kernel on stderr
    int trace = 0;                    // 1: per-evaluation LM trace on stderr
    int dbg_fit = 0;                  // 1: in-kernel timeline of the persistent fit (vp_debug_timeline reads it)
    int batch_slots = 8;              // independent-batch kernel: problems in flight per CTA
    int pool_mb = 1024;               // cap of the idle device-buffer cache
    int fit_warps = 8;                // fused / work-queue kernels: preferred warps per CTA among the tilings that cover m
    int max_ctas = 0;                 // > 0: cap on the grid of the fused / streaming kernels (problems created later);
                                      // lets several contexts share one GPU concurrently
};

struct vp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t launches = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    BufferPool dev_pool, host_pool;
    CtxOptions opt;
    std::vector<vp_ctx *> workers; // contexts of vp_fit_host_batch's worker threads (same device, own streams)
};

struct vp_model {
    vp_ctx *ctx = nullptr;
    int dtype = VP_F64;
    vp::ModelDesc md{};
    void *x_dev = nullptr;
    int ld = 0; // padded row count (multiple of 16/sizeof(T))
    // host-evaluated model (vp_model_create_hosteval)
    bool hosteval = false;
    vp_host_eval_fn eval_fn = nullptr;
    void *eval_user = nullptr;
    double *pre_host = nullptr; // pinned m x (n+p)
    double *pre_dev = nullptr;  // the unweighted [Phi | D] of the last callback
};

// Column-sharded global fit across the GPUs of one box: this rank's mailbox (device memory, exported through
// CUDA IPC for one-process-per-GPU setups) plus the peers' mailboxes mapped over NVLink.
struct vp_comm {
    vp_ctx *ctx = nullptr;
    int world = 1, rank = 0;
    bool connected = false;
    bool ipc_peers = false; // peer[] came from cudaIpcOpenMemHandle (must be closed)
    vp::CommMailbox *local_box = nullptr;
    unsigned long long *epoch = nullptr; // device
    int *error = nullptr;                // device
    void *peer[vp::COMM_MAX_WORLD] = {nullptr};
    vp::CommArgs args{};
};

struct vp_problem {
    vp_ctx *ctx = nullptr;
    vp_comm *comm = nullptr;
    vp_model *model = nullptr;
    int64_t S = 0;
    void *Yw = nullptr;    // ld x S
    void *w_dev = nullptr; // m or null
    double svd_eps = 0.0;  // the builder's epsilon (absolute threshold)
    int rank_policy = 0;   // VP_RANK_ABSOLUTE / VP_RANK_RELATIVE
    double rank_tol = 0.0; // what the kernels get (rank_policy.cuh): >= 0 absolute eps; < 0: -(m * eps_machine), relative
    double alpha[VP_MAX_Q] = {0};
    double *alpha_dev = nullptr; // = &fit_dev->st.x_trial[0]: the parameters the next evaluation is made at
    vp::FitDevice *fit_dev = nullptr;  // device-resident LM state
    vp::FitDevice *fit_host = nullptr; // pinned staging copy (VP_FIT_BLOCK_BYTES: the state and the control word)
    cudaGraph_t fit_graph = nullptr;
    cudaGraphExec_t fit_exec = nullptr;
    cudaGraphConditionalHandle fit_cond = 0;
    void *Pq = nullptr; // panel [Q | E | 0], (n+p+1) columns of ldp rows
    int ldp = 0;
    vp::PanelSmall *small = nullptr;
    void *C[2] = {nullptr, nullptr}; // coefficient buffers (n x S); C[cur] belongs to `alpha`
    int cur = 0;
    double *partials = nullptr;
    int red_stride = 0;
    int max_grid = 0;
    unsigned int *ticket = nullptr;
    vp::EvalOut *out_dev = nullptr;
    vp::EvalOut *out_host = nullptr; // pinned
    double *alpha_stage = nullptr;   // pinned: VP_MAX_Q doubles + a FitCtl
    double *phi_scratch = nullptr;   // m x n (best_fit)
    unsigned long long *dbg = nullptr; // optional in-kernel timeline (vp_debug_timeline)
    vp::LmEval eval{};               // reduction at `alpha`
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
    int fit_grid = 0, fit_nst = 0, fit_lds = 0;
    size_t fit_smem = 0;
    vp::FitCtl *fit_ctl = nullptr;
    int jac_full = 0;       // 1: full Golub-Pereyra Jacobian (vp_problem_set_jacobian); 0: Kaufman (the reference)
    double *Pq64 = nullptr; // f64 panel buffer used by the work-queue kernel for fp32 problems (lazily allocated)
    int ldp64 = 0;
};

// ----------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------
int vp_fail(vp_ctx *ctx, int code, const std::string &msg);
#define VP_CUDA(ctx, expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return vp_fail((ctx), (_e == cudaErrorMemoryAllocation) ? VP_ERR_OUT_OF_MEMORY : VP_ERR_CUDA, \
                           std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)

inline double vp_rank_tol(int policy, double svd_eps, int m, int dtype)
{
    return policy == VP_RANK_RELATIVE ? -(double)m * (dtype == VP_F32 ? (double)FLT_EPSILON : DBL_EPSILON) : svd_eps;
}
inline size_t vp_esize(int dtype) { return dtype == VP_F32 ? 4 : 8; }
inline int vp_vec_of(int dtype) { return dtype == VP_F32 ? 4 : 2; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel shared by every problem and host
// thread of the process: the limit may only ever be raised (a later, smaller request must not lower it under a launch
// that was planned with the larger value). Cached per (device, function).
cudaError_t vp_ensure_dynamic_smem(int device, const void *fn, size_t bytes);

// ----------------------------------------------------------------------------
// kernel tables (inst.cu translation units, see kernel_tables.h)
// ----------------------------------------------------------------------------
struct KernelTables {
    std::vector<StreamKernelEntry> stream;
    std::vector<DmmaKernelEntry> dmma;
    std::vector<PanelHHEntry> panel;
    std::vector<FitKernelEntry> fit;
    std::vector<BatchKernelEntry> batch;
    std::vector<QueueKernelEntry> queue;
};
const KernelTables &vp_kernel_tables(); // thread-safe, built on first use

// Padded column stride (elements) of a tile slot in shared memory: the smallest value >= n whose fragment loads are
// bank-conflict free, 4 (mod 16) doubles / 8 (mod 32) floats (see stream_kernel_dmma.cuh).
inline int vp_pad_lds(int dtype, int n)
{
    if (dtype == VP_F64) { while (n % 16 != 4) n += 2; }
    else { while (n % 32 != 8) n += 4; }
    return n;
}
// The stride to use with a (rows, exact) tiling for a problem with leading dimension ld, or -1 if the tiling does not
// apply. The unpredicated EXACT tile code needs every fragment row inside the slot (rows <= lds): when the rows of
// the tiling exceed ld by at most 1/8, the slots are simply made that much taller (their extra rows are zero), so that
// e.g. m = 1000 runs the EXACT code of the 1024-row tiling.
inline int vp_tile_lds(int dtype, int ld, int rows, int exact)
{
    const int lds = vp_pad_lds(dtype, ld);
    if (!exact || rows <= lds) return lds;
    if ((long long)rows * 8 <= (long long)ld * 9) return vp_pad_lds(dtype, rows);
    return -1;
}

// device / pinned block of a problem's LM state: [FitDevice | FitCtl]
constexpr size_t VP_FIT_CTL_OFFSET = (sizeof(vp::FitDevice) + 15) / 16 * 16;
constexpr size_t VP_FIT_BLOCK_BYTES = VP_FIT_CTL_OFFSET + sizeof(vp::FitCtl);

// Row tilings of the DMMA-based kernels: (ksteps, nwarps) covers 4*ksteps*nwarps rows. Better = fewer rows (less
// padding work), then the unpredicated EXACT variant, then the preferred warp count.
inline bool vp_better_tiling(int ks, int nw, int exact, int ks0, int nw0, int exact0, int prefer_warps)
{
    const int rows = 4 * ks * nw, rows0 = 4 * ks0 * nw0;
    if (rows != rows0) return rows < rows0;
    if ((exact != 0) != (exact0 != 0)) return exact != 0;
    return nw == prefer_warps && nw0 != prefer_warps;
}

// ----------------------------------------------------------------------------
// internal entry points shared between the translation units
// ----------------------------------------------------------------------------
int vp_launch_panel(vp_problem *pr);
int vp_launch_stream(vp_problem *pr, int cdst, bool graph_mode = false);
int vp_launch_fused(vp_problem *pr, int cdst, bool fit_mode);
int vp_launch_eval(vp_problem *pr, int cdst);
int vp_comm_check(vp_problem *pr);
int vp_evaluate_sync(vp_problem *pr, const double *alpha, int cdst);
void vp_evalout_to_lm(const vp::EvalOut &o, int q, vp::LmEval &ev);
void vp_lm_config_from_options(int dtype, int q, const vp_lm_options *opt, vp::LmConfig &cfg);
int vp_refresh_cached_evaluation(vp_problem *pr); // re-evaluate at pr->alpha into the other coefficient buffer
