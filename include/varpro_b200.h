/*
 * varpro_b200.h -- C ABI of the B200-native variable-projection engine.
 *
 * This is the drop-in boundary for the hot path of geo-ant/varpro v0.13.3
 * (file:line below are relative to the reference repository):
 *
 *   per LM iteration:  Phi(alpha) / dPhi evaluation  ->  inner linear LSQ for
 *   many right-hand sides  ->  Kaufman Jacobian  ->  Levenberg-Marquardt step
 *
 * i.e. `impl LeastSquaresProblem for SeparableProblem`
 * (src/solvers/levmar/mod.rs:22-202), the state in `SeparableProblem` /
 * `CachedCalculations` (src/problem.rs:57-107), basis-function evaluation
 * (src/model/mod.rs:441-512) and the external `levenberg-marquardt` loop
 * called at src/solvers/levmar/mod.rs:247.
 *
 * The reference has no FFI; INTEGRATION.md shows the Rust `extern "C"` block a
 * maintainer adds to bind these symbols behind the unchanged trait surface.
 *
 * Conventions: plain pointers and sizes only; every function returns a
 * vp_status (0 = ok); no exceptions cross the ABI; the library owns all device
 * memory and streams; one handle is used by one host thread at a time (mirrors
 * `&mut self`). All matrices are COLUMN-MAJOR like nalgebra's. Host pointers
 * unless the name says `_device`.
 *
 * There is NO CPU fallback: every entry point that computes needs a CUDA
 * device and fails with VP_ERR_CUDA otherwise.
 */
#ifndef VARPRO_B200_H
#define VARPRO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_ABI_VERSION 2
#define VP_MAX_BASIS_PARAMS 4 /* parameters one basis function may depend on */
#define VP_MAX_N 8            /* basis functions (linear coefficients) */
#define VP_MAX_Q 8            /* nonlinear parameters */
#define VP_MAX_P 12           /* non-zero columns over all dPhi/dalpha_k */

typedef enum {
    VP_OK = 0,
    /* SeparableProblemBuilderError (src/problem/builder.rs:15-46) */
    VP_ERR_Y_DATA_MISSING = 1,
    VP_ERR_INVALID_LENGTH_OF_DATA = 2,
    VP_ERR_ZERO_LENGTH_VECTOR = 3,
    VP_ERR_INVALID_PARAMETER_COUNT = 4,
    VP_ERR_INVALID_LENGTH_OF_WEIGHTS = 5,
    /* ModelError / ModelBuildError (src/model/errors.rs:5-42, src/model/builder/error.rs) */
    VP_ERR_PARAMETER_NOT_IN_MODEL = 10,
    VP_ERR_DERIVATIVE_INDEX_OUT_OF_BOUNDS = 11,
    VP_ERR_INCORRECT_PARAMETER_COUNT = 12,
    VP_ERR_EMPTY_MODEL = 13,          /* ModelBuildError::EmptyModel: no basis function */
    VP_ERR_UNUSED_PARAMETER = 14,     /* ModelBuildError::UnusedParameter */
    VP_ERR_UNSUPPORTED_BASIS = 15,
    VP_ERR_MODEL_TOO_LARGE = 16,      /* exceeds VP_MAX_N / VP_MAX_Q / VP_MAX_P */
    /* cache is None (src/solvers/levmar/mod.rs:43-45,70-72): evaluation failed */
    VP_ERR_NO_CACHED_CALCULATION = 20,
    /* FitStatistics errors (src/statistics/mod.rs) */
    VP_ERR_UNDERDETERMINED = 30,
    VP_ERR_MATRIX_INVERSION = 31,
    /* library */
    VP_ERR_INVALID_ARGUMENT = 40,
    VP_ERR_CUDA = 41,
    VP_ERR_OUT_OF_MEMORY = 42,
    VP_ERR_COMM = 43
} vp_status;

typedef enum { VP_F64 = 0, VP_F32 = 1 } vp_dtype;

/* Built-in basis-function kinds evaluated on the device.
 * The reference's basis functions are boxed CPU closures
 * (src/model/model_basis_function.rs:11-12) which a kernel cannot call; models
 * are therefore described by a table of these kinds plus the parameter-index
 * map the reference builds in create_index_mapping (src/model/detail.rs:60-78). */
typedef enum {
    VP_BASIS_EXP_DECAY = 0,    /* exp(-x/tau); d/dtau = exp(-x/tau)*x/tau^2
                                  shared_test_code/src/lib.rs:101-114            */
    VP_BASIS_CONSTANT = 1,     /* 1 (invariant)  shared_test_code/src/lib.rs:123 */
    VP_BASIS_EXP_RATE_COS = 2, /* exp(-a x)cos(b x), params (a,b)
                                  shared_test_code/src/models.rs:321-322,362-385 */
    VP_BASIS_SIN_PHASE = 3,    /* sin(omega x + phi), params (omega,phi)
                                  src/test_helpers/mod.rs:27-51                  */
    VP_BASIS_LINEAR_X = 4,     /* scale*x (invariant) src/model/builder/test.rs:97,101 */
    VP_BASIS_HOST = 100        /* column supplied by the host callback (vp_model_create_hosteval) */
} vp_basis_kind;

typedef struct {
    int32_t kind;                           /* vp_basis_kind */
    int32_t n_params;                       /* 0..VP_MAX_BASIS_PARAMS */
    int32_t param_idx[VP_MAX_BASIS_PARAMS]; /* index of each function parameter in alpha */
    double scale;                           /* VP_BASIS_LINEAR_X only */
} vp_basis_desc;

/* TerminationReason of levenberg-marquardt 0.14. */
typedef enum {
    VP_TERM_USER = 0,
    VP_TERM_NUMERICAL = 1,
    VP_TERM_RESIDUALS_ZERO = 2,
    VP_TERM_ORTHOGONAL = 3,
    VP_TERM_CONVERGED_FTOL = 4,
    VP_TERM_CONVERGED_XTOL = 5,
    VP_TERM_CONVERGED_FTOL_XTOL = 6,
    VP_TERM_NO_IMPROVEMENT_POSSIBLE = 7,
    VP_TERM_LOST_PATIENCE = 8,
    VP_TERM_NO_PARAMETERS = 9,
    VP_TERM_NO_RESIDUALS = 10,
    VP_TERM_WRONG_DIMENSIONS = 11
} vp_termination;

/* LevenbergMarquardt::new().with_*() knobs (src/solvers/levmar/mod.rs:221).
 * A NEGATIVE value (or NaN) selects the crate default: ftol = xtol = gtol =
 * 30*eps(dtype), stepbound = 100, patience = 100 (maxfev = patience*(q+1)),
 * scale_diag = true. ftol / xtol / gtol = 0 are legal and pass through (they
 * disable the respective convergence criterion, as in the crate); stepbound and
 * patience must be positive, so 0 selects their defaults as well. */
typedef struct {
    double ftol, xtol, gtol;
    double stepbound;
    int32_t patience;
    int32_t scale_diag;
} vp_lm_options;

/* MinimizationReport (src/fit.rs:28) */
typedef struct {
    int32_t termination;           /* vp_termination */
    int32_t number_of_evaluations; /* residual evaluations, as MINPACK counts nfev */
    double objective_function;     /* 0.5*||r_w||^2 */
    int32_t successful;            /* ResidualsZero | Orthogonal | Converged* */
    int32_t reserved;
} vp_fit_report;

/* Everything the LM step needs from one evaluation: these are
 * functions of the reference's residuals() and jacobian() outputs:
 * rnorm2 = ||vec(R)||^2, g = J^T r, H = J^T J (q x q, column-major). */
typedef struct {
    double rnorm2;
    double g[VP_MAX_Q];
    double H[VP_MAX_Q * VP_MAX_Q];
    int32_t finite; /* 0 if Phi/dPhi or the sums were not finite */
    int32_t q;
} vp_reduced;

typedef struct vp_ctx vp_ctx;
typedef struct vp_model vp_model;
typedef struct vp_problem vp_problem;

/* ---- lifetime ---------------------------------------------------------- */
int vp_abi_version(void);
const char *vp_status_string(int status);
int vp_ctx_create(int device_ordinal, vp_ctx **out);
int vp_ctx_destroy(vp_ctx *ctx);
/* text of the last error raised through this context (never NULL) */
const char *vp_last_error(const vp_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t vp_ctx_kernel_launches(const vp_ctx *ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
void *vp_ctx_stream(const vp_ctx *ctx);
/* Tunables of a context (strings; unknown keys / values -> VP_ERR_INVALID_ARGUMENT). Defaults are read from
 * the environment once, in vp_ctx_create (VP_FIT_MODE, VP_EVAL_KERNEL, ...: the upper-cased key with a VP_
 * prefix); nothing on the fit path reads the environment. Keys:
 *   fit_mode        auto (default: persistent kernel / work queue / CUDA-graph loop by availability) | host | graph
 *   eval_kernel     fused (default) | split (K1 panel kernel + K2 streaming kernel); applies to problems created later
 *   stream_kernel   auto (default) | simt | generic; applies to problems created later
 *   panel_generic, stream_stages, stream_ct, stream_occ, queue_items_per_cta, batch_slots, pool_mb,
 *   max_ctas, fit_warps (integers; max_ctas caps the grid of problems created later so that contexts can share a GPU)
 *   trace, queue_dbg, batch_dbg, dbg_fit (0 | 1: diagnostics on stderr / in-kernel timelines) */
int vp_ctx_set_option(vp_ctx *ctx, const char *key, const char *value);
/* Return the context's idle cached device / pinned buffers to the CUDA allocator (the cache is capped at
 * pool_mb, default 1024 MiB; co-resident frameworks may want the memory back). Synchronises the stream. */
int vp_ctx_trim(vp_ctx *ctx);

/* ---- model: replaces SeparableModel / SeparableNonlinearModel evaluation --
 * (src/model/mod.rs:239-363 trait, :441-512 eval / eval_partial_deriv).
 * x_host: m values of the independent variable in `dtype`. alpha has q entries;
 * every parameter must be used by some basis function
 * (src/model/builder/mod.rs:547-557) and n >= 1 (:538). */
int vp_model_create(vp_ctx *ctx, int dtype, int64_t m, const void *x_host, int32_t q, int32_t n,
                    const vp_basis_desc *basis, vp_model **out);
/* Host-evaluated model: keeps ARBITRARY SeparableNonlinearModel implementations (user closures,
 * hand-rolled models: src/model/mod.rs:239-363, src/model/detail.rs:96-127) working while the
 * O(m*S) work stays on the GPU. The library calls `eval` on the host once per evaluation with the
 * q parameters; it must fill phi_out (m x n, column-major, f64: model.eval(), :441-471) and
 * dphi_out (m x p, column-major, f64): the p NON-ZERO columns of the derivative matrices
 * (eval_partial_deriv, :473-512), column e being d(basis ind[2e]) / d(param ind[2e+1]) -- the
 * `Ind` table of the original MATLAB code (matlab/varpro.m:147-189). Return 0, or non-zero to
 * signal a model error (the cache becomes None, src/solvers/levmar/mod.rs:43-45).
 * Such problems are fitted with the host-driven LM loop (one callback + one streaming pass per
 * evaluation); the fused / persistent kernels need the built-in device basis functions. */
typedef int (*vp_host_eval_fn)(void *user, const double *alpha, double *phi_out, double *dphi_out);
int vp_model_create_hosteval(vp_ctx *ctx, int dtype, int64_t m, int32_t q, int32_t n, int32_t p, const int32_t *ind,
                             vp_host_eval_fn eval, void *user, vp_model **out);
int vp_model_destroy(vp_model *model);

/* ---- problem: replaces SeparableProblemBuilder::build + SeparableProblem ---
 * (src/problem/builder.rs:278-324, src/problem.rs:57-107).
 * Y_host: m x S column-major with leading dimension ldY >= m, in the model's
 * dtype. w_host: m weights or NULL for unit weights. svd_eps: |eps| is used;
 * pass a negative value for the default (machine epsilon of dtype,
 * builder.rs:282). alpha0: q initial parameters (f64). Copies Y to the device,
 * forms Y_w = W*Y once (:307) and runs the first evaluation (:321). */
int vp_problem_create(vp_ctx *ctx, vp_model *model, int64_t S, const void *Y_host, int64_t ldY,
                      const void *w_host, double svd_eps, const double *alpha0, vp_problem **out);
/* Same, but Y is already resident in HBM (device pointer, same layout). The
 * data is copied into the library's own buffer (device-to-device). */
int vp_problem_create_device(vp_ctx *ctx, vp_model *model, int64_t S, const void *Y_device,
                             int64_t ldY, const void *w_host, double svd_eps, const double *alpha0,
                             vp_problem **out);
int vp_problem_destroy(vp_problem *problem);

/* ---- trait-mirroring entry points (LeastSquaresProblem) ------------------- */
/* set_params: src/solvers/levmar/mod.rs:42-73 */
int vp_set_params(vp_problem *problem, const double *alpha);
/* params: :80-82 */
int vp_params(const vp_problem *problem, double *alpha_out);
/* residuals: :91-95 -> vec(R_w), m*S values in dtype, RHS-major (row = s*m+i) */
int vp_residuals(vp_problem *problem, void *out_host);
/* jacobian: :101-201 -> (m*S) x q column-major in dtype (Kaufman approximation) */
int vp_jacobian(vp_problem *problem, void *out_host);
/* SeparableProblem::linear_coefficients (src/problem.rs:142-150): n x S in dtype */
int vp_linear_coefficients(vp_problem *problem, void *out_host);
/* FitResult::best_fit (src/fit.rs:55-59): Phi(alpha)*C with the unweighted Phi, m x S */
int vp_best_fit(vp_problem *problem, void *out_host);
/* The same three outputs written into a caller-owned DEVICE buffer of the same layout (SURVEY.md
 * 8f row 3): one write-bound streaming kernel, no device-to-host copy. */
int vp_residuals_device(vp_problem *problem, void *out_device);
int vp_jacobian_device(vp_problem *problem, void *out_device);
int vp_best_fit_device(vp_problem *problem, void *out_device);
/* Jacobian used by residual-derivative consumers (vp_jacobian, vp_reduce's J^T J, vp_fit):
 * VP_JACOBIAN_KAUFMAN (default) = what the reference implements (src/solvers/levmar/mod.rs:101-201);
 * VP_JACOBIAN_FULL adds the second Golub-Pereyra term the reference leaves as a TODO (:188-190;
 * matlab/varpro.m:696-731): exact derivative of the projected residual, faster convergence on
 * large-residual problems. The fused kernels (fp64 models of the compiled shapes) accumulate the
 * extra p x p sum in the same single pass over Y, so vp_fit / vp_fit_many / column-sharded fits run
 * at full speed in either mode; other problems pay one more pass per evaluation and use the
 * host-driven LM loop. Re-evaluates at the current parameters. */
enum { VP_JACOBIAN_KAUFMAN = 0, VP_JACOBIAN_FULL = 1 };
int vp_problem_set_jacobian(vp_problem *problem, int mode);
/* Which singular values of Phi_w count as zero in the inner solve C = Phi_w^+ Y_w:
 * VP_RANK_ABSOLUTE (default) = the reference: sigma_i <= svd_eps is truncated (src/solvers/levmar/mod.rs:52-54;
 * svd_eps = SeparableProblemBuilder::epsilon, default machine epsilon), the residual is formed with that C (:57-59)
 * and the Jacobian keeps the UNtruncated projector (:123-124).
 * VP_RANK_RELATIVE = the original MATLAB rule sigma_i <= m * eps_machine * sigma_1 (matlab/varpro.m:642-643), the
 * robust choice for nearly collinear basis functions (two equal decay times, tau -> infinity next to a constant).
 * The decision is made on the n x n triangle of the panel's QR factorisation; full-rank panels pay nothing.
 * Re-evaluates at the current parameters. */
enum { VP_RANK_ABSOLUTE = 0, VP_RANK_RELATIVE = 1 };
int vp_problem_set_rank_policy(vp_problem *problem, int policy);
/* ||r||^2, J^T r and J^T J of the current parameters without materialising r or J */
int vp_reduce(vp_problem *problem, vp_reduced *out);

/* ---- multi-GPU: column-sharded global fit ------------------------------------
 * BASELINE config 5 / SURVEY.md 8(e): the S right-hand sides of ONE global fit
 * (shared alpha) are partitioned contiguously over the GPUs of a box, one
 * process per GPU. Every rank builds a vp_problem from ITS columns and attaches
 * the communicator; from then on every evaluation (vp_set_params, vp_fit, ...)
 * is collective: each GPU reduces its columns to (||r||^2, J^T r, J^T J), the
 * <= 74 doubles are exchanged in one shot through NVLink peer mappings inside
 * the evaluation kernel (no host round trip, no extra launch) and summed in
 * rank order, so all ranks take bitwise identical LM steps. All ranks must make
 * the same sequence of calls. The reference has no distributed code; this is
 * the natural sharding of src/solvers/levmar/mod.rs:42-201 over columns of Y.
 *
 * Setup: (1) every rank calls vp_comm_create and receives the 64-byte CUDA IPC
 * handle of its mailbox; (2) the host all-gathers the handles in rank order
 * (torch.distributed / MPI / any transport); (3) every rank calls
 * vp_comm_connect with the world*64 bytes. world == 1 needs no connect.
 * ONE process driving several GPUs (one context and one host thread per GPU, as a Rust caller would)
 * skips the IPC handles: create the `world` communicators and pass them in rank order to
 * vp_comm_connect_local (peer access is enabled between the devices; several contexts on ONE device
 * work too -- that is how the test-suite exercises world = 2 on a one-GPU box). */
#define VP_COMM_HANDLE_BYTES 64
typedef struct vp_comm vp_comm;
int vp_comm_create(vp_ctx *ctx, int rank, int world, vp_comm **out, void *local_handle_out);
int vp_comm_connect(vp_comm *comm, const void *all_handles);
int vp_comm_connect_local(vp_comm **comms_in_rank_order, int world);
int vp_comm_destroy(vp_comm *comm);
/* comm == NULL detaches. Re-evaluates at the current parameters (collective). */
int vp_problem_set_comm(vp_problem *problem, vp_comm *comm);

/* ---- solve: replaces LevMarSolver::fit (src/solvers/levmar/mod.rs:238-254) */
int vp_fit(vp_problem *problem, const vp_lm_options *options, vp_fit_report *report);
/* Fit n independent problems of one context (throughput mode): what a caller of the reference
 * does in a loop over LevMarSolver::fit (e.g. benches/multiple_right_hand_sides.rs:97-101 per
 * criterion iteration). Like-shaped problems are fitted by ONE persistent kernel whose CTAs take
 * (fit, group-of-columns) work items from a device-side queue, so the latency-bound phases of
 * one fit (panel, reduction, LM step) overlap the HBM streaming of the others and the load is
 * balanced dynamically. Results are BITWISE those of n calls of vp_fit (same partition of the
 * columns into partial sums, same fold, same LM code): same parameters, same number of
 * evaluations. `reserved` is ignored (ABI v1 had a concurrency hint there). reports: n entries.
 * Returns the first error, VP_OK otherwise. */
int vp_fit_many(vp_problem **problems, int64_t n, const vp_lm_options *options, vp_fit_report *reports,
                int32_t reserved);

/* Fit n_problems problems of ONE model whose observations live in HOST memory (pinned memory for full PCIe
 * speed), pipelined: `workers` (<= 0: 3) threads of the library, each with its own stream on ctx's device, build
 * (vp_problem_create: the host-to-device copy), fit (vp_fit) and read back (vp_params -> alpha_out + i*q,
 * vp_linear_coefficients -> C_outs[i], n x S in dtype; C_outs or entries of it may be NULL) one problem after the
 * other, so the copy of one problem overlaps the fit of another. Equivalent to that loop over the reference API
 * (src/problem/builder.rs:116-324 + src/solvers/levmar/mod.rs:238-254 per data set); results are those of vp_fit.
 * Y_hosts: n_problems pointers to m x S column-major matrices with leading dimension ldY. alpha0: q shared initial
 * parameters. Returns the first error. */
int vp_fit_host_batch(vp_ctx *ctx, int dtype, int64_t m, const void *x_host, int32_t q, int32_t n,
                      const vp_basis_desc *basis, int64_t n_problems, int64_t S, const void *const *Y_hosts,
                      int64_t ldY, const void *w_host, double svd_eps, const double *alpha0,
                      const vp_lm_options *options, int32_t workers, vp_fit_report *reports, double *alpha_out,
                      void *const *C_outs);

/* ---- fit statistics: replaces FitStatistics::try_calculate (src/statistics/mod.rs:352-441) ----
 * The reference computes statistics for a single right-hand side only
 * (src/solvers/levmar/mod.rs:269-278); here the same calculation is applied to EVERY
 * column s with the shared nonlinear parameters (BASELINE config 4):
 *   cov_out          (n+q)^2 doubles per column, column-major blocks, ordering (c..., alpha...)
 *                    as in the reference (:66-76, :507-510): chi2_s * (H_s^T H_s)^-1
 *   reduced_chi2_out S doubles or NULL: ||r_w,s||^2 / (m - n - q)
 *   conf_sigma_out   m x S doubles or NULL: sqrt(j_i^T Cov_s j_i) (:415-430); the caller
 *                    scales by the Student-t quantile for a confidence band (:285-288)
 * Errors: VP_ERR_UNDERDETERMINED (m <= n+q, :377-379), VP_ERR_MATRIX_INVERSION (:399). */
int vp_statistics(vp_problem *problem, double *cov_out, double *reduced_chi2_out, double *conf_sigma_out);

/* ---- independent batch (BASELINE config 3) -------------------------------------
 * P independent single-RHS problems that share the model structure, x and the
 * weights but have their own observations (column p of Y), nonlinear parameters
 * (column p of alpha0, q x P column-major) and linear coefficients: the loop
 *   for p in 0..P { LevMarSolver::fit(SeparableProblemBuilder::new(model_p).observations(y_p).build()) }
 * over the reference API (src/problem/builder.rs:116-324, src/solvers/levmar/mod.rs:238-254),
 * executed as ONE kernel launch: one CTA fits one problem from start to finish (y_p is
 * read from HBM once, Phi(alpha_p) is regenerated on the fly, per-problem LM state) and
 * then takes the next. fp64 models. vp_batch_fit starts from the current parameters
 * (alpha0 at first, the previous result afterwards; vp_batch_set_params overrides);
 * reports (P entries) may be NULL. */
typedef struct vp_batch vp_batch;
int vp_batch_create(vp_ctx *ctx, vp_model *model, int64_t P, const void *Y_host, int64_t ldY, const void *w_host,
                    double svd_eps, const double *alpha0, vp_batch **out);
int vp_batch_create_device(vp_ctx *ctx, vp_model *model, int64_t P, const void *Y_device, int64_t ldY,
                           const void *w_host, double svd_eps, const double *alpha0, vp_batch **out);
int vp_batch_destroy(vp_batch *batch);
int vp_batch_fit(vp_batch *batch, const vp_lm_options *options, vp_fit_report *reports);
int vp_batch_params(vp_batch *batch, double *alpha_out);            /* q x P */
int vp_batch_set_params(vp_batch *batch, const double *alpha);      /* q x P */
int vp_batch_set_rank_policy(vp_batch *batch, int policy);          /* VP_RANK_* */
int vp_batch_linear_coefficients(vp_batch *batch, double *C_out);   /* n x P */

/* ---- diagnostics ---------------------------------------------------------
 * Device time (CUDA events on the context's stream, microseconds, averaged over
 * `iters`) of ONE evaluation at the current parameters. Problems with a fused
 * evaluation kernel (fit_kernel_dmma: panel + streaming reduce in one launch) report
 * it in stream_us with panel_us ~ 0; otherwise panel_us is the panel kernel (K1) and
 * stream_us the Y-streaming reduce (K2). If flush_bytes > 0 a scratch buffer of that
 * size is overwritten before every iteration so that the observations are read from
 * HBM, not from L2. Also reports the streaming kernel's grid size and dynamic shared
 * memory. */
int vp_profile_evaluation(vp_problem *problem, int iters, int64_t flush_bytes, double *panel_us,
                          double *stream_us, int64_t *stream_grid, int64_t *stream_smem);

/* fp64 ALU peaks of the context's GPU measured on its stream: dependent-chain-free DFMA throughput (TFLOP/s) and
 * double-precision exp() throughput (Gexp/s). The denominators of the fp64-ALU roofline of the independent-batch
 * kernel (bench.py); MEASURED_PEAKS.json has no fp64 entry. */
int vp_measure_fp64_peaks(vp_ctx *ctx, double *dfma_tflops, double *dexp_gexps);

/* One evaluation with the in-kernel timeline enabled (diagnostics): `out`
 * receives (grid+1)*16 %globaltimer stamps in ns relative to the earliest one,
 * one row per CTA of K2 (start, panel slice loaded, first tile landed, loop
 * done, publish begin, published, finalize done, -) and a last row for K1. */
int vp_debug_timeline(vp_problem *problem, long long *out, int64_t capacity, int64_t *grid_out);

#ifdef __cplusplus
}
#endif
#endif /* VARPRO_B200_H */
