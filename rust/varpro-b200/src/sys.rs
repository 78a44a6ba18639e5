//! Raw bindings: one declaration per symbol of `include/varpro_b200.h` (VP_ABI_VERSION 2), in header order.
//! Each group cites the reference interface it replaces (file:line in geo-ant/varpro v0.13.3).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_longlong, c_void};

pub const VP_ABI_VERSION: c_int = 2;
pub const VP_MAX_BASIS_PARAMS: usize = 4;
pub const VP_MAX_N: usize = 8;
pub const VP_MAX_Q: usize = 8;
pub const VP_MAX_P: usize = 12;
pub const VP_COMM_HANDLE_BYTES: usize = 64;

// vp_status
pub const VP_OK: c_int = 0;
pub const VP_ERR_Y_DATA_MISSING: c_int = 1;
pub const VP_ERR_INVALID_LENGTH_OF_DATA: c_int = 2;
pub const VP_ERR_ZERO_LENGTH_VECTOR: c_int = 3;
pub const VP_ERR_INVALID_PARAMETER_COUNT: c_int = 4;
pub const VP_ERR_INVALID_LENGTH_OF_WEIGHTS: c_int = 5;
pub const VP_ERR_PARAMETER_NOT_IN_MODEL: c_int = 10;
pub const VP_ERR_DERIVATIVE_INDEX_OUT_OF_BOUNDS: c_int = 11;
pub const VP_ERR_INCORRECT_PARAMETER_COUNT: c_int = 12;
pub const VP_ERR_EMPTY_MODEL: c_int = 13;
pub const VP_ERR_UNUSED_PARAMETER: c_int = 14;
pub const VP_ERR_UNSUPPORTED_BASIS: c_int = 15;
pub const VP_ERR_MODEL_TOO_LARGE: c_int = 16;
pub const VP_ERR_NO_CACHED_CALCULATION: c_int = 20;
pub const VP_ERR_UNDERDETERMINED: c_int = 30;
pub const VP_ERR_MATRIX_INVERSION: c_int = 31;
pub const VP_ERR_INVALID_ARGUMENT: c_int = 40;
pub const VP_ERR_CUDA: c_int = 41;
pub const VP_ERR_OUT_OF_MEMORY: c_int = 42;
pub const VP_ERR_COMM: c_int = 43;

// vp_dtype, vp_basis_kind, Jacobian modes
pub const VP_F64: c_int = 0;
pub const VP_F32: c_int = 1;
pub const VP_BASIS_EXP_DECAY: i32 = 0;
pub const VP_BASIS_CONSTANT: i32 = 1;
pub const VP_BASIS_EXP_RATE_COS: i32 = 2;
pub const VP_BASIS_SIN_PHASE: i32 = 3;
pub const VP_BASIS_LINEAR_X: i32 = 4;
pub const VP_BASIS_HOST: i32 = 100;
pub const VP_JACOBIAN_KAUFMAN: c_int = 0;
pub const VP_JACOBIAN_FULL: c_int = 1;
pub const VP_RANK_ABSOLUTE: c_int = 0;
pub const VP_RANK_RELATIVE: c_int = 1;

#[repr(C)] pub struct vp_ctx { _p: [u8; 0] }
#[repr(C)] pub struct vp_model { _p: [u8; 0] }
#[repr(C)] pub struct vp_problem { _p: [u8; 0] }
#[repr(C)] pub struct vp_comm { _p: [u8; 0] }
#[repr(C)] pub struct vp_batch { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct vp_basis_desc {
    pub kind: i32,
    pub n_params: i32,
    pub param_idx: [i32; VP_MAX_BASIS_PARAMS],
    pub scale: f64,
}

/// Negative (or NaN) = crate default; 0 is a legal tolerance (disables the criterion).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct vp_lm_options {
    pub ftol: f64,
    pub xtol: f64,
    pub gtol: f64,
    pub stepbound: f64,
    pub patience: i32,
    pub scale_diag: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct vp_fit_report {
    pub termination: i32,
    pub number_of_evaluations: i32,
    pub objective_function: f64,
    pub successful: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct vp_reduced {
    pub rnorm2: f64,
    pub g: [f64; VP_MAX_Q],
    pub h: [f64; VP_MAX_Q * VP_MAX_Q],
    pub finite: i32,
    pub q: i32,
}

/// Host evaluation callback of `vp_model_create_hosteval`: fill Phi (m x n) and the p non-zero derivative columns.
pub type vp_host_eval_fn =
    unsafe extern "C" fn(user: *mut c_void, alpha: *const f64, phi_out: *mut f64, dphi_out: *mut f64) -> c_int;

#[link(name = "varpro_b200")]
extern "C" {
    // ---- lifetime -----------------------------------------------------------------------------------
    pub fn vp_abi_version() -> c_int;
    pub fn vp_status_string(status: c_int) -> *const c_char;
    pub fn vp_ctx_create(device_ordinal: c_int, out: *mut *mut vp_ctx) -> c_int;
    pub fn vp_ctx_destroy(ctx: *mut vp_ctx) -> c_int;
    pub fn vp_last_error(ctx: *const vp_ctx) -> *const c_char;
    pub fn vp_ctx_kernel_launches(ctx: *const vp_ctx) -> i64;
    pub fn vp_ctx_stream(ctx: *const vp_ctx) -> *mut c_void;
    pub fn vp_ctx_set_option(ctx: *mut vp_ctx, key: *const c_char, value: *const c_char) -> c_int;
    pub fn vp_ctx_trim(ctx: *mut vp_ctx) -> c_int;

    // ---- model: SeparableNonlinearModel evaluation (src/model/mod.rs:239-363, :441-512) ------------
    pub fn vp_model_create(ctx: *mut vp_ctx, dtype: c_int, m: i64, x_host: *const c_void, q: i32, n: i32,
                           basis: *const vp_basis_desc, out: *mut *mut vp_model) -> c_int;
    pub fn vp_model_create_hosteval(ctx: *mut vp_ctx, dtype: c_int, m: i64, q: i32, n: i32, p: i32, ind: *const i32,
                                    eval: vp_host_eval_fn, user: *mut c_void, out: *mut *mut vp_model) -> c_int;
    pub fn vp_model_destroy(model: *mut vp_model) -> c_int;

    // ---- problem: SeparableProblemBuilder::build + SeparableProblem (src/problem/builder.rs:278-324,
    //      src/problem.rs:57-107) --------------------------------------------------------------------
    pub fn vp_problem_create(ctx: *mut vp_ctx, model: *mut vp_model, s: i64, y_host: *const c_void, ld_y: i64,
                             w_host: *const c_void, svd_eps: f64, alpha0: *const f64, out: *mut *mut vp_problem) -> c_int;
    pub fn vp_problem_create_device(ctx: *mut vp_ctx, model: *mut vp_model, s: i64, y_device: *const c_void, ld_y: i64,
                                    w_host: *const c_void, svd_eps: f64, alpha0: *const f64, out: *mut *mut vp_problem) -> c_int;
    pub fn vp_problem_destroy(problem: *mut vp_problem) -> c_int;

    // ---- impl LeastSquaresProblem for SeparableProblem (src/solvers/levmar/mod.rs:42,80,91,101) ----
    pub fn vp_set_params(problem: *mut vp_problem, alpha: *const f64) -> c_int;
    pub fn vp_params(problem: *const vp_problem, alpha_out: *mut f64) -> c_int;
    pub fn vp_residuals(problem: *mut vp_problem, out_host: *mut c_void) -> c_int;
    pub fn vp_jacobian(problem: *mut vp_problem, out_host: *mut c_void) -> c_int;
    pub fn vp_linear_coefficients(problem: *mut vp_problem, out_host: *mut c_void) -> c_int; // src/problem.rs:142-150
    pub fn vp_best_fit(problem: *mut vp_problem, out_host: *mut c_void) -> c_int;            // src/fit.rs:55-59
    pub fn vp_residuals_device(problem: *mut vp_problem, out_device: *mut c_void) -> c_int;
    pub fn vp_jacobian_device(problem: *mut vp_problem, out_device: *mut c_void) -> c_int;
    pub fn vp_best_fit_device(problem: *mut vp_problem, out_device: *mut c_void) -> c_int;
    pub fn vp_problem_set_jacobian(problem: *mut vp_problem, mode: c_int) -> c_int; // :188-190 TODO of the reference
    pub fn vp_problem_set_rank_policy(problem: *mut vp_problem, policy: c_int) -> c_int; // :52-54; matlab/varpro.m:642-643
    pub fn vp_reduce(problem: *mut vp_problem, out: *mut vp_reduced) -> c_int;

    // ---- column-sharded global fit over the GPUs of one box ---------------------------------------
    pub fn vp_comm_create(ctx: *mut vp_ctx, rank: c_int, world: c_int, out: *mut *mut vp_comm,
                          local_handle_out: *mut c_void) -> c_int;
    pub fn vp_comm_connect(comm: *mut vp_comm, all_handles: *const c_void) -> c_int;
    pub fn vp_comm_connect_local(comms_in_rank_order: *mut *mut vp_comm, world: c_int) -> c_int;
    pub fn vp_comm_destroy(comm: *mut vp_comm) -> c_int;
    pub fn vp_problem_set_comm(problem: *mut vp_problem, comm: *mut vp_comm) -> c_int;

    // ---- solve: LevMarSolver::fit (src/solvers/levmar/mod.rs:238-254) --------------------------------
    pub fn vp_fit(problem: *mut vp_problem, options: *const vp_lm_options, report: *mut vp_fit_report) -> c_int;
    pub fn vp_fit_many(problems: *mut *mut vp_problem, n: i64, options: *const vp_lm_options,
                       reports: *mut vp_fit_report, reserved: i32) -> c_int;
    /// Host-resident data sets of one model, pipelined on library worker threads (copy of one problem overlaps the
    /// fit of another): the loop over SeparableProblemBuilder::build + LevMarSolver::fit per data set.
    pub fn vp_fit_host_batch(ctx: *mut vp_ctx, dtype: c_int, m: i64, x_host: *const c_void, q: i32, n: i32,
                             basis: *const vp_basis_desc, n_problems: i64, s: i64, y_hosts: *const *const c_void,
                             ld_y: i64, w_host: *const c_void, svd_eps: f64, alpha0: *const f64,
                             options: *const vp_lm_options, workers: i32, reports: *mut vp_fit_report,
                             alpha_out: *mut f64, c_outs: *const *mut c_void) -> c_int;

    // ---- FitStatistics::try_calculate per right-hand side (src/statistics/mod.rs:352-441) ----------
    pub fn vp_statistics(problem: *mut vp_problem, cov_out: *mut f64, reduced_chi2_out: *mut f64,
                         conf_sigma_out: *mut f64) -> c_int;

    // ---- independent batch (BASELINE config 3) ---------------------------------------------------------
    pub fn vp_batch_create(ctx: *mut vp_ctx, model: *mut vp_model, p: i64, y_host: *const c_void, ld_y: i64,
                           w_host: *const c_void, svd_eps: f64, alpha0: *const f64, out: *mut *mut vp_batch) -> c_int;
    pub fn vp_batch_create_device(ctx: *mut vp_ctx, model: *mut vp_model, p: i64, y_device: *const c_void, ld_y: i64,
                                  w_host: *const c_void, svd_eps: f64, alpha0: *const f64, out: *mut *mut vp_batch) -> c_int;
    pub fn vp_batch_destroy(batch: *mut vp_batch) -> c_int;
    pub fn vp_batch_fit(batch: *mut vp_batch, options: *const vp_lm_options, reports: *mut vp_fit_report) -> c_int;
    pub fn vp_batch_params(batch: *mut vp_batch, alpha_out: *mut f64) -> c_int;
    pub fn vp_batch_set_params(batch: *mut vp_batch, alpha: *const f64) -> c_int;
    pub fn vp_batch_set_rank_policy(batch: *mut vp_batch, policy: c_int) -> c_int;
    pub fn vp_batch_linear_coefficients(batch: *mut vp_batch, c_out: *mut f64) -> c_int;

    // ---- diagnostics -------------------------------------------------------------------------------------
    pub fn vp_profile_evaluation(problem: *mut vp_problem, iters: c_int, flush_bytes: i64, panel_us: *mut f64,
                                 stream_us: *mut f64, stream_grid: *mut i64, stream_smem: *mut i64) -> c_int;
    pub fn vp_measure_fp64_peaks(ctx: *mut vp_ctx, dfma_tflops: *mut f64, dexp_gexps: *mut f64) -> c_int;
    pub fn vp_debug_timeline(problem: *mut vp_problem, out: *mut c_longlong, capacity: i64, grid_out: *mut i64) -> c_int;
}
