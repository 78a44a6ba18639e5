//! `LevMarSolver` (src/solvers/levmar/mod.rs:208-315). The external `levenberg_marquardt::minimize` loop (:247) is
//! NOT called: the same MINPACK lmder state machine runs inside the fit kernels on the q x q system.
use crate::error::{check, Error};
use crate::fit::{report_from, FitResult};
use crate::model::OnDevice;
use crate::problem::{MultiRhs, RhsType, SeparableProblem, SingleRhs};
use crate::statistics::FitStatistics;
use crate::sys;
use std::marker::PhantomData;

/// The knobs of `levenberg_marquardt::LevenbergMarquardt` (whose fields are private, so they cannot be read back
/// from the crate's own struct): same builder method names, same defaults (ftol = xtol = gtol = 30 eps,
/// stepbound 100, patience 100, scale_diag on). Unset knobs travel as -1 = "crate default".
#[derive(Clone, Copy, Debug)]
pub struct LevenbergMarquardt {
    pub(crate) options: sys::vp_lm_options,
}
impl Default for LevenbergMarquardt {
    fn default() -> Self {
        Self { options: sys::vp_lm_options { ftol: -1.0, xtol: -1.0, gtol: -1.0, stepbound: -1.0, patience: -1, scale_diag: -1 } }
    }
}
impl LevenbergMarquardt {
    pub fn new() -> Self { Self::default() }
    pub fn with_ftol(mut self, v: f64) -> Self { self.options.ftol = v; self }
    pub fn with_xtol(mut self, v: f64) -> Self { self.options.xtol = v; self }
    pub fn with_gtol(mut self, v: f64) -> Self { self.options.gtol = v; self }
    pub fn with_tol(self, v: f64) -> Self { self.with_ftol(v).with_xtol(v).with_gtol(v) }
    pub fn with_stepbound(mut self, v: f64) -> Self { self.options.stepbound = v; self }
    pub fn with_patience(mut self, v: usize) -> Self { self.options.patience = v as i32; self }
    pub fn with_scale_diag(mut self, v: bool) -> Self { self.options.scale_diag = v as i32; self }
}

pub struct LevMarSolver<Model: OnDevice> {
    solver: LevenbergMarquardt,
    phantom: PhantomData<Model>,
}

impl<Model: OnDevice> Default for LevMarSolver<Model> {
    /// :307-315
    fn default() -> Self {
        Self::with_solver(LevenbergMarquardt::default())
    }
}

impl<Model: OnDevice> LevMarSolver<Model> {
    /// :221-223
    pub fn with_solver(solver: LevenbergMarquardt) -> Self {
        Self { solver, phantom: PhantomData }
    }

    /// :238-254. The problem is moved in and returned inside the result; `Ok` / `Err` from the report.
    /// ABI-level failures (CUDA errors, time-outs) come back as `Err(FitResult)` with a `User` termination,
    /// like a `None` from `residuals()` does in the reference.
    #[allow(clippy::result_large_err)]
    pub fn fit<Rhs: RhsType>(&self, mut problem: SeparableProblem<Model, Rhs>) -> Result<FitResult<Model, Rhs>, FitResult<Model, Rhs>> {
        let mut rep = sys::vp_fit_report::default();
        let st = unsafe { sys::vp_fit(problem.handle, &self.solver.options, &mut rep) };
        if st != sys::VP_OK {
            rep = sys::vp_fit_report { termination: 0, number_of_evaluations: 0, objective_function: f64::NAN, successful: 0, reserved: 0 };
        }
        problem.sync_model_params();
        let result = FitResult::new(problem, report_from(&rep));
        if result.was_successful() { Ok(result) } else { Err(result) }
    }

    /// Many independent problems together (`vp_fit_many`): the loop callers write around `fit`, on one persistent
    /// grid with a device-side work queue. Bitwise the results of `fit` per problem.
    #[allow(clippy::type_complexity)]
    pub fn fit_many<Rhs: RhsType>(&self, mut problems: Vec<SeparableProblem<Model, Rhs>>)
        -> Result<Vec<Result<FitResult<Model, Rhs>, FitResult<Model, Rhs>>>, Error> {
        if problems.is_empty() {
            return Ok(vec![]);
        }
        let mut handles: Vec<*mut sys::vp_problem> = problems.iter().map(|p| p.handle).collect();
        let mut reps = vec![sys::vp_fit_report::default(); problems.len()];
        let ctx = problems[0].ctx.raw();
        check(unsafe { sys::vp_fit_many(handles.as_mut_ptr(), handles.len() as i64, &self.solver.options, reps.as_mut_ptr(), 0) }, ctx)?;
        Ok(problems.drain(..).zip(reps).map(|(mut p, rep)| {
            p.sync_model_params();
            let r = FitResult::new(p, report_from(&rep));
            if r.was_successful() { Ok(r) } else { Err(r) }
        }).collect())
    }

    /// :275-304 for a single right-hand side
    #[allow(clippy::result_large_err, clippy::type_complexity)]
    pub fn fit_with_statistics(&self, problem: SeparableProblem<Model, SingleRhs>)
        -> Result<(FitResult<Model, SingleRhs>, FitStatistics), FitResult<Model, SingleRhs>> {
        let result = self.fit(problem)?;
        match FitStatistics::calculate_all(&result.problem, true) {
            Ok(mut all) => Ok((result, all.remove(0))),
            Err(_) => Err(result),
        }
    }

    /// The same for EVERY right-hand side of an MRHS problem with the shared nonlinear parameters (the reference
    /// returns an error for MultiRhs, :269-278; BASELINE config 4 asks for it).
    #[allow(clippy::result_large_err, clippy::type_complexity)]
    pub fn fit_with_statistics_mrhs(&self, problem: SeparableProblem<Model, MultiRhs>, confidence_sigma: bool)
        -> Result<(FitResult<Model, MultiRhs>, Vec<FitStatistics>), FitResult<Model, MultiRhs>> {
        let result = self.fit(problem)?;
        match FitStatistics::calculate_all(&result.problem, confidence_sigma) {
            Ok(all) => Ok((result, all)),
            Err(_) => Err(result),
        }
    }
}
