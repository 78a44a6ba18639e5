//! `FitResult` (src/fit.rs:15-123): the final problem plus the minimization report.
use crate::model::OnDevice;
use crate::problem::{MultiRhs, RhsType, SeparableProblem, SingleRhs};
use crate::sys;
use levenberg_marquardt::{MinimizationReport, TerminationReason};
use nalgebra::{DMatrix, DVector};
use varpro::model::SeparableNonlinearModel;

pub struct FitResult<Model: OnDevice, Rhs: RhsType> {
    /// final state of the problem (regardless of success)
    pub problem: SeparableProblem<Model, Rhs>,
    /// the report of the LM loop, in the `levenberg_marquardt` crate's own type
    pub minimization_report: MinimizationReport<f64>,
}

/// `vp_fit_report` -> `MinimizationReport`; the termination codes are the crate's variants in declaration order.
pub(crate) fn report_from(rep: &sys::vp_fit_report) -> MinimizationReport<f64> {
    let termination = match rep.termination {
        0 => TerminationReason::User("residuals() or jacobian() returned None"),
        1 => TerminationReason::Numerical("non-finite value during the minimization"),
        2 => TerminationReason::ResidualsZero,
        3 => TerminationReason::Orthogonal,
        4 => TerminationReason::Converged { ftol: true, xtol: false },
        5 => TerminationReason::Converged { ftol: false, xtol: true },
        6 => TerminationReason::Converged { ftol: true, xtol: true },
        7 => TerminationReason::NoImprovementPossible("machine precision reached"),
        8 => TerminationReason::LostPatience,
        9 => TerminationReason::NoParameters,
        10 => TerminationReason::NoResiduals,
        _ => TerminationReason::WrongDimensions("unexpected termination code"),
    };
    MinimizationReport { termination, number_of_evaluations: rep.number_of_evaluations as usize, objective_function: rep.objective_function }
}

impl<Model: OnDevice, Rhs: RhsType> FitResult<Model, Rhs> {
    pub(crate) fn new(problem: SeparableProblem<Model, Rhs>, minimization_report: MinimizationReport<f64>) -> Self {
        Self { problem, minimization_report }
    }
    /// src/fit.rs:113-115
    pub fn nonlinear_parameters(&self) -> DVector<f64> {
        self.problem.model().params()
    }
    /// src/fit.rs:120-122
    pub fn was_successful(&self) -> bool {
        self.minimization_report.termination.was_successful()
    }
}

impl<Model: OnDevice> FitResult<Model, MultiRhs> {
    /// src/fit.rs:45-47
    pub fn linear_coefficients(&self) -> Option<DMatrix<f64>> {
        self.problem.linear_coefficients()
    }
    /// src/fit.rs:55-59
    pub fn best_fit(&self) -> Option<DMatrix<f64>> {
        self.problem.best_fit_matrix()
    }
}
impl<Model: OnDevice> FitResult<Model, SingleRhs> {
    /// src/fit.rs:72-80
    pub fn linear_coefficients(&self) -> Option<DVector<f64>> {
        self.problem.linear_coefficients()
    }
    /// src/fit.rs:87-91
    pub fn best_fit(&self) -> Option<DVector<f64>> {
        self.problem.best_fit_matrix().map(|b| b.column(0).into_owned())
    }
}
