//! # varpro-b200
//!
//! B200-native drop-in for the variable-projection hot path of [`varpro`] v0.13.3:
//! `SeparableProblemBuilder -> SeparableProblem -> LevMarSolver::fit -> FitResult` keep their names, argument
//! meaning and `Ok` / `Err(FitResult)` behaviour, while `set_params`, `residuals`, `jacobian` and the whole
//! Levenberg-Marquardt loop run on the GPU behind the C ABI of `libvarpro_b200.so` (`include/varpro_b200.h`).
//!
//! | reference (file:line)                                           | here                                  |
//! |-----------------------------------------------------------------|---------------------------------------|
//! | `trait SeparableNonlinearModel` `src/model/mod.rs:239-363`      | re-exported unchanged; [`model`] maps implementations onto the device |
//! | `SeparableProblemBuilder` `src/problem/builder.rs:116-324`      | [`problem::SeparableProblemBuilder`]  |
//! | `SeparableProblem` `src/problem.rs:57-213`, `impl LeastSquaresProblem` `src/solvers/levmar/mod.rs:22-202` | [`problem::SeparableProblem`] |
//! | `LevMarSolver` `src/solvers/levmar/mod.rs:208-315`              | [`solver::LevMarSolver`]              |
//! | `FitResult` `src/fit.rs:15-123`                                 | [`fit::FitResult`]                    |
//! | `FitStatistics` `src/statistics/mod.rs:60-345`                  | [`statistics::FitStatistics`]         |
//!
//! There is no CPU fallback: every computing call needs a CUDA device.
pub mod batch;
pub mod comm;
pub mod context;
pub mod error;
pub mod fit;
pub mod model;
pub mod problem;
pub mod solver;
pub mod statistics;
pub mod sys;

pub use varpro::model::SeparableNonlinearModel;

pub mod prelude {
    pub use crate::context::Context;
    pub use crate::fit::FitResult;
    pub use crate::model::{BuiltinBasis, DeviceModel, DeviceModelBuilder};
    pub use crate::problem::{MultiRhs, RhsType, SeparableProblem, SeparableProblemBuilder, SingleRhs};
    pub use crate::solver::{LevMarSolver, LevenbergMarquardt};
    pub use crate::statistics::FitStatistics;
    pub use varpro::model::SeparableNonlinearModel;
}
