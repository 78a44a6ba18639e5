"""varpro_b200 -- B200-native variable-projection hot path behind the trait
surface of geo-ant/varpro. Importing the package does not load
the CUDA library; the first use of a problem/solver does, and fails loudly if
the extension is missing."""
from .api import (  # noqa: F401
    BatchFitResult, Constant, ExpDecay, ExpRateCos, FitError, FitResult, FitStatistics, HostFunction, IndependentBatch, LevenbergMarquardt, LevMarSolver, LinearX,
    MinimizationReport, ModelBuildError, ModelError, SeparableModel, SeparableModelBuilder,
    SeparableProblem, SeparableProblemBuilder, SeparableProblemBuilderError, SinPhase,
    TerminationReason, VarproError, kernel_launches, set_option,
)

__all__ = [n for n in dir() if not n.startswith("_")]
