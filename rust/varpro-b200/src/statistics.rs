//! `FitStatistics` (src/statistics/mod.rs:60-345) filled by `vp_statistics` (try_calculate, :352-441, per column).
use crate::error::{check, Error};
use crate::model::OnDevice;
use crate::problem::{RhsType, SeparableProblem};
use crate::sys;
use nalgebra::{DMatrix, DVector};
use varpro::model::SeparableNonlinearModel;

pub struct FitStatistics {
    /// ordered (linear coefficients..., nonlinear parameters...) as in the reference (:66-76, :507-510)
    covariance_matrix: DMatrix<f64>,
    reduced_chi2: f64,
    degrees_of_freedom: usize,
    linear_coefficient_count: usize,
    /// sqrt(j_i^T Cov j_i) per sample (:415-430); scale by the Student-t quantile for a band (:285-288)
    unscaled_confidence_sigma: Option<DVector<f64>>,
}

impl FitStatistics {
    pub(crate) fn calculate_all<Model: OnDevice, Rhs: RhsType>(problem: &SeparableProblem<Model, Rhs>, confidence_sigma: bool)
        -> Result<Vec<FitStatistics>, Error> {
        let (n, q) = (problem.model().base_function_count(), problem.model().parameter_count());
        let (t, m, s) = (n + q, problem.m, problem.s);
        let mut cov = vec![0.0f64; t * t * s];
        let mut chi2 = vec![0.0f64; s];
        let mut conf = if confidence_sigma { vec![0.0f64; m * s] } else { vec![] };
        check(unsafe {
            sys::vp_statistics(problem.handle, cov.as_mut_ptr(), chi2.as_mut_ptr(),
                               if confidence_sigma { conf.as_mut_ptr() } else { std::ptr::null_mut() })
        }, problem.ctx.raw())?;
        Ok((0..s).map(|k| FitStatistics {
            covariance_matrix: DMatrix::from_column_slice(t, t, &cov[k * t * t..(k + 1) * t * t]),
            reduced_chi2: chi2[k],
            degrees_of_freedom: m - t,
            linear_coefficient_count: n,
            unscaled_confidence_sigma: confidence_sigma.then(|| DVector::from_column_slice(&conf[k * m..(k + 1) * m])),
        }).collect())
    }
    /// :129-131
    pub fn covariance_matrix(&self) -> &DMatrix<f64> { &self.covariance_matrix }
    /// :147-158
    pub fn calculate_correlation_matrix(&self) -> DMatrix<f64> {
        let d: Vec<f64> = self.covariance_matrix.diagonal().iter().map(|v| v.sqrt()).collect();
        DMatrix::from_fn(d.len(), d.len(), |i, j| self.covariance_matrix[(i, j)] / (d[i] * d[j]))
    }
    /// :174-179
    pub fn regression_standard_error(&self) -> f64 { self.reduced_chi2.sqrt() }
    /// :183-185
    pub fn reduced_chi2(&self) -> f64 { self.reduced_chi2 }
    /// :206-214
    pub fn linear_coefficients_variance(&self) -> DVector<f64> {
        DVector::from_iterator(self.linear_coefficient_count, self.covariance_matrix.diagonal().iter().take(self.linear_coefficient_count).copied())
    }
    /// :190-200
    pub fn nonlinear_parameters_variance(&self) -> DVector<f64> {
        let n = self.linear_coefficient_count;
        DVector::from_iterator(self.covariance_matrix.nrows() - n, self.covariance_matrix.diagonal().iter().skip(n).copied())
    }
    pub fn degrees_of_freedom(&self) -> usize { self.degrees_of_freedom }
    /// :271-296: radius of the confidence band at `probability`, given the Student-t quantile for
    /// `degrees_of_freedom()` (the caller supplies it; the reference uses `distrs::StudentsT::ppf`).
    pub fn confidence_band_radius(&self, student_t_quantile: f64) -> Option<DVector<f64>> {
        self.unscaled_confidence_sigma.as_ref().map(|s| s * student_t_quantile)
    }
}
