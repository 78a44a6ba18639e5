//! Independent batch (BASELINE config 3): P single-RHS problems with their own observations, nonlinear parameters
//! and coefficients -- the loop `for p { LevMarSolver::fit(SeparableProblemBuilder::new(model_p)...) }` as ONE launch.
use crate::context::Context;
use crate::error::{check, Error};
use crate::model::{DeviceModel, OnDevice};
use crate::solver::LevenbergMarquardt;
use crate::sys;
use nalgebra::{DMatrix, DVector};
use std::os::raw::c_void;
use varpro::model::SeparableNonlinearModel;

pub struct IndependentBatch {
    ctx: Context,
    handle: *mut sys::vp_batch,
    model_handle: *mut sys::vp_model,
    p: usize,
    n: usize,
    q: usize,
}

impl Drop for IndependentBatch {
    fn drop(&mut self) {
        unsafe {
            sys::vp_batch_destroy(self.handle);
            sys::vp_model_destroy(self.model_handle);
        }
    }
}

impl IndependentBatch {
    /// `observations`: m x P (column p = problem p); `initial_parameters`: q x P; `weights`: shared by all problems.
    pub fn new(ctx: Context, model: &DeviceModel, observations: &DMatrix<f64>, initial_parameters: &DMatrix<f64>,
               weights: Option<&DVector<f64>>, epsilon: Option<f64>) -> Result<Self, Error> {
        let (n, q, p) = (model.base_function_count(), model.parameter_count(), observations.ncols());
        if initial_parameters.nrows() != q || initial_parameters.ncols() != p {
            return Err(Error::InvalidParameterCount);
        }
        let model_handle = unsafe { model.create_handle(&ctx, std::ptr::null_mut())? };
        let mut handle = std::ptr::null_mut();
        let st = unsafe {
            sys::vp_batch_create(ctx.raw(), model_handle, p as i64, observations.as_ptr() as *const c_void, observations.nrows() as i64,
                                 weights.map_or(std::ptr::null(), |w| w.as_ptr() as *const c_void), epsilon.unwrap_or(-1.0),
                                 initial_parameters.as_ptr(), &mut handle)
        };
        if let Err(e) = check(st, ctx.raw()) {
            unsafe { sys::vp_model_destroy(model_handle) };
            return Err(e);
        }
        Ok(Self { ctx, handle, model_handle, p, n, q })
    }
    /// Fit every problem from its current parameters; one report per problem.
    pub fn fit(&mut self, solver: &LevenbergMarquardt) -> Result<Vec<sys::vp_fit_report>, Error> {
        let mut reps = vec![sys::vp_fit_report::default(); self.p];
        check(unsafe { sys::vp_batch_fit(self.handle, &solver.options, reps.as_mut_ptr()) }, self.ctx.raw())?;
        Ok(reps)
    }
    /// q x P
    pub fn nonlinear_parameters(&self) -> Result<DMatrix<f64>, Error> {
        let mut a = DMatrix::zeros(self.q, self.p);
        check(unsafe { sys::vp_batch_params(self.handle, a.as_mut_ptr()) }, self.ctx.raw())?;
        Ok(a)
    }
    /// n x P
    pub fn linear_coefficients(&self) -> Result<DMatrix<f64>, Error> {
        let mut c = DMatrix::zeros(self.n, self.p);
        check(unsafe { sys::vp_batch_linear_coefficients(self.handle, c.as_mut_ptr()) }, self.ctx.raw())?;
        Ok(c)
    }
    pub fn set_parameters(&mut self, parameters: &DMatrix<f64>) -> Result<(), Error> {
        check(unsafe { sys::vp_batch_set_params(self.handle, parameters.as_ptr()) }, self.ctx.raw())
    }
}
