//! `SeparableProblemBuilder` / `SeparableProblem` on the device
//! (src/problem/builder.rs:116-324, src/problem.rs:57-213, impl LeastSquaresProblem src/solvers/levmar/mod.rs:22-202).
use crate::context::Context;
use crate::error::{check, Error};
use crate::model::OnDevice;
use crate::sys;
use levenberg_marquardt::LeastSquaresProblem;
use nalgebra::{DMatrix, DVector, Dyn};
use std::marker::PhantomData;
use std::os::raw::c_void;

/// src/problem.rs:16-25
pub trait RhsType {}
pub struct SingleRhs;
pub struct MultiRhs;
impl RhsType for SingleRhs {}
impl RhsType for MultiRhs {}

/// Which singular values of Phi_w count as zero in the inner solve (see `vp_problem_set_rank_policy`).
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum RankPolicy {
    /// the reference: sigma_i <= epsilon (absolute; src/solvers/levmar/mod.rs:52-54)
    Absolute,
    /// the original MATLAB code: sigma_i <= m * eps * sigma_1 (matlab/varpro.m:642-643)
    RelativeMatlab,
}

/// Builder with the reference's method names. `new` / `mrhs` pick the right-hand-side kind.
#[allow(non_snake_case)]
pub struct SeparableProblemBuilder<Model: OnDevice, Rhs: RhsType> {
    Y: Option<DMatrix<f64>>,
    model: Model,
    epsilon: Option<f64>,
    weights: Option<DVector<f64>>,
    ctx: Option<Context>,
    phantom: PhantomData<Rhs>,
}

impl<Model: OnDevice> SeparableProblemBuilder<Model, SingleRhs> {
    /// src/problem/builder.rs:116-130
    pub fn new(model: Model) -> Self {
        Self { Y: None, model, epsilon: None, weights: None, ctx: None, phantom: PhantomData }
    }
    /// :142-150
    pub fn observations(self, observed: DVector<f64>) -> Self {
        let n = observed.nrows();
        Self { Y: Some(observed.reshape_generic(Dyn(n), Dyn(1))), ..self }
    }
}

impl<Model: OnDevice> SeparableProblemBuilder<Model, MultiRhs> {
    /// :194-206
    pub fn mrhs(model: Model) -> Self {
        Self { Y: None, model, epsilon: None, weights: None, ctx: None, phantom: PhantomData }
    }
    /// :220-228: the right-hand sides are the columns
    pub fn observations(self, observed: DMatrix<f64>) -> Self {
        Self { Y: Some(observed), ..self }
    }
}

impl<Model: OnDevice, Rhs: RhsType> SeparableProblemBuilder<Model, Rhs> {
    /// :246-251 (the absolute value is used)
    pub fn epsilon(self, eps: f64) -> Self {
        Self { epsilon: Some(eps.abs()), ..self }
    }
    /// :261-266
    pub fn weights(self, weights: DVector<f64>) -> Self {
        Self { weights: Some(weights), ..self }
    }
    /// The GPU context to build on (default: a fresh context on device 0).
    pub fn context(self, ctx: Context) -> Self {
        Self { ctx: Some(ctx), ..self }
    }

    /// :278-324: validate, copy Y to the device, form Y_w = W Y once (:307), run the first evaluation (:321).
    #[allow(non_snake_case)]
    pub fn build(self) -> Result<SeparableProblem<Model, Rhs>, Error> {
        let Y = self.Y.ok_or(Error::YDataMissing)?;
        let x_len = self.model.output_len();
        if x_len == 0 || Y.is_empty() {
            return Err(Error::ZeroLengthVector);
        }
        if x_len != Y.nrows() {
            return Err(Error::InvalidLengthOfData(format!("Given x length = {} and y length = {}", x_len, Y.nrows())));
        }
        if let Some(w) = &self.weights {
            if w.len() != Y.nrows() {
                return Err(Error::InvalidLengthOfWeights);
            }
        }
        let ctx = match self.ctx {
            Some(c) => c,
            None => Context::new(0)?,
        };
        let mut model = Box::new(self.model); // boxed: the host-evaluation trampoline keeps its address
        let alpha0: Vec<f64> = model.params().iter().copied().collect();
        let model_handle = unsafe { model.create_handle(&ctx, &mut *model as *mut Model as *mut c_void)? };
        let mut handle = std::ptr::null_mut();
        let st = unsafe {
            sys::vp_problem_create(ctx.raw(), model_handle, Y.ncols() as i64, Y.as_ptr() as *const c_void, Y.nrows() as i64,
                                   self.weights.as_ref().map_or(std::ptr::null(), |w| w.as_ptr() as *const c_void),
                                   self.epsilon.unwrap_or(-1.0), alpha0.as_ptr(), &mut handle)
        };
        if let Err(e) = check(st, ctx.raw()) {
            unsafe { sys::vp_model_destroy(model_handle) };
            return Err(e);
        }
        Ok(SeparableProblem { ctx, handle, model_handle, model, weights: self.weights, m: Y.nrows(), s: Y.ncols(), phantom: PhantomData })
    }
}

/// The fitting problem: observations, weights, the model and the cached calculations live on the device.
pub struct SeparableProblem<Model: OnDevice, Rhs: RhsType> {
    pub(crate) ctx: Context,
    pub(crate) handle: *mut sys::vp_problem,
    model_handle: *mut sys::vp_model,
    pub(crate) model: Box<Model>,
    weights: Option<DVector<f64>>,
    pub(crate) m: usize,
    pub(crate) s: usize,
    phantom: PhantomData<Rhs>,
}

impl<Model: OnDevice, Rhs: RhsType> Drop for SeparableProblem<Model, Rhs> {
    fn drop(&mut self) {
        unsafe {
            sys::vp_problem_destroy(self.handle);
            sys::vp_model_destroy(self.model_handle);
        }
    }
}

impl<Model: OnDevice, Rhs: RhsType> SeparableProblem<Model, Rhs> {
    /// src/problem.rs:205-207
    pub fn model(&self) -> &Model {
        &self.model
    }
    /// :210-212 (`None` = unit weights)
    pub fn weights(&self) -> Option<&DVector<f64>> {
        self.weights.as_ref()
    }
    fn n(&self) -> usize {
        self.model.base_function_count()
    }
    fn q(&self) -> usize {
        self.model.parameter_count()
    }
    pub(crate) fn coefficient_matrix(&self) -> Option<DMatrix<f64>> {
        let mut c = DMatrix::zeros(self.n(), self.s);
        let st = unsafe { sys::vp_linear_coefficients(self.handle, c.as_mut_ptr() as *mut c_void) };
        (st == sys::VP_OK).then_some(c)
    }
    /// FitResult::best_fit (src/fit.rs:55-59, 87-91): Phi(alpha) C with the unweighted Phi
    pub(crate) fn best_fit_matrix(&self) -> Option<DMatrix<f64>> {
        let mut b = DMatrix::zeros(self.m, self.s);
        let st = unsafe { sys::vp_best_fit(self.handle, b.as_mut_ptr() as *mut c_void) };
        (st == sys::VP_OK).then_some(b)
    }
    /// after a fit: pull the parameters the library ended at into the host-side model
    pub(crate) fn sync_model_params(&mut self) {
        let mut a = vec![0.0; self.q()];
        unsafe { sys::vp_params(self.handle, a.as_mut_ptr()) };
        self.model.adopt_params(&a);
    }
    /// Kaufman approximation (the reference) or the full Golub-Pereyra Jacobian (its TODO at levmar/mod.rs:188-190).
    pub fn set_full_jacobian(&mut self, full: bool) -> Result<(), Error> {
        check(unsafe { sys::vp_problem_set_jacobian(self.handle, if full { sys::VP_JACOBIAN_FULL } else { sys::VP_JACOBIAN_KAUFMAN }) }, self.ctx.raw())
    }
    pub fn set_rank_policy(&mut self, policy: RankPolicy) -> Result<(), Error> {
        let p = match policy { RankPolicy::Absolute => sys::VP_RANK_ABSOLUTE, RankPolicy::RelativeMatlab => sys::VP_RANK_RELATIVE };
        check(unsafe { sys::vp_problem_set_rank_policy(self.handle, p) }, self.ctx.raw())
    }
    /// ||r||^2, J^T r and J^T J at the current parameters without materialising r or J
    pub fn reduce(&mut self) -> Option<(f64, DVector<f64>, DMatrix<f64>)> {
        let mut r: sys::vp_reduced = unsafe { std::mem::zeroed() };
        let st = unsafe { sys::vp_reduce(self.handle, &mut r) };
        if st != sys::VP_OK { return None; }
        let q = r.q as usize;
        Some((r.rnorm2, DVector::from_column_slice(&r.g[..q]), DMatrix::from_column_slice(q, q, &r.h[..q * q])))
    }
}

impl<Model: OnDevice> SeparableProblem<Model, MultiRhs> {
    /// src/problem.rs:142-150: one coefficient vector per right-hand side (columns)
    pub fn linear_coefficients(&self) -> Option<DMatrix<f64>> {
        self.coefficient_matrix()
    }
}
impl<Model: OnDevice> SeparableProblem<Model, SingleRhs> {
    /// src/problem.rs:173-181
    pub fn linear_coefficients(&self) -> Option<DVector<f64>> {
        self.coefficient_matrix().map(|c| c.column(0).into_owned())
    }
}

/// The lower seam (src/solvers/levmar/mod.rs:22-202): kept so that callers driving the `levenberg_marquardt`
/// crate themselves still work -- every call is one streaming pass on the GPU. `LevMarSolver::fit` does NOT go
/// through here (the lmder state machine runs inside the fit kernel).
impl<Model: OnDevice, Rhs: RhsType> LeastSquaresProblem<f64, Dyn, Dyn> for SeparableProblem<Model, Rhs> {
    type ResidualStorage = nalgebra::VecStorage<f64, Dyn, nalgebra::U1>;
    type JacobianStorage = nalgebra::VecStorage<f64, Dyn, Dyn>;
    type ParameterStorage = nalgebra::VecStorage<f64, Dyn, nalgebra::U1>;

    /// :42-73
    fn set_params(&mut self, params: &DVector<f64>) {
        unsafe { sys::vp_set_params(self.handle, params.as_ptr()) };
        self.model.adopt_params(params.as_slice());
    }
    /// :80-82
    fn params(&self) -> DVector<f64> {
        let mut a = DVector::zeros(self.q());
        unsafe { sys::vp_params(self.handle, a.as_mut_ptr()) };
        a
    }
    /// :91-95: vec(R_w), `None` when the cache is `None`
    fn residuals(&self) -> Option<DVector<f64>> {
        let mut r = DVector::zeros(self.m * self.s);
        let st = unsafe { sys::vp_residuals(self.handle, r.as_mut_ptr() as *mut c_void) };
        (st == sys::VP_OK).then_some(r)
    }
    /// :101-201: (m*S) x q, column-major
    fn jacobian(&self) -> Option<DMatrix<f64>> {
        let mut j = DMatrix::zeros(self.m * self.s, self.q());
        let st = unsafe { sys::vp_jacobian(self.handle, j.as_mut_ptr() as *mut c_void) };
        (st == sys::VP_OK).then_some(j)
    }
}
