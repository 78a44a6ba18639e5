//! Models on the device.
//!
//! The reference's basis functions are boxed CPU closures (src/model/model_basis_function.rs:11-12) which a kernel
//! cannot call. Two routes keep the `SeparableNonlinearModel` surface:
//!
//! * [`DeviceModel`]: a table of built-in basis kinds (exactly the functions the reference's tests and benches
//!   use) plus the parameter-index map of `create_index_mapping` (src/model/detail.rs:60-78). It implements
//!   `SeparableNonlinearModel` itself (host evaluation from the same formulas, used by `best_fit` style helpers
//!   of callers) and runs entirely inside the fused kernels.
//! * ANY other `impl SeparableNonlinearModel<ScalarType = f64>` (closure models from `SeparableModelBuilder`,
//!   hand-rolled models): [`HostEvalBridge`] hands the library a trampoline that calls `set_params`, `eval` and
//!   `eval_partial_deriv` once per evaluation (`vp_model_create_hosteval`); the O(m*S) work stays on the GPU.
use crate::context::Context;
use crate::error::{check, Error};
use crate::sys;
use nalgebra::{DMatrix, DVector};
use std::convert::Infallible;
use std::os::raw::{c_int, c_void};
use varpro::model::SeparableNonlinearModel;

/// Built-in basis kinds; formulas exactly as the reference writes them.
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum BuiltinBasis {
    /// `exp(-x/tau)`; shared_test_code/src/lib.rs:101-114
    ExpDecay,
    /// `1`; shared_test_code/src/lib.rs:123
    Constant,
    /// `exp(-a x) cos(b x)`, parameters (a, b); shared_test_code/src/models.rs:321-322
    ExpRateCos,
    /// `sin(omega x + phi)`, parameters (omega, phi); src/test_helpers/mod.rs:27-51
    SinPhase,
    /// `scale * x`; src/model/builder/test.rs:97,101
    LinearX(f64),
}

impl BuiltinBasis {
    fn kind(&self) -> i32 {
        match self {
            Self::ExpDecay => sys::VP_BASIS_EXP_DECAY,
            Self::Constant => sys::VP_BASIS_CONSTANT,
            Self::ExpRateCos => sys::VP_BASIS_EXP_RATE_COS,
            Self::SinPhase => sys::VP_BASIS_SIN_PHASE,
            Self::LinearX(_) => sys::VP_BASIS_LINEAR_X,
        }
    }
    pub fn arity(&self) -> usize {
        match self {
            Self::ExpDecay => 1,
            Self::Constant | Self::LinearX(_) => 0,
            Self::ExpRateCos | Self::SinPhase => 2,
        }
    }
    /// value and the partial derivatives w.r.t. the (up to two) parameters at `x`
    fn eval(&self, x: f64, a: &[f64]) -> (f64, [f64; 2]) {
        match *self {
            Self::ExpDecay => { let e = (-x / a[0]).exp(); (e, [e * x / (a[0] * a[0]), 0.0]) }
            Self::Constant => (1.0, [0.0, 0.0]),
            Self::ExpRateCos => { let e = (-a[0] * x).exp(); let (s, c) = (a[1] * x).sin_cos(); (e * c, [-x * (e * c), -x * e * s]) }
            Self::SinPhase => { let (s, c) = (a[0] * x + a[1]).sin_cos(); (s, [x * c, c]) }
            Self::LinearX(scale) => (scale * x, [0.0, 0.0]),
        }
    }
}

/// `SeparableModelBuilder` for built-in functions: `.function(&["tau1"], ExpDecay)` replaces
/// `.function(["tau1"], exp_decay).partial_deriv("tau1", exp_decay_dtau)` (src/model/builder/mod.rs:252-525).
pub struct DeviceModelBuilder {
    names: Vec<String>,
    functions: Vec<(BuiltinBasis, Vec<usize>)>,
    x: Option<DVector<f64>>,
    initial: Option<Vec<f64>>,
    error: Option<Error>,
}

impl DeviceModelBuilder {
    pub fn new<S: AsRef<str>>(parameter_names: &[S]) -> Self {
        Self { names: parameter_names.iter().map(|s| s.as_ref().to_owned()).collect(), functions: vec![], x: None, initial: None, error: None }
    }
    pub fn invariant_function(mut self, f: BuiltinBasis) -> Self {
        self.functions.push((f, vec![]));
        self
    }
    pub fn function<S: AsRef<str>>(mut self, params: &[S], f: BuiltinBasis) -> Self {
        let mut idx = Vec::new();
        for p in params {
            match self.names.iter().position(|n| n == p.as_ref()) {
                Some(i) => idx.push(i),
                None => self.error = Some(Error::Model { status: sys::VP_ERR_PARAMETER_NOT_IN_MODEL, message: format!("Parameter '{}' is not in model", p.as_ref()) }),
            }
        }
        if idx.len() != f.arity() {
            self.error = Some(Error::Model { status: sys::VP_ERR_INCORRECT_PARAMETER_COUNT, message: format!("basis function expects {} parameters, but got {}", f.arity(), idx.len()) });
        }
        self.functions.push((f, idx));
        self
    }
    pub fn independent_variable(mut self, x: DVector<f64>) -> Self {
        self.x = Some(x);
        self
    }
    pub fn initial_parameters(mut self, p: Vec<f64>) -> Self {
        self.initial = Some(p);
        self
    }
    pub fn build(self) -> Result<DeviceModel, Error> {
        if let Some(e) = self.error {
            return Err(e);
        }
        let x = self.x.ok_or(Error::Model { status: -1, message: "MissingX".into() })?;
        let params = self.initial.ok_or(Error::Model { status: -1, message: "MissingInitialParameters".into() })?;
        if params.len() != self.names.len() {
            return Err(Error::InvalidParameterCount);
        }
        if self.functions.is_empty() {
            return Err(Error::Model { status: sys::VP_ERR_EMPTY_MODEL, message: "EmptyModel".into() });
        }
        Ok(DeviceModel { names: self.names, functions: self.functions, x, params: DVector::from_vec(params) })
    }
}

/// A model the kernels evaluate themselves.
#[derive(Clone, Debug)]
pub struct DeviceModel {
    names: Vec<String>,
    functions: Vec<(BuiltinBasis, Vec<usize>)>,
    x: DVector<f64>,
    params: DVector<f64>,
}

impl DeviceModel {
    pub fn parameters(&self) -> &[String] {
        &self.names
    }
    pub(crate) fn descriptors(&self) -> Vec<sys::vp_basis_desc> {
        self.functions.iter().map(|(f, idx)| {
            let mut d = sys::vp_basis_desc { kind: f.kind(), n_params: idx.len() as i32, param_idx: [0; sys::VP_MAX_BASIS_PARAMS], scale: 1.0 };
            for (s, &k) in idx.iter().enumerate() { d.param_idx[s] = k as i32; }
            if let BuiltinBasis::LinearX(scale) = f { d.scale = *scale; }
            d
        }).collect()
    }
    pub(crate) fn x(&self) -> &DVector<f64> {
        &self.x
    }
    pub(crate) fn set_params_unchecked(&mut self, p: &[f64]) {
        self.params = DVector::from_column_slice(p);
    }
}

impl SeparableNonlinearModel for DeviceModel {
    type ScalarType = f64;
    type Error = Infallible;
    fn parameter_count(&self) -> usize { self.names.len() }
    fn base_function_count(&self) -> usize { self.functions.len() }
    fn output_len(&self) -> usize { self.x.len() }
    fn set_params(&mut self, parameters: DVector<f64>) -> Result<(), Infallible> {
        self.params = parameters;
        Ok(())
    }
    fn params(&self) -> DVector<f64> { self.params.clone() }
    fn eval(&self) -> Result<DMatrix<f64>, Infallible> {
        let mut phi = DMatrix::zeros(self.x.len(), self.functions.len());
        for (j, (f, idx)) in self.functions.iter().enumerate() {
            let a: Vec<f64> = idx.iter().map(|&k| self.params[k]).collect();
            for i in 0..self.x.len() { phi[(i, j)] = f.eval(self.x[i], &a).0; }
        }
        Ok(phi)
    }
    fn eval_partial_deriv(&self, derivative_index: usize) -> Result<DMatrix<f64>, Infallible> {
        let mut d = DMatrix::zeros(self.x.len(), self.functions.len());
        for (j, (f, idx)) in self.functions.iter().enumerate() {
            let a: Vec<f64> = idx.iter().map(|&k| self.params[k]).collect();
            for (s, &k) in idx.iter().enumerate() {
                if k == derivative_index {
                    for i in 0..self.x.len() { d[(i, j)] += f.eval(self.x[i], &a).1[s]; }
                }
            }
        }
        Ok(d)
    }
}

/// How a model reaches the device: implemented for [`DeviceModel`] (descriptor table) and, through
/// [`HostEvalBridge`], for every other `SeparableNonlinearModel<ScalarType = f64>`.
pub trait OnDevice: SeparableNonlinearModel<ScalarType = f64> {
    /// Create the `vp_model`. `self_ptr` is the address the model will keep for the lifetime of the handle
    /// (the problem boxes its model so that the host-evaluation trampoline can call back into it).
    ///
    /// # Safety
    /// `self_ptr` must stay valid and unaliased during library calls until the handle is destroyed.
    unsafe fn create_handle(&self, ctx: &Context, self_ptr: *mut c_void) -> Result<*mut sys::vp_model, Error>;
    /// after a fit: adopt the parameters the library ended at
    fn adopt_params(&mut self, p: &[f64]);
}

impl OnDevice for DeviceModel {
    unsafe fn create_handle(&self, ctx: &Context, _self_ptr: *mut c_void) -> Result<*mut sys::vp_model, Error> {
        let descs = self.descriptors();
        let mut h = std::ptr::null_mut();
        check(sys::vp_model_create(ctx.raw(), sys::VP_F64, self.x().len() as i64, self.x().as_ptr() as *const c_void,
                                   self.parameter_count() as i32, descs.len() as i32, descs.as_ptr(), &mut h), ctx.raw())?;
        Ok(h)
    }
    fn adopt_params(&mut self, p: &[f64]) {
        self.set_params_unchecked(p);
    }
}

/// Wraps any `SeparableNonlinearModel<ScalarType = f64>` (closure models of `varpro::model::builder`, hand-rolled
/// implementations) for host evaluation. `ind` lists the non-zero derivative columns as (basis j, parameter k)
/// pairs -- the reference's per-function derivative map (src/model/mod.rs:497-510), MATLAB's `Ind`
/// (matlab/varpro.m:147-189); `HostEvalBridge::dense` assumes every basis function depends on every parameter.
pub struct HostEvalBridge<M> {
    pub model: M,
    ind: Vec<(usize, usize)>,
}

impl<M: SeparableNonlinearModel<ScalarType = f64>> HostEvalBridge<M> {
    pub fn new(model: M, nonzero_derivative_columns: Vec<(usize, usize)>) -> Self {
        Self { model, ind: nonzero_derivative_columns }
    }
    pub fn dense(model: M) -> Self {
        let ind = (0..model.base_function_count()).flat_map(|j| (0..model.parameter_count()).map(move |k| (j, k))).collect();
        Self { model, ind }
    }
}

/// Called by the library once per evaluation: model.set_params(alpha); Phi = model.eval(); the listed columns
/// of model.eval_partial_deriv(k). A model error makes the cache `None` (src/solvers/levmar/mod.rs:43-45).
unsafe extern "C" fn eval_trampoline<M: SeparableNonlinearModel<ScalarType = f64>>(
    user: *mut c_void, alpha: *const f64, phi_out: *mut f64, dphi_out: *mut f64) -> c_int {
    let bridge = &mut *(user as *mut HostEvalBridge<M>);
    let q = bridge.model.parameter_count();
    let a = DVector::from_column_slice(std::slice::from_raw_parts(alpha, q));
    if bridge.model.set_params(a).is_err() {
        return 1;
    }
    let Ok(phi) = bridge.model.eval() else { return 1 };
    let m = bridge.model.output_len();
    std::ptr::copy_nonoverlapping(phi.as_ptr(), phi_out, phi.len()); // m x n, column-major like nalgebra
    let mut cache: Vec<Option<DMatrix<f64>>> = vec![None; q];
    for (e, &(j, k)) in bridge.ind.iter().enumerate() {
        if cache[k].is_none() {
            let Ok(d) = bridge.model.eval_partial_deriv(k) else { return 1 };
            cache[k] = Some(d);
        }
        let d = cache[k].as_ref().unwrap();
        std::ptr::copy_nonoverlapping(d.column(j).as_ptr(), dphi_out.add(e * m), m);
    }
    0
}

impl<M: SeparableNonlinearModel<ScalarType = f64>> SeparableNonlinearModel for HostEvalBridge<M> {
    type ScalarType = f64;
    type Error = M::Error;
    fn parameter_count(&self) -> usize { self.model.parameter_count() }
    fn base_function_count(&self) -> usize { self.model.base_function_count() }
    fn output_len(&self) -> usize { self.model.output_len() }
    fn set_params(&mut self, p: DVector<f64>) -> Result<(), M::Error> { self.model.set_params(p) }
    fn params(&self) -> DVector<f64> { self.model.params() }
    fn eval(&self) -> Result<DMatrix<f64>, M::Error> { self.model.eval() }
    fn eval_partial_deriv(&self, k: usize) -> Result<DMatrix<f64>, M::Error> { self.model.eval_partial_deriv(k) }
}

impl<M: SeparableNonlinearModel<ScalarType = f64>> OnDevice for HostEvalBridge<M> {
    unsafe fn create_handle(&self, ctx: &Context, self_ptr: *mut c_void) -> Result<*mut sys::vp_model, Error> {
        let ind: Vec<i32> = self.ind.iter().flat_map(|&(j, k)| [j as i32, k as i32]).collect();
        let mut h = std::ptr::null_mut();
        check(sys::vp_model_create_hosteval(ctx.raw(), sys::VP_F64, self.output_len() as i64, self.parameter_count() as i32,
                                            self.base_function_count() as i32, self.ind.len() as i32, ind.as_ptr(),
                                            eval_trampoline::<M>, self_ptr, &mut h), ctx.raw())?;
        Ok(h)
    }
    fn adopt_params(&mut self, p: &[f64]) {
        let _ = self.model.set_params(DVector::from_column_slice(p));
    }
}
