//! Status codes of the ABI mapped back onto the reference's error enums.
use crate::sys;
use std::ffi::CStr;
use thiserror::Error;

/// `SeparableProblemBuilderError` (src/problem/builder.rs:15-46) plus the library's own failures.
#[derive(Debug, Clone, Error, PartialEq)]
pub enum Error {
    #[error("Right hand side(s) not provided")]
    YDataMissing,
    #[error("Vectors x and y must have same lengths: {0}")]
    InvalidLengthOfData(String),
    #[error("x or y must have nonzero number of elements.")]
    ZeroLengthVector,
    #[error("Initial guess vector must have same length as parameters")]
    InvalidParameterCount,
    #[error("The weights must have the same length as the data y.")]
    InvalidLengthOfWeights,
    /// `ModelError` / `ModelBuildError` conditions detected by the library (src/model/errors.rs:5-42)
    #[error("model error (status {status}): {message}")]
    Model { status: i32, message: String },
    /// cache is `None` (src/solvers/levmar/mod.rs:43-45, 70-72)
    #[error("no cached calculation: the last evaluation failed")]
    NoCachedCalculation,
    /// `FitStatistics` errors (src/statistics/mod.rs:24-58)
    #[error("problem is underdetermined")]
    Underdetermined,
    #[error("matrix inversion failed")]
    MatrixInversion,
    #[error("library error (status {status}): {message}")]
    Library { status: i32, message: String },
}

/// Turn a `vp_status` into `Result`, fetching the message of the context (or of the calling thread).
pub(crate) fn check(status: i32, ctx: *const sys::vp_ctx) -> Result<(), Error> {
    if status == sys::VP_OK {
        return Ok(());
    }
    let message = unsafe {
        let p = sys::vp_last_error(ctx);
        let s = if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() };
        if s.is_empty() { CStr::from_ptr(sys::vp_status_string(status)).to_string_lossy().into_owned() } else { s }
    };
    Err(match status {
        sys::VP_ERR_Y_DATA_MISSING => Error::YDataMissing,
        sys::VP_ERR_INVALID_LENGTH_OF_DATA => Error::InvalidLengthOfData(message),
        sys::VP_ERR_ZERO_LENGTH_VECTOR => Error::ZeroLengthVector,
        sys::VP_ERR_INVALID_PARAMETER_COUNT => Error::InvalidParameterCount,
        sys::VP_ERR_INVALID_LENGTH_OF_WEIGHTS => Error::InvalidLengthOfWeights,
        10..=16 => Error::Model { status, message },
        sys::VP_ERR_NO_CACHED_CALCULATION => Error::NoCachedCalculation,
        sys::VP_ERR_UNDERDETERMINED => Error::Underdetermined,
        sys::VP_ERR_MATRIX_INVERSION => Error::MatrixInversion,
        _ => Error::Library { status, message },
    })
}
