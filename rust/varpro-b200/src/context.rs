//! `vp_ctx`: one CUDA device + stream + buffer pool. Used by one host thread at a time (mirrors `&mut self`).
use crate::error::{check, Error};
use crate::sys;
use std::ffi::CString;
use std::rc::Rc;

pub(crate) struct CtxHandle(pub(crate) *mut sys::vp_ctx);
impl Drop for CtxHandle {
    fn drop(&mut self) {
        unsafe { sys::vp_ctx_destroy(self.0) };
    }
}

/// Cheap to clone (reference counted); models, problems and batches keep their context alive.
#[derive(Clone)]
pub struct Context(pub(crate) Rc<CtxHandle>);

impl Context {
    /// `vp_ctx_create`: fails with `Error::Library { status: VP_ERR_CUDA, .. }` when there is no CUDA device.
    pub fn new(device_ordinal: i32) -> Result<Self, Error> {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::vp_ctx_create(device_ordinal, &mut h) }, std::ptr::null())?;
        assert_eq!(unsafe { sys::vp_abi_version() }, sys::VP_ABI_VERSION, "libvarpro_b200.so / binding version mismatch");
        Ok(Self(Rc::new(CtxHandle(h))))
    }
    pub(crate) fn raw(&self) -> *mut sys::vp_ctx {
        self.0 .0
    }
    /// Tunables (`fit_mode`, `eval_kernel`, `queue_items_per_cta`, ...): see `vp_ctx_set_option` in the header.
    pub fn set_option(&self, key: &str, value: &str) -> Result<(), Error> {
        let (k, v) = (CString::new(key).unwrap(), CString::new(value).unwrap());
        check(unsafe { sys::vp_ctx_set_option(self.raw(), k.as_ptr(), v.as_ptr()) }, self.raw())
    }
    /// Return the idle cached device buffers to the CUDA allocator.
    pub fn trim(&self) -> Result<(), Error> {
        check(unsafe { sys::vp_ctx_trim(self.raw()) }, self.raw())
    }
    /// Kernels launched so far through this context.
    pub fn kernel_launches(&self) -> i64 {
        unsafe { sys::vp_ctx_kernel_launches(self.raw()) }
    }
}
