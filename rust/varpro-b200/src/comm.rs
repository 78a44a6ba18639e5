//! Column-sharded global fit over the GPUs of one box (BASELINE config 5): every rank builds a problem from ITS
//! columns and attaches the communicator; from then on every evaluation is collective (the per-GPU sums are
//! exchanged through NVLink peer mappings inside the evaluation kernel).
use crate::context::Context;
use crate::error::{check, Error};
use crate::model::OnDevice;
use crate::problem::{RhsType, SeparableProblem};
use crate::sys;

pub struct Communicator {
    ctx: Context,
    handle: *mut sys::vp_comm,
}
impl Drop for Communicator {
    fn drop(&mut self) {
        unsafe { sys::vp_comm_destroy(self.handle) };
    }
}

impl Communicator {
    /// One process per GPU: returns the communicator and the 64-byte CUDA IPC handle of this rank's mailbox; the host
    /// all-gathers the handles in rank order (MPI, sockets, ...) and calls [`Communicator::connect`].
    pub fn create(ctx: Context, rank: i32, world: i32) -> Result<(Self, [u8; sys::VP_COMM_HANDLE_BYTES]), Error> {
        let mut h = std::ptr::null_mut();
        let mut ipc = [0u8; sys::VP_COMM_HANDLE_BYTES];
        check(unsafe { sys::vp_comm_create(ctx.raw(), rank, world, &mut h, ipc.as_mut_ptr() as *mut _) }, ctx.raw())?;
        Ok((Self { ctx, handle: h }, ipc))
    }
    pub fn connect(&mut self, all_handles_in_rank_order: &[u8]) -> Result<(), Error> {
        check(unsafe { sys::vp_comm_connect(self.handle, all_handles_in_rank_order.as_ptr() as *const _) }, self.ctx.raw())
    }
    /// ONE process driving all GPUs (one context + one host thread per GPU): no IPC handles needed.
    pub fn local_group(contexts: Vec<Context>) -> Result<Vec<Self>, Error> {
        let world = contexts.len() as i32;
        let mut comms = Vec::new();
        for (r, ctx) in contexts.into_iter().enumerate() {
            comms.push(Self::create(ctx, r as i32, world)?.0);
        }
        let mut raw: Vec<*mut sys::vp_comm> = comms.iter().map(|c| c.handle).collect();
        check(unsafe { sys::vp_comm_connect_local(raw.as_mut_ptr(), world) }, comms[0].ctx.raw())?;
        Ok(comms)
    }
    /// Make `problem` (built from this rank's columns) a shard of the global fit. Collective.
    pub fn attach<M: OnDevice, R: RhsType>(&self, problem: &mut SeparableProblem<M, R>) -> Result<(), Error> {
        check(unsafe { sys::vp_problem_set_comm(problem.handle, self.handle) }, self.ctx.raw())
    }
}
