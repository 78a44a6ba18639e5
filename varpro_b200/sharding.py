"""Column sharding of a global fit over the GPUs of one box (BASELINE config 5).

The S right-hand sides of one MRHS problem (shared nonlinear parameters) are partitioned
contiguously over the ranks; every rank builds a SeparableProblem from its own columns and attaches
a `Communicator`. From then on each evaluation is collective: the per-GPU reductions
(||r||^2, J^T r, J^T J) are exchanged through NVLink peer mappings inside the evaluation kernel
(vp_comm in include/varpro_b200.h) and every rank takes the identical LM step. The reference has
no distributed code; the host side here only partitions columns and moves the 64-byte CUDA IPC
handles between the processes (torch.distributed, any backend -- gloo on CPU works).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib
from .api import SeparableProblem, VarproError, _check, _Ctx

HANDLE_BYTES = 64


def shard_columns(S: int, world: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of the columns rank `rank` owns: contiguous, sizes differ by at most one,
    the first S % world ranks hold the extra column."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, rem = divmod(int(S), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_handles(local: bytes, world: int, rank: int, group=None) -> bytes:
    """All-gather the ranks' IPC handles in rank order over torch.distributed."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    mine = torch.frombuffer(bytearray(local), dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        mine = mine.cuda()
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


class Communicator:
    """vp_comm: this rank's mailbox plus the peers' mailboxes mapped over NVLink."""

    def __init__(self, rank: int, world: int, device: int = 0, group=None, ctx_slot: int = 0, connect: bool = True):
        lib = _lib.load()
        self._ctx = _Ctx.get(device, ctx_slot)
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        handle = (C.c_ubyte * HANDLE_BYTES)()
        _check(lib.vp_comm_create(self._ctx.h, rank, world, C.byref(self._h), handle), self._ctx.h)
        if world > 1 and connect:
            allh = gather_handles(bytes(handle), world, rank, group)
            if len(allh) != world * HANDLE_BYTES:
                raise VarproError("handle exchange returned the wrong number of bytes")
            buf = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
            _check(lib.vp_comm_connect(self._h, buf), self._ctx.h)

    @classmethod
    def local_group(cls, world: int, devices=None, ctx_slots=None):
        """The `world` communicators of ONE process that drives several GPUs (vp_comm_connect_local): rank r
        lives on (devices[r], ctx_slots[r]); no IPC handles, peer access between the devices. Several contexts
        of one device work too. Each rank's collective calls (attach, set_params, fit) must then be made from
        its own host thread, concurrently with the other ranks'."""
        devices = list(devices) if devices is not None else [0] * world
        ctx_slots = list(ctx_slots) if ctx_slots is not None else list(range(world))
        comms = [cls(r, world, device=devices[r], ctx_slot=ctx_slots[r], connect=False) for r in range(world)]
        if world > 1:
            arr = (C.c_void_p * world)(*[c._h for c in comms])
            _check(_lib.load().vp_comm_connect_local(arr, world), comms[0]._ctx.h)
        return comms

    def attach(self, problem: SeparableProblem) -> SeparableProblem:
        """Make `problem` (built from this rank's columns) a shard of the global fit. Collective."""
        _check(_lib.load().vp_problem_set_comm(problem._h, self._h), self._ctx.h)
        problem._comm = self  # keep the mailbox alive as long as the problem
        return problem

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.load().vp_comm_destroy(self._h)
            self._h = C.c_void_p()


def shard_observations(Y: np.ndarray, world: int, rank: int) -> np.ndarray:
    """This rank's columns of the m x S observation matrix (column-major copy)."""
    b, e = shard_columns(Y.shape[1], world, rank)
    return np.asfortranarray(Y[:, b:e])
