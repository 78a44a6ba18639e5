"""N>1 host-side logic of the column-sharded global fit, on CPU with the gloo backend (world size 2).

Covers (no GPU): the column partition, the all-gather of the per-rank 64-byte mailbox handles in rank
order, and the algebra the device exchange relies on -- (||r||^2, J^T r, J^T J) of the whole problem
are the SUMS of the per-shard values -- checked with the CPU oracle on each rank's shard and
torch.distributed.all_reduce standing in for the NVLink mailbox. The device-side exchange itself is
covered on the GPU by tests/test_gpu_parity.py (world size 1) and scripts/multi_gpu_check.py (N GPUs).
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def test_shard_columns_partition():
    from varpro_b200.sharding import shard_columns
    for S in (1, 2, 7, 8, 4096, 1048576, 1048577):
        for world in (1, 2, 3, 8):
            spans = [shard_columns(S, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_columns(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist
        import workloads as W
        from varpro_b200 import sharding
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        # 1. handle exchange in rank order
        local = bytes([rank + 1]) * sharding.HANDLE_BYTES
        allh = sharding.gather_handles(local, world, rank)
        assert len(allh) == world * sharding.HANDLE_BYTES
        for r in range(world):
            assert allh[r * 64:(r + 1) * 64] == bytes([r + 1]) * 64
        # 2. per-shard reductions sum to the whole problem's
        wl = W.c2(S=37)
        Yl = sharding.shard_observations(wl["Y"], world, rank)
        b, e = sharding.shard_columns(37, world, rank)
        assert Yl.shape == (1024, e - b) and Yl.flags["F_CONTIGUOUS"]
        alpha = [1.7, 4.2]
        op = W.make_oracle(wl, alpha0=alpha, Y=Yl)
        r, J = op.residuals(), op.jacobian()
        vec = np.concatenate([[r @ r], J.T @ r, (J.T @ J).ravel()])
        t = torch.from_numpy(vec.copy())
        dist.all_reduce(t)
        whole = W.make_oracle(wl, alpha0=alpha)
        rw, Jw = whole.residuals(), whole.jacobian()
        ref = np.concatenate([[rw @ rw], Jw.T @ rw, (Jw.T @ Jw).ravel()])
        err = np.max(np.abs(t.numpy() - ref) / np.maximum(np.abs(ref), 1e-300))
        assert err <= 1e-9, err
        # 3. coefficients are sharded like the columns
        Cl = op.linear_coefficients()
        assert np.allclose(Cl, whole.linear_coefficients()[:, b:e], rtol=1e-9, atol=0)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc() + str(ex)))


def test_sharded_reduction_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
