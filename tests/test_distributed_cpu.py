"""world_size-2 gloo test (CPU) of the multi-GPU path's host logic: shard -> scatter ->
per-rank solve -> gather.  The per-rank solver is a stand-in (closed-form for a trivial QP
family), because the real one needs a GPU; the sharding/collective plumbing is identical."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _standin_solver(P, q, G, h):
    # unconstrained minimiser of 1/2 x'Px + q'x, slack of the (inactive) constraints
    B, n = q.shape
    x = np.stack([np.linalg.solve(P[k], -q[k]) for k in range(B)]) if B else np.zeros((0, n))
    s = h - np.einsum("bmn,bn->bm", G, x) if B else np.zeros((0, h.shape[1]))
    return {"x": x, "s": s, "z": np.zeros_like(s), "status_code": np.ones(B, np.int32),
            "iterations": np.arange(B, dtype=np.int32), "primal objective": 0.5 * np.einsum("bn,bn->b", q, x),
            "dual objective": 0.5 * np.einsum("bn,bn->b", q, x)}


def _worker(rank, world, port, nprob, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cvxopt_b200.batch import qp_batch_distributed
    rng = np.random.default_rng(7)
    n, m = 6, 9
    A = rng.standard_normal((nprob, n, n))
    P = np.einsum("bij,bkj->bik", A, A) + np.eye(n)
    q = rng.standard_normal((nprob, n))
    G = rng.standard_normal((nprob, m, n))
    h = 100.0 + rng.standard_normal((nprob, m))
    args = (P, q, G, h) if rank == 0 else (None, None, None, None)
    res = qp_batch_distributed(*args, solver=_standin_solver)
    if rank == 0:
        want = _standin_solver(P, q, G, h)
        ok = np.allclose(res["all"]["x"], want["x"]) and np.allclose(res["all"]["s"], want["s"]) \
            and len(res["all"]["status"]) == nprob
        ret["ok"] = bool(ok)
        ret["shard0"] = res["x"].shape[0]
    dist.destroy_process_group()


@pytest.mark.parametrize("nprob", [7, 2, 1])
def test_scatter_solve_gather_gloo(nprob):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nprob, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret["ok"]
    assert ret["shard0"] == (nprob + 1) // 2


def test_shard_bounds():
    from cvxopt_b200.batch import shard_bounds
    assert shard_bounds(512, 8) == [(64 * r, 64 * (r + 1)) for r in range(8)]
    b = shard_bounds(10, 4)
    assert b == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_bounds(1, 2) == [(0, 1), (1, 1)]
