"""Batch of independent dense QPs solved in lock-step on the device (BASELINE config 4).

    minimize 1/2 x'P x + q'x   subject to   G x <= h        (one 'l' cone of m rows, no A)

The per-problem algorithm is coneprog.coneqp restricted to dims={'l': m} (reference
src/python/coneprog.py:1998-2547) — same start, stopping rule, Mehrotra steps — so every
problem converges in the same number of iterations as `solvers.qp(P, q, G, h)` does.
The reference has no batch API; its counterpart is a Python loop over `solvers.qp`.
"""
import ctypes as C

import numpy as np

from . import _lib

STATUS = {0: "running", 1: "optimal", 2: "unknown", 3: "unknown"}
DEFAULTS = dict(maxiters=100, abstol=1e-7, reltol=1e-6, feastol=1e-7)   # coneprog.py:436-456


def _stack(P, q, G, h):
    """-> contiguous (B,n,n), (B,n), (B,n,m) [= m x n column-major per problem], (B,m)"""
    P = np.ascontiguousarray(np.asarray(P, dtype=np.float64))
    q = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    G = np.asarray(G, dtype=np.float64)
    h = np.ascontiguousarray(np.asarray(h, dtype=np.float64))
    if P.ndim != 3 or P.shape[1] != P.shape[2]:
        raise TypeError("P must have shape (B, n, n)")
    B, n = P.shape[0], P.shape[1]
    if q.shape != (B, n):
        raise TypeError("q must have shape (B, n)")
    if G.ndim != 3 or G.shape[0] != B or G.shape[2] != n:
        raise TypeError("G must have shape (B, m, n)")
    m = G.shape[1]
    if h.shape != (B, m):
        raise TypeError("h must have shape (B, m)")
    # P is symmetric: its row-major image equals its column-major image as far as tril goes
    Pcm = np.ascontiguousarray(np.transpose(P, (0, 2, 1)))
    Gcm = np.ascontiguousarray(np.transpose(G, (0, 2, 1)))      # (B, n, m): column-major m x n
    return Pcm, q, Gcm, h, B, n, m


class QPBatch:
    def __init__(self, nprob, n, m, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.B, self.n, self.m = int(nprob), int(n), int(m)
        _lib.check(self._lib.cvxb_batch_create(C.byref(self._h), self.B, self.n, self.m, device), "batch")

    def load(self, P, q, G, h):
        Pcm, q, Gcm, h, B, n, m = _stack(P, q, G, h)
        if (B, n, m) != (self.B, self.n, self.m):
            raise TypeError("problem shapes do not match the batch")
        rc = self._lib.cvxb_batch_load(self._h, Pcm.ctypes.data, q.ctypes.data, Gcm.ctypes.data,
                                       h.ctypes.data, _lib.HOST)
        _lib.check(rc, "batch_load")

    def load_ptr(self, P, q, G, h, space=_lib.DEVICE):
        """raw addresses of already laid-out buffers (device-resident callers)"""
        _lib.check(self._lib.cvxb_batch_load(self._h, P, q, G, h, space), "batch_load")

    def solve(self, **options):
        o = dict(DEFAULTS)
        o.update(options)
        rc = self._lib.cvxb_batch_solve(self._h, int(o["maxiters"]), float(o["abstol"]),
                                        float(o["reltol"]), float(o["feastol"]))
        if rc == _lib.E_ARG and "Rank(" in _lib.last_error():
            raise ValueError(_lib.last_error())       # coneprog.py:2065-2067
        _lib.check(rc, "batch_solve")

    def results(self):
        B, n, m = self.B, self.n, self.m
        x, s, z = np.zeros((B, n)), np.zeros((B, m)), np.zeros((B, m))
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        pobj, dobj = np.zeros(B), np.zeros(B)
        rc = self._lib.cvxb_batch_results(self._h, x.ctypes.data, s.ctypes.data, z.ctypes.data,
                                          status.ctypes.data, iters.ctypes.data, pobj.ctypes.data,
                                          dobj.ctypes.data, _lib.HOST)
        _lib.check(rc, "batch_results")
        return {"x": x, "s": s, "z": z, "status": [STATUS[int(k)] for k in status],
                "status_code": status, "iterations": iters, "primal objective": pobj,
                "dual objective": dobj}

    def stats(self):
        ms, it = C.c_double(), C.c_int()
        self._lib.cvxb_batch_stats(self._h, C.byref(ms), C.byref(it))
        return {"solve_ms": ms.value, "lockstep_iterations": it.value}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cvxb_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def qp_batch(P, q, G, h, device=0, **options):
    """Solve B independent dense QPs on one GPU.  P (B,n,n), q (B,n), G (B,m,n), h (B,m)."""
    P = np.asarray(P)
    G = np.asarray(G)
    b = QPBatch(P.shape[0], P.shape[1], G.shape[1], device)
    try:
        b.load(P, q, G, h)
        b.solve(**options)
        out = b.results()
        out.update(b.stats())
        return out
    finally:
        b.close()


# ---------------------------------------------------------------------------------------
# multi-GPU: problems are independent -> shard them across ranks, no data-path collective.
# One scatter of (P, q, G, h) from rank 0, one gather of (x, s, z, status, iters, objectives).

def shard_bounds(nprob, world):
    """contiguous block partition: rank r owns [lo, hi)"""
    base, extra = divmod(nprob, world)
    bounds, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def qp_batch_distributed(P, q, G, h, solver=None, group=None, **options):
    """Rank 0 passes the full batch (other ranks pass None); every rank returns its shard's
    results and rank 0 additionally gets the gathered batch under key 'all'.

    Collectives: scatter_object-free — plain tensor scatter/gather over the process group
    (NCCL over NVLink on GPUs; gloo in the CPU tests, where `solver` is a stand-in)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta = torch.zeros(3, dtype=torch.int64, device=dev)
    if rank == 0:
        P, q, G, h = (np.asarray(a, dtype=np.float64) for a in (P, q, G, h))
        meta = torch.tensor([P.shape[0], P.shape[1], G.shape[1]], dtype=torch.int64, device=dev)
    dist.broadcast(meta, 0, group=group)
    Btot, n, m = (int(v) for v in meta.tolist())
    bounds = shard_bounds(Btot, world)
    lo, hi = bounds[rank]
    cap = max(b[1] - b[0] for b in bounds)            # equal-sized scatter slots (padded)

    def scatter(full, tail):
        out = torch.zeros((cap,) + tail, dtype=torch.float64, device=dev)
        if rank == 0:
            chunks = []
            for (a, b) in bounds:
                c = torch.zeros((cap,) + tail, dtype=torch.float64)
                c[: b - a] = torch.from_numpy(np.ascontiguousarray(full[a:b]))
                chunks.append(c.to(dev))
            dist.scatter(out, chunks, src=0, group=group)
        else:
            dist.scatter(out, None, src=0, group=group)
        return out[: hi - lo].cpu().numpy()
    Ps, qs, Gs, hs = scatter(P, (n, n)), scatter(q, (n,)), scatter(G, (m, n)), scatter(h, (m,))
    if solver is None:
        local_dev = torch.cuda.current_device() if backend == "nccl" else 0
        solver = lambda a, b_, c, d: qp_batch(a, b_, c, d, device=local_dev, **options)   # noqa: E731
    res = solver(Ps, qs, Gs, hs) if hi > lo else {
        "x": np.zeros((0, n)), "s": np.zeros((0, m)), "z": np.zeros((0, m)),
        "status_code": np.zeros(0, np.int32), "iterations": np.zeros(0, np.int32),
        "primal objective": np.zeros(0), "dual objective": np.zeros(0)}

    def gather(local, tail, dtype=torch.float64):
        buf = torch.zeros((cap,) + tail, dtype=dtype, device=dev)
        buf[: hi - lo] = torch.from_numpy(np.ascontiguousarray(local)).to(dtype).to(dev)
        if rank == 0:
            outs = [torch.zeros_like(buf) for _ in range(world)]
            dist.gather(buf, outs, dst=0, group=group)
            return np.concatenate([o[: b - a].cpu().numpy() for o, (a, b) in zip(outs, bounds)])
        dist.gather(buf, None, dst=0, group=group)
        return None
    full = {"x": gather(res["x"], (n,)), "s": gather(res["s"], (m,)), "z": gather(res["z"], (m,)),
            "status_code": gather(res["status_code"], (), torch.int64),
            "iterations": gather(res["iterations"], (), torch.int64),
            "primal objective": gather(res["primal objective"], ()),
            "dual objective": gather(res["dual objective"], ())}
    if rank == 0:
        full["status"] = [STATUS[int(k)] for k in full["status_code"]]
        res = dict(res)
        res["all"] = full
    return res
