"""config-4-style batch on one GPU: time against the number of concurrently solved sub-batches (QPBatchGroup)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cvxopt_b200  # noqa: E402

n, m = 512, 1024
for B in (512, 64):
    P, q, G, h = (np.empty((B, n, n)), np.empty((B, n)), np.empty((B, m, n)), np.empty((B, m)))
    for k in range(B):
        P[k], q[k], G[k], h[k] = bench.make_qp(n, m, k * (512 // B))
    for nsub in (1, 2, 4, 8):
        g = cvxopt_b200.QPBatchGroup(B, n, m, 0, nsub)
        g.load(P, q, G, h)
        g.solve()
        t0 = time.perf_counter()
        g.solve()
        wall = (time.perf_counter() - t0) * 1e3
        r, st = g.results(), g.stats()
        g.close()
        print("B=%3d nsub=%d: wall %.1f ms, max kernel-event %.1f ms, iterations total %d, lockstep per sub-batch %s" % (
            B, nsub, wall, st["solve_ms"], int(r["iterations"].sum()), st["lockstep_iterations_per_subbatch"]), flush=True)
