"""cProfile of the unmodified reference driver (oracle/_ref solvers.coneqp) with the device kktsolver and the
device G/P operators at BASELINE config 2: where the host-side time per iteration goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import bench  # noqa: E402
import cvxopt_b200  # noqa: E402
from cvxopt import matrix, solvers  # noqa: E402

solvers.options["show_progress"] = False
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = 2 * n
P, q, G, h = bench.make_qp(n, m, 1234)
Pm, qm, Gm, hm = matrix(P), matrix(q), matrix(G), matrix(h)
dims = {"l": m, "q": [], "s": []}
f = cvxopt_b200.kkt_chol(Gm, dims, None, H=Pm)


def Gop(u, v, alpha=1.0, beta=0.0, trans="N"):
    f.G(u, v, alpha, beta, trans)


def Pop(u, v, alpha=1.0, beta=0.0):
    f.P(u, v, alpha, beta)


solvers.coneqp(Pop, qm, Gop, hm, dims, kktsolver=lambda W: f(W))
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
sol = solvers.coneqp(Pop, qm, Gop, hm, dims, kktsolver=lambda W: f(W))
pr.disable()
print("total %.3f s, %d iterations" % (time.perf_counter() - t0, sol["iterations"]))
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)

# ---- is the factor slow because the GPU idles between calls (clock ramp) or because of the call itself? ----
import numpy as np  # noqa: E402
rng = np.random.Generator(np.random.PCG64(1))
d = 10.0 ** rng.uniform(-2, 2, m)
W = {"d": d, "di": 1.0 / d, "v": [], "beta": [], "r": [], "rti": []}
for gap in (0.0, 0.002, 0.005, 0.02):
    ev, wall = [], []
    for _ in range(8):
        t0 = time.perf_counter()
        f(W)
        wall.append((time.perf_counter() - t0) * 1e3)
        ev.append(f.last_ms()[0])
        br = f.last_breakdown()
        if gap:
            time.sleep(gap)
    print("idle gap %5.1f ms: factor wall %.2f ms, CUDA-event %.2f ms (syrk %.2f potrf %.2f scale %.2f)" % (
        gap * 1e3, float(np.median(wall)), float(np.median(ev)), br["syrk_ms"], br["potrf_ms"], br["scale_ms"]))


# ---- per-call GPU-event vs wall time of factor() INSIDE a solver run ----
rec = []


def ks(Wc):
    t0 = time.perf_counter()
    g = f(Wc)
    rec.append(((time.perf_counter() - t0) * 1e3, f.last_ms()[0], f.last_breakdown()))
    return g


solvers.coneqp(Pop, qm, Gop, hm, dims, kktsolver=ks)
for i, (wall, evms, br) in enumerate(rec):
    print("in-solver factor %2d: wall %7.2f ms, CUDA-event %7.2f ms (syrk %.2f potrf %.2f scale %.2f)" % (
        i, wall, evms, br["syrk_ms"], br["potrf_ms"], br["scale_ms"]))
