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
