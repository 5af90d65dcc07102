"""End-to-end timing of the UNMODIFIED CVXOPT driver (oracle/_ref acts as the host application)
with (a) its own kktsolver='chol', (b) cvxopt_b200's kktsolver, (c) cvxopt_b200's kktsolver plus
device-resident G/P operators.  Usage: python tests/e2e_cvxopt_timing.py [n]   (needs a GPU)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "_ref")):
    sys.path.insert(0, p)
from problems import dense_qp


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    m = 2 * n
    import cvxopt_b200
    from cvxopt import matrix, solvers
    solvers.options["show_progress"] = False
    P, q, G, h = dense_qp(n, m, seed=1234)
    Pm, qm, Gm, hm = matrix(P), matrix(q), matrix(G), matrix(h)
    dims = {"l": m, "q": [], "s": []}
    out = {"n": n, "m": m, "host_cores": os.cpu_count()}
    t0 = time.perf_counter()
    r = solvers.coneqp(Pm, qm, Gm, hm, dims, kktsolver="chol")
    out["reference_chol_s"] = time.perf_counter() - t0
    out["iterations"] = r["iterations"]
    f = cvxopt_b200.kkt_chol(Gm, dims, None, H=Pm)
    t0 = time.perf_counter()
    a = solvers.coneqp(Pm, qm, Gm, hm, dims, kktsolver=lambda W: f(W))
    out["b200_kktsolver_s"] = time.perf_counter() - t0
    Gop = lambda x, y, alpha=1.0, beta=0.0, trans="N": f.G(x, y, alpha, beta, trans)   # noqa: E731
    Pop = lambda x, y, alpha=1.0, beta=0.0: f.P(x, y, alpha, beta)                     # noqa: E731
    t0 = time.perf_counter()
    b = solvers.coneqp(Pop, qm, Gop, hm, dims, kktsolver=lambda W: f(W))
    out["b200_kktsolver_and_operators_s"] = time.perf_counter() - t0
    assert a["iterations"] == b["iterations"] == r["iterations"]
    assert abs(a["primal objective"] - r["primal objective"]) <= 1e-8 * abs(r["primal objective"])
    assert abs(b["primal objective"] - r["primal objective"]) <= 1e-8 * abs(r["primal objective"])
    bt = cvxopt_b200.qp_batch(P[None], q[None], G[None], h[None])
    out["b200_device_ipm_s"] = bt["solve_ms"] * 1e-3
    assert int(bt["iterations"][0]) == r["iterations"]
    for k in ("b200_kktsolver_s", "b200_kktsolver_and_operators_s", "b200_device_ipm_s"):
        out["speedup_" + k[:-2]] = out["reference_chol_s"] / out[k]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
