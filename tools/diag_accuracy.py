"""Diagnostic: accuracy of the GPU factor/solve vs LAPACK on late-IPM (ill-conditioned) scalings."""
import sys, os
import numpy as np
import scipy.linalg as sla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cvxopt_b200
from problems import dense_qp

n, m = 200, 400
P, q, G, h = dense_qp(n, m, seed=1)
rng = np.random.Generator(np.random.PCG64(3))
for spread in (0, 2, 4, 6, 8):
    d = 10.0 ** rng.uniform(-spread / 2, spread / 2, m)
    # a few tiny slacks only (active constraints), like a late IPM iterate
    di = 1.0 / d
    W = {"d": d, "di": di, "v": [], "beta": [], "r": [], "rti": []}
    dims = {"l": m, "q": [], "s": []}
    fac = cvxopt_b200.kkt_chol(G, dims, H=P)
    solve = fac(W)
    K = P + G.T @ (di[:, None] ** 2 * G)
    L = fac.get_L()
    Lref = np.linalg.cholesky(K)
    bx, bz = rng.standard_normal(n), rng.standard_normal(m)
    x, z = bx.copy(), bz.copy()
    solve(x, None, z)
    # reference through LAPACK on identical K
    rhs = bx + G.T @ (di * (di * bz))
    xr = sla.cho_solve((Lref, True), rhs)
    # extended precision "truth"

    print("spread 1e%d cond(K)=%.1e  |K-LL'|/|K| gpu=%.1e lapack=%.1e   |L-Lref|/|Lref|=%.1e  |x-xr|/|xr|=%.1e  resid gpu=%.1e lapack=%.1e" % (
        spread, np.linalg.cond(K), np.linalg.norm(K - L @ L.T) / np.linalg.norm(K),
        np.linalg.norm(K - Lref @ Lref.T) / np.linalg.norm(K), np.linalg.norm(L - Lref) / np.linalg.norm(Lref),
        np.linalg.norm(x - xr) / np.linalg.norm(xr), np.linalg.norm(K @ x - rhs) / np.linalg.norm(rhs),
        np.linalg.norm(K @ xr - rhs) / np.linalg.norm(rhs)))
    fac.close()
