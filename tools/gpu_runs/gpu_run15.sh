cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle/_ref")
import numpy as np
import cvxopt_b200
from problems import cone_lp
from cvxopt import matrix, solvers
solvers.options["show_progress"] = False
dims = {"l": 0, "q": [64] * 64, "s": []}
c, G, h = cone_lp(2048, dims, seed=11)
f = cvxopt_b200.kkt_chol(matrix(G), dims)
ref = solvers.conelp(matrix(c), matrix(G), matrix(h), dims, kktsolver=lambda W: f(W))
xr, sr, zr = (np.array(ref[k]).ravel() for k in ("x", "s", "z"))
print("ref iterations", ref["iterations"], "dres", ref["dual infeasibility"])
resx0 = max(1.0, np.linalg.norm(c))
def hook(it, x, s, z, tau, kappa, rx, rz):
    xh, sh, zh = x.cpu().numpy() / tau, s.cpu().numpy() / tau, z.cpu().numpy() / tau
    rxh = -G.T @ zh - c
    line = "it %d tau %.6e  dres(dev rx) %.4e  dres(host recompute) %.4e" % (
        it, tau, float(np.linalg.norm(rx.cpu().numpy())) / tau / resx0, np.linalg.norm(rxh) / resx0)
    if it == ref["iterations"]:
        line += "  | vs ref final: dx %.2e ds %.2e dz %.2e" % (
            np.linalg.norm(xh - xr) / np.linalg.norm(xr), np.linalg.norm(sh - sr) / np.linalg.norm(sr),
            np.linalg.norm(zh - zr) / np.linalg.norm(zr))
        rxr = -G.T @ zr - c
        line += "  ref dres recomputed %.4e" % (np.linalg.norm(rxr) / resx0)
    print(line)
r = cvxopt_b200.conelp(c, G, h, dims, debug_hook=hook)
print(r["iterations"])
PY
