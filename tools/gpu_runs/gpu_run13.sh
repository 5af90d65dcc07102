cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle/_ref")
import numpy as np
import cvxopt_b200
from problems import cone_lp
from cvxopt import matrix, solvers
dims = {"l": 0, "q": [64] * 64, "s": []}
c, G, h = cone_lp(2048, dims, seed=11)
r = cvxopt_b200.conelp(c, G, h, dims, show_progress=True)
print(r["status"], r["iterations"], r["primal objective"])
f = cvxopt_b200.kkt_chol(matrix(G), dims)
solvers.options["show_progress"] = True
s = solvers.conelp(matrix(c), matrix(G), matrix(h), dims, kktsolver=lambda W: f(W))
print(s["iterations"], s["dual infeasibility"])
# smaller instance of the same family for a quick A/B
dims = {"l": 0, "q": [64] * 8, "s": []}
c, G, h = cone_lp(256, dims, seed=11)
r = cvxopt_b200.conelp(c, G, h, dims, show_progress=True)
s = solvers.conelp(matrix(c), matrix(G), matrix(h), dims, kktsolver="chol")
print(r["iterations"], s["iterations"])
PY
