cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle/_ref")
import numpy as np
import cvxopt_b200
import cvxopt_b200._misc_solvers as ms
from problems import cone_lp
from cvxopt import matrix, solvers, misc
solvers.options["show_progress"] = False
dims = {"l": 0, "q": [64] * 64, "s": []}
c, G, h = cone_lp(2048, dims, seed=11)
cm, Gm, hm = matrix(c), matrix(G), matrix(h)
f = cvxopt_b200.kkt_chol(Gm, dims)
saved = {n: getattr(misc, n) for n in ms.__all__}
for subset in ([], ["scale"], ["scale2"], ["sprod"], ["sinv"], ["sdot"], ["max_step"], ["pack", "unpack"], list(ms.__all__)):
    for n in subset:
        setattr(misc, n, getattr(ms, n))
    f.reset()
    s = solvers.conelp(cm, Gm, hm, dims, kktsolver=lambda W: f(W))
    for n in subset:
        setattr(misc, n, saved[n])
    print(subset if len(subset) < 5 else "ALL", s["iterations"], "dres %.6e pres %.6e" % (s["dual infeasibility"], s["primal infeasibility"]))
PY
