"""Whole-solve results of the REFERENCE (oracle/_ref = cvxopt built from /root/reference) on the BASELINE
configurations at their stated sizes (SURVEY.md §8d inputs), written to tests/golden/config_runs.json:

  cfg2  solvers.coneqp  dense QP n=4096 m=8192 seed 1234, kktsolver='chol'
  cfg3  solvers.conelp  SOCP n=2048, 64 cones of 64, seed 11, kktsolver='chol'
  cfg4  solvers.qp      512 dense QPs n=512 m=1024, seeds 0..511: iterations + objectives of every one
  cfg5  solvers.conelp  SDP one 512x512 block, n=512, seed 11, kktsolver='chol'

Run where the reference exists (takes ~15 min on 8 cores):   python tests/golden/make_config_golden.py [cfg ...]
The committed JSON is what the -m gpu tests compare against on the GPU box (no /root/reference there)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cvxopt import matrix, solvers  # noqa: E402
from problems import cone_lp, dense_qp  # noqa: E402

solvers.options["show_progress"] = False
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_runs.json")


def summary(sol, t):
    return {"status": sol["status"], "iterations": int(sol["iterations"]),
            "primal objective": float(sol["primal objective"]), "dual objective": float(sol["dual objective"]),
            "gap": float(sol["gap"]), "seconds": round(t, 2)}


def main():
    want = sys.argv[1:] or ["cfg2", "cfg3", "cfg4", "cfg5"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    if "cfg2" in want:
        P, q, G, h = dense_qp(4096, 8192, seed=1234)
        t = time.time()
        sol = solvers.coneqp(matrix(P), matrix(q), matrix(G), matrix(h), {"l": 8192, "q": [], "s": []}, kktsolver="chol")
        res["cfg2"] = summary(sol, time.time() - t)
        print("cfg2", res["cfg2"], flush=True)
    if "cfg3" in want:
        dims = {"l": 0, "q": [64] * 64, "s": []}
        c, G, h = cone_lp(2048, dims, seed=11)
        t = time.time()
        sol = solvers.conelp(matrix(c), matrix(G), matrix(h), dims, kktsolver="chol")
        res["cfg3"] = summary(sol, time.time() - t)
        print("cfg3", res["cfg3"], flush=True)
    if "cfg4" in want:
        its, pobj, dobj, st = [], [], [], []
        t = time.time()
        for k in range(512):
            P, q, G, h = dense_qp(512, 1024, seed=k)
            sol = solvers.qp(matrix(P), matrix(q), matrix(G), matrix(h), kktsolver="chol")
            its.append(int(sol["iterations"])); pobj.append(float(sol["primal objective"]))
            dobj.append(float(sol["dual objective"])); st.append(sol["status"])
        res["cfg4"] = {"iterations": its, "primal objective": pobj, "dual objective": dobj,
                       "all_optimal": all(s == "optimal" for s in st), "iterations_total": int(sum(its)),
                       "seconds": round(time.time() - t, 2)}
        print("cfg4 total iterations", sum(its), "all optimal", res["cfg4"]["all_optimal"], flush=True)
    if "cfg5" in want:
        dims = {"l": 0, "q": [], "s": [512]}
        c, G, h = cone_lp(512, dims, seed=11)
        t = time.time()
        sol = solvers.conelp(matrix(c), matrix(G), matrix(h), dims, kktsolver="chol")
        res["cfg5"] = summary(sol, time.time() - t)
        print("cfg5", res["cfg5"], flush=True)
    json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
