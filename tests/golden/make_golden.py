"""Generates tests/golden/kkt_golden.npz from the REFERENCE itself (oracle/_ref = cvxopt built from
/root/reference by oracle/build_ref.sh): inputs and outputs of misc.kkt_chol factor/solve,
misc_solvers.scale / pack / scale2 / sprod / sinv / sdot / max_step, misc.compute_scaling, and the
iteration counts / objectives of small solver runs.  Run where the reference exists:

    python tests/golden/make_golden.py

The committed .npz is what the tests use on the GPU box (where /root/reference is absent)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cvxopt import matrix, misc, solvers  # noqa: E402
from problems import cone_dim, cone_lp, cone_point, dense_qp  # noqa: E402

solvers.options["show_progress"] = False
CASES = {
    "l": {"l": 9, "q": [], "s": []},
    "q": {"l": 0, "q": [5, 3, 8], "s": []},
    "s": {"l": 0, "q": [], "s": [4, 6]},
    "mixed": {"l": 4, "q": [6], "s": [3, 5]},
}


def arr(m):
    return np.array(m, dtype=float)


def main():
    out = {}
    for name, dims in CASES.items():
        rng = np.random.Generator(np.random.PCG64(len(name) + 100))
        K = cone_dim(dims)
        n, p = 5, 2
        nl = dims["l"] + sum(dims["q"]) + sum(dims["s"])
        s, z = cone_point(dims, rng), cone_point(dims, rng)
        lm = matrix(0.0, (nl, 1))
        W = misc.compute_scaling(matrix(s), matrix(z), lm, dims)
        out[name + "/s"], out[name + "/z"], out[name + "/lmbda"] = s, z, arr(lm).ravel()
        out[name + "/d"], out[name + "/di"] = arr(W["d"]).ravel(), arr(W["di"]).ravel()
        out[name + "/beta"] = np.array(list(W["beta"]), dtype=float)
        for k, v in enumerate(W["v"]):
            out["%s/v%d" % (name, k)] = arr(v).ravel()
        for k, (r, rti) in enumerate(zip(W["r"], W["rti"])):
            out["%s/r%d" % (name, k)], out["%s/rti%d" % (name, k)] = arr(r), arr(rti)
        G = rng.standard_normal((K, n))
        A = rng.standard_normal((p, n))
        B = rng.standard_normal((n, n))
        H = B @ B.T + np.eye(n)
        out[name + "/G"], out[name + "/A"], out[name + "/H"] = G, A, H
        for tag, Am in (("p0", matrix(0.0, (0, n))), ("p2", matrix(A))):
            f = misc.kkt_chol(matrix(G), dims, Am)(W, matrix(H))
            bx, by, bz = rng.standard_normal(n), rng.standard_normal(Am.size[0]), rng.standard_normal(K)
            x, y, zz = matrix(bx), matrix(by), matrix(bz)
            f(x, y, zz)
            for key, val in (("bx", bx), ("by", by), ("bz", bz), ("ux", arr(x).ravel()),
                             ("uy", arr(y).ravel()), ("uz", arr(zz).ravel())):
                out["%s/%s/%s" % (name, tag, key)] = val
        X = rng.standard_normal((K, 3))
        out[name + "/scale_in"] = X
        for tr in "NT":
            for inv in "NI":
                Xm = matrix(X)
                misc.scale(Xm, W, trans=tr, inverse=inv)
                out["%s/scale_%s%s" % (name, tr, inv)] = arr(Xm)
        v = cone_point(dims, rng)
        out[name + "/vec"] = v
        for inv in "NI":
            vm = matrix(v)
            misc.scale2(lm, vm, dims, inverse=inv)
            out["%s/scale2_%s" % (name, inv)] = arr(vm).ravel()
        w = cone_point(dims, rng)
        out[name + "/vec2"] = w
        vm = matrix(v); misc.sprod(vm, matrix(w), dims); out[name + "/sprod"] = arr(vm).ravel()
        vm = matrix(v); misc.sprod(vm, lm, dims, diag="D"); out[name + "/sprod_D"] = arr(vm).ravel()
        vm = matrix(v); misc.sinv(vm, lm, dims); out[name + "/sinv"] = arr(vm).ravel()
        out[name + "/sdot"] = np.array([misc.sdot(matrix(v), matrix(w), dims)])
        out[name + "/max_step"] = np.array([misc.max_step(matrix(v), dims)])
        _, _, _, cp = (0, 0, 0, dims["l"] + sum(dims["q"]) + sum(k * (k + 1) // 2 for k in dims["s"]))
        pk = matrix(0.0, (cp, 1)); misc.pack(matrix(v), pk, dims); out[name + "/pack"] = arr(pk).ravel()
    # whole-solver runs (kktsolver='chol'): iteration counts and objectives
    P, q, G, h = dense_qp(40, 90, seed=7)
    r = solvers.coneqp(matrix(P), matrix(q), matrix(G), matrix(h), {"l": 90, "q": [], "s": []}, kktsolver="chol")
    out["qp40/iterations"] = np.array([r["iterations"]])
    out["qp40/pobj"] = np.array([r["primal objective"]]); out["qp40/x"] = arr(r["x"]).ravel()
    dims = {"l": 6, "q": [5, 4], "s": [4]}
    c, G2, h2 = cone_lp(12, dims, seed=3)
    r = solvers.conelp(matrix(c), matrix(G2), matrix(h2), dims, kktsolver="chol")
    out["clp12/iterations"] = np.array([r["iterations"]])
    out["clp12/pobj"] = np.array([r["primal objective"]]); out["clp12/x"] = arr(r["x"]).ravel()
    path = os.path.join(ROOT, "tests", "golden", "kkt_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
