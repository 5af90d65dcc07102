"""CPU check of the driver logic of cvxopt_b200/conelp.py: the same restatement of coneprog.conelp, with every
device closure replaced by the REFERENCE's own function (cvxopt.misc / misc.kkt_chol from oracle/_ref) acting on CPU
torch tensors.  If this twin reproduces the reference's iteration counts, the restatement of the driver is right and
any difference of the device run is arithmetic.   python tools/conelp_host_twin.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "_ref")):
    sys.path.insert(0, p)
from cvxopt import blas, matrix, misc, solvers  # noqa: E402
from problems import cone_lp  # noqa: E402


def host_backend(torch, dims, G):
    Gm = matrix(G)
    n = G.shape[1]
    W = {}
    state = {}
    factory = misc.kkt_chol(Gm, dims, matrix(0.0, (0, n)))

    def m_of(t):
        return matrix(t.numpy().copy())

    def back(t, m):
        t.copy_(torch.from_numpy(np.array(m).ravel().copy()))

    def set_identity_scaling():
        W.clear()
        W["d"] = matrix(1.0, (dims["l"], 1)); W["di"] = matrix(1.0, (dims["l"], 1))
        W["v"] = [matrix(0.0, (m, 1)) for m in dims["q"]]
        W["beta"] = len(dims["q"]) * [1.0]
        for v in W["v"]:
            v[0] = 1.0
        W["r"] = [matrix(0.0, (m, m)) for m in dims["s"]]
        W["rti"] = [matrix(0.0, (m, m)) for m in dims["s"]]
        for r in W["r"]:
            r[::r.size[0] + 1] = 1.0
        for r in W["rti"]:
            r[::r.size[0] + 1] = 1.0

    def factor():
        state["f"] = factory(W)

    def f3(xx, zz):
        x, z = m_of(xx), m_of(zz)
        state["f"](x, matrix(0.0, (0, 1)), z)
        back(xx, x); back(zz, z)

    def scale(xx, trans="N", inverse="N"):
        x = m_of(xx); misc.scale(x, W, trans=trans, inverse=inverse); back(xx, x)

    def scale2(lm, xx, inverse="N"):
        x = m_of(xx); misc.scale2(m_of(lm), x, dims, inverse=inverse); back(xx, x)

    def sprod(xx, yy):
        x = m_of(xx); misc.sprod(x, m_of(yy), dims); back(xx, x)

    def sinv(xx, yy):
        x = m_of(xx); misc.sinv(x, m_of(yy), dims); back(xx, x)

    def sdot(xx, yy):
        return misc.sdot(m_of(xx), m_of(yy), dims)

    def max_step(xx, sigma=None):
        x = m_of(xx)
        if sigma is not None and sigma.numel():
            sg = m_of(sigma)
            t = misc.max_step(x, dims, sigma=sg)
            back(sigma, sg); back(xx, x)
            return t
        return misc.max_step(x, dims)

    def symm_blocks(xx):
        x = m_of(xx)
        ind = dims["l"] + sum(dims["q"])
        for m in dims["s"]:
            misc.symm(x, m, ind)
            ind += m * m
        back(xx, x)

    def Gf(xx, yy, alpha=1.0, beta=0.0, trans="N"):
        x, y = m_of(xx), m_of(yy)
        misc.sgemv(Gm, x, y, dims, trans=trans, alpha=alpha, beta=beta)
        back(yy, y)

    def compute_scaling(ss, zz, lm):
        lmm = m_of(lm)
        Wn = misc.compute_scaling(m_of(ss), m_of(zz), lmm, dims, mnl=0)
        W.clear(); W.update(Wn)
        back(lm, lmm)

    def update_scaling(lm, dss, dzz):
        lmm, a, b = m_of(lm), m_of(dss), m_of(dzz)
        misc.update_scaling(W, lmm, a, b)
        back(lm, lmm); back(dss, a); back(dzz, b)

    return (set_identity_scaling, factor, f3, scale, scale2, sprod, sinv, sdot, max_step, symm_blocks, Gf, compute_scaling,
            update_scaling)


def twin_conelp(dc, c, G, h, dims, **options):
    """cvxopt_b200.conelp._conelp_core on CPU tensors with the reference's functions as its 13 closures"""
    import torch
    o = dict(dc.DEFAULTS)
    o.update(options)
    G = np.asarray(G, dtype=np.float64)
    dims = {"l": int(dims["l"]), "q": [int(k) for k in dims["q"]], "s": [int(k) for k in dims["s"]]}
    ops = host_backend(torch, dims, G)
    return dc._conelp_core(torch, torch.device("cpu"), np.asarray(c, dtype=np.float64).reshape(-1),
                           np.asarray(h, dtype=np.float64).reshape(-1), G.shape[1], dims, ops, o)


if __name__ == "__main__":
    import importlib
    dc = importlib.import_module("cvxopt_b200.conelp")
    solvers.options["show_progress"] = False
    cases = [({"l": 30, "q": [], "s": []}, 12, 1), ({"l": 0, "q": [16] * 6, "s": []}, 40, 11),
             ({"l": 10, "q": [7, 3], "s": [6, 9]}, 25, 12), ({"l": 0, "q": [], "s": [24]}, 30, 13),
             ({"l": 0, "q": [64] * 16, "s": []}, 512, 11)]
    if len(sys.argv) > 1:
        cases.append(({"l": 0, "q": [64] * 64, "s": []}, 2048, 11))
    for dims, n, seed in cases:
        c, G, h = cone_lp(n, dims, seed)
        want = solvers.conelp(matrix(c), matrix(G), matrix(h), dims, kktsolver="chol")
        got = twin_conelp(dc, c, G, h, dims)
        print(dims if len(str(dims)) < 60 else "cfg-like", "reference %d iterations, twin %d; pobj %.12e vs %.12e; dres %.4e vs %.4e" % (
            want["iterations"], got["iterations"], want["primal objective"], got["primal objective"],
            want["dual infeasibility"], got["dual infeasibility"]))
