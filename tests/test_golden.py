"""Golden vectors generated from the reference itself (tests/golden/make_golden.py, committed as
tests/golden/kkt_golden.npz).  CPU: the numpy oracle reproduces them.  GPU: the CUDA path through
the C-ABI reproduces them — these run on the GPU box where /root/reference does not exist."""
import os

import numpy as np
import pytest

import kkt_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "kkt_golden.npz"))
CASES = {
    "l": {"l": 9, "q": [], "s": []},
    "q": {"l": 0, "q": [5, 3, 8], "s": []},
    "s": {"l": 0, "q": [], "s": [4, 6]},
    "mixed": {"l": 4, "q": [6], "s": [3, 5]},
}


def scaling(name, dims):
    return {"d": GOLD[name + "/d"], "di": GOLD[name + "/di"], "beta": list(GOLD[name + "/beta"]),
            "v": [GOLD["%s/v%d" % (name, k)] for k in range(len(dims["q"]))],
            "r": [np.asfortranarray(GOLD["%s/r%d" % (name, k)]) for k in range(len(dims["s"]))],
            "rti": [np.asfortranarray(GOLD["%s/rti%d" % (name, k)]) for k in range(len(dims["s"]))]}


def lower_mask(dims):
    K = dims["l"] + sum(dims["q"]) + sum(k * k for k in dims["s"])
    mask = np.ones(K, bool)
    off = dims["l"] + sum(dims["q"])
    for k in dims["s"]:
        M = np.ones((k, k), bool)
        M[np.triu_indices(k, 1)] = False
        mask[off:off + k * k] = M.reshape(-1, order="F")
        off += k * k
    return mask


def close(a, b, tol=1e-10):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) <= tol * max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(name):
    dims = CASES[name]
    W, mask = scaling(name, dims), lower_mask(dims)
    lm = np.zeros_like(GOLD[name + "/lmbda"])
    Wc = ko.compute_scaling(GOLD[name + "/s"].copy(), GOLD[name + "/z"].copy(), lm, dims)
    assert close(lm, GOLD[name + "/lmbda"], 1e-11) and close(Wc["d"], W["d"], 1e-13)
    for tag, A in (("p0", None), ("p2", GOLD[name + "/A"])):
        f = ko.KktChol(GOLD[name + "/G"], dims, A).factor(W, GOLD[name + "/H"])
        x, y, z = (GOLD["%s/%s/%s" % (name, tag, k)].copy() for k in ("bx", "by", "bz"))
        f(x, y, z)
        assert close(x, GOLD["%s/%s/ux" % (name, tag)], 1e-9)
        assert close(y, GOLD["%s/%s/uy" % (name, tag)], 1e-9)
        assert close(z[mask], GOLD["%s/%s/uz" % (name, tag)][mask], 1e-9)
    for tr in "NT":
        for inv in "NI":
            X = np.asfortranarray(GOLD[name + "/scale_in"].copy())
            ko.scale(X, W, tr, inv)
            assert close(X[mask], GOLD["%s/scale_%s%s" % (name, tr, inv)][mask], 1e-11)
    lmg = GOLD[name + "/lmbda"]
    for inv in "NI":
        v = GOLD[name + "/vec"].copy()
        ko.scale2(lmg, v, dims, inverse=inv)
        assert close(v, GOLD["%s/scale2_%s" % (name, inv)], 1e-12)
    v = GOLD[name + "/vec"].copy(); ko.sprod(v, GOLD[name + "/vec2"], dims)
    assert close(v[mask], GOLD[name + "/sprod"][mask], 1e-12)
    v = GOLD[name + "/vec"].copy(); ko.sinv(v, lmg, dims)
    assert close(v[mask], GOLD[name + "/sinv"][mask], 1e-12)
    assert abs(ko.sdot(GOLD[name + "/vec"], GOLD[name + "/vec2"], dims) - GOLD[name + "/sdot"][0]) < 1e-11
    # max_step on the reference reads only the lower triangles too
    assert abs(ko.max_step(GOLD[name + "/vec"].copy(), dims) - GOLD[name + "/max_step"][0]) < 1e-9
    pk = np.zeros_like(GOLD[name + "/pack"]); ko.pack(GOLD[name + "/vec"], pk, dims)
    assert np.array_equal(pk, GOLD[name + "/pack"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_reproduces_golden(name):
    import cvxopt_b200
    from cvxopt_b200 import misc_solvers as ms
    dims = CASES[name]
    W, mask = scaling(name, dims), lower_mask(dims)
    G, H = np.asfortranarray(GOLD[name + "/G"]), np.asfortranarray(GOLD[name + "/H"])
    for tag, A in (("p0", None), ("p2", np.asfortranarray(GOLD[name + "/A"]))):
        fac = cvxopt_b200.kkt_chol(G, dims, A)
        solve = fac(W, H)
        x, y, z = (GOLD["%s/%s/%s" % (name, tag, k)].copy() for k in ("bx", "by", "bz"))
        solve(x, y if A is not None else None, z)
        assert close(x, GOLD["%s/%s/ux" % (name, tag)], 1e-9)
        if A is not None:
            assert close(y, GOLD["%s/%s/uy" % (name, tag)], 1e-9)
        assert close(z[mask], GOLD["%s/%s/uz" % (name, tag)][mask], 1e-9)
        fac.close()
    for tr in "NT":
        for inv in "NI":
            X = np.asfortranarray(GOLD[name + "/scale_in"].copy())
            ms.scale(X, W, tr, inv)
            assert close(X[mask], GOLD["%s/scale_%s%s" % (name, tr, inv)][mask], 1e-11)
    lmg = GOLD[name + "/lmbda"]
    for inv in "NI":
        v = GOLD[name + "/vec"].copy(); ms.scale2(lmg, v, dims, inverse=inv)
        assert close(v, GOLD["%s/scale2_%s" % (name, inv)], 1e-12)
    v = GOLD[name + "/vec"].copy(); ms.sprod(v, GOLD[name + "/vec2"].copy(), dims)
    assert close(v[mask], GOLD[name + "/sprod"][mask], 1e-12)
    v = GOLD[name + "/vec"].copy(); ms.sprod(v, lmg, dims, diag="D")
    assert close(v[mask], GOLD[name + "/sprod_D"][mask], 1e-12)
    v = GOLD[name + "/vec"].copy(); ms.sinv(v, lmg, dims)
    assert close(v[mask], GOLD[name + "/sinv"][mask], 1e-12)
    assert abs(ms.sdot(GOLD[name + "/vec"].copy(), GOLD[name + "/vec2"].copy(), dims) - GOLD[name + "/sdot"][0]) < 1e-11
    pk = np.zeros_like(GOLD[name + "/pack"]); ms.pack(GOLD[name + "/vec"].copy(), pk, dims)
    assert np.array_equal(pk, GOLD[name + "/pack"])
    assert abs(ms.max_step(GOLD[name + "/vec"].copy(), dims) - GOLD[name + "/max_step"][0]) < 1e-11


@pytest.mark.gpu
def test_device_ipm_reproduces_golden_qp_run():
    """cvxopt_b200.qp_batch vs the reference's solvers.coneqp(kktsolver='chol') run stored in the fixtures."""
    import cvxopt_b200
    from problems import dense_qp
    P, q, G, h = dense_qp(40, 90, seed=7)
    r = cvxopt_b200.qp_batch(P[None], q[None], G[None], h[None])
    assert int(r["iterations"][0]) == int(GOLD["qp40/iterations"][0])
    assert abs(r["primal objective"][0] - GOLD["qp40/pobj"][0]) <= 1e-8 * abs(GOLD["qp40/pobj"][0])
    assert close(r["x"][0], GOLD["qp40/x"], 1e-6)
