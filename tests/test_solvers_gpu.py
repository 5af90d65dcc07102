"""Drop-in tests: the UNMODIFIED reference drivers (cvxopt.solvers.* from oracle/_ref, acting as
the host application) run with cvxopt_b200's kktsolver and must reproduce the reference's own
`kktsolver='chol'` run: same iteration count, objectives to rtol 1e-8 (north_star), and every
search direction to 1e-10 while the KKT system is reasonably conditioned."""
import numpy as np
import pytest

from problems import cone_lp, dense_qp

pytestmark = pytest.mark.gpu


EPS = 2.2e-16


def _spread(W):
    """conditioning proxy of the scaled system: (max/min of the scaling)^2 over all cones"""
    vals = [np.array(W["di"]).ravel()]
    # W_q = beta (2vv' - J) has eigenvalues beta (v0 + |v1|)^{+-2}   (v'Jv = 1)
    for b, v in zip(W["beta"], W["v"]):
        v = np.array(v).ravel()
        t = (abs(v[0]) + np.linalg.norm(v[1:])) ** 2
        vals.append(np.array([1.0 / (b * t), t / b]))
    for r in W["rti"]:
        sv = np.linalg.svd(np.array(r), compute_uv=False)
        vals.append(sv)
    v = np.concatenate([x for x in vals if x.size]) if any(x.size for x in vals) else np.ones(1)
    return float((v.max() / v.min()) ** 2)


def _paired(ref_factory, gpu_factory, log):
    """kktsolver that runs BOTH solvers on identical inputs at every call, logs the relative
    difference of the search direction, and hands the reference's result back to the driver."""
    def kktsolver(W):
        fr, fg = ref_factory(W), gpu_factory(W)
        sp = _spread(W)

        def g(x, y, z):
            from cvxopt import matrix
            xg, zg = matrix(x), matrix(z)          # copies of the right-hand side
            fr(x, y, z)
            fg(xg, y, zg)
            xa, xb = np.array(xg).ravel(), np.array(x).ravel()
            log.append((np.linalg.norm(xa - xb) / max(np.linalg.norm(xb), 1e-300), sp))
        return g
    return kktsolver


def _check_log(log):
    """north_star: 1e-10 on the search direction.  Two backward-stable Cholesky solves of the same
    system agree to ~cond*eps, so the 1e-10 bar is enforced while the scaling spread keeps
    cond*eps below it and a cond-proportional bar (the reference's own forward-error scale,
    see DESIGN.md §parity) beyond."""
    assert log
    for err, sp in log:
        assert err < max(1e-10, 200 * EPS * sp), (err, sp, [("%.1e" % e, "%.1e" % c) for e, c in log])
    assert any(sp < 1e4 for _, sp in log)


def test_qp_config1_matches_reference_chol(ref):
    """BASELINE config 1: dense QP n=200 m=400."""
    import cvxopt_b200
    from cvxopt import matrix, misc, solvers
    P, q, G, h = dense_qp(200, 400, seed=1)
    Pm, qm, Gm, hm = matrix(P), matrix(q), matrix(G), matrix(h)
    dims = {"l": 400, "q": [], "s": []}
    log = []
    fref = misc.kkt_chol(Gm, dims, matrix(0.0, (0, 200)))
    fgpu = cvxopt_b200.kkt_chol(Gm, dims, None, H=Pm)
    sol_ref = solvers.coneqp(Pm, qm, Gm, hm, dims, kktsolver="chol")
    sol_gpu = solvers.coneqp(Pm, qm, Gm, hm, dims, kktsolver=lambda W: fgpu(W))
    solvers.coneqp(Pm, qm, Gm, hm, dims, kktsolver=_paired(lambda W: fref(W, Pm), lambda W: fgpu(W), log))
    assert sol_ref["status"] == sol_gpu["status"] == "optimal"
    assert sol_ref["iterations"] == sol_gpu["iterations"]
    np.testing.assert_allclose(sol_gpu["primal objective"], sol_ref["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(sol_gpu["dual objective"], sol_ref["dual objective"], rtol=1e-8)
    np.testing.assert_allclose(np.array(sol_gpu["x"]), np.array(sol_ref["x"]), rtol=1e-7, atol=1e-9)
    _check_log(log)


def test_solvers_qp_entry_point(ref):
    """solvers.qp(P, q, G, h, kktsolver=...) — the call a CVXOPT user makes."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    P, q, G, h = dense_qp(150, 321, seed=3)
    Pm, qm, Gm, hm = matrix(P), matrix(q), matrix(G), matrix(h)
    f = cvxopt_b200.kkt_chol(Gm, {"l": 321, "q": [], "s": []}, None, H=Pm)
    a = solvers.qp(Pm, qm, Gm, hm, kktsolver=lambda W: f(W))
    b = solvers.qp(Pm, qm, Gm, hm, kktsolver="chol")
    assert a["status"] == b["status"] == "optimal" and a["iterations"] == b["iterations"]
    np.testing.assert_allclose(a["primal objective"], b["primal objective"], rtol=1e-8)


@pytest.mark.parametrize("dims,n,seed", [
    ({"l": 0, "q": [16] * 6, "s": []}, 40, 11),          # SOCP (config 3 in miniature)
    ({"l": 10, "q": [7, 3], "s": [6, 9]}, 25, 12),       # mixed
    ({"l": 0, "q": [], "s": [24]}, 30, 13),              # SDP (config 5 in miniature)
])
def test_conelp_matches_reference_chol(ref, dims, n, seed):
    import cvxopt_b200
    from cvxopt import matrix, misc, solvers
    c, G, h = cone_lp(n, dims, seed)
    cm, Gm, hm = matrix(c), matrix(G), matrix(h)
    log = []
    fref = misc.kkt_chol(Gm, dims, matrix(0.0, (0, n)))
    fgpu = cvxopt_b200.kkt_chol(Gm, dims)
    sol_ref = solvers.conelp(cm, Gm, hm, dims, kktsolver="chol")
    sol_gpu = solvers.conelp(cm, Gm, hm, dims, kktsolver=lambda W: fgpu(W))
    solvers.conelp(cm, Gm, hm, dims, kktsolver=_paired(lambda W: fref(W), lambda W: fgpu(W), log))
    assert sol_ref["status"] == sol_gpu["status"] == "optimal"
    assert sol_ref["iterations"] == sol_gpu["iterations"]
    np.testing.assert_allclose(sol_gpu["primal objective"], sol_ref["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(sol_gpu["dual objective"], sol_ref["dual objective"], rtol=1e-8)
    _check_log(log)


def test_socp_sdp_wrappers(ref):
    """solvers.socp / solvers.sdp stack their blocks and dispatch to conelp (coneprog.py:3538, 4126)."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    # chap8 socp example (reference examples/doc/chap8/socp.py)
    c = matrix([-2., 1., 5.])
    Gq = [matrix([[12., 13., 12.], [6., -3., -12.], [-5., -5., 6.]])]
    Gq += [matrix([[3., 3., -1., 1.], [-6., -6., -9., 19.], [10., -2., -2., -3.]])]
    hq = [matrix([-12., -3., -2.]), matrix([27., 0., 3., -42.])]
    dims = {"l": 0, "q": [3, 4], "s": []}
    Gstack = matrix([Gq[0], Gq[1]])
    f = cvxopt_b200.kkt_chol(Gstack, dims)
    a = solvers.socp(c, Gq=Gq, hq=hq, kktsolver=lambda W: f(W))
    b = solvers.socp(c, Gq=Gq, hq=hq, kktsolver="chol")
    assert a["status"] == b["status"] == "optimal" and a["iterations"] == b["iterations"]
    np.testing.assert_allclose(np.array(a["x"]), np.array(b["x"]), rtol=1e-7, atol=1e-9)


def test_device_operators_G_P(ref):
    """function-valued G / P (coneprog.py:1682-1711) served by the resident copies."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    P, q, G, h = dense_qp(120, 260, seed=9)
    Pm, qm, Gm, hm = matrix(P), matrix(q), matrix(G), matrix(h)
    dims = {"l": 260, "q": [], "s": []}
    f = cvxopt_b200.kkt_chol(Gm, dims, None, H=Pm)

    def Gop(x, y, alpha=1.0, beta=0.0, trans="N"):
        f.G(x, y, alpha, beta, trans)

    def Pop(x, y, alpha=1.0, beta=0.0):
        f.P(x, y, alpha, beta)
    a = solvers.coneqp(Pop, qm, Gop, hm, dims, kktsolver=lambda W: f(W))
    b = solvers.coneqp(Pm, qm, Gm, hm, dims, kktsolver="chol")
    assert a["status"] == b["status"] == "optimal" and a["iterations"] == b["iterations"]
    np.testing.assert_allclose(a["primal objective"], b["primal objective"], rtol=1e-8)


def test_qp_with_equality_constraints(ref):
    """solvers.qp(P, q, G, h, A, b, kktsolver=...) — p > 0 through the plugin."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, m, p = 90, 200, 12
    P, q, G, h = dense_qp(n, m, seed=21)
    rng = np.random.Generator(np.random.PCG64(5))
    A = rng.standard_normal((p, n))
    x0 = np.linalg.lstsq(G, h - 1.0, rcond=None)[0]
    b = A @ np.zeros(n)        # x = 0 is feasible for A x = b; G 0 = 0 <= h requires h > 0
    h = np.abs(h) + 1.0
    Pm, qm, Gm, hm, Am, bm = matrix(P), matrix(q), matrix(G), matrix(h), matrix(A), matrix(b)
    f = cvxopt_b200.kkt_chol(Gm, {"l": m, "q": [], "s": []}, Am, H=Pm)
    a = solvers.qp(Pm, qm, Gm, hm, Am, bm, kktsolver=lambda W: f(W))
    r = solvers.qp(Pm, qm, Gm, hm, Am, bm, kktsolver="chol")
    assert a["status"] == r["status"] == "optimal" and a["iterations"] == r["iterations"]
    np.testing.assert_allclose(a["primal objective"], r["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(np.array(a["x"]), np.array(r["x"]), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(np.array(a["y"]), np.array(r["y"]), rtol=1e-5, atol=1e-7)
