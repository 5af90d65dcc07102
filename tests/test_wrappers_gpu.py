"""The reference's problem-class wrappers through the `kktsolver` plugin: solvers.lp (coneprog.py:2550),
solvers.sdp (:3566), solvers.socp (:3013), and the function-valued A operator (:1682-1711) served by the
resident equality-constraint matrix.  Reference runs ('chol') from oracle/_ref."""
import numpy as np
import pytest

from problems import cone_lp, dense_qp

pytestmark = pytest.mark.gpu


def _same(a, b, xtol=1e-6):
    assert a["status"] == b["status"] == "optimal"
    assert a["iterations"] == b["iterations"]
    np.testing.assert_allclose(a["primal objective"], b["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(a["dual objective"], b["dual objective"], rtol=1e-8)
    np.testing.assert_allclose(np.array(a["x"]), np.array(b["x"]), rtol=xtol, atol=1e-8)


def test_solvers_lp_through_the_plugin(ref):
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, m = 60, 150
    dims = {"l": m, "q": [], "s": []}
    c, G, h = cone_lp(n, dims, seed=31)
    cm, Gm, hm = matrix(c), matrix(G), matrix(h)
    f = cvxopt_b200.kkt_chol(Gm, dims)
    a = solvers.lp(cm, Gm, hm, kktsolver=lambda W: f(W))
    b = solvers.lp(cm, Gm, hm, kktsolver="chol")
    f.close()
    _same(a, b)


def test_solvers_lp_with_equalities_through_the_plugin(ref):
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, m, p = 50, 120, 7
    dims = {"l": m, "q": [], "s": []}
    c, G, h = cone_lp(n, dims, seed=32)
    rng = np.random.Generator(np.random.PCG64(2))
    A = rng.standard_normal((p, n))
    x0 = np.linalg.lstsq(G, h - 0.5, rcond=None)[0]            # roughly central point: keep A x = b feasible
    b = A @ x0
    cm, Gm, hm, Am, bm = matrix(c), matrix(G), matrix(h), matrix(A), matrix(b)
    ref_sol = solvers.lp(cm, Gm, hm, Am, bm, kktsolver="chol")
    f = cvxopt_b200.kkt_chol(Gm, dims, Am)
    got = solvers.lp(cm, Gm, hm, Am, bm, kktsolver=lambda W: f(W))
    f.close()
    assert got["status"] == ref_sol["status"]
    if ref_sol["status"] == "optimal":
        _same(got, ref_sol, xtol=1e-5)


def test_solvers_sdp_through_the_plugin(ref):
    """solvers.sdp stacks [Gl; Gs...] and calls conelp (coneprog.py:4083-4126): the factory gets the same stack."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n = 12
    dims = {"l": 4, "q": [], "s": [6, 9]}
    c, G, h = cone_lp(n, dims, seed=33)
    cm, Gm, hm = matrix(c), matrix(G), matrix(h)
    Gl, hl = Gm[:4, :], hm[:4]
    Gs = [Gm[4:40, :], Gm[40:121, :]]
    hs = [matrix(hm[4:40], (6, 6)), matrix(hm[40:121], (9, 9))]
    f = cvxopt_b200.kkt_chol(Gm, dims)
    a = solvers.sdp(cm, Gl, hl, Gs, hs, kktsolver=lambda W: f(W))
    b = solvers.sdp(cm, Gl, hl, Gs, hs, kktsolver="chol")
    f.close()
    _same(a, b)
    for k in range(2):
        np.testing.assert_allclose(np.array(a["zs"][k]), np.array(b["zs"][k]), rtol=1e-5, atol=1e-7)


def test_solvers_socp_through_the_plugin(ref):
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n = 20
    dims = {"l": 5, "q": [8, 3, 12], "s": []}
    c, G, h = cone_lp(n, dims, seed=34)
    cm, Gm, hm = matrix(c), matrix(G), matrix(h)
    Gq = [Gm[5:13, :], Gm[13:16, :], Gm[16:28, :]]
    hq = [hm[5:13], hm[13:16], hm[16:28]]
    f = cvxopt_b200.kkt_chol(Gm, dims)
    a = solvers.socp(cm, Gm[:5, :], hm[:5], Gq, hq, kktsolver=lambda W: f(W))
    b = solvers.socp(cm, Gm[:5, :], hm[:5], Gq, hq, kktsolver="chol")
    f.close()
    _same(a, b)


def test_device_operator_A(ref):
    """function-valued G, A and P all served by the device copies (requires a custom kktsolver,
    coneprog.py:1820-1833)."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, m, p = 70, 160, 9
    P, q, G, h = dense_qp(n, m, seed=41)
    rng = np.random.Generator(np.random.PCG64(9))
    A = rng.standard_normal((p, n))
    h = np.abs(h) + 1.0                     # x = 0 strictly feasible for G x <= h
    b = np.zeros(p)
    Pm, qm, Gm, hm, Am, bm = matrix(P), matrix(q), matrix(G), matrix(h), matrix(A), matrix(b)
    dims = {"l": m, "q": [], "s": []}
    f = cvxopt_b200.kkt_chol(Gm, dims, Am, H=Pm)
    # y := alpha*A*x + beta*y against numpy
    x, y = rng.standard_normal(n), rng.standard_normal(p)
    want = 0.7 * A @ x - 0.3 * y
    f.A(x, y, 0.7, -0.3, "N")
    np.testing.assert_allclose(y, want, rtol=1e-12, atol=1e-12)
    xt, yt = rng.standard_normal(p), rng.standard_normal(n)
    want = 2.0 * A.T @ xt + yt
    f.A(xt, yt, 2.0, 1.0, "T")
    np.testing.assert_allclose(yt, want, rtol=1e-12, atol=1e-12)

    def Gop(u, v, alpha=1.0, beta=0.0, trans="N"):
        f.G(u, v, alpha, beta, trans)

    def Aop(u, v, alpha=1.0, beta=0.0, trans="N"):
        f.A(u, v, alpha, beta, trans)

    def Pop(u, v, alpha=1.0, beta=0.0):
        f.P(u, v, alpha, beta)
    got = solvers.coneqp(Pop, qm, Gop, hm, dims, Aop, bm, kktsolver=lambda W: f(W))
    want = solvers.coneqp(Pm, qm, Gm, hm, dims, Am, bm, kktsolver="chol")
    f.close()
    assert got["status"] == want["status"] == "optimal" and got["iterations"] == want["iterations"]
    np.testing.assert_allclose(got["primal objective"], want["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(np.array(got["x"]), np.array(want["x"]), rtol=1e-6, atol=1e-8)
