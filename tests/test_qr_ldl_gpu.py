"""The other two factorisation routes of the reference through the same plugin boundary (SURVEY.md §8 rows a10, a11,
f3): misc.kkt_qr (QR of A' and of W^{-T} G Q2; the drivers' default for 'q'/'s' cones) and misc.kkt_ldl2 (LDL' of
the 2x2 system with Bunch-Kaufman pivoting), against the reference's own factories on identical inputs."""
import numpy as np
import pytest

from problems import cone_dim, cone_lp, cone_point

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _problem(dims, n, p, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    K = cone_dim(dims)
    G = np.asfortranarray(rng.standard_normal((K, n)))
    off = dims["l"] + sum(dims["q"])
    for k in dims["s"]:
        for j in range(n):
            M = G[off:off + k * k, j].reshape(k, k, order="F")
            G[off:off + k * k, j] = ((M + M.T) / 2).reshape(-1, order="F")
        off += k * k
    A = np.asfortranarray(rng.standard_normal((p, n)))
    return G, A, rng


def _rhs(dims, n, p, rng):
    K = cone_dim(dims)
    x, y, z = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(K)
    off = dims["l"] + sum(dims["q"])
    for k in dims["s"]:
        M = z[off:off + k * k].reshape(k, k, order="F")
        z[off:off + k * k] = ((M + M.T) / 2).reshape(-1, order="F")
        off += k * k
    return x, y, z


def _scaling(ref, dims, rng, spread=1.0):
    from cvxopt import matrix, misc
    s, z = cone_point(dims, rng), cone_point(dims, rng)
    if spread > 1.0 and dims["l"]:
        s[:dims["l"]] *= 10.0 ** rng.uniform(-np.log10(spread), np.log10(spread), dims["l"])
    lm = matrix(0.0, (dims["l"] + sum(dims["q"]) + sum(dims["s"]), 1))
    return misc.compute_scaling(matrix(s), matrix(z), lm, dims)


def _compare(f_ref, f_gpu, dims, n, p, rng, tol):
    from cvxopt import matrix, misc
    for rep in range(2):
        x0, y0, z0 = _rhs(dims, n, p, rng)
        xr, yr, zr = matrix(x0), matrix(y0, (p, 1)), matrix(z0)
        xg, yg, zg = matrix(x0), matrix(y0, (p, 1)), matrix(z0)
        f_ref(xr, yr, zr)
        f_gpu(xg, yg, zg)
        assert relerr(np.array(xg).ravel(), np.array(xr).ravel()) < tol
        if p:
            assert relerr(np.array(yg).ravel(), np.array(yr).ravel()) < 10 * tol
        pr, pg = matrix(0.0, zr.size), matrix(0.0, zg.size)
        misc.pack(zr, pr, dims); misc.pack(zg, pg, dims)
        assert relerr(np.array(pg).ravel(), np.array(pr).ravel()) < 10 * tol


@pytest.mark.parametrize("dims,n,p", [
    ({"l": 60, "q": [], "s": []}, 25, 0),
    ({"l": 20, "q": [9, 30], "s": [7]}, 40, 6),
    ({"l": 0, "q": [16] * 8, "s": []}, 50, 11),
    ({"l": 5, "q": [], "s": [12, 6]}, 30, 0),
    ({"l": 300, "q": [40], "s": []}, 200, 140),        # p spans two 128-blocks of reflectors
])
def test_kkt_qr_matches_reference(ref, dims, n, p):
    import cvxopt_b200
    from cvxopt import matrix, misc
    G, A, rng = _problem(dims, n, p, seed=n + p)
    W = _scaling(ref, dims, rng)
    Gm, Am = matrix(G), matrix(A) if p else matrix(0.0, (0, n))
    f_ref = misc.kkt_qr(Gm, dims, Am)(W)
    fac = cvxopt_b200.kkt_qr(Gm, dims, Am if p else None)
    f_gpu = fac(W)
    assert fac.qr_passes() == 2
    _compare(f_ref, f_gpu, dims, n, p, rng, 1e-10)
    fac.close()


def test_kkt_qr_ill_conditioned_takes_the_shifted_path(ref):
    """cond(W^{-T} G) = 1e9: cond^2 is beyond 1/eps, the plain Cholesky-QR breaks down (non-positive pivot) and the
    shifted three-pass variant takes over; the direction still matches the reference's Householder QR to ~cond*eps"""
    import cvxopt_b200
    from cvxopt import matrix, misc
    dims, n, p = {"l": 400, "q": [], "s": []}, 120, 0
    rng = np.random.Generator(np.random.PCG64(3))
    U, _ = np.linalg.qr(rng.standard_normal((400, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    G = np.asfortranarray((U * np.logspace(0, -9, n)[None, :]) @ V.T)
    W = _scaling(ref, dims, rng)
    Gm = matrix(G)
    f_ref = misc.kkt_qr(Gm, dims, matrix(0.0, (0, n)))(W)
    fac = cvxopt_b200.kkt_qr(Gm, dims, None)
    f_gpu = fac(W)
    assert fac.qr_passes() == 3
    _compare(f_ref, f_gpu, dims, n, p, rng, 1e-5)
    fac.close()


@pytest.mark.parametrize("dims,n,seed", [({"l": 0, "q": [16] * 6, "s": []}, 40, 11),
                                         ({"l": 10, "q": [7, 3], "s": [6, 9]}, 25, 12)])
def test_conelp_with_the_default_qr_route(ref, dims, n, seed):
    """solvers.conelp's default solver for these cones is 'qr' (coneprog.py:458-462): same run with the device QR"""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    c, G, h = cone_lp(n, dims, seed)
    cm, Gm, hm = matrix(c), matrix(G), matrix(h)
    want = solvers.conelp(cm, Gm, hm, dims)               # kktsolver=None -> 'qr'
    f = cvxopt_b200.kkt_qr(Gm, dims)
    got = solvers.conelp(cm, Gm, hm, dims, kktsolver=lambda W: f(W))
    f.close()
    assert want["status"] == got["status"] == "optimal" and want["iterations"] == got["iterations"]
    np.testing.assert_allclose(got["primal objective"], want["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(got["dual objective"], want["dual objective"], rtol=1e-8)


@pytest.mark.parametrize("dims,n,p,with_H", [
    ({"l": 90, "q": [], "s": []}, 45, 11, True),
    ({"l": 30, "q": [8, 5], "s": [6]}, 45, 9, True),
    ({"l": 20, "q": [], "s": []}, 50, 35, False),          # S singular: only the pivoted 2x2 system is regular
    ({"l": 200, "q": [], "s": []}, 260, 150, False),       # interchanges across many columns, N = 410
])
def test_kkt_ldl2_matches_reference(ref, dims, n, p, with_H):
    import cvxopt_b200
    from cvxopt import matrix, misc
    G, A, rng = _problem(dims, n, p, seed=2 * n + p)
    W = _scaling(ref, dims, rng)
    H = None
    if with_H:
        B = rng.standard_normal((n, n))
        H = matrix(B @ B.T / n + np.eye(n))
    Gm, Am = matrix(G), matrix(A)
    f_ref = misc.kkt_ldl2(Gm, dims, Am)(W, H)
    fac = cvxopt_b200.kkt_ldl2(Gm, dims, Am)
    f_gpu = fac(W, H)
    _compare(f_ref, f_gpu, dims, n, p, rng, 1e-9)
    fac.close()


def test_qp_with_ldl2_through_the_driver(ref):
    import cvxopt_b200
    from cvxopt import matrix, solvers
    from problems import dense_qp
    n, m, p = 60, 140, 8
    P, q, G, h = dense_qp(n, m, seed=17)
    rng = np.random.Generator(np.random.PCG64(4))
    A = rng.standard_normal((p, n))
    h = np.abs(h) + 1.0
    b = np.zeros(p)
    Pm, qm, Gm, hm, Am, bm = matrix(P), matrix(q), matrix(G), matrix(h), matrix(A), matrix(b)
    dims = {"l": m, "q": [], "s": []}
    want = solvers.coneqp(Pm, qm, Gm, hm, dims, Am, bm, kktsolver="ldl2")
    f = cvxopt_b200.kkt_ldl2(Gm, dims, Am, H=Pm)
    got = solvers.coneqp(Pm, qm, Gm, hm, dims, Am, bm, kktsolver=lambda W: f(W))
    f.close()
    assert want["status"] == got["status"] == "optimal" and want["iterations"] == got["iterations"]
    np.testing.assert_allclose(got["primal objective"], want["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(np.array(got["x"]), np.array(want["x"]), rtol=1e-6, atol=1e-8)
