"""The nonlinear-solver route: mnl > 0 nonlinear rows Df with the 'dnl'/'dnli' part of
the scaling (reference misc.py:45-71, 1268-1270), the `kktsolver(x, z, W)` signature of cvxprog.cpl / cp
(cvxprog.py:277-293, 526-537, 1876-1887), against the reference itself (oracle/_ref)."""
import numpy as np
import pytest

from problems import cone_dim, cone_point

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("dims,n,mnl,p", [
    ({"l": 40, "q": [], "s": []}, 25, 3, 0),
    ({"l": 12, "q": [6, 3], "s": [4]}, 20, 5, 0),
    ({"l": 30, "q": [5], "s": []}, 22, 2, 4),
    ({"l": 0, "q": [], "s": []}, 16, 7, 0),            # nonlinear rows only
])
def test_factor_with_nonlinear_rows_matches_reference_kkt_chol(ref, dims, n, mnl, p):
    """factor(W, H, Df)(x, y, z) with W['dnl'], W['dnli'] vs misc.kkt_chol(G, dims, A, mnl) — 1e-10."""
    import cvxopt_b200
    from cvxopt import matrix, misc
    rng = np.random.Generator(np.random.PCG64(n + mnl))
    K = cone_dim(dims)
    G = np.asfortranarray(rng.standard_normal((K, n)))
    off = dims["l"] + sum(dims["q"])
    for k in dims["s"]:                                   # symmetric 's' columns
        for j in range(n):
            M = G[off:off + k * k, j].reshape(k, k, order="F")
            G[off:off + k * k, j] = ((M + M.T) / 2).reshape(-1, order="F")
        off += k * k
    A = np.asfortranarray(rng.standard_normal((p, n)))
    Df = np.asfortranarray(rng.standard_normal((mnl, n)))
    B = rng.standard_normal((n, n))
    H = np.asfortranarray(B @ B.T / n + 0.5 * np.eye(n))
    # NT scaling of a random interior pair, computed by the reference (dnl block included)
    s = np.concatenate([rng.uniform(0.5, 2.0, mnl), cone_point(dims, rng)])
    z = np.concatenate([rng.uniform(0.5, 2.0, mnl), cone_point(dims, rng)])
    lmbda = matrix(0.0, (mnl + dims["l"] + sum(dims["q"]) + sum(dims["s"]), 1))
    W = misc.compute_scaling(matrix(s), matrix(z), lmbda, dims, mnl)
    assert "dnl" in W and len(W["dnl"]) == mnl
    Gm, Am = matrix(G), matrix(A) if p else matrix(0.0, (0, n))
    f_ref = misc.kkt_chol(Gm, dims, Am, mnl)(W, matrix(H), matrix(Df))
    fac = cvxopt_b200.kkt_chol(Gm, dims, Am if p else None, mnl)
    f_gpu = fac(W, matrix(H), matrix(Df))
    for rep in range(2):
        x0, y0, z0 = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(mnl + K)
        off = mnl + dims["l"] + sum(dims["q"])
        for k in dims["s"]:
            M = z0[off:off + k * k].reshape(k, k, order="F")
            z0[off:off + k * k] = ((M + M.T) / 2).reshape(-1, order="F")
            off += k * k
        xr, yr, zr = matrix(x0), matrix(y0, (p, 1)), matrix(z0)
        xg, yg, zg = matrix(x0), matrix(y0, (p, 1)), matrix(z0)
        f_ref(xr, yr, zr)
        f_gpu(xg, yg, zg)
        assert relerr(np.array(xg).ravel(), np.array(xr).ravel()) < 1e-10
        if p:
            assert relerr(np.array(yg).ravel(), np.array(yr).ravel()) < 1e-9
        # compare the significant (lower-triangular) part of the 's' blocks through pack
        pr, pg = matrix(0.0, zr.size), matrix(0.0, zg.size)
        misc.pack(zr, pr, dims, mnl); misc.pack(zg, pg, dims, mnl)
        assert relerr(np.array(pg).ravel(), np.array(pr).ravel()) < 1e-10
    fac.close()


def _lse_problem(n, mnl, dims, seed):
    """minimize  lse(B0 x + g0) + 1/2 |x|^2   s.t.  lse(Bk x + gk) <= 0 (k = 1..mnl),  G x <=_K h.
    x = 0 is strictly feasible.  Returns (F, G, h) in cvxopt types."""
    from cvxopt import matrix
    rng = np.random.Generator(np.random.PCG64(seed))
    r = 6
    Bs = [rng.standard_normal((r, n)) / np.sqrt(n) for _ in range(mnl + 1)]
    gs = [rng.standard_normal(r) for _ in range(mnl + 1)]
    for k in range(1, mnl + 1):
        gs[k] -= np.log(np.exp(gs[k]).sum()) + 1.0          # lse(gk) = -1 < 0
    K = cone_dim(dims)
    G = np.asfortranarray(rng.standard_normal((K, n)))
    off = dims["l"] + sum(dims["q"])
    for k in dims["s"]:
        for j in range(n):
            M = G[off:off + k * k, j].reshape(k, k, order="F")
            G[off:off + k * k, j] = ((M + M.T) / 2).reshape(-1, order="F")
        off += k * k
    h = cone_point(dims, rng)                               # G*0 + s = h with s interior

    def F(x=None, z=None):
        if x is None:
            return mnl, matrix(0.0, (n, 1))
        xv = np.array(x).ravel()
        f = np.zeros(mnl + 1)
        Df = np.zeros((mnl + 1, n))
        Hs = np.zeros((n, n))
        for k in range(mnl + 1):
            y = Bs[k] @ xv + gs[k]
            ymax = y.max()
            e = np.exp(y - ymax)
            f[k] = ymax + np.log(e.sum())
            pk = e / e.sum()
            Df[k] = Bs[k].T @ pk
            if z is not None:
                Hs += float(z[k]) * (Bs[k].T @ (np.diag(pk) - np.outer(pk, pk)) @ Bs[k])
        f[0] += 0.5 * xv @ xv
        Df[0] += xv
        if z is None:
            return matrix(f), matrix(Df)
        Hs += float(z[0]) * np.eye(n)
        return matrix(f), matrix(Df), matrix(Hs)
    return F, matrix(G), matrix(h)


@pytest.mark.parametrize("dims,n,mnl", [
    ({"l": 40, "q": [], "s": []}, 30, 3),
    ({"l": 15, "q": [6], "s": [4]}, 24, 2),
])
def test_cp_with_device_kktsolver_matches_reference(ref, dims, n, mnl):
    """solvers.cp (unmodified reference driver) with cvxopt_b200.cp_kktsolver(x, z, W) vs kktsolver='chol'."""
    import cvxopt_b200
    from cvxopt import solvers
    from cvxopt import matrix, misc
    F, G, h = _lse_problem(n, mnl, dims, seed=3 * n + mnl)
    calls = {"ref": 0, "gpu": 0}
    # the reference's own 'chol' route, spelled out as cp does it (cvxprog.py:1876-1887) so calls can be counted
    fref = misc.kkt_chol(G, dims, matrix(0.0, (0, n)), mnl)

    def ks_ref(x, z, W):
        calls["ref"] += 1
        f, Df, H = F(x, z)
        return fref(W, H, Df[1:, :])
    want = solvers.cp(F, G, h, dims, kktsolver=ks_ref)
    ks = cvxopt_b200.cp_kktsolver(F, G, dims, None, mnl)

    def ks_gpu(x, z, W):
        calls["gpu"] += 1
        return ks(x, z, W)
    before = cvxopt_b200.launch_count()
    got = solvers.cp(F, G, h, dims, kktsolver=ks_gpu)
    assert cvxopt_b200.launch_count() > before
    ks.factory.close()
    assert want["status"] == got["status"] == "optimal"
    assert calls["ref"] == calls["gpu"] > 3          # same number of interior-point iterations
    byname = solvers.cp(F, G, h, dims, kktsolver="chol")
    np.testing.assert_allclose(byname["primal objective"], want["primal objective"], rtol=1e-12)
    np.testing.assert_allclose(got["primal objective"], want["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(got["dual objective"], want["dual objective"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(np.array(got["x"]), np.array(want["x"]), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(np.array(got["znl"]), np.array(want["znl"]), rtol=1e-5, atol=1e-8)


def test_cpl_with_device_kktsolver_matches_reference(ref):
    """solvers.cpl: linear objective, nonlinear constraints, kktsolver(x, z, W) = factor(W, H, Df)."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, mnl, dims = 20, 4, {"l": 25, "q": [5], "s": []}
    Fcp, G, h = _lse_problem(n, mnl, dims, seed=77)

    def F(x=None, z=None):              # drop the objective row of the cp problem: constraints only
        if x is None:
            return mnl, matrix(0.0, (n, 1))
        if z is None:
            f, Df = Fcp(x)
            return f[1:], Df[1:, :]
        zz = matrix(0.0, (mnl + 1, 1))
        zz[1:] = z
        f, Df, H = Fcp(x, zz)
        return f[1:], Df[1:, :], H
    c = matrix(np.random.Generator(np.random.PCG64(5)).standard_normal(n))
    # bounded: add box rows through the 'l' block of G? keep it bounded with the quadratic-free lse rows + cone rows
    want = solvers.cpl(c, F, G, h, dims, kktsolver="chol")
    ks = cvxopt_b200.cpl_kktsolver(F, G, dims, None, mnl)
    got = solvers.cpl(c, F, G, h, dims, kktsolver=ks)
    ks.factory.close()
    assert want["status"] == got["status"] == "optimal"
    np.testing.assert_allclose(got["primal objective"], want["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(got["dual objective"], want["dual objective"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(np.array(got["x"]), np.array(want["x"]), rtol=1e-6, atol=1e-8)
