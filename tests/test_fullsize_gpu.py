"""GPU parity at the BASELINE shapes (SURVEY.md §8d inputs), with the library's DEFAULT kernel selection —
i.e. the headline path is the tested path:

  * cfg 2 (n=4096, m=8192) and N* (n=8192, m=16384): one factor + two solves against the oracle
    (oracle/kkt_oracle.py, pinned to the reference) for scaling spreads 1e0 / 1e4 / 1e8, once with the
    default 'l'-row SYRK (int8 slices on tcgen05, asserted) and once with CVXB_OZAKI=0 (fp64 DMMA, asserted);
  * m = 40000 > 32768: the int32-overflow drain split of the int8-slice kernel (nrange > 1);
  * cfg 3 exactly (n=2048, 64 second-order cones of 64) and cfg 5 exactly (one 512x512 'sdp' block, n=512);
  * cfg 2 as a whole solve through the unmodified solvers.coneqp: 14 iterations, pobj 3.534265721964e+03
    (SURVEY.md §8d probe of the reference), and through the device-resident IPM;
  * cfg 4: 16 of the 512 problems against the reference's solvers.qp loop + the batch's total iteration count.

Tolerance on the search direction: north_star's 1e-10 while cond(K)*eps allows it; two backward-stable
factorisations of the same K differ by ~cond(K)*eps, so beyond that the bar is cond_1(K)*eps with cond_1 from
LAPACK's dpocon on the ORACLE's factor — and test_reference_own_solver_spread shows, on the same inputs, that
the reference's own solvers ('chol' vs 'ldl') differ by as much or more.
"""
import json
import os

import numpy as np
import pytest
import scipy.linalg as sla

import kkt_oracle as ko
from problems import cone_dim, cone_lp, dense_qp, random_scaling

pytestmark = pytest.mark.gpu
EPS = 2.220446049250313e-16
# whole-solve results of the reference at the BASELINE shapes (tests/golden/make_config_golden.py)
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_runs.json")))


def relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cond1(L, K1norm):
    """1-norm condition number of K = L L' from its Cholesky factor (LAPACK dpocon)."""
    rcond, info = sla.lapack.dpocon(L, K1norm, uplo='L')
    assert info == 0
    return 1.0 / max(rcond, 1e-300)


def spread_scaling(m, spread, rng):
    """'l'-cone NT scaling whose entries span `spread` (log-uniform, centred on 1)"""
    h = 0.5 * np.log10(spread)
    d = 10.0 ** rng.uniform(-h, h, m) if spread > 1.0 else np.ones(m)
    return {"d": d, "di": 1.0 / d, "v": [], "beta": [], "r": [], "rti": []}


_CACHE = {}


def l_problem(n):
    """SURVEY §8(d) dense QP data (P, G) of size n, m = 2n, seed 1234; cached across the parametrised cases"""
    if n not in _CACHE:
        _CACHE.clear()                       # one size resident at a time (G is 1.07 GB at n=8192)
        rng = np.random.Generator(np.random.PCG64(1234))
        A0 = rng.standard_normal((n, n))
        P = sla.blas.dsyrk(1.0 / n, A0, trans=1, lower=1)
        P = P + np.tril(P, -1).T + np.eye(n)
        del A0
        G = np.asfortranarray(rng.standard_normal((n, 2 * n)).T)
        _CACHE[n] = (np.asfortranarray(P), G)
    return _CACHE[n]


def oracle_l(G, P, W, rhs):
    """oracle factor + solves for an 'l'-only problem; returns solutions, cond_1(K)"""
    m, n = G.shape
    dims = {"l": m, "q": [], "s": []}
    f = ko.KktChol(G, dims)
    solve = f.factor(W, P)
    L, anorm = f.L, f.K1norm
    out = []
    for (x, z) in rhs:
        xo, zo = x.copy(), z.copy()
        solve(xo, None, zo)
        out.append((xo, zo))
    return out, cond1(L, anorm)


@pytest.mark.parametrize("n", [4096, 8192])
@pytest.mark.parametrize("spread", [1e0, 1e4, 1e8])
def test_l_cone_at_baseline_size_default_and_dmma(n, spread, monkeypatch):
    import cvxopt_b200
    P, G = l_problem(n)
    m = 2 * n
    dims = {"l": m, "q": [], "s": []}
    rng = np.random.Generator(np.random.PCG64(int(np.log10(spread)) + n))
    W = spread_scaling(m, spread, rng)
    rhs = [(rng.standard_normal(n), rng.standard_normal(m)) for _ in range(2)]
    want, c1 = oracle_l(G, P, W, rhs)
    bar = max(1e-10, c1 * EPS)
    got = {}
    for env, path in ((None, "int8"), ("0", "dmma")):
        if env is None:
            monkeypatch.delenv("CVXB_OZAKI", raising=False)
        else:
            monkeypatch.setenv("CVXB_OZAKI", env)
        fac = cvxopt_b200.kkt_chol(G, dims, None, H=P)
        solve = fac(W)
        assert fac.syrk_path() == path          # the default at this size IS the int8-slice kernel
        errs = []
        for (x, z), (xo, zo) in zip(rhs, want):
            xg, zg = x.copy(), z.copy()
            solve(xg, None, zg)
            errs.append((relerr(xg, xo), relerr(zg, zo)))
            got[(path, len(errs))] = xg
        fac.close()
        for ex, ez in errs:
            assert ex < bar and ez < 10 * bar, (path, n, spread, errs, c1)
    # the two tensor paths against each other: same K to fp64 level -> same bar
    assert relerr(got[("int8", 1)], got[("dmma", 1)]) < bar


def test_int8_slice_syrk_drain_split_m_40000():
    """m = 40000 > 32768 rows: the int32 accumulators are drained twice per pass (nrange = 2), default path."""
    import cvxopt_b200
    n, m = 4096, 40000
    rng = np.random.Generator(np.random.PCG64(4))
    G = np.asfortranarray(rng.standard_normal((n, m)).T)
    P = np.asfortranarray(np.eye(n) * 2.0)
    dims = {"l": m, "q": [], "s": []}
    W = spread_scaling(m, 1e6, rng)
    rhs = [(rng.standard_normal(n), rng.standard_normal(m))]
    want, c1 = oracle_l(G, P, W, rhs)
    bar = max(1e-10, c1 * EPS)
    os.environ.pop("CVXB_OZAKI", None)
    fac = cvxopt_b200.kkt_chol(G, dims, None, H=P)
    solve = fac(W)
    assert fac.syrk_path() == "int8"
    x, z = rhs[0][0].copy(), rhs[0][1].copy()
    solve(x, None, z)
    fac.close()
    assert relerr(x, want[0][0]) < bar and relerr(z, want[0][1]) < 10 * bar, (relerr(x, want[0][0]), c1)


def test_int8_slice_syrk_drain_split_ragged_forced(monkeypatch):
    """same split with ragged sizes (n not a multiple of 128, m not a multiple of 32), forced at small n"""
    import cvxopt_b200
    monkeypatch.setenv("CVXB_OZAKI", "2")
    n, m = 333, 32768 + 4097
    rng = np.random.Generator(np.random.PCG64(5))
    G = np.asfortranarray(rng.standard_normal((n, m)).T)
    dims = {"l": m, "q": [], "s": []}
    W = spread_scaling(m, 1e3, rng)
    rhs = [(rng.standard_normal(n), rng.standard_normal(m))]
    want, c1 = oracle_l(G, None, W, rhs)
    fac = cvxopt_b200.kkt_chol(G, dims, None)
    solve = fac(W)
    assert fac.syrk_path() == "int8"
    x, z = rhs[0][0].copy(), rhs[0][1].copy()
    solve(x, None, z)
    fac.close()
    assert relerr(x, want[0][0]) < 1e-10 and relerr(z, want[0][1]) < 1e-10


def _packed(z, dims):
    cp = ko.cone_sizes(dims)[4]
    out = np.zeros(cp)
    ko.pack(z.copy(), out, dims)
    return out


@pytest.mark.parametrize("name,dims,n", [
    ("cfg3_socp", {"l": 0, "q": [64] * 64, "s": []}, 2048),
    ("cfg5_sdp", {"l": 0, "q": [], "s": [512]}, 512),
])
def test_config3_and_config5_factor_solve_exact_shapes(name, dims, n):
    """BASELINE configs 3 and 5 at their stated shapes: factor + 2 solves vs the oracle, 1e-10."""
    import cvxopt_b200
    c, G, h = cone_lp(n, dims, seed=11)
    K = cone_dim(dims)
    W, _ = random_scaling(dims, seed=12)
    rng = np.random.Generator(np.random.PCG64(13))
    fac = cvxopt_b200.kkt_chol(G, dims, None)
    solve = fac(W)
    f_or = ko.KktChol(G, dims).factor(W)
    for rep in range(2):
        x, z = rng.standard_normal(n), rng.standard_normal(K)
        if dims["s"]:       # right-hand sides of the solver are symmetric 's' blocks
            for k in dims["s"]:
                M = z[-k * k:].reshape(k, k, order="F")
                z[-k * k:] = ((M + M.T) / 2).reshape(-1, order="F")
        xo, zo = x.copy(), z.copy()
        solve(x, None, z)
        f_or(xo, None, zo)
        assert relerr(x, xo) < 1e-10, (name, relerr(x, xo))
        assert relerr(_packed(z, dims), _packed(zo, dims)) < 1e-10, name
    fac.close()


def test_config2_whole_solve_through_unmodified_coneqp(ref):
    """cfg 2: solvers.coneqp (unmodified reference driver) + this kktsolver: 14 iterations,
    pobj 3.534265721964e+03 to rtol 1e-8 — the reference's own run (SURVEY.md §8d probe); and the same
    numbers from the device-resident IPM, whose factor now uses the same int8-slice SYRK."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, m = 4096, 8192
    P, q, G, h = dense_qp(n, m, seed=1234)
    Pm, qm, Gm, hm = matrix(P), matrix(q), matrix(G), matrix(h)
    dims = {"l": m, "q": [], "s": []}
    os.environ.pop("CVXB_OZAKI", None)
    f = cvxopt_b200.kkt_chol(Gm, dims, None, H=Pm)
    sol = solvers.coneqp(Pm, qm, Gm, hm, dims, kktsolver=lambda W: f(W))
    assert f.syrk_path() == "int8"
    f.close()
    assert sol["status"] == GOLD["cfg2"]["status"] == "optimal"
    assert sol["iterations"] == GOLD["cfg2"]["iterations"] == 14
    np.testing.assert_allclose(sol["primal objective"], 3.534265721964e+03, rtol=1e-8)
    np.testing.assert_allclose(sol["primal objective"], GOLD["cfg2"]["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(sol["dual objective"], GOLD["cfg2"]["dual objective"], rtol=1e-8)
    b = cvxopt_b200.QPBatch(1, n, m)
    b.load(P[None], q[None], G[None], h[None])
    b.solve()
    r, st = b.results(), b.stats()
    b.close()
    assert st["syrk_path"] == "int8"
    assert r["status"][0] == "optimal" and int(r["iterations"][0]) == 14
    np.testing.assert_allclose(r["primal objective"][0], sol["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(r["x"][0], np.array(sol["x"]).ravel(), rtol=1e-6, atol=1e-8)


def test_config3_whole_solve_matches_reference(ref):
    """cfg 3 exactly: solvers.conelp with 64 second-order cones of 64, n=2048: reference 'chol' vs the plugin."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, dims = 2048, {"l": 0, "q": [64] * 64, "s": []}
    c, G, h = cone_lp(n, dims, seed=11)
    cm, Gm, hm = matrix(c), matrix(G), matrix(h)
    f = cvxopt_b200.kkt_chol(Gm, dims)
    a = solvers.conelp(cm, Gm, hm, dims, kktsolver=lambda W: f(W))
    f.close()
    b = GOLD["cfg3"]                      # the reference's own run of this problem with kktsolver='chol'
    assert a["status"] == b["status"] == "optimal" and a["iterations"] == b["iterations"]
    np.testing.assert_allclose(a["primal objective"], b["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(a["dual objective"], b["dual objective"], rtol=1e-8)


def test_config5_whole_solve_matches_reference(ref):
    """cfg 5 exactly: one 512x512 'sdp' block, n=512 (G is 262144 x 512): unmodified solvers.conelp + plugin vs
    the reference's own kktsolver='chol' run (committed golden: 10 iterations)."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    n, dims = 512, {"l": 0, "q": [], "s": [512]}
    c, G, h = cone_lp(n, dims, seed=11)
    cm, Gm, hm = matrix(c), matrix(G), matrix(h)
    f = cvxopt_b200.kkt_chol(Gm, dims)
    a = solvers.conelp(cm, Gm, hm, dims, kktsolver=lambda W: f(W))
    f.close()
    b = GOLD["cfg5"]
    assert a["status"] == b["status"] == "optimal" and a["iterations"] == b["iterations"]
    np.testing.assert_allclose(a["primal objective"], b["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(a["dual objective"], b["dual objective"], rtol=1e-8)


def test_config4_batch_subset_vs_reference_and_total_iterations(ref):
    """cfg 4: the 512 QPs (n=512, m=1024, seeds 0..511) on one GPU: all optimal, total iteration count pinned
    (5648 = sum over the reference's solvers.qp runs), 16 problems spread over the batch compared one by one."""
    import cvxopt_b200
    from cvxopt import matrix, solvers
    B, n, m = 512, 512, 1024
    Ps, qs, Gs, hs = (np.empty((B, n, n)), np.empty((B, n)), np.empty((B, m, n)), np.empty((B, m)))
    for k in range(B):
        Ps[k], qs[k], Gs[k], hs[k] = dense_qp(n, m, seed=k)
    got = cvxopt_b200.qp_batch(Ps, qs, Gs, hs)
    assert all(s == "optimal" for s in got["status"])
    g4 = GOLD["cfg4"]                     # the reference's solvers.qp on each of the 512 problems
    assert g4["all_optimal"] and g4["iterations_total"] == 5648
    assert int(got["iterations"].sum()) == 5648
    assert list(map(int, got["iterations"])) == g4["iterations"]
    np.testing.assert_allclose(got["primal objective"], g4["primal objective"], rtol=1e-8)
    np.testing.assert_allclose(got["dual objective"], g4["dual objective"], rtol=1e-8)
    for k in range(0, B, 32):
        w = solvers.qp(matrix(Ps[k]), matrix(qs[k]), matrix(Gs[k]), matrix(hs[k]), kktsolver="chol")
        assert w["status"] == "optimal" and got["iterations"][k] == w["iterations"], k
        np.testing.assert_allclose(got["primal objective"][k], w["primal objective"], rtol=1e-8)
        np.testing.assert_allclose(got["dual objective"][k], w["dual objective"], rtol=1e-8)
        np.testing.assert_allclose(got["x"][k], np.array(w["x"]).ravel(), rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("spread", [1e0, 1e4, 1e8])
def test_reference_own_solver_spread(ref, spread):
    """The yardstick for the cond-proportional bar: on identical inputs, the reference's OWN factorisations
    ('chol' = Cholesky of the reduced system, 'ldl' = LDL' of the 3x3 system; misc.py:1213, :1055) differ from
    each other by e_ref; this library's direction must be within max(1e-10, 2 e_ref) of the reference's chol."""
    import cvxopt_b200
    from cvxopt import matrix, misc
    n, m = 512, 1024
    P, q, G, h = dense_qp(n, m, seed=7)
    dims = {"l": m, "q": [], "s": []}
    rng = np.random.Generator(np.random.PCG64(3))
    Wn = spread_scaling(m, spread, rng)
    W = {"d": matrix(Wn["d"]), "di": matrix(Wn["di"]), "v": [], "beta": [], "r": [], "rti": []}
    Gm, Pm, A0 = matrix(G), matrix(P), matrix(0.0, (0, n))
    x0, z0 = rng.standard_normal(n), rng.standard_normal(m)
    sols = {}
    for name in ("kkt_chol", "kkt_chol2", "kkt_ldl"):
        f = getattr(misc, name)(Gm, dims, A0)(W, Pm)
        x, z = matrix(x0), matrix(z0)
        f(x, matrix(0.0, (0, 1)), z)
        sols[name] = np.array(x).ravel()
    e_ref = max(relerr(sols["kkt_ldl"], sols["kkt_chol"]), relerr(sols["kkt_chol2"], sols["kkt_chol"]))
    for env in (None, "2"):
        if env is None:
            os.environ.pop("CVXB_OZAKI", None)
        else:
            os.environ["CVXB_OZAKI"] = env
        try:
            fac = cvxopt_b200.kkt_chol(G, dims, None, H=P)
            x, z = x0.copy(), z0.copy()
            fac(Wn)(x, None, z)
            fac.close()
        finally:
            os.environ.pop("CVXB_OZAKI", None)
        e = relerr(x, sols["kkt_chol"])
        assert e < max(1e-10, 2.0 * e_ref), (spread, env, e, e_ref)
