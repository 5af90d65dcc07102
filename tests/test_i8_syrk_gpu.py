"""The experimental fp64 SYRK on the int8 tensor path (csrc/ozaki_syrk.cu, C-ABI cvxb_syrk_scaled_i8):
nine radix-2^7 slices per entry, exact int32 products, fp64-level result.  Checked against an 80-bit
long double evaluation of the same sums and against the DMMA kernel (cvxb_syrk_scaled), and through
the KKT path (CVXB_OZAKI=2 forces it at any size) against the oracle."""
import os

import numpy as np
import pytest

import kkt_oracle as ko
from problems import cone_dim, random_scaling

pytestmark = pytest.mark.gpu


def _ref_long(G, d, H):
    Gs = (G * d[:, None]).astype(np.longdouble)          # fl(d*g) in fp64 first, like the kernels
    C = Gs.T @ Gs + H.astype(np.longdouble)
    mag = np.abs(Gs).T @ np.abs(Gs) + np.abs(H).astype(np.longdouble)
    return C, mag


@pytest.mark.parametrize("n,m,spread", [(1, 1, 0.0), (130, 77, 1.0), (257, 1000, 3.0), (384, 4100, 6.0)])
def test_i8_slice_syrk_matches_long_double(n, m, spread):
    import torch
    from cvxopt_b200 import _lib
    lib = _lib.load()
    rng = np.random.Generator(np.random.PCG64(100 + n))
    G = np.asfortranarray(rng.standard_normal((m, n)))
    d = np.exp(spread * rng.standard_normal(m))           # late-IPM-like row scaling spread
    B = rng.standard_normal((n, n))
    H = np.asfortranarray(B @ B.T / n + np.eye(n))
    dG = torch.from_numpy(np.ascontiguousarray(G.T)).cuda()        # column-major m x n
    dd = torch.from_numpy(d.copy()).cuda()
    dH = torch.from_numpy(np.ascontiguousarray(H.T)).cuda()
    dC = torch.full((n, n), float("nan"), dtype=torch.float64, device="cuda")
    rc = lib.cvxb_syrk_scaled_i8(n, m, dG.data_ptr(), m, dd.data_ptr(), dH.data_ptr(), n, dC.data_ptr(), n, 9, 0)
    assert rc == 0, _lib.last_error()
    C = dC.cpu().numpy().T                                 # lower triangle significant
    ref, mag = _ref_long(G, d, H)
    il = np.tril_indices(n)
    err = np.abs(C[il].astype(np.longdouble) - ref[il]) / mag[il]
    assert float(err.max()) < 1e-15, float(err.max())
    # the DMMA kernel on the same data (weights d^2): both are fp64-accurate, so they agree to ~1e-15
    dC2 = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    dw = torch.from_numpy((d * d).copy()).cuda()
    rc = lib.cvxb_syrk_scaled(n, m, dG.data_ptr(), m, dw.data_ptr(), dH.data_ptr(), n, dC2.data_ptr(), n, 0)
    assert rc == 0, _lib.last_error()
    C2 = dC2.cpu().numpy().T
    assert float((np.abs(C[il] - C2[il]) / np.asarray(mag[il], dtype=np.float64)).max()) < 1e-14


def test_fewer_slices_lose_precision_gracefully():
    """s slices keep 6 + 7(s-1) bits below each column maximum: the error shrinks by ~2^-7 per slice."""
    import torch
    from cvxopt_b200 import _lib
    lib = _lib.load()
    rng = np.random.Generator(np.random.PCG64(5))
    n, m = 200, 300
    G = np.asfortranarray(rng.standard_normal((m, n)))
    d = np.ones(m)
    H = np.zeros((n, n))
    ref, mag = _ref_long(G, d, H)
    dG = torch.from_numpy(np.ascontiguousarray(G.T)).cuda()
    dd = torch.from_numpy(d).cuda()
    il = np.tril_indices(n)
    prev = None
    for s in (2, 4, 6, 9):
        dC = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        rc = lib.cvxb_syrk_scaled_i8(n, m, dG.data_ptr(), m, dd.data_ptr(), None, n, dC.data_ptr(), n, s, 0)
        assert rc == 0, _lib.last_error()
        err = float((np.abs(dC.cpu().numpy().T[il].astype(np.longdouble) - ref[il]) / mag[il]).max())
        assert err < max(2.0 ** (-(6 + 7 * (s - 1)) + 4), 1e-15)
        if prev is not None:
            assert err < prev
        prev = err
    assert lib.cvxb_syrk_scaled_i8(n, m, dG.data_ptr(), m, dd.data_ptr(), None, n, dC.data_ptr(), n, 10, 0) != 0


def test_kkt_path_on_int8_slices(monkeypatch):
    """CVXB_OZAKI=2: the factor's 'l'-row SYRK runs on the int8 path; directions vs the oracle."""
    import cvxopt_b200
    monkeypatch.setenv("CVXB_OZAKI", "2")
    for (ml, n, seed, extra) in [(700, 300, 3, {}), (1500, 260, 4, {"q": [9, 30], "s": [6]})]:
        dims = {"l": ml, "q": extra.get("q", []), "s": extra.get("s", [])}
        rng = np.random.Generator(np.random.PCG64(seed))
        K = cone_dim(dims)
        G = np.asfortranarray(rng.standard_normal((K, n)))
        B = rng.standard_normal((n, n))
        H = np.asfortranarray(B @ B.T / n + np.eye(n))
        W, _ = random_scaling(dims, seed=seed + 1)
        fac = cvxopt_b200.kkt_chol(G, dims, None)
        solve = fac(W, H)
        f_or = ko.KktChol(G, dims).factor(W, H)
        x, z = rng.standard_normal(n), rng.standard_normal(K)
        xo, zo = x.copy(), z.copy()
        solve(x, None, z)
        f_or(xo, None, zo)
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
        _, _, _, _, cp = ko.cone_sizes(dims)
        a, b = np.zeros(cp), np.zeros(cp)
        ko.pack(z.copy(), a, dims); ko.pack(zo.copy(), b, dims)
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-10
        fac.close()
