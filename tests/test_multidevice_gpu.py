"""Two devices used from ONE process through the C-ABI's `device` argument: function attributes (dynamic
shared-memory opt-in) are per device, so every kernel family must work on device 1 after first use on device 0
(and the other way round).  Skipped on a one-GPU box."""
import numpy as np
import pytest

import kkt_oracle as ko
from problems import random_scaling

pytestmark = pytest.mark.gpu


def test_two_devices_in_one_process(monkeypatch):
    import cvxopt_b200
    if cvxopt_b200.device_count() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("CVXB_OZAKI", "2")        # also the int8-slice kernel (222 KB dynamic shared memory)
    n, dims = 300, {"l": 700, "q": [9], "s": [5]}
    rng = np.random.Generator(np.random.PCG64(1))
    K = 700 + 9 + 25
    G = np.asfortranarray(rng.standard_normal((K, n)))
    B = rng.standard_normal((n, n))
    H = np.asfortranarray(B @ B.T / n + np.eye(n))
    W, _ = random_scaling(dims, seed=2)
    f_or = ko.KktChol(G, dims).factor(W, H)
    for dev in (0, 1, 0, 1):
        fac = cvxopt_b200.kkt_chol(G, dims, None, H=H, device=dev)
        solve = fac(W)
        x, z = rng.standard_normal(n), rng.standard_normal(K)
        xo, zo = x.copy(), z.copy()
        solve(x, None, z)
        f_or(xo, None, zo)
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10, dev
        fac.close()
    # the batch path on device 1 as well
    from problems import dense_qp
    P, q, Gq, h = dense_qp(60, 130, seed=3)
    a = cvxopt_b200.qp_batch(P[None], q[None], Gq[None], h[None], device=0)
    b = cvxopt_b200.qp_batch(P[None], q[None], Gq[None], h[None], device=1)
    assert a["iterations"][0] == b["iterations"][0] and a["status"][0] == "optimal"
    np.testing.assert_allclose(a["x"], b["x"], rtol=1e-12, atol=1e-14)
