"""The opt-in TMA / mbarrier SYRK kernel (CVXB_TMA=1) must stay parity-green.  The switch is read
once per process, so the check runs in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import cvxopt_b200, kkt_oracle as ko
from problems import random_scaling
for (n, m) in ((513, 1100), (300, 2048)):        # ragged n, K tail and K multiple of 16
    rng = np.random.Generator(np.random.PCG64(n))
    G = np.asfortranarray(rng.standard_normal((m, n)))
    B = rng.standard_normal((n, n)); H = np.asfortranarray(B @ B.T / n + np.eye(n))
    dims = {"l": m, "q": [], "s": []}
    W, _ = random_scaling(dims, 3)
    f = cvxopt_b200.kkt_chol(G, dims, H=H)
    solve = f(W)
    x, z = rng.standard_normal(n), rng.standard_normal(m)
    xo, zo = x.copy(), z.copy()
    solve(x, None, z)
    ko.KktChol(G, dims).factor(W, H)(xo, None, zo)
    err = np.linalg.norm(x - xo) / np.linalg.norm(xo)
    assert err < 1e-10, err
print("TMA_PATH_OK")
""" % (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"))


def test_tma_syrk_variant_parity():
    env = dict(os.environ, CVXB_TMA="1")
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    assert "TMA_PATH_OK" in out.stdout, out.stdout + out.stderr
