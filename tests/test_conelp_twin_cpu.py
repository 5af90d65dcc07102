"""CPU check of the DRIVER LOGIC of cvxopt_b200/conelp.py (the device-resident restatement of coneprog.conelp): the
same code with every device closure replaced by the reference's own function (tools/conelp_host_twin.py) must take
the reference's iterates: same status, iteration count and objectives on 'l', 'q', 's' and mixed cones."""
import importlib
import os
import sys

import numpy as np
import pytest

from problems import cone_lp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("dims,n,seed", [
    ({"l": 30, "q": [], "s": []}, 12, 1),
    ({"l": 0, "q": [16] * 6, "s": []}, 40, 11),
    ({"l": 10, "q": [7, 3], "s": [6, 9]}, 25, 12),
    ({"l": 0, "q": [], "s": [24]}, 30, 13),
])
def test_driver_logic_matches_reference_on_cpu(ref, dims, n, seed):
    from cvxopt import matrix, solvers
    import conelp_host_twin as twin
    dc = importlib.import_module("cvxopt_b200.conelp")
    c, G, h = cone_lp(n, dims, seed)
    want = solvers.conelp(matrix(c), matrix(G), matrix(h), dims, kktsolver="chol")
    got = twin.twin_conelp(dc, c, G, h, dims)
    assert want["status"] == got["status"] == "optimal"
    assert want["iterations"] == got["iterations"]
    np.testing.assert_allclose(got["primal objective"], want["primal objective"], rtol=1e-9)
    np.testing.assert_allclose(got["dual objective"], want["dual objective"], rtol=1e-9)
    np.testing.assert_allclose(got["x"], np.array(want["x"]).ravel(), rtol=1e-6, atol=1e-9)
