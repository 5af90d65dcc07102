"""Time cvxopt_b200.misc_solvers.max_step on 's' blocks (Jacobi eigensolver) against numpy's LAPACK
eigh on the host.  Usage: python tools/eig_bench.py [orders...]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from cvxopt_b200 import misc_solvers as ms

for spec in (sys.argv[1:] or ["8x64", "64x16", "128", "256", "512"]):
    if "x" in spec:
        mk, cnt = (int(v) for v in spec.split("x"))
    else:
        mk, cnt = int(spec), 1
    dims = {"l": 0, "q": [], "s": [mk] * cnt}
    rng = np.random.Generator(np.random.PCG64(1))
    x = rng.standard_normal(mk * mk * cnt)
    sig = np.zeros(mk * cnt)
    ms.max_step(x.copy(), dims)
    t0 = time.perf_counter(); t = ms.max_step(x.copy(), dims); t1 = time.perf_counter()
    xx = x.copy(); ms.max_step(xx, dims, sigma=sig); t2 = time.perf_counter()
    X = x[:mk * mk].reshape(mk, mk, order="F"); X = np.tril(X) + np.tril(X, -1).T
    t3 = time.perf_counter(); w = np.linalg.eigvalsh(X); t4 = time.perf_counter()
    print("s=[%d]x%d  values-only %.2f ms  with-vectors %.2f ms  (numpy eigvalsh of ONE block %.2f ms)  err %.2e"
          % (mk, cnt, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t4 - t3), np.abs(sig[:mk] - w).max() / np.abs(w).max()))
