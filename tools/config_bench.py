"""KKT factor + solve timings at BASELINE configs 3 (SOCP) and 5 (SDP) through the plugin
boundary (cvxopt_b200.kkt_chol), with a parity spot-check against the numpy oracle."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import cvxopt_b200
import kkt_oracle as ko
from problems import cone_dim, cone_lp, random_scaling


def run(name, dims, n, check, route="chol"):
    K = cone_dim(dims)
    c, G, h = cone_lp(n, dims, seed=11)
    W, _ = random_scaling(dims, seed=5)
    fac = cvxopt_b200.kkt_qr(G, dims) if route == "qr" else cvxopt_b200.kkt_chol(G, dims)
    rng = np.random.Generator(np.random.PCG64(1))
    solve = fac(W)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        solve = fac(W)
        tf = time.perf_counter() - t0
        x, z = rng.standard_normal(n), rng.standard_normal(K)
        t0 = time.perf_counter()
        solve(x, None, z)
        tsol = time.perf_counter() - t0
        ts.append((tf, tsol, fac.last_ms(), fac.last_breakdown()))
    tf = min(t[0] for t in ts) * 1e3
    tsol = min(t[1] for t in ts) * 1e3
    _, _, _, _, kp = ko.cone_sizes(dims)
    f_cong = sum(3.0 * s ** 3 * n for s in dims["s"])
    f_fac = float(n) * n * kp + n ** 3 / 3.0 + f_cong
    out = {"config": name, "route": route, "n": n, "cdim": K, "cdim_pckd": kp, "factor_ms_wall": tf, "solve_ms_wall": tsol,
           "factor_ms_dev": ts[-1][2][0], "solve_ms_dev": ts[-1][2][1], "breakdown": ts[-1][3],
           "factor_tflops": f_fac / (ts[-1][2][0] * 1e-3) * 1e-12}
    if check:
        x, z = rng.standard_normal(n), rng.standard_normal(K)
        xo, zo = x.copy(), z.copy()
        solve(x, None, z)
        t0 = time.perf_counter()
        fo = ko.KktChol(G, dims).factor(W)
        out["oracle_factor_s"] = time.perf_counter() - t0
        fo(xo, None, zo)
        out["dx_rel_vs_oracle"] = float(np.linalg.norm(x - xo) / np.linalg.norm(xo))
    fac.close()
    print(json.dumps(out))


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "3", "3qr", "5"]
    if "2" in which:
        run("2: QP n=4096, l=8192 (KKT step without P)", {"l": 8192, "q": [], "s": []}, 4096, True)
    if "3qr" in which:
        run("3: SOCP n=2048, 64 x q64", {"l": 0, "q": [64] * 64, "s": []}, 2048, False, route="qr")
    if "3" in which:
        run("3: SOCP n=2048, 64 x q64", {"l": 0, "q": [64] * 64, "s": []}, 2048, True)
    if "5" in which:
        run("5: SDP one 512x512 block, n=512", {"l": 0, "q": [], "s": [512]}, 512, False)
    if "5s" in which:
        run("5 (reduced): SDP one 128x128 block, n=128", {"l": 0, "q": [], "s": [128]}, 128, True)
