"""CPU study for the round-2 plan: fp64 SYRK  K = A'A  (A = diag(di) G) emulated with s signed 7-bit
integer slices per column (Ozaki splitting, every slice product exact in int32) — how many slices
are needed to match the fp64 DMMA result on IPM-like scalings?

Error measure: max_ij |K_emul - K_exact| / (|A|'|A|)_ij  (componentwise, the bound fp64 summation
itself satisfies with ~m*2^-53), K_exact in 80-bit long double."""
import sys
import numpy as np


def slices(A, s, w=7):
    """A (m x n) -> (e_j, S[p] int8-valued arrays): A[:, j] ~ 2^e_j * sum_p S[p][:, j] 2^(-w p)"""
    amax = np.max(np.abs(A), axis=0)
    e = np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1          # |scaled| < 1/2
    t = A * np.exp2(-e)[None, :]
    S = []
    for _ in range(s):
        t = t * (1 << w)
        sp = np.rint(t)
        t = t - sp
        assert np.max(np.abs(sp)) <= (1 << (w - 1)), np.max(np.abs(sp))
        S.append(sp.astype(np.int64))
    return e, S


def syrk_emulated(A, s, w=7):
    e, S = slices(A, s, w)
    n = A.shape[1]
    C = np.zeros((n, n))
    # anti-diagonals d = p + q (1-based), smallest weights first; each S_p'S_q is exact (int64 here,
    # < 2^31 on the tensor cores for m <= 2^17)
    for d in range(s + 1, 1, -1):
        acc = np.zeros((n, n), dtype=np.int64)
        for p in range(1, d):
            q = d - p
            if p <= s and q <= s:
                acc += S[p - 1].T @ S[q - 1]
        C += acc.astype(np.float64) * np.exp2(-w * d)
    return C * np.exp2(e)[:, None] * np.exp2(e)[None, :]


def main():
    m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 48)
    rng = np.random.Generator(np.random.PCG64(0))
    G = rng.standard_normal((m, n))
    print("| scaling spread | fp64 dot | " + " | ".join("s=%d (%d products)" % (s, s * (s + 1) // 2) for s in range(6, 11)) + " |")
    print("|---|---|" + "---|" * 5)
    for spread in (0, 4, 8, 12):
        di = 10.0 ** rng.uniform(-spread / 2, spread / 2, m)
        if spread >= 8:                      # late IPM: a handful of active constraints dominate
            di[rng.choice(m, 8, replace=False)] *= 10.0 ** (spread / 2)
        A = di[:, None] * G
        Al = A.astype(np.longdouble)
        K_exact = Al.T @ Al
        bound = (np.abs(Al).T @ np.abs(Al)).astype(np.float64)
        err = lambda K: float(np.max(np.abs(K.astype(np.longdouble) - K_exact).astype(np.float64) / bound))   # noqa: E731
        row = ["1e%d" % spread, "%.1e" % err(A.T @ A)]
        for s in range(6, 11):
            row.append("%.1e" % err(syrk_emulated(A, s)))
        print("| " + " | ".join(row) + " |")


if __name__ == "__main__":
    main()
