"""CPU study for the round-2 plan: left-looking blocked Cholesky whose block-column
updates  K[j:, j] -= L[j:, :j] L[j, :j]'  are evaluated from int8 digit slices of the ROWS of L (exact integer
products, 9 radix-2^7 digits, row exponent fixed in advance by |L[r,c]| <= sqrt(K[r,r])).  Question: is the
factorisation as backward-stable as the fp64 one on IPM-like normal-equation matrices?

Usage: python tools/chol_i8_study.py [n] [W]"""
import sys
import numpy as np


def row_slices(Lblk, e, s=9):
    """digits of the rows of Lblk (rows x cols) with per-row exponent e: L[r,c] = 2^e_r sum_t q_t 2^-(6+7t)"""
    x = Lblk * np.ldexp(64.0, -e)[:, None]
    out = []
    for _ in range(s):
        q = np.rint(x)
        x = (x - q) * 128.0
        out.append(q.astype(np.int64))
    assert max(int(np.max(np.abs(q), initial=0)) for q in out) <= 64
    return out


def sliced_product(QA, QB, eA, eB, s=9):
    """sum over levels d <= s-1 of exact integer products, combined like the kernel's epilogue"""
    acc = np.zeros((QA[0].shape[0], QB[0].shape[0]))
    for d in range(s - 1, -1, -1):
        lev = np.zeros(acc.shape, dtype=np.int64)
        for a in range(d + 1):
            lev += QA[a] @ QB[d - a].T
        acc += lev.astype(np.float64) * 2.0 ** (-12 - 7 * d)
    return acc * np.ldexp(1.0, eA)[:, None] * np.ldexp(1.0, eB)[None, :]


def chol_i8(K, W, s=9):
    n = K.shape[0]
    A = np.tril(K).copy()
    _, e = np.frexp(np.sqrt(np.diag(K)))               # |L[r, c]| <= sqrt(K[r, r]) < 2^e_r
    e = e.astype(np.int64)
    Q = [np.zeros((n, 0), dtype=np.int64) for _ in range(s)]
    for j in range(0, n, W):
        w = min(W, n - j)
        if j > 0:
            QA = [q[j:, :] for q in Q]
            QB = [q[j:j + w, :] for q in Q]
            A[j:, j:j + w] -= sliced_product(QA, QB, e[j:], e[j:j + w], s)
        # the block column itself in fp64 (the existing right-looking code)
        D = A[j:j + w, j:j + w]
        D = np.tril(D) + np.tril(D, -1).T
        Ljj = np.linalg.cholesky(D)
        A[j:j + w, j:j + w] = Ljj
        if j + w < n:
            A[j + w:, j:j + w] = np.linalg.solve(Ljj, A[j + w:, j:j + w].T).T
        new = row_slices(np.tril(A[:, j:j + w], 0) if False else A[:, j:j + w] * (np.arange(n)[:, None] >= j), e, s)
        Q = [np.concatenate([q, t], axis=1) for q, t in zip(Q, new)]
    return np.tril(A)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rng = np.random.Generator(np.random.PCG64(0))
    print("| scaling spread | cond(K) | ||K-LL'||/||K|| fp64 | int8-slice updates | componentwise (|K-LL'| / |L||L'|) fp64 | int8 | solve rel diff |")
    print("|---|---|---|---|---|---|---|")
    for spread in (0, 4, 8):
        m = 2 * n
        G = rng.standard_normal((m, n))
        di = 10.0 ** rng.uniform(-spread / 2, spread / 2, m)
        B = rng.standard_normal((n, n))
        K = (G * di[:, None]).T @ (G * di[:, None]) + B @ B.T / n + np.eye(n)
        K = (K + K.T) / 2
        L0 = np.linalg.cholesky(K)
        L1 = chol_i8(K, W)
        nk = np.linalg.norm(K)
        Kl = K.astype(np.longdouble)
        def berr(L):
            Ll = L.astype(np.longdouble)
            R = np.abs(Kl - Ll @ Ll.T)
            return float(np.linalg.norm(R.astype(np.float64)) / nk), float((R / (np.abs(Ll) @ np.abs(Ll).T)).max())
        b0, c0 = berr(L0)
        b1, c1 = berr(L1)
        rhs = rng.standard_normal(n)
        x0 = np.linalg.solve(L0.T, np.linalg.solve(L0, rhs))
        x1 = np.linalg.solve(L1.T, np.linalg.solve(L1, rhs))
        print("| 1e%d | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (spread, np.linalg.cond(K), b0, b1, c0, c1,
              np.linalg.norm(x0 - x1) / np.linalg.norm(x0)))


if __name__ == "__main__":
    main()
