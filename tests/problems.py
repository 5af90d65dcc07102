"""Synthetic inputs of SURVEY.md §8(d): numpy Generator(PCG64(seed)) only."""
import numpy as np


def dense_qp(n, m=None, seed=1):
    """min 1/2 x'Px + q'x  s.t. Gx <= h, strictly feasible.  P = A0'A0/n + I."""
    m = 2 * n if m is None else m
    rng = np.random.Generator(np.random.PCG64(seed))
    A0 = rng.standard_normal((n, n))
    P = np.asfortranarray(A0.T @ A0 / n + np.eye(n))
    q = rng.standard_normal(n)
    G = np.asfortranarray(rng.standard_normal((m, n)))
    x0 = rng.standard_normal(n)
    h = G @ x0 + rng.uniform(0.1, 1.1, m)
    return P, q, G, h


def cone_point(dims, rng):
    """a strictly interior point of the cone product (unpacked 's' storage)."""
    parts = [rng.uniform(0.5, 2.0, dims["l"])]
    for k in dims["q"]:
        u = rng.standard_normal(k)
        u[0] = np.linalg.norm(u[1:]) + 1.0
        parts.append(u)
    for k in dims["s"]:
        B = rng.standard_normal((k, k))
        S = B @ B.T / max(k, 1) + np.eye(k)
        parts.append(S.reshape(-1, order="F"))
    return np.concatenate(parts) if parts else np.zeros(0)


def cone_dim(dims):
    return dims["l"] + sum(dims["q"]) + sum(k * k for k in dims["s"])


def cone_lp(n, dims, seed=11):
    """min c'x s.t. Gx + s = h, s in K; primal and dual strictly feasible."""
    rng = np.random.Generator(np.random.PCG64(seed))
    K = cone_dim(dims)
    G = rng.standard_normal((K, n))
    # symmetrise the 's' columns so every column of G is a symmetric matrix
    off = dims["l"] + sum(dims["q"])
    for k in dims["s"]:
        for j in range(n):
            M = G[off:off + k * k, j].reshape(k, k, order="F")
            G[off:off + k * k, j] = ((M + M.T) / 2).reshape(-1, order="F")
        off += k * k
    G = np.asfortranarray(G)
    x0 = rng.standard_normal(n)
    s0 = cone_point(dims, rng)
    z0 = cone_point(dims, rng)
    h = G @ x0 + s0
    c = -sgemv_t(G, z0, dims)
    return c, G, h


def sgemv_t(G, z, dims):
    """G' z in the trace inner product (lower triangles of 's' blocks count twice off-diagonal),
    i.e. misc.sgemv(trans='T') — reference src/python/misc.py:801-832."""
    z = z.copy()
    off = dims["l"] + sum(dims["q"])
    for k in dims["s"]:
        M = z[off:off + k * k].reshape(k, k, order="F")
        M = np.tril(M) + np.tril(M, -1).T
        z[off:off + k * k] = M.reshape(-1, order="F")
        off += k * k
    # with symmetric columns of G the plain dot product equals the trace inner product
    return G.T @ z


def random_scaling(dims, seed=0):
    """(W, lmbda) from a random interior pair (s, z) through the oracle's compute_scaling."""
    import kkt_oracle
    rng = np.random.Generator(np.random.PCG64(seed))
    s = cone_point(dims, rng)
    z = cone_point(dims, rng)
    lm = np.zeros(dims["l"] + sum(dims["q"]) + sum(dims["s"]))
    W = kkt_oracle.compute_scaling(s, z, lm, dims)
    return W, lm
