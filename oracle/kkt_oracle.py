"""TEST INFRASTRUCTURE — CPU restatement (numpy/scipy) of the reference's KKT hot path.

This module is the *oracle*: a plain restatement of what CVXOPT computes on the
path  NT scaling -> normal equations -> Cholesky -> solves.  It is imported only
by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline arm; nothing under
cvxopt_b200/ may import it.  Parity pinning: tests/test_oracle_vs_reference.py
checks every function here against the reference itself (oracle/_ref, built from
/root/reference by oracle/build_ref.sh) and against the committed golden vectors
in tests/golden/ (generated from the reference by tests/golden/make_golden.py).

All citations are relative to /root/reference.  Vectors/matrices are numpy fp64,
column-major; a "cone vector" is laid out [mnl | l | q blocks | s blocks (n^2 each)].
"""
import math

import numpy as np
import scipy.linalg as sla


def cone_sizes(dims, mnl=0):
    ml, q, s = int(dims["l"]), list(dims["q"]), list(dims["s"])
    cdim = mnl + ml + sum(q) + sum(k * k for k in s)
    cdim_pckd = mnl + ml + sum(q) + sum(k * (k + 1) // 2 for k in s)
    return ml, q, s, cdim, cdim_pckd


# --------------------------------------------------------------------------- scale
def scale(x, W, trans="N", inverse="N"):
    """In place x := W x, W' x, W^{-1} x, W^{-T} x.  src/C/misc_solvers.c:85-244
    (python twin src/python/misc.py:30-164).  x: (rows, cols) F-ordered ndarray."""
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    ind = 0
    if "dnl" in W:                                      # misc_solvers.c:117-124
        w = np.asarray(W["dnl"] if inverse == "N" else W["dnli"]).reshape(-1)
        x[ind:ind + w.size, :] *= w[:, None]
        ind += w.size
    w = np.asarray(W["d"] if inverse == "N" else W["di"]).reshape(-1)   # :132-141
    x[ind:ind + w.size, :] *= w[:, None]
    ind += w.size
    for k, vk in enumerate(W["v"]):                     # :158-183
        v = np.asarray(vk).reshape(-1)
        m = v.size
        blk = x[ind:ind + m, :]
        if inverse == "I":
            blk[0, :] *= -1.0
        w = v @ blk
        blk[0, :] *= -1.0
        blk += 2.0 * np.outer(v, w)
        if inverse == "I":
            blk[0, :] *= -1.0
            a = 1.0 / float(W["beta"][k])
        else:
            a = float(W["beta"][k])
        blk *= a
        ind += m
    rs = W["r"] if inverse == "N" else W["rti"]         # :203-240
    for rk in rs:
        r = np.asarray(rk)
        n = r.shape[0]
        right = (inverse == "N" and trans == "T") or (inverse == "I" and trans == "N")
        for i in range(x.shape[1]):
            X = x[ind:ind + n * n, i].reshape(n, n, order="F")
            Lx = np.tril(X).copy()
            Lx[np.diag_indices(n)] *= 0.5               # :219
            if right:                                   # wrk = r * tril(x);  x = r wrk' + wrk r'
                wrk = r @ Lx
                Y = r @ wrk.T + wrk @ r.T
            else:                                       # wrk = tril(x) * r;  x = r' wrk + wrk' r
                wrk = Lx @ r
                Y = r.T @ wrk + wrk.T @ r
            il = np.tril_indices(n)
            X[il] = Y[il]                               # dsyr2k 'L': upper triangle untouched
            x[ind:ind + n * n, i] = X.reshape(-1, order="F")
        ind += n * n
    return x


# --------------------------------------------------------------------------- pack / unpack
def pack(x, y, dims, mnl=0, offsetx=0, offsety=0):
    """misc_solvers.c:412-465 (vector version; diagonal goes through /sqrt2 then *sqrt2)."""
    ml, q, s, _, _ = cone_sizes(dims, mnl)
    nlq = mnl + ml + sum(q)
    y[offsety:offsety + nlq] = x[offsetx:offsetx + nlq]
    iu, ip = offsetx + nlq, offsety + nlq
    ip0 = ip
    for n in s:
        for k in range(n):
            ln = n - k
            y[ip:ip + ln] = x[iu + k * (n + 1): iu + k * (n + 1) + ln]
            y[ip] /= math.sqrt(2.0)
            ip += ln
        iu += n * n
    y[ip0:ip] *= math.sqrt(2.0)
    return y


def pack2(x, dims, mnl=0):
    """misc_solvers.c:476-541: in place on the columns of x (rows compacted)."""
    ml, q, s, _, _ = cone_sizes(dims, mnl)
    nlq = mnl + ml + sum(q)
    iu = ip = nlq
    a = math.sqrt(2.0)
    for n in s:
        for k in range(n):
            ln = n - k
            wrk = x[iu + k * (n + 1): iu + k * (n + 1) + ln, :].copy()
            wrk[1:, :] *= a
            x[ip:ip + ln, :] = wrk
            ip += ln
        iu += n * n
    return x


def unpack(x, y, dims, mnl=0, offsetx=0, offsety=0):
    """misc_solvers.c:552-601."""
    ml, q, s, _, _ = cone_sizes(dims, mnl)
    m = mnl + ml + sum(q)
    y[offsety:offsety + m] = x[offsetx:offsetx + m]
    ip, iu = offsetx + m, offsety + m
    a = 1.0 / math.sqrt(2.0)
    for n in s:
        for k in range(n):
            ln = n - k
            y[iu + k * (n + 1): iu + k * (n + 1) + ln] = x[ip:ip + ln]
            ip += ln
            y[iu + k * (n + 1) + 1: iu + k * (n + 1) + ln] *= a
        iu += n * n
    return y


def symm(x, n, offset=0):
    """misc_solvers.c:610-625: fill the upper triangle from the lower one."""
    X = x[offset:offset + n * n].reshape(n, n, order="F")
    il = np.tril_indices(n, -1)
    X.T[il] = X[il]
    return x


# --------------------------------------------------------------------------- hyperbolic helpers
def jnrm2(x, n=None, offset=0):
    """sqrt(x' J x), J = diag(1,-I).  src/python/misc.py:848-856"""
    if n is None:
        n = len(x)
    a = np.linalg.norm(x[offset + 1: offset + n])
    return math.sqrt(x[offset] - a) * math.sqrt(x[offset] + a)


def jdot(x, y, n=None, offsetx=0, offsety=0):
    """x' J y.  misc.py:835-845"""
    if n is None:
        n = len(x)
    return x[offsetx] * y[offsety] - float(
        np.dot(x[offsetx + 1: offsetx + n], y[offsety + 1: offsety + n]))


# --------------------------------------------------------------------------- NT scaling
def compute_scaling(s, z, lmbda, dims, mnl=None):
    """misc.py:250-419.  Returns the dict W; writes lmbda."""
    W = {}
    if mnl is None:
        mnl = 0
    else:
        W["dnl"] = np.sqrt(s[:mnl] / z[:mnl])
        W["dnli"] = W["dnl"] ** -1
        lmbda[:mnl] = np.sqrt(s[:mnl] * z[:mnl])
    m = dims["l"]
    W["d"] = np.sqrt(s[mnl:mnl + m] / z[mnl:mnl + m])
    W["di"] = W["d"] ** -1
    lmbda[mnl:mnl + m] = np.sqrt(s[mnl:mnl + m] * z[mnl:mnl + m])
    ind = mnl + dims["l"]
    W["v"] = [np.zeros(k) for k in dims["q"]]
    W["beta"] = len(dims["q"]) * [0.0]
    for k in range(len(dims["q"])):
        m = dims["q"][k]
        v = W["v"][k]
        aa = jnrm2(s, offset=ind, n=m)
        bb = jnrm2(z, offset=ind, n=m)
        W["beta"][k] = math.sqrt(aa / bb)
        cc = math.sqrt((float(np.dot(s[ind:ind + m], z[ind:ind + m])) / aa / bb + 1.0) / 2.0)
        v[:] = z[ind:ind + m]
        v *= -1.0 / bb
        v[0] *= -1.0
        v += (1.0 / aa) * s[ind:ind + m]
        v *= 1.0 / 2.0 / cc
        v[0] += 1.0
        v *= 1.0 / math.sqrt(2.0 * v[0])
        lmbda[ind] = cc
        dd = 2 * cc + s[ind] / aa + z[ind] / bb
        lmbda[ind + 1: ind + m] = s[ind + 1: ind + m]
        lmbda[ind + 1: ind + m] *= (cc + z[ind] / bb) / dd / aa
        lmbda[ind + 1: ind + m] += ((cc + s[ind] / aa) / dd / bb) * z[ind + 1: ind + m]
        lmbda[ind: ind + m] *= math.sqrt(aa * bb)
        ind += m
    W["r"] = [np.zeros((m, m), order="F") for m in dims["s"]]
    W["rti"] = [np.zeros((m, m), order="F") for m in dims["s"]]
    ind2 = ind
    for k in range(len(dims["s"])):
        m = dims["s"][k]
        Ls = np.linalg.cholesky(_symL(s[ind2:ind2 + m * m], m))      # :386-391
        Lz = np.linalg.cholesky(_symL(z[ind2:ind2 + m * m], m))
        U, sv, _ = np.linalg.svd(Lz.T @ Ls)                          # :397-399
        lmbda[ind:ind + m] = sv
        r = sla.solve_triangular(Lz, U, lower=True, trans="T")       # :402-403
        rti = Lz @ U                                                 # :406-407
        a = np.sqrt(sv)
        W["r"][k][:, :] = r * a[None, :]
        W["rti"][k][:, :] = rti / a[None, :]
        ind += m
        ind2 += m * m
    return W


def _symL(x, m):
    X = np.array(x, dtype=float).reshape(m, m, order="F")
    return np.tril(X) + np.tril(X, -1).T


def update_scaling(W, lmbda, s, z):
    """misc.py:422-634.  In place on W (numpy arrays / lists of floats), lmbda, s, z.  The 's' components of s, z
    hold the Cholesky factors Ls, Lz of the new iterates in the current scaling (zero above the diagonal)."""
    mnl = len(W["dnl"]) if "dnl" in W else 0
    ml = len(W["d"])
    m = mnl + ml
    s[:m] = np.sqrt(s[:m])                                            # :450-451
    z[:m] = np.sqrt(z[:m])
    if "dnl" in W:                                                    # :454-457
        W["dnl"][:] = W["dnl"] * s[:mnl] / z[:mnl]
        W["dnli"][:] = W["dnl"] ** -1
    W["d"][:] = W["d"] * s[mnl:m] / z[mnl:m]                          # :458-460
    W["di"][:] = W["d"] ** -1
    lmbda[:m] = s[:m] * z[:m]                                         # :463-464
    ind = m
    for k in range(len(W["v"])):                                      # :504-573
        v = W["v"][k]
        mk = len(v)
        aa = jnrm2(s, offset=ind, n=mk)
        s[ind:ind + mk] *= 1.0 / aa
        bb = jnrm2(z, offset=ind, n=mk)
        z[ind:ind + mk] *= 1.0 / bb
        sk, zk = s[ind:ind + mk], z[ind:ind + mk]
        cc = math.sqrt((1.0 + float(np.dot(sk, zk))) / 2.0)
        vs = float(np.dot(v, sk))
        vz = jdot(v, z, offsety=ind, n=mk)
        vq = (vs + vz) / 2.0 / cc
        vu = vs - vz
        lmbda[ind] = cc
        wk0 = 2 * v[0] * vq - (sk[0] + zk[0]) / 2.0 / cc
        dd = (v[0] * vu - sk[0] / 2.0 + zk[0] / 2.0) / (wk0 + 1.0)
        lmbda[ind + 1: ind + mk] = v[1:] * (2.0 * (-dd * vq + 0.5 * vu))
        lmbda[ind + 1: ind + mk] += 0.5 * (1.0 - dd / cc) * sk[1:]
        lmbda[ind + 1: ind + mk] += 0.5 * (1.0 + dd / cc) * zk[1:]
        lmbda[ind: ind + mk] *= math.sqrt(aa * bb)
        v *= 2.0 * vq
        v[0] -= sk[0] / 2.0 / cc
        v[1:] += (0.5 / cc) * sk[1:]
        v += (-0.5 / cc) * zk
        v[0] += 1.0
        v *= 1.0 / math.sqrt(2.0 * v[0])
        W["beta"][k] *= math.sqrt(aa / bb)
        ind += mk
    ind2 = ind
    for k in range(len(W["r"])):                                      # :592-634
        r, rti = W["r"][k], W["rti"][k]
        mk = r.shape[0]
        Ls = s[ind2:ind2 + mk * mk].reshape(mk, mk, order="F").copy()
        Lz = z[ind2:ind2 + mk * mk].reshape(mk, mk, order="F").copy()
        r[:, :] = r @ Ls
        rti[:, :] = rti @ Lz
        U, sv, Vt = np.linalg.svd(Lz.T @ Ls)
        lmbda[ind:ind + mk] = sv
        s[ind2:ind2 + mk * mk] = U.reshape(-1, order="F")             # U in sk, V' in zk (:611-613)
        z[ind2:ind2 + mk * mk] = Vt.reshape(-1, order="F")
        a = 1.0 / np.sqrt(sv)
        r[:, :] = (r @ Vt.T) * a[None, :]
        rti[:, :] = (rti @ U) * a[None, :]
        ind += mk
        ind2 += mk * mk


# --------------------------------------------------------------------------- kkt_chol
class KktChol:
    """misc.kkt_chol (misc.py:1213-1349): factor(W, H, Df) -> solve(x, y, z).

    p == 0 follows kkt_chol line by line.  For p > 0 the equality constraints are eliminated
    with the Schur complement K = A S^{-1} A' of kkt_chol2 (misc.py:1464-1565, including its
    `S += A'A` fallback when S is singular, :1433-1447) instead of kkt_chol's QR of A'
    (:1244-1250, 1278-1339): the two are the same linear system; tests pin this class to the
    reference's QR-based kkt_chol."""

    def __init__(self, G, dims, A=None, mnl=0):
        self.G = np.asfortranarray(np.asarray(G, dtype=float))
        self.dims = dims
        self.mnl = mnl
        self.n = self.G.shape[1]
        _, _, _, self.cdim, self.cdim_pckd = cone_sizes(dims, mnl)
        self.A = None
        if A is not None and np.asarray(A).shape[0] > 0:
            self.A = np.asarray(A, dtype=float)
        self.singular = False
        self.first = True

    def factor(self, W, H=None, Df=None):
        n, mnl = self.n, self.mnl
        Gs = np.zeros((self.cdim, n), order="F")                     # :1252
        if mnl:
            Gs[:mnl, :] = Df
        Gs[mnl:, :] = self.G                                         # :1270
        scale(Gs, W, trans="T", inverse="I")                         # :1271
        pack2(Gs, self.dims, mnl)                                    # :1272
        Gp = Gs[:self.cdim_pckd, :]
        # blas.syrk(Gs, K, trans='T', k=cdim_pckd): lower triangle only  (:1275)
        K = sla.blas.dsyrk(1.0, Gp, trans=1, lower=1) if Gp.shape[0] else np.zeros((n, n), order="F")
        if H is not None:
            K += np.tril(np.asarray(H))                              # K[:n,:n] += H  (:1276, lower part matters)
        K = K + np.tril(K, -1).T                                     # misc.symm  (:1277)
        self.K1norm = float(np.abs(K).sum(axis=0).max()) if n else 0.0    # for condition estimates in tests
        if self.A is not None and self.singular:
            K = K + self.A.T @ self.A
        try:
            L = np.linalg.cholesky(K)                                # lapack.potrf  :1282
        except np.linalg.LinAlgError:
            if self.A is not None and self.first and not self.singular:
                self.singular = True                                 # misc.py:1433-1447
                K = K + self.A.T @ self.A
                try:
                    L = np.linalg.cholesky(K)
                except np.linalg.LinAlgError:
                    raise ArithmeticError("potrf: not positive definite")
            else:
                raise ArithmeticError("potrf: not positive definite")
        self.first = False
        if self.A is not None:
            self.Asct = sla.solve_triangular(L, self.A.T, lower=True)    # :1470
            try:
                self.Lp = np.linalg.cholesky(self.Asct.T @ self.Asct)    # :1471-1472
            except np.linalg.LinAlgError:
                raise ArithmeticError("potrf: not positive definite")
        self.Gs, self.L, self.W = Gp, L, W
        return self.solve

    def solve(self, x, y, z):
        """In place (bx, by, bz) -> (ux, uy, W uz).  misc.py:1284-1345 / 1489-1563"""
        zc = z.reshape(-1, 1)
        scale(zc, self.W, trans="T", inverse="I")                    # :1306
        bzp = np.zeros(self.cdim_pckd)
        pack(z, bzp, self.dims, self.mnl)                            # :1307
        x += self.Gs.T @ bzp                                         # :1311
        if self.A is None:
            x[:] = sla.cho_solve((self.L, True), x)                  # :1327
        else:
            if self.singular:
                x += self.A.T @ y                                    # misc.py:1526-1527
            x[:] = sla.solve_triangular(self.L, x, lower=True)       # :1529
            y[:] = self.Asct.T @ x - y                               # :1541
            y[:] = sla.cho_solve((self.Lp, True), y)                 # :1543
            x -= self.Asct @ y                                       # :1553
            x[:] = sla.solve_triangular(self.L, x, lower=True, trans="T")   # :1555
        bzp = self.Gs @ x - bzp                                      # :1344
        unpack(bzp, z, self.dims, self.mnl)                          # :1345


# --------------------------------------------------------------------------- IPM-side cone algebra
def _nlq(dims, mnl=0):
    return mnl + dims["l"] + sum(dims["q"])


def scale2(lmbda, x, dims, mnl=0, inverse="N"):
    """misc_solvers.c:256-401 (python twin misc.py:170-247)."""
    m = mnl + dims["l"]
    if inverse == "N":
        x[:m] /= lmbda[:m]
    else:
        x[:m] *= lmbda[:m]
    for mk in dims["q"]:
        nrm = np.linalg.norm(lmbda[m + 1:m + mk])
        a = math.sqrt(lmbda[m] + nrm) * math.sqrt(lmbda[m] - nrm)
        if inverse == "N":
            lx = (lmbda[m] * x[m] - float(np.dot(lmbda[m + 1:m + mk], x[m + 1:m + mk]))) / a
        else:
            lx = float(np.dot(lmbda[m:m + mk], x[m:m + mk])) / a
        x0 = x[m]
        x[m] = lx
        b = (x0 + lx) / (lmbda[m] / a + 1.0) / a
        if inverse == "N":
            b *= -1.0
        x[m + 1:m + mk] += b * lmbda[m + 1:m + mk]
        x[m:m + mk] *= (1.0 / a) if inverse == "N" else a
        m += mk
    ind2 = m
    for mk in dims["s"]:
        sql = np.sqrt(lmbda[ind2:ind2 + mk])
        for j in range(mk):
            c = sql * math.sqrt(lmbda[ind2 + j])
            if inverse == "N":
                x[m + j * mk:m + (j + 1) * mk] /= c
            else:
                x[m + j * mk:m + (j + 1) * mk] *= c
        m += mk * mk
        ind2 += mk
    return x


def sprod(x, y, dims, mnl=0, diag="N"):
    """x := y o x.  misc_solvers.c:634-767"""
    ind = mnl + dims["l"]
    x[:ind] *= y[:ind]
    for mk in dims["q"]:
        a = float(np.dot(y[ind:ind + mk], x[ind:ind + mk]))
        x0 = x[ind]
        x[ind + 1:ind + mk] = y[ind] * x[ind + 1:ind + mk] + x0 * y[ind + 1:ind + mk]
        x[ind] = a
        ind += mk
    if diag == "N":
        for mk in dims["s"]:
            A = _symL(x[ind:ind + mk * mk], mk)
            Y = _symL(y[ind:ind + mk * mk], mk)
            R = 0.5 * (A @ Y + Y @ A)
            X = x[ind:ind + mk * mk].reshape(mk, mk, order="F")
            il = np.tril_indices(mk)
            X[il] = R[il]
            x[ind:ind + mk * mk] = X.reshape(-1, order="F")
            ind += mk * mk
    else:
        ind2 = ind
        for mk in dims["s"]:
            yk = y[ind2:ind2 + mk]
            for k in range(mk):
                x[ind + k * (mk + 1): ind + k * (mk + 1) + mk - k] *= 0.5 * (yk[k:] + yk[k])
            ind += mk * mk
            ind2 += mk
    return x


def sinv(x, y, dims, mnl=0):
    """x := y o\\ x.  misc_solvers.c:775-878"""
    ind = mnl + dims["l"]
    x[:ind] /= y[:ind]
    for mk in dims["q"]:
        nrm = np.linalg.norm(y[ind + 1:ind + mk])
        a = (y[ind] + nrm) * (y[ind] - nrm)
        c = x[ind]
        d = float(np.dot(x[ind + 1:ind + mk], y[ind + 1:ind + mk]))
        x[ind] = c * y[ind] - d
        x[ind + 1:ind + mk] *= a / y[ind]
        x[ind + 1:ind + mk] += (d / y[ind] - c) * y[ind + 1:ind + mk]
        x[ind:ind + mk] *= 1.0 / a
        ind += mk
    ind2 = ind
    for mk in dims["s"]:
        yk = y[ind2:ind2 + mk]
        for k in range(mk):
            x[ind + k * (mk + 1): ind + k * (mk + 1) + mk - k] /= 0.5 * (yk[k:] + yk[k])
        ind += mk * mk
        ind2 += mk
    return x


def trisc(x, dims, offset=0):
    """misc_solvers.c:887-935"""
    ox = offset + dims["l"] + sum(dims["q"])
    for nk in dims["s"]:
        X = x[ox:ox + nk * nk].reshape(nk, nk, order="F")
        X[np.triu_indices(nk, 1)] = 0.0
        X[np.tril_indices(nk, -1)] *= 2.0
        x[ox:ox + nk * nk] = X.reshape(-1, order="F")
        ox += nk * nk
    return x


def triusc(x, dims, offset=0):
    """misc_solvers.c:940-986"""
    ox = offset + dims["l"] + sum(dims["q"])
    for nk in dims["s"]:
        X = x[ox:ox + nk * nk].reshape(nk, nk, order="F")
        X[np.tril_indices(nk, -1)] *= 0.5
        x[ox:ox + nk * nk] = X.reshape(-1, order="F")
        ox += nk * nk
    return x


def sdot(x, y, dims, mnl=0):
    """misc_solvers.c:991-1039"""
    m = _nlq(dims, mnl)
    a = float(np.dot(x[:m], y[:m]))
    for nk in dims["s"]:
        X = x[m:m + nk * nk].reshape(nk, nk, order="F")
        Y = y[m:m + nk * nk].reshape(nk, nk, order="F")
        a += float(np.dot(np.diag(X), np.diag(Y)))
        il = np.tril_indices(nk, -1)
        a += 2.0 * float(np.dot(X[il], Y[il]))
        m += nk * nk
    return a


def max_step(x, dims, mnl=0, sigma=None):
    """min {t | x + t e >= 0}, misc_solvers.c:1052-1153.  With sigma: eigenvalues of the 's' blocks
    -> sigma, eigenvectors -> the 's' blocks of x (dsyevd_ 'V', :1132-1136)."""
    ind = mnl + dims["l"]
    t = -np.finfo(np.float32).max
    if ind:
        t = max(t, float(np.max(-x[:ind])))
    for mk in dims["q"]:
        t = max(t, float(np.linalg.norm(x[ind + 1:ind + mk]) - x[ind]))
        ind += mk
    ind2 = 0
    for mk in dims["s"]:
        if mk:
            if sigma is None:
                w = np.linalg.eigvalsh(_symL(x[ind:ind + mk * mk], mk))
            else:
                w, Q = np.linalg.eigh(_symL(x[ind:ind + mk * mk], mk))
                sigma[ind2:ind2 + mk] = w
                x[ind:ind + mk * mk] = Q.reshape(-1, order="F")
            t = max(t, -float(w[0]))
        ind += mk * mk
        ind2 += mk
    return t if ind else 0.0
