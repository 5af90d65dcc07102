"""Nesterov-Todd scaling on the device: mirrors of the reference's `misc.compute_scaling(s, z, lmbda, dims, mnl)`
and `misc.update_scaling(W, lmbda, s, z)` (reference src/python/misc.py:250-419, :422-634) with the same
signatures and the same in-place semantics, for every cone type ('l', 'q' and 's'; 's' blocks: Cholesky and a
one-sided Jacobi SVD on the device instead of lapack.potrf / lapack.gesvd).

Swap into an unmodified CVXOPT:   cvxopt.misc.compute_scaling = cvxopt_b200.scaling.compute_scaling
                                  cvxopt.misc.update_scaling  = cvxopt_b200.scaling.update_scaling
(coneprog resolves both as misc.<name> at call time: coneprog.py:1033, 1395, 2243, 2519).

W is a dict as in the reference (coneprog.py:327-334); its entries may be cvxopt matrices or F-ordered numpy
arrays — anything writable through the buffer protocol.  compute_scaling returns W made of the type given by
`new_matrix` (default: numpy arrays; pass cvxopt.matrix to get the reference's types)."""
import ctypes as C

import numpy as np

from . import _lib
from .kkt import make_dims


def _flat_view(a, n, name):
    arr = np.asarray(a)
    if arr.dtype != np.float64:
        raise TypeError("%s must be a 'd' matrix" % name)
    flat = arr.reshape(-1, order="F") if arr.ndim > 1 else arr
    if flat.size < n or not np.shares_memory(flat, arr) or not flat.flags.writeable:
        raise TypeError("%s must be a writable contiguous buffer of at least %d entries" % (name, n))
    return flat


def _default_new(rows, cols):
    return np.zeros((rows, cols), order="F")


def compute_scaling(s, z, lmbda, dims, mnl=None, new_matrix=None):
    """Returns the Nesterov-Todd scaling W at the points s and z and stores the scaled variable in lmbda:
    W*z = W^{-T}*s = lmbda.  misc.py:250-419."""
    lib = _lib.load()
    nonlinear = mnl is not None
    mnl = int(mnl) if nonlinear else 0
    ml, q, sd = int(dims["l"]), [int(k) for k in dims["q"]], [int(k) for k in dims["s"]]
    cd, keep, cdim, _ = make_dims({"l": ml, "q": q, "s": sd}, mnl)
    nlam = mnl + ml + sum(q) + sum(sd)
    sv, zv = _flat_view(s, cdim, "s"), _flat_view(z, cdim, "z")
    lv = _flat_view(lmbda, nlam, "lmbda")
    # flat outputs, scattered into the W dict afterwards
    out = {k: np.zeros(max(1, n)) for k, n in (("dnl", mnl), ("dnli", mnl), ("d", ml), ("di", ml), ("v", sum(q)),
                                               ("beta", len(q)), ("r", sum(k * k for k in sd)),
                                               ("rti", sum(k * k for k in sd)))}
    sc = _lib.Scaling(*(out[k].ctypes.data for k in ("dnl", "dnli", "d", "di", "v", "beta", "r", "rti")))
    rc = lib.cvxb_compute_scaling(sv.ctypes.data, zv.ctypes.data, lv.ctypes.data, C.byref(cd), C.byref(sc), _lib.HOST)
    _lib.check(rc, "compute_scaling")
    new = new_matrix or _default_new

    def mat(flat, rows, cols):
        m = new(rows, cols)
        np.asarray(m).reshape(-1, order="F")[:] = flat
        return m
    W = {}
    if nonlinear:
        W["dnl"], W["dnli"] = mat(out["dnl"][:mnl], mnl, 1), mat(out["dnli"][:mnl], mnl, 1)
    W["d"], W["di"] = mat(out["d"][:ml], ml, 1), mat(out["di"][:ml], ml, 1)
    W["v"], W["beta"], o = [], [], 0
    for k, m in enumerate(q):
        W["v"].append(mat(out["v"][o:o + m], m, 1))
        W["beta"].append(float(out["beta"][k]))
        o += m
    W["r"], W["rti"], o = [], [], 0
    for m in sd:
        W["r"].append(mat(out["r"][o:o + m * m], m, m))
        W["rti"].append(mat(out["rti"][o:o + m * m], m, m))
        o += m * m
    return W


def update_scaling(W, lmbda, s, z):
    """Updates W and lmbda in place from the new iterates in the current scaling (s, z: nonlinear/'l'/'q' rows;
    Cholesky factors Ls, Lz in the 's' blocks), overwriting s and z as the reference does.  misc.py:422-634."""
    lib = _lib.load()
    mnl = np.asarray(W["dnl"]).size if "dnl" in W else 0
    ml = np.asarray(W["d"]).size
    q = [np.asarray(v).size for v in W["v"]]
    sd = [np.asarray(r).shape[0] for r in W["r"]]
    cd, keep, cdim, _ = make_dims({"l": ml, "q": q, "s": sd}, mnl)
    nlam = mnl + ml + sum(q) + sum(sd)
    sv, zv = _flat_view(s, cdim, "s"), _flat_view(z, cdim, "z")
    lv = _flat_view(lmbda, nlam, "lmbda")

    def cat(items):
        parts = [np.asarray(m, dtype=np.float64).reshape(-1, order="F") for m in items]
        return np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(1)
    flat = {"dnl": cat([W["dnl"]]) if mnl else np.zeros(1), "dnli": cat([W["dnli"]]) if mnl else np.zeros(1),
            "d": cat([W["d"]]) if ml else np.zeros(1), "di": cat([W["di"]]) if ml else np.zeros(1),
            "v": cat(W["v"]) if q else np.zeros(1),
            "beta": np.array([float(b) for b in W["beta"]] or [0.0]),
            "r": cat(W["r"]) if sd else np.zeros(1), "rti": cat(W["rti"]) if sd else np.zeros(1)}
    sc = _lib.Scaling(*(flat[k].ctypes.data for k in ("dnl", "dnli", "d", "di", "v", "beta", "r", "rti")))
    rc = lib.cvxb_update_scaling(C.byref(sc), lv.ctypes.data, sv.ctypes.data, zv.ctypes.data, C.byref(cd), _lib.HOST)
    _lib.check(rc, "update_scaling")

    def put(dst, src):
        np.asarray(dst).reshape(-1, order="F")[:] = src
    if mnl:
        put(W["dnl"], flat["dnl"][:mnl]); put(W["dnli"], flat["dnli"][:mnl])
    if ml:
        put(W["d"], flat["d"][:ml]); put(W["di"], flat["di"][:ml])
    o = 0
    for k, m in enumerate(q):
        put(W["v"][k], flat["v"][o:o + m])
        W["beta"][k] = float(flat["beta"][k])
        o += m
    o = 0
    for k, m in enumerate(sd):
        put(W["r"][k], flat["r"][o:o + m * m]); put(W["rti"][k], flat["rti"][o:o + m * m])
        o += m * m
