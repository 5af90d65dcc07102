"""Host-side mirror of the reference's `cvxopt.misc_solvers` C extension
(reference src/C/misc_solvers.c:1155-1173): same names, keyword lists and in-place
semantics, executed by the CUDA library on the buffers of the arguments.

Swap into an unmodified CVXOPT with e.g. `cvxopt.misc.scale = cvxopt_b200.misc_solvers.scale`
(misc.kkt_* and coneprog resolve `misc.<name>` at call time).
"""
import ctypes as C

import numpy as np

from . import _lib
from .kkt import make_dims, make_scaling


def _buf(x, name="x"):
    a = np.asarray(x)
    if a.dtype != np.float64:
        raise TypeError("%s must be a 'd' matrix" % name)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if not a.flags.f_contiguous or not a.flags.writeable:
        raise TypeError("%s must be a writable column-major buffer" % name)
    return a


def _dims_of_W(W):
    """the reference's scale() takes no `dims`: sizes come from W itself (misc_solvers.c:117-214)"""
    mnl = len(np.asarray(W["dnl"]).reshape(-1)) if "dnl" in W else 0
    if "d" not in W or "di" not in W:
        raise KeyError("missing item W['d'] or W['di']")
    ml = np.asarray(W["d"]).size
    q = [np.asarray(v).size for v in W["v"]]
    s = [np.asarray(r).shape[0] for r in W["r"]]
    return {"l": ml, "q": q, "s": s}, mnl


def scale(x, W, trans="N", inverse="N"):
    """x := W*x, W'*x, W^{-1}*x, W^{-T}*x in place.  misc_solvers.c:85-244"""
    lib = _lib.load()
    a = _buf(x)
    dims, mnl = _dims_of_W(W)
    cd, keep, cdim, _ = make_dims(dims, mnl)
    sc, keep2 = make_scaling(W, dims["l"], dims["q"], dims["s"], mnl)
    if a.shape[0] < cdim:
        raise ValueError("x has fewer rows than the cone dimension")
    rc = lib.cvxb_scale(a.ctypes.data, a.shape[0], a.shape[1], C.byref(cd), C.byref(sc),
                        ord(trans), ord(inverse), _lib.HOST)
    _lib.check(rc, "scale")


def pack(x, y, dims, mnl=0, offsetx=0, offsety=0):
    """misc_solvers.c:412-465"""
    lib = _lib.load()
    cd, keep, cdim, cp = make_dims(dims, mnl)
    xa = _buf(x).reshape(-1, order="F")[offsetx:offsetx + cdim]
    ya = _buf(y, "y").reshape(-1, order="F")[offsety:offsety + cp]
    if xa.size != cdim or ya.size != cp:
        raise ValueError("pack: buffer too short")
    _lib.check(lib.cvxb_pack(xa.ctypes.data, ya.ctypes.data, C.byref(cd), _lib.HOST), "pack")


def unpack(x, y, dims, mnl=0, offsetx=0, offsety=0):
    """misc_solvers.c:552-601"""
    lib = _lib.load()
    cd, keep, cdim, cp = make_dims(dims, mnl)
    xa = _buf(x).reshape(-1, order="F")[offsetx:offsetx + cp]
    ya = _buf(y, "y").reshape(-1, order="F")[offsety:offsety + cdim]
    if xa.size != cp or ya.size != cdim:
        raise ValueError("unpack: buffer too short")
    _lib.check(lib.cvxb_unpack(xa.ctypes.data, ya.ctypes.data, C.byref(cd), _lib.HOST), "unpack")


def pack2(x, dims, mnl=0):
    """misc_solvers.c:476-541"""
    lib = _lib.load()
    a = _buf(x)
    cd, keep, cdim, cp = make_dims(dims, mnl)
    if a.shape[0] < cdim:
        raise ValueError("pack2: x has fewer rows than the cone dimension")
    _lib.check(lib.cvxb_pack2(a.ctypes.data, a.shape[0], a.shape[1], C.byref(cd), _lib.HOST), "pack2")


def symm(x, n, offset=0):
    """misc_solvers.c:610-625"""
    lib = _lib.load()
    a = _buf(x).reshape(-1, order="F")[offset:offset + n * n]
    if a.size != n * n:
        raise ValueError("symm: buffer too short")
    _lib.check(lib.cvxb_symm(a.ctypes.data, n, _lib.HOST), "symm")


def _nlam(dims, mnl):
    return mnl + int(dims["l"]) + sum(int(k) for k in dims["q"]) + sum(int(k) for k in dims["s"])


def _vec(x, n, name, offset=0):
    a = _buf(x, name).reshape(-1, order="F")[offset:offset + n]
    if a.size != n:
        raise ValueError("%s: buffer too short" % name)
    return a


def scale2(lmbda, x, dims, mnl=0, inverse="N"):
    """x := H(lambda^{1/2}) x or its inverse.  misc_solvers.c:256-401"""
    lib = _lib.load()
    cd, keep, cdim, _ = make_dims(dims, mnl)
    la, xa = _vec(lmbda, _nlam(dims, mnl), "lmbda"), _vec(x, cdim, "x")
    _lib.check(lib.cvxb_scale2(la.ctypes.data, xa.ctypes.data, C.byref(cd), ord(inverse), _lib.HOST), "scale2")


def sprod(x, y, dims, mnl=0, diag="N"):
    """x := y o x.  misc_solvers.c:634-767"""
    lib = _lib.load()
    cd, keep, cdim, _ = make_dims(dims, mnl)
    xa = _vec(x, cdim, "x")
    ya = _vec(y, _nlam(dims, mnl) if diag == "D" else cdim, "y")
    _lib.check(lib.cvxb_sprod(xa.ctypes.data, ya.ctypes.data, C.byref(cd), ord(diag), _lib.HOST), "sprod")


def sinv(x, y, dims, mnl=0):
    """x := y o\\ x ('s' components of y diagonal).  misc_solvers.c:775-878"""
    lib = _lib.load()
    cd, keep, cdim, _ = make_dims(dims, mnl)
    xa, ya = _vec(x, cdim, "x"), _vec(y, _nlam(dims, mnl), "y")
    _lib.check(lib.cvxb_sinv(xa.ctypes.data, ya.ctypes.data, C.byref(cd), _lib.HOST), "sinv")


def trisc(x, dims, offset=0):
    """misc_solvers.c:887-935"""
    lib = _lib.load()
    cd, keep, cdim, _ = make_dims(dims, 0)
    xa = _vec(x, cdim, "x", offset)
    _lib.check(lib.cvxb_trisc(xa.ctypes.data, C.byref(cd), _lib.HOST), "trisc")


def triusc(x, dims, offset=0):
    """misc_solvers.c:940-986"""
    lib = _lib.load()
    cd, keep, cdim, _ = make_dims(dims, 0)
    xa = _vec(x, cdim, "x", offset)
    _lib.check(lib.cvxb_triusc(xa.ctypes.data, C.byref(cd), _lib.HOST), "triusc")


def sdot(x, y, dims, mnl=0):
    """misc_solvers.c:991-1039"""
    lib = _lib.load()
    cd, keep, cdim, _ = make_dims(dims, mnl)
    xa, ya = _vec(x, cdim, "x"), _vec(y, cdim, "y")
    out = C.c_double()
    _lib.check(lib.cvxb_sdot(xa.ctypes.data, ya.ctypes.data, C.byref(cd), C.byref(out), _lib.HOST), "sdot")
    return out.value


def max_step(x, dims, mnl=0, sigma=None):
    """min {t | x + t*e >= 0}.  misc_solvers.c:1052-1153.  With `sigma` (a 'd' buffer of length
    sum(dims['s'])) the eigenvalues of the 's' blocks are returned in sigma (ascending per block) and
    their eigenvectors overwrite the 's' blocks of x, as the reference's dsyevd_ call does
    (:1132-1136); eigenvector signs are the eigensolver's choice there and here."""
    lib = _lib.load()
    cd, keep, cdim, _ = make_dims(dims, mnl)
    xa = _vec(x, cdim, "x")
    sp = None
    if sigma is not None:
        sa = _vec(sigma, sum(int(k) for k in dims["s"]), "sigma")
        sp = sa.ctypes.data_as(_lib.c_double_p)
    out = C.c_double()
    _lib.check(lib.cvxb_max_step(xa.ctypes.data, C.byref(cd), sp, C.byref(out), _lib.HOST), "max_step")
    return out.value
