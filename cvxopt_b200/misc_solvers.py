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


def _not_yet(name):
    def f(*a, **k):
        raise NotImplementedError(
            "cvxopt_b200.misc_solvers.%s: O(cdim) IPM-side cone algebra stays on the host in this "
            "round (use cvxopt.misc_solvers.%s)" % (name, name))
    f.__name__ = name
    return f


scale2 = _not_yet("scale2")
sprod = _not_yet("sprod")
sinv = _not_yet("sinv")
trisc = _not_yet("trisc")
triusc = _not_yet("triusc")
sdot = _not_yet("sdot")
max_step = _not_yet("max_step")
