"""Device-resident `conelp`: the reference's cone-LP interior-point driver (`coneprog.conelp`, reference
src/python/coneprog.py:586-1436) restated with every vector resident in HBM.

    minimize    c'x
    subject to  G x + s = h,   s in K = 'l' x 'q' cones x 's' cones            (no equality constraints here)

It is the same algorithm, statement by statement — self-dual embedding with (tau, kappa), the starting point of
:655-845, the stopping rule of :902-1030, Nesterov-Todd scaling, the Mehrotra predictor/corrector of :1256-1330 with
STEP = 0.99 and EXPON = 3 — so it takes the same iterates as `solvers.conelp(c, G, h, dims, kktsolver='chol')`.  What
differs is where the data lives: `x, s, z, ...` are torch tensors on the GPU and every operation on them is either a
torch elementwise op or a call into the B200 library with DEVICE pointers (cvxb_scale / scale2 / sprod / sinv / sdot /
max_step / compute_scaling / update_scaling, the KKT factor/solve of cvxb_kkt_*, the G operator of cvxb_kkt_gemv_G).
Per iteration only scalars (inner products, step lengths) cross PCIe.  Iterative refinement (options['refinement'])
is 0, as in the reference's default for dense problems.

PyTorch is used for device memory and elementwise vector arithmetic only."""
import ctypes as C
import math

import numpy as np

from . import _lib
from .kkt import KKTChol, make_dims

STEP, EXPON = 0.99, 3            # coneprog.py:423-424
DEFAULTS = dict(maxiters=100, abstol=1e-7, reltol=1e-6, feastol=1e-7)   # coneprog.py:436-456


def conelp(c, G, h, dims=None, kktsolver="chol", device=0, **options):
    """Solve the cone LP on the device.  c (n), G (cdim x n, 's' blocks as unpacked columns), h (cdim) are host
    arrays (cvxopt matrices or numpy); returns the reference's result dictionary (numpy arrays for x, s, z)."""
    import torch
    o = dict(DEFAULTS)
    o.update(options)
    lib = _lib.load()
    Gm = np.asarray(G, dtype=np.float64)
    if Gm.ndim != 2:
        raise TypeError("G must be a 'd' matrix")
    cdim_in, n = Gm.shape
    if dims is None:
        dims = {"l": cdim_in, "q": [], "s": []}
    ml, q, sd = int(dims["l"]), [int(k) for k in dims["q"]], [int(k) for k in dims["s"]]
    dims = {"l": ml, "q": q, "s": sd}
    cdim = ml + sum(q) + sum(k * k for k in sd)
    cdim_diag = ml + sum(q) + sum(sd)
    nlq = ml + sum(q)
    if cdim_in != cdim:
        raise TypeError("G must have %d rows" % cdim)
    cv = np.asarray(c, dtype=np.float64).reshape(-1)
    hv = np.asarray(h, dtype=np.float64).reshape(-1)
    if cv.size != n or hv.size != cdim:
        raise TypeError("c / h have the wrong size")
    if cdim == 0:
        raise ValueError("at least one cone constraint is required")
    if kktsolver not in ("chol", "qr"):
        raise ValueError("kktsolver must be 'chol' or 'qr'")
    dev = torch.device("cuda", device)
    f64 = torch.float64
    cd, keep, _, _ = make_dims(dims, 0)
    kkt = KKTChol(Gm, dims, None, 0, None, device, method="chol" if kktsolver == "chol" else "qr")

    def sync():
        torch.cuda.current_stream(dev).synchronize()

    def chk(rc, what):
        _lib.check(rc, what)

    # ---- the scaling W as one device buffer laid out like cvxb_scaling ----
    sq, s2 = sum(q), sum(k * k for k in sd)
    Wbuf = torch.zeros(2 * ml + sq + len(q) + 2 * s2 + 8, dtype=f64, device=dev)
    off = 0

    def take(cnt):
        nonlocal off
        v = Wbuf[off:off + cnt]
        off += cnt
        return v
    Wd, Wdi, Wv, Wbeta, Wr, Wrti = take(ml), take(ml), take(sq), take(len(q)), take(s2), take(s2)
    Wsc = _lib.Scaling(None, None, Wd.data_ptr() or None, Wdi.data_ptr() or None, Wv.data_ptr() or None,
                       Wbeta.data_ptr() or None, Wr.data_ptr() or None, Wrti.data_ptr() or None)

    def set_identity_scaling():                 # coneprog.py:667-676
        Wbuf.zero_()
        Wd.fill_(1.0); Wdi.fill_(1.0); Wbeta.fill_(1.0)
        o2 = 0
        for m in q:
            Wv[o2] = 1.0
            o2 += m
        o2 = 0
        for m in sd:
            Wr[o2:o2 + m * m].view(m, m).diagonal().fill_(1.0)
            Wrti[o2:o2 + m * m].view(m, m).diagonal().fill_(1.0)
            o2 += m * m

    def factor():
        sync()
        kkt.factor_ptr(d=Wd.data_ptr() if ml else 0, di=Wdi.data_ptr() if ml else 0, v=Wv.data_ptr() if sq else 0,
                       beta=Wbeta.data_ptr() if q else 0, r=Wr.data_ptr() if s2 else 0,
                       rti=Wrti.data_ptr() if s2 else 0, use_resident_H=False, space=_lib.DEVICE)

    def f3(xx, zz):
        sync()
        kkt.solve_ptr(xx.data_ptr(), zz.data_ptr(), space=_lib.DEVICE)

    # ---- cone algebra on device vectors (misc_solvers mirror, DEVICE space) ----
    def scale(xx, trans="N", inverse="N"):
        sync()
        chk(lib.cvxb_scale(xx.data_ptr(), cdim, 1, C.byref(cd), C.byref(Wsc), ord(trans), ord(inverse), _lib.DEVICE), "scale")

    def scale2(lm, xx, inverse="N"):
        sync()
        chk(lib.cvxb_scale2(lm.data_ptr(), xx.data_ptr(), C.byref(cd), ord(inverse), _lib.DEVICE), "scale2")

    def sprod(xx, yy):
        sync()
        chk(lib.cvxb_sprod(xx.data_ptr(), yy.data_ptr(), C.byref(cd), ord("N"), _lib.DEVICE), "sprod")

    def sinv(xx, yy):
        sync()
        chk(lib.cvxb_sinv(xx.data_ptr(), yy.data_ptr(), C.byref(cd), _lib.DEVICE), "sinv")

    def sdot(xx, yy):
        sync()
        out = C.c_double()
        chk(lib.cvxb_sdot(xx.data_ptr(), yy.data_ptr(), C.byref(cd), C.byref(out), _lib.DEVICE), "sdot")
        return out.value

    def max_step(xx, sigma=None):
        sync()
        out = C.c_double()
        sp = C.cast(sigma.data_ptr(), _lib.c_double_p) if (sigma is not None and sigma.numel()) else None
        rc = lib.cvxb_max_step(xx.data_ptr(), C.byref(cd), sp, C.byref(out), _lib.DEVICE)
        if rc > 0:
            raise ArithmeticError("max_step: eigensolver failure")
        chk(rc, "max_step")
        return out.value

    def symm_blocks(xx):                        # misc.symm on every 's' block (coneprog.py:968-972)
        sync()
        o2 = nlq
        for m in sd:
            if m > 1:
                chk(lib.cvxb_symm(xx.data_ptr() + 8 * o2, m, _lib.DEVICE), "symm")
            o2 += m * m

    def Gf(xx, yy, alpha=1.0, beta=0.0, trans="N"):   # misc.sgemv(G, ...)  (misc.py:801-832)
        sync()
        if trans == "T" and sd and alpha:
            chk(lib.cvxb_trisc(xx.data_ptr(), C.byref(cd), _lib.DEVICE), "trisc")
        chk(lib.cvxb_kkt_gemv_G(kkt._h, xx.data_ptr(), yy.data_ptr(), float(alpha), float(beta), ord(trans),
                                _lib.DEVICE), "G")
        if trans == "T" and sd and alpha:
            chk(lib.cvxb_triusc(xx.data_ptr(), C.byref(cd), _lib.DEVICE), "triusc")

    def compute_scaling(ss, zz, lm):
        sync()
        chk(lib.cvxb_compute_scaling(ss.data_ptr(), zz.data_ptr(), lm.data_ptr(), C.byref(cd), C.byref(Wsc),
                                     _lib.DEVICE), "compute_scaling")

    def update_scaling(lm, dss, dzz):
        sync()
        chk(lib.cvxb_update_scaling(C.byref(Wsc), lm.data_ptr(), dss.data_ptr(), dzz.data_ptr(), C.byref(cd),
                                    _lib.DEVICE), "update_scaling")

    ops = (set_identity_scaling, factor, f3, scale, scale2, sprod, sinv, sdot, max_step, symm_blocks, Gf, compute_scaling,
           update_scaling)
    prof = o.get("profile")
    if isinstance(prof, dict):
        # options['profile'] = {}: wall time (device-synchronised) and call count per operation, 'total' for the solve
        import time

        def timed(name, f):
            def g(*a, **k):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                try:
                    return f(*a, **k)
                finally:
                    torch.cuda.synchronize(dev)
                    e = prof.setdefault(name, [0.0, 0])
                    e[0] += time.perf_counter() - t0
                    e[1] += 1
            return g
        names = ("set_identity_scaling", "factor", "f3", "scale", "scale2", "sprod", "sinv", "sdot", "max_step",
                 "symm_blocks", "Gf", "compute_scaling", "update_scaling")
        ops = tuple(timed(nm, f) for nm, f in zip(names, ops))
        t_all = time.perf_counter()
    try:
        sol = _conelp_core(torch, dev, cv, hv, n, dims, ops, o)
        if isinstance(prof, dict):
            torch.cuda.synchronize(dev)
            prof["total"] = [time.perf_counter() - t_all, 1]
        return sol
    finally:
        kkt.close()


def _conelp_core(torch, dev, cv, hv, n, dims, ops, o):
    """The driver itself: coneprog.conelp (coneprog.py:586-1436) on tensors that live on `dev`, every cone / KKT
    operation through the 13 closures in `ops`.  `conelp` passes closures that call the B200 library with DEVICE
    pointers; tests/test_conelp_twin_cpu.py passes closures made of the reference's own functions (CPU tensors) to
    check this restatement of the driver logic where no GPU is available."""
    (set_identity_scaling, factor, f3, scale, scale2, sprod, sinv, sdot, max_step, symm_blocks, Gf, compute_scaling,
     update_scaling) = ops
    MAXITERS, ABSTOL, RELTOL, FEASTOL = int(o["maxiters"]), float(o["abstol"]), float(o["reltol"]), float(o["feastol"])
    show_progress = bool(o.get("show_progress", False))
    debug_hook = o.get("debug_hook")
    f64 = torch.float64
    ml, q, sd = dims["l"], dims["q"], dims["s"]
    cdim = ml + sum(q) + sum(k * k for k in sd)
    cdim_diag = ml + sum(q) + sum(sd)
    nlq = ml + sum(q)

    def snrm2(xx):
        return math.sqrt(sdot(xx, xx))

    def zeros(k):
        return torch.zeros(k, dtype=f64, device=dev)

    def xdot(a, b):
        return float(torch.dot(a, b))

    # index helpers for the 'q' cones: first entries, and the cone each 'q' row belongs to
    qfirst = torch.tensor(np.cumsum([0] + q[:-1], dtype=np.int64) + ml, dtype=torch.int64, device=dev) if q else None
    qseg = torch.tensor(np.repeat(np.arange(len(q)), q), dtype=torch.int64, device=dev) if q else None

    def ssqr(xx, yy):                           # x := y o y   (misc.py:945-959), diag storage
        xx[:ml] = yy[:ml] * yy[:ml]
        if q:
            yq = yy[ml:nlq]
            nrm2 = torch.zeros(len(q), dtype=f64, device=dev).index_add_(0, qseg, yq * yq)
            y0 = yy[qfirst]
            xx[ml:nlq] = 2.0 * y0[qseg] * yq
            xx[qfirst] = nrm2
        ns_ = sum(sd)
        xx[nlq:nlq + ns_] = yy[nlq:nlq + ns_] * yy[nlq:nlq + ns_]

    def add_to_cone_identity(v, a):             # v += a*e   (coneprog.py:816-822)
        v[:ml] += a
        if q:
            v[qfirst] += a
        o2 = nlq
        for m in sd:
            v[o2:o2 + m * m].view(m, m).diagonal().add_(a)
            o2 += m * m

    def diag_from(lm_part, v):                  # 's' blocks of v := diag(lm_part), 'l','q' := lm_part  (:1266-1275)
        v[:nlq] = lm_part[:nlq]
        v[nlq:].zero_()
        o2, o3 = nlq, nlq
        for m in sd:
            v[o2:o2 + m * m].view(m, m).diagonal().copy_(lm_part[o3:o3 + m])
            o2 += m * m
            o3 += m


    ct = torch.from_numpy(cv).to(dev)
    ht = torch.from_numpy(hv).to(dev)
    resx0 = max(1.0, math.sqrt(xdot(ct, ct)))
    resz0 = max(1.0, snrm2(ht))
    resy0 = 1.0

    x, s, z = zeros(n), zeros(cdim), zeros(cdim)
    dx, ds, dz = zeros(n), zeros(cdim), zeros(cdim)
    try:
        # ---- starting point (coneprog.py:655-845) ----
        set_identity_scaling()
        try:
            factor()
        except ArithmeticError:
            raise ValueError("Rank(A) < p or Rank([G; A]) < n")
        # primal: minimize ||G x - h||^2
        x.zero_()
        s.copy_(ht)
        try:
            f3(x, s)
        except ArithmeticError:
            raise ValueError("Rank(A) < p or Rank([G; A]) < n")
        s.mul_(-1.0)
        ts = max_step(s)
        # dual: minimize ||z||^2 s.t. G'z + c = 0
        dx.copy_(ct).mul_(-1.0)
        z.zero_()
        f3(dx, z)
        tz = max_step(z)
        nrms, nrmz = snrm2(s), snrm2(z)
        gap = sdot(s, z)
        pcost = xdot(ct, x)
        dcost = -sdot(ht, z)
        if pcost < 0.0:
            relgap = gap / -pcost
        elif dcost > 0.0:
            relgap = gap / dcost
        else:
            relgap = None

        def result(status, iters, **kw):
            out = {"x": x.cpu().numpy().copy() if kw.get("x", True) else None, "y": np.zeros(0),
                   "s": s.cpu().numpy().copy() if kw.get("s", True) else None,
                   "z": z.cpu().numpy().copy() if kw.get("z", True) else None,
                   "status": status, "iterations": iters}
            out.update(kw.get("fields", {}))
            return out

        if ts <= 0 and tz <= 0 and (gap <= ABSTOL or (relgap is not None and relgap <= RELTOL)):
            # the constructed starting points are feasible and optimal (:776-811)
            symm_blocks(s); symm_blocks(z)
            rxv = ct.clone()
            Gf(z, rxv, beta=1.0, trans="T")
            resx = math.sqrt(xdot(rxv, rxv))
            rzv = zeros(cdim)
            Gf(x, rzv)
            rzv += s
            rzv -= ht
            resz = snrm2(rzv)
            return result("optimal", 0, fields={
                "gap": gap, "relative gap": relgap, "primal objective": xdot(ct, x),
                "dual objective": -sdot(ht, z), "primal infeasibility": resz / resz0, "primal slack": -ts,
                "dual slack": -tz, "dual infeasibility": resx / resx0,
                "residual as primal infeasibility certificate": None,
                "residual as dual infeasibility certificate": None})
        if ts >= -1e-8 * max(nrms, 1.0):
            add_to_cone_identity(s, 1.0 + ts)
        if tz >= -1e-8 * max(nrmz, 1.0):
            add_to_cone_identity(z, 1.0 + tz)

        tau, kappa = 1.0, 1.0
        rx, hrx = zeros(n), zeros(n)
        rz, hrz = zeros(cdim), zeros(cdim)
        ns_ = sum(sd)
        sigs, sigz = zeros(ns_), zeros(ns_)
        lmbda, lmbdasq = zeros(cdim_diag + 1), zeros(cdim_diag + 1)
        x1, z1, th, ws3 = zeros(n), zeros(cdim), zeros(cdim), zeros(cdim)
        wkappa3 = 0.0
        dg = dgi = 1.0
        gap = sdot(s, z)

        for iters in range(MAXITERS + 1):
            # hrx = -G'z ; rx = hrx - c*tau          (:861-870)
            Gf(z, hrx, alpha=-1.0, beta=0.0, trans="T")
            hresx = math.sqrt(xdot(hrx, hrx))
            rx.copy_(hrx).add_(ct, alpha=-tau)
            resx = math.sqrt(xdot(rx, rx)) / tau
            hresy, resy = 0.0, 0.0
            # hrz = s + G x ; rz = hrz - h*tau       (:883-893)
            Gf(x, hrz)
            hrz += s
            hresz = snrm2(hrz)
            rz.copy_(hrz).add_(ht, alpha=-tau)
            resz = snrm2(rz) / tau
            cx, by, hz = xdot(ct, x), 0.0, sdot(ht, z)
            rt = kappa + cx + by + hz
            pcost, dcost = cx / tau, -(by + hz) / tau
            if pcost < 0.0:
                relgap = gap / -pcost
            elif dcost > 0.0:
                relgap = gap / dcost
            else:
                relgap = None
            pres = max(resy / resy0, resz / resz0)
            dres = resx / resx0
            pinfres = hresx / resx0 / (-hz - by) if hz + by < 0.0 else None
            dinfres = max(hresy / resy0, hresz / resz0) / (-cx) if cx < 0.0 else None
            if debug_hook is not None:
                debug_hook(iters, x, s, z, tau, kappa, rx, rz)
            if show_progress:                    # the reference's progress line (:926-932), more digits
                if iters == 0:
                    print("% 10s% 12s% 10s% 8s% 7s % 5s" % ("pcost", "dcost", "gap", "pres", "dres", "k/t"))
                print("%2d: % 8.4e % 8.4e % 4.0e% 7.1e% 7.1e% 7.0e" % (iters, pcost, dcost, gap, pres, dres, kappa / tau))

            if (pres <= FEASTOL and dres <= FEASTOL and (gap <= ABSTOL or (relgap is not None and relgap <= RELTOL))) \
                    or iters == MAXITERS:
                x.mul_(1.0 / tau); s.mul_(1.0 / tau); z.mul_(1.0 / tau)
                symm_blocks(s); symm_blocks(z)
                ts, tz = max_step(s), max_step(z)
                opt = iters != MAXITERS            # the reference reports 'unknown' at MAXITERS (:974-993)
                return result("optimal" if opt else "unknown", iters, fields={
                    "gap": gap, "relative gap": relgap, "primal objective": pcost, "dual objective": dcost,
                    "primal infeasibility": pres, "dual infeasibility": dres, "primal slack": -ts, "dual slack": -tz,
                    "residual as primal infeasibility certificate": None if opt else pinfres,
                    "residual as dual infeasibility certificate": None if opt else dinfres})
            elif pinfres is not None and pinfres <= FEASTOL:
                z.mul_(1.0 / (-hz - by))
                symm_blocks(z)
                tz = max_step(z)
                return result("primal infeasible", iters, x=False, s=False, fields={
                    "gap": None, "relative gap": None, "primal objective": None, "dual objective": 1.0,
                    "primal infeasibility": None, "dual infeasibility": None, "primal slack": None, "dual slack": -tz,
                    "residual as primal infeasibility certificate": pinfres,
                    "residual as dual infeasibility certificate": None})
            elif dinfres is not None and dinfres <= FEASTOL:
                x.mul_(1.0 / (-cx)); s.mul_(1.0 / (-cx))
                symm_blocks(s)
                ts = max_step(s)
                return result("dual infeasible", iters, z=False, fields={
                    "gap": None, "relative gap": None, "primal objective": -1.0, "dual objective": None,
                    "primal infeasibility": None, "dual infeasibility": None, "primal slack": -ts, "dual slack": None,
                    "residual as primal infeasibility certificate": None,
                    "residual as dual infeasibility certificate": dinfres})

            if iters == 0:                       # (:1033-1044)
                compute_scaling(s, z, lmbda)
                dg = math.sqrt(kappa / tau)
                dgi = math.sqrt(tau / kappa)
                lmbda[-1] = math.sqrt(tau * kappa)
            ssqr(lmbdasq, lmbda)
            lmbdasq[-1] = lmbda[-1] * lmbda[-1]
            lmbdag = float(lmbda[-1])

            try:                                 # (:1066-1078)
                factor()
                x1.copy_(ct).mul_(-1.0)
                z1.copy_(ht)
                f3(x1, z1)
                x1.mul_(dgi); z1.mul_(dgi)
            except ArithmeticError:
                x.mul_(1.0 / tau); s.mul_(1.0 / tau); z.mul_(1.0 / tau)
                symm_blocks(s); symm_blocks(z)
                ts, tz = max_step(s), max_step(z)
                return result("unknown", iters, fields={
                    "gap": gap, "relative gap": relgap, "primal objective": pcost, "dual objective": dcost,
                    "primal infeasibility": pres, "dual infeasibility": dres, "primal slack": -ts, "dual slack": -tz,
                    "residual as primal infeasibility certificate": pinfres,
                    "residual as dual infeasibility certificate": dinfres})
            th.copy_(ht)                         # th = W^{-T} h   (:1124-1126)
            scale(th, trans="T", inverse="I")
            z1z1 = sdot(z1, z1)

            def f6_no_ir(xx, zz, tau_, ss, kappa_):        # (:1130-1192), p = 0
                sinv(ss, lmbda)
                ss.mul_(-1.0)
                ws3t = ss.clone()
                scale(ws3t, trans="T")
                zz += ws3t
                zz.mul_(-1.0)
                f3(xx, zz)
                kappa_ = -kappa_ / lmbdag
                tau_ = tau_ + kappa_ / dgi
                tau_ = dgi * (tau_ + xdot(ct, xx) + sdot(th, zz)) / (1.0 + z1z1)
                xx.add_(x1, alpha=tau_)
                zz.add_(z1, alpha=tau_)
                ss.sub_(zz)
                kappa_ -= tau_
                return tau_, kappa_

            mu = xdot(lmbda, lmbda) / (1 + cdim_diag)
            sigma = 0.0
            step = 1.0
            tt = tk = 0.0
            for i in (0, 1):                     # (:1256-1330)
                diag_from(lmbdasq, ds)
                dkappa = float(lmbdasq[-1])
                if i == 1:
                    ds += ws3
                    add_to_cone_identity(ds, -sigma * mu)
                    dkappa += wkappa3 - sigma * mu
                dx.copy_(rx).mul_(1.0 - sigma)
                dz.copy_(rz).mul_(1.0 - sigma)
                dtau = (1.0 - sigma) * rt
                dtau, dkappa = f6_no_ir(dx, dz, dtau, ds, dkappa)
                if i == 0:
                    ws3.copy_(ds)
                    sprod(ws3, dz)
                    wkappa3 = dtau * dkappa
                scale2(lmbda, ds)
                scale2(lmbda, dz)
                if i == 0:
                    ts, tz = max_step(ds), max_step(dz)
                else:
                    ts, tz = max_step(ds, sigma=sigs), max_step(dz, sigma=sigz)
                tt = -dtau / lmbdag
                tk = -dkappa / lmbdag
                t = max([0.0, ts, tz, tt, tk])
                if t == 0.0:
                    step = 1.0
                else:
                    step = min(1.0, 1.0 / t) if i == 0 else min(1.0, STEP / t)
                if i == 0:
                    sigma = (1.0 - step) ** EXPON

            x.add_(dx, alpha=step)               # (:1334)
            # ds, dz := updated variables in the current scaling ('l','q'), factors Ls, Lz ('s')   (:1346-1391)
            ds[:nlq].mul_(step); dz[:nlq].mul_(step)
            ds[:ml] += 1.0; dz[:ml] += 1.0
            if q:
                ds[qfirst] += 1.0; dz[qfirst] += 1.0
            scale2(lmbda, ds, inverse="I")
            scale2(lmbda, dz, inverse="I")
            if ns_:
                sigs.mul_(step).add_(1.0).div_(lmbda[nlq:nlq + ns_])
                sigz.mul_(step).add_(1.0).div_(lmbda[nlq:nlq + ns_])
                o2, o3 = nlq, 0
                for m in sd:
                    ds[o2:o2 + m * m].view(m, m).mul_(torch.sqrt(sigs[o3:o3 + m]).unsqueeze(1))    # column i *= sqrt(sig_i)
                    dz[o2:o2 + m * m].view(m, m).mul_(torch.sqrt(sigz[o3:o3 + m]).unsqueeze(1))
                    o2 += m * m
                    o3 += m
            update_scaling(lmbda, ds, dz)
            dg *= math.sqrt(1.0 - step * tk) / math.sqrt(1.0 - step * tt)      # (:1403-1405)
            dgi = 1.0 / dg
            lmbda[-1] = lmbdag * math.sqrt(1.0 - step * tt) * math.sqrt(1.0 - step * tk)
            # unscale s, z, tau, kappa             (:1411-1436)
            diag_from(lmbda, s)
            scale(s, trans="T")
            diag_from(lmbda, z)
            scale(z, inverse="I")
            lg = float(lmbda[-1])
            kappa, tau = lg / dgi, lg * dgi
            gap = (float(torch.linalg.vector_norm(lmbda[:-1])) / tau) ** 2
        raise RuntimeError("unreachable")
    finally:
        pass    # (the caller owns the KKT factory and closes it)
