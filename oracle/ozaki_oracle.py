"""TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md section 4): a numpy restatement of the arithmetic of
cvxopt_b200/csrc/ozaki_syrk.cu -- the fp64 SYRK  C = A' diag(d)^2 A + H  (reference: blas.syrk(Gs, K, trans='T')
in misc.kkt_chol.factor, src/python/misc.py:1275, on Gs = W^{-T} G scaled at misc.py:1268-1272) evaluated as
exact integer products of radix-2^7 digit slices.  Same exponent rule, digit rule, level grouping and order of
the fp64 operations as the kernels, so it states what "parity" means for that path; the GPU tests compare the
kernel with an 80-bit evaluation, the CPU tests compare this restatement with the same."""
import numpy as np

SLICES = 9
GROUPS = ((0, 1), (2, 4), (5, 8))          # levels kept in TMEM together, in launch order (oz_choose_groups)


def column_exponents(Gs):
    """e_j with |Gs[:, j]| < 2^e_j (frexp of the column maximum; 0 for a zero column)  -- oz_colscale_kernel"""
    amax = np.max(np.abs(Gs), axis=0) if Gs.shape[0] else np.zeros(Gs.shape[1])
    _, e = np.frexp(amax)
    return np.where(amax > 0, e, 0).astype(np.int64)


def slices(Gs, s=SLICES):
    """digits q_0..q_{s-1} (int64 arrays, |q| <= 64) with Gs[k,j] = 2^e_j sum_t q_t[k,j] 2^-(6+7t) + O(2^(e_j-6-7s))
    -- oz_slice_kernel: x = Gs * 64 * 2^-e; q = rint(x); x = (x - q) * 128 (all exact in fp64)"""
    e = column_exponents(Gs)
    x = Gs * np.ldexp(64.0, -e)[None, :]
    out = []
    for _ in range(s):
        q = np.rint(x)
        x = (x - q) * 128.0
        out.append(q.astype(np.int64))
    return e, out


def syrk(G, d, H=None, s=SLICES, groups=None):
    """lower triangle significant, like the kernel; returns the full symmetric product for convenience"""
    Gs = G * d[:, None] if d is not None else G            # fl(d*g): the scaled entry is rounded to fp64 first
    e, q = slices(Gs, s)
    n = G.shape[1]
    if groups is None:
        groups = GROUPS if s == 9 else tuple((a, min(a + 3, s - 1)) for a in range(0, s, 4))
    cs = np.ldexp(1.0, e)
    C = None
    for (d0, d1) in groups:
        v = np.zeros((n, n))
        for lev in range(d0, d1 + 1):                       # Horner over the levels of the pass: exact (< 2^53)
            acc = np.zeros((n, n), dtype=np.int64)
            for a in range(0, lev + 1):
                b = lev - a
                if a < s and b < s:
                    acc += q[a].T @ q[b]
            assert np.max(np.abs(acc), initial=0) < 2 ** 31 or G.shape[0] > 32768
            v = v * 128.0 + acc.astype(np.float64)
        v = ((v * np.ldexp(1.0, -12 - 7 * d1)) * cs[:, None]) * cs[None, :]
        if C is None:
            C = v + (H if H is not None else 0.0)
        else:
            C = v + C
    return C
