"""CPU checks of the arithmetic behind the int8-slice SYRK (oracle/ozaki_oracle.py restates
csrc/ozaki_syrk.cu): the digit expansion, the exactness bounds and the fp64-level accuracy claim."""
import numpy as np
import pytest

import ozaki_oracle as oz


def test_digits_reconstruct_the_entries():
    rng = np.random.Generator(np.random.PCG64(1))
    A = rng.standard_normal((300, 17)) * np.exp(4 * rng.standard_normal((300, 1)))
    A[:, 3] = 0.0
    A[5, 4] = 0.0
    e, q = oz.slices(A, 9)
    assert all(np.max(np.abs(x)) <= 64 for x in q)           # int8 range, products <= 4096
    rec = np.zeros_like(A, dtype=np.longdouble)
    for t, x in enumerate(q):
        rec += x.astype(np.longdouble) * np.longdouble(2.0) ** (-(6 + 7 * t))
    rec *= np.ldexp(1.0, e)[None, :].astype(np.longdouble)
    err = np.abs(rec - A.astype(np.longdouble)) / np.ldexp(1.0, e)[None, :]
    assert float(err.max()) <= 2.0 ** -63                    # half a unit of the last digit: 2^-(6+56) / 2
    assert np.all(q[0][:, 3] == 0) and e[3] == 0


def test_int32_accumulators_cannot_overflow():
    # level d sums (d+1) slice products of K rows, each |q q'| <= 64*64: the kernel drains every 32768 rows
    assert 9 * 4096 * 32768 < 2 ** 31
    # and the Horner combination of four levels stays exact in fp64
    assert (2 ** 31) * 128 ** 3 + (2 ** 31) * 128 ** 2 + (2 ** 31) * 128 + 2 ** 31 < 2 ** 53


@pytest.mark.parametrize("spread", [0.0, 2.0, 6.0])
def test_nine_slices_give_an_fp64_accurate_syrk(spread):
    rng = np.random.Generator(np.random.PCG64(int(10 * spread) + 3))
    m, n = 700, 60
    G = rng.standard_normal((m, n))
    d = np.exp(spread * rng.standard_normal(m))
    B = rng.standard_normal((n, n))
    H = B @ B.T / n + np.eye(n)
    C = oz.syrk(G, d, H)
    Gs = (G * d[:, None]).astype(np.longdouble)
    ref = Gs.T @ Gs + H.astype(np.longdouble)
    mag = np.abs(Gs).T @ np.abs(Gs) + np.abs(H).astype(np.longdouble)
    err = float((np.abs(C.astype(np.longdouble) - ref) / mag).max())
    assert err < 5e-16
    # plain fp64 evaluation of the same sums is no better
    C64 = (G * d[:, None]).T @ (G * d[:, None]) + H
    err64 = float((np.abs(C64.astype(np.longdouble) - ref) / mag).max())
    assert err < 4 * max(err64, 1.2e-16)


def test_eight_slices_are_not_enough_and_groups_do_not_matter():
    rng = np.random.Generator(np.random.PCG64(9))
    m, n = 2000, 40
    G = rng.standard_normal((m, n))
    d = np.exp(2.0 * rng.standard_normal(m))                 # a modest NT-scaling spread already shows it
    Gs = (G * d[:, None]).astype(np.longdouble)
    ref = Gs.T @ Gs
    mag = np.abs(Gs).T @ np.abs(Gs)
    e8 = float((np.abs(oz.syrk(G, d, None, s=8).astype(np.longdouble) - ref) / mag).max())
    e9 = float((np.abs(oz.syrk(G, d, None, s=9).astype(np.longdouble) - ref) / mag).max())
    assert e9 < 4e-16 and 2e-15 < e8 < 1e-12
    a = oz.syrk(G, d, None, groups=((0, 3), (4, 7), (8, 8)))
    b = oz.syrk(G, d, None, groups=((0, 0), (1, 4), (5, 8)))
    assert float(np.abs(a - b).max() / np.abs(a).max()) < 4e-16
