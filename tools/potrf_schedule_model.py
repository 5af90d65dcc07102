"""Discrete-event model of the look-ahead Cholesky schedule in csrc/chol.cu (non-pair mode), calibrated with the timelines
in profiles/r02v_potrf_chain_gap_experiment.txt (n=8192, NB=128, one B200).  CPU only; prints the predicted potrf time
for the present code and for the levers discussed in DESIGN.md section 7.

Streams / kernels per block step j (nb = n/128 steps, k = nb-1-j trailing block rows):
  D(j)  potf2_inv            after D(j-1), C0(j-2), R(j-2)               duration c (+ gap between dependent launches)
  Tr(j) panel TRSM           after D(j), C0(j-1) [stream order], R(j-2)  duration tr(k)
  C0(j) next block column    after Tr(j), R(j-1)                         duration c0(k)
  R(j)  bulk trailing update after Tr(j), R(j-1) [stream order]          duration tiles(k) * 4.19 MFLOP / rate
"""
import sys


def model(nb=64, c=88.5, gap=4.5, rate=24.4, r_min=25.0, tr0=13.0, tr_slope=0.48, c0_0=15.0, c0_slope=0.33,
          launch=3.5, verbose=False):
    D = [0.0] * nb; Tr = [0.0] * nb; C0 = [0.0] * nb; R = [0.0] * nb

    def get(a, i):
        return a[i] if i >= 0 else 0.0
    for j in range(nb):
        k = nb - 1 - j                       # block rows below the diagonal block
        d_start = max(get(D, j - 1) + gap, get(C0, j - 2), get(R, j - 2))
        D[j] = d_start + c
        if k == 0:
            break
        tr_start = max(D[j], get(C0, j - 1), get(R, j - 2)) + launch
        Tr[j] = tr_start + tr0 + tr_slope * k
        c0_start = max(Tr[j], get(R, j - 1)) + launch
        C0[j] = c0_start + (c0_0 + c0_slope * k if k > 1 else 0.0)
        tiles = (k - 1) * k / 2.0            # 128x128 tiles of block columns >= j+2
        # measured rate of the K=128 update: 23.9 TF/s at 1953 tiles, 21.4 at 741, 19.8 at 378, 18.6 at 120
        # (profiles/r01_potrf_timeline.md, r02v): rate_k = rate * (0.735 + 0.245 * sqrt(tiles / 1953))
        rk = rate * (0.735 + 0.245 * min(1.0, (tiles / 1953.0) ** 0.5)) if tiles > 0 else rate
        r_dur = max(r_min, tiles * 4.194304e6 / (rk * 1e6)) if tiles > 0 else 0.0   # us
        r_start = max(Tr[j], get(R, j - 1)) + launch
        R[j] = r_start + r_dur if tiles > 0 else get(R, j - 1)
        if verbose and j in (0, 1, 8, 16, 24, 32, 35, 40, 48, 56):
            print("  step %2d: D %8.1f-%8.1f  Tr -%8.1f  C0 -%8.1f  R %8.1f-%8.1f" %
                  (j, d_start, D[j], Tr[j], C0[j], r_start, R[j]))
    return max(D[nb - 1], max(R), max(C0), max(Tr)) * 1e-3


if __name__ == "__main__":
    base = model(verbose="-v" in sys.argv)
    print("present code (c = 88.5 us, bulk 24.4 TF/s at step 0)  : %.2f ms   (measured 10.0-10.3)" % base)
    print("chain kernel 70 us                                   : %.2f ms" % model(c=70.0))
    print("chain kernel 50 us                                   : %.2f ms" % model(c=50.0))
    print("bulk update 1.35x faster                             : %.2f ms" % model(rate=33.0))
    print("bulk 1.35x + chain 70 us                             : %.2f ms" % model(rate=33.0, c=70.0))
    print("bulk 1.35x + chain 50 us                             : %.2f ms" % model(rate=33.0, c=50.0))
    print("bulk 2x (long-K / int8 far updates) + chain 50 us   : %.2f ms" % model(rate=48.8, c=50.0, r_min=20.0))
    print("chain alone (64 x (c + gap))                         : %.2f ms   (measured 6.0 with the waits removed)"
          % (64 * (88.5 + 4.5) * 1e-3))
