"""Efficiency of the DMMA SYRK kernel vs contraction length K (n fixed): separates per-tile
prologue/epilogue overhead from main-loop efficiency."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvxopt_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8064
for K in (128, 256, 512, 1024, 2048):
    A = torch.randn(n, K, dtype=torch.float64, device="cuda")          # column-major K x n (ld K)
    H = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.empty(n, n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for useH in (0, 1):
        for _ in range(3):
            lib.cvxb_syrk_scaled(n, K, A.data_ptr(), K, None, H.data_ptr() if useH else None, n, C.data_ptr(), n, 0)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            lib.cvxb_syrk_scaled(n, K, A.data_ptr(), K, None, H.data_ptr() if useH else None, n, C.data_ptr(), n, 0)
        dt = (time.perf_counter() - t0) / reps
        fl = float(n) * n * K
        print("n=%d K=%5d H=%d: %8.1f us  %5.2f TF/s (%.0f%% of 37.2)" % (n, K, useH, dt * 1e6, fl / dt * 1e-12, fl / dt * 1e-12 / 37.2 * 100))
