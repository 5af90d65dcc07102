import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvxopt_b200 import _lib
lib = _lib.load()
n, K = 8192, 16384
A = torch.randn(n, K, dtype=torch.float64, device="cuda")
H = torch.randn(n, n, dtype=torch.float64, device="cuda")
C = torch.empty(n, n, dtype=torch.float64, device="cuda")
w = torch.rand(K, dtype=torch.float64, device="cuda") + 0.5
torch.cuda.synchronize()
for name, wp in (("no scaling", None), ("fused di^2", w.data_ptr())):
    for _ in range(2):
        lib.cvxb_syrk_scaled(n, K, A.data_ptr(), K, wp, H.data_ptr(), n, C.data_ptr(), n, 0)
    t0 = time.perf_counter()
    for _ in range(5):
        lib.cvxb_syrk_scaled(n, K, A.data_ptr(), K, wp, H.data_ptr(), n, C.data_ptr(), n, 0)
    dt = (time.perf_counter() - t0) / 5
    print("%s: %.2f ms  %.2f TF/s (%.1f%%)" % (name, dt * 1e3, float(n) * n * K / dt * 1e-12, float(n) * n * K / dt * 1e-12 / 37.2 * 100))
