"""Full device-resident IPM solves at the BASELINE shapes (configs 2 and N*), printing iterations
and objectives (survey probe for cfg 2, seed 1234: 14 iterations, pobj 3.534265721964e+03)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
for n in (int(a) for a in sys.argv[1:] or ["4096"]):
    print(json.dumps(bench.run_ipm(n, 2 * n, 1234, 0)))
