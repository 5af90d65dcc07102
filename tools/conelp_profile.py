"""Where a device-resident conelp solve spends its time (options['profile']): BASELINE configs 3 and 5."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cvxopt_b200
from problems import cone_lp

for name, n, dims in (("cfg3", 2048, {"l": 0, "q": [64] * 64, "s": []}), ("cfg5", 512, {"l": 0, "q": [], "s": [512]}),
                      ("sdp 4 x s128", 256, {"l": 0, "q": [], "s": [128] * 4})):
    c, G, h = cone_lp(n, dims, seed=11)
    cvxopt_b200.conelp(c, G, h, dims, maxiters=2)
    prof = {}
    sol = cvxopt_b200.conelp(c, G, h, dims, profile=prof)
    tot = prof.pop("total")[0]
    rows = sorted(prof.items(), key=lambda kv: -kv[1][0])
    acc = sum(v[0] for _, v in rows)
    print("%s: %d iterations, %.3f s total; in the 13 library operations %.3f s, driver (torch vector ops, scalars) %.3f s"
          % (name, sol["iterations"], tot, acc, tot - acc))
    for k, (t, cnt) in rows:
        print("   %-22s %8.1f ms  %5d calls  %8.3f ms/call" % (k, t * 1e3, cnt, t * 1e3 / cnt))
