"""Stage timing of the Nystrom driver (gpurun helper, not a pytest file):
python tests/bench_nystrom.py [n] [d] [s] [K]      (defaults: 1000000 3 2000 200; one bandwidth, fixed pars)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402
from flgp_b200.datasets import make  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
n, d, s, K = (a + [1_000_000, 3, 2000, 200][len(a):])[:4]
m = 5000
X, Y, _ = make("C4", 7, n=n)
init = F.default_init(n, s, 3)
ctx = F.default_ctx()
ctx.set_timing(True)
for it in range(3):
    ctx.stage_reset()
    t0 = time.perf_counter()
    res = F.fit_nystrom_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, K, a2s=[1.0], pars=(10.0, 0.01), init_idx=init,
                                           iter_max=10)
    t1 = time.perf_counter()
    agg = {}
    for q in ctx.stages():
        e = agg.setdefault(q["name"], [0.0, 0.0])
        e[0] += q["ms"]
        e[1] += q["flops"]
    print("run %d: wall %.1f ms; " % (it, (t1 - t0) * 1e3) +
          ", ".join("%s %.2f ms (%.1f TFLOP/s)" % (k, v[0], v[1] / v[0] / 1e9 if v[0] else 0) for k, v in agg.items()),
          flush=True)
rmse = float(np.sqrt(np.mean((res["Y_pred"]["test"][:100000] - Y[m:m + 100000]) ** 2)))
print("test rmse (first 100k rows) %.4f" % rmse)
