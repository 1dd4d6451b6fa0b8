"""Wall time of fit_se_regression_gp_rcpp on BASELINE config 2 (spiral n=4000, d=2, m=200, s=500, r=3, K=100) with the
10-point bandwidth grid and with a single bandwidth (gpurun helper, not a pytest file)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402
from flgp_b200.datasets import make  # noqa: E402

X, Y, cfg = make("C2")
m, s, r, K = cfg["m"], cfg["s"], cfg["r"], cfg["K"]
init = F.default_init(len(X), s, 1)
ctx = F.default_ctx()
grid = np.exp(np.linspace(np.log(0.1), np.log(10.0), 10))
for name, a2s in (("grid of 10", grid), ("single a2", grid[4:5])):
    for rep in range(3):
        ctx.set_timing(True)
        ctx.stage_reset()
        t0 = time.perf_counter()
        res = F.fit_se_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, a2s=a2s, init_idx=init)
        dt = time.perf_counter() - t0
        st = {}
        for d in ctx.stages():
            st[d["name"]] = round(st.get(d["name"], 0.0) + d["ms"], 2)
    print("%s: wall %.1f ms, a2 %.3f, pars %s, stages %s" % (name, dt * 1e3, res["a2"], np.round(res["pars"], 4), st))
