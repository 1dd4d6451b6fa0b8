"""Stage timing of subsample "minibatchkmeans" at the C4 shape (gpurun helper, not a pytest file):
python tests/bench_minibatch.py [n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402
from flgp_b200.datasets import make  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, Y, cfg = make("C4", 1234, n=n)
s, r, K = cfg["s"], cfg["r"], cfg["K"]
init = F.default_init(n, s, 1)
ctx = F.default_ctx()
for models in (dict(subsample="minibatchkmeans"), dict(subsample="kmeans")):
    for rep in range(2):
        ctx.set_timing(True)
        ctx.stage_reset()
        t0 = time.perf_counter()
        ep = F.heat_kernel_spectrum_cpp(X[:5000], X[5000:], s, r, K, models=models, init_idx=init, seed=1)
        wall = time.perf_counter() - t0
        st = ctx.stages()
        iters = ep.kmeans_iters
        ep.close()
    tot = {}
    for d in st:
        tot[d["name"]] = tot.get(d["name"], 0.0) + d["ms"]
    print(models["subsample"], "iters", iters, "wall %.1f ms" % (1e3 * wall), {k: round(v, 3) for k, v in tot.items()})
