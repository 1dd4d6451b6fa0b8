"""Per-pass timing of the C4 k-means stage (gpurun helper, not a pytest file): python tests/bench_kmeans.py [n] [iter_max]
FLGP_KMEANS_PROF=1 adds the survivors of the bound test per pass."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402
from flgp_b200.datasets import make  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
iter_max = int(sys.argv[2]) if len(sys.argv) > 2 else 100
X, Y, cfg = make("C4", 1234, n=n)
s, r, K = cfg["s"], cfg["r"], cfg["K"]
init = F.default_init(n, s, 1)
ctx = F.default_ctx()
models = dict(subsample="kmeans", kernel="lae", gl="cluster-normalized", root=True)
for rep in range(2):
    ctx.set_timing(True)
    ctx.stage_reset()
    ep = F.heat_kernel_spectrum_cpp(X[:5000], X[5000:], s, r, K, models=models, init_idx=init, iter_max=iter_max)
    st = ctx.stages()
    ep.close()
passes = [d["ms"] for d in st if d["name"] == "kmeans_pruned_pass"]
tot = {}
for d in st:
    tot[d["name"]] = tot.get(d["name"], 0.0) + d["ms"]
print({k: round(v, 3) for k, v in tot.items()})
print("pruned passes: n=%d total %.2f ms" % (len(passes), sum(passes)))
print("ms per pass:", " ".join("%.3f" % x for x in passes))
