"""Stage timing of the large-d path on BASELINE config 3's shape (gpurun helper, not a pytest file):
python tests/bench_c3.py [n] [d] [s] [r] [K] [iter_max]      (defaults: 70000 784 1000 5 200 20)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
n, d, s, r, K, iter_max = (a + [70000, 784, 1000, 5, 200, 20][len(a):])[:6]
m = 1000
rng = np.random.default_rng(3)
if d == 16:  # BASELINE config 5's data: a 2-torus embedded in 16 dimensions + N(0, 0.01^2)
    a_, b_ = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    Q_, _ = np.linalg.qr(rng.standard_normal((d, 4)))
    X = np.asfortranarray(np.c_[np.cos(a_), np.sin(a_), np.cos(b_), np.sin(b_)] @ Q_.T + 0.01 * rng.standard_normal((n, d)))
else:  # BASELINE config 3's data: 10 Gaussian classes
    means = 3.0 * rng.standard_normal((10, d))
    lab = rng.integers(0, 10, n)
    X = np.asfortranarray(means[lab] + rng.standard_normal((n, d)))
init = np.sort(rng.choice(n, s, replace=False)).astype(np.int32)
ctx = F.default_ctx()
ctx.set_timing(True)
for it in range(3):
    ctx.stage_reset()
    t0 = time.perf_counter()
    ep = F.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=iter_max)
    t1 = time.perf_counter()
    st = {q["name"]: (round(q["ms"], 3), q["launches"]) for q in ctx.stages()}
    print("run %d: wall %.1f ms, kmeans iters %d, stages %s" % (it, (t1 - t0) * 1e3, ep.kmeans_iters, st), flush=True)
fl = 2.0 * n * s * d
print("distance work per k-means pass / KNN: %.1f GFLOP" % (fl / 1e9))
