"""Stand-alone check + timing of the top-K eigensolver on a given symmetric matrix (gpurun helper, not a pytest file):
python tests/bench_chfsi.py [file.npy | s] [K]     FLGP_EIGH_DIRECT=1 selects the direct route for comparison."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402

arg = sys.argv[1] if len(sys.argv) > 1 else "2000"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
if arg.endswith(".npy"):
    A = np.asfortranarray(np.load(arg))
else:
    s = int(arg)
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((s, s)))
    lam = 1.0 / (1.0 + 0.02 * np.arange(s)) ** 2
    A = np.asfortranarray((Q * lam) @ Q.T)
A = (A + A.T) / 2
s = A.shape[0]
ctx = F.default_ctx()
ctx.set_timing(True)
for it in range(3):
    ctx.stage_reset()
    t0 = time.perf_counter()
    res = F.eigs_sym(A, K)
    t1 = time.perf_counter()
    st = {}
    for d in ctx.stages():
        st[d["name"]] = round(st.get(d["name"], 0.0) + d["ms"], 3)
    print("run %d: wall %.2f ms, stages %s" % (it, (t1 - t0) * 1e3, st))
w, V = np.linalg.eigh(A)
w, V = w[::-1], V[:, ::-1]
Y = res["vectors"]
print("max |dlam| = %.3e" % np.abs(res["values"] - w[:K]).max())
print("resid = %.3e, orth = %.3e" % (np.abs(A @ Y - Y * res["values"]).max(), np.abs(Y.T @ Y - np.eye(K)).max()))
print("gap at K = %.3e, subspace err = %.3e" % (w[K - 1] - w[K], np.abs(Y @ Y.T - V[:, :K] @ V[:, :K].T).max()))
