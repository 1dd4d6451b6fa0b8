"""Generates tests/golden/oracle_small.npz from the oracle (python tests/golden/make_golden.py).
The reference has no golden vectors and cannot be executed in this environment (no R / Rcpp / Eigen);
these fixtures pin the ORACLE (and through it the CUDA path) against drift."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from conftest import spiral  # noqa: E402

meta = dict(n=1500, seed=11, s=48, r=3, K=10)
X, Y = spiral(meta["n"], meta["seed"])
init = np.sort(np.random.default_rng(5).choice(meta["n"], meta["s"], replace=False)).astype(np.int32)
U, assign, iters = O.kmeans_lloyd(X, meta["s"], init)
meta["iters"] = iters
ind = O.knn(X, U[:, :2], meta["r"])
Zj, Zx = O.cross_similarity_lae(X, U, meta["r"], "cluster-normalized")
values, V = O.spectrum_from_Z(Zj, Zx, meta["s"], meta["K"], True)
H = O.hk_from_spectrum(V, values, meta["K"], 2.0, np.arange(20, dtype=np.int32), np.arange(20, dtype=np.int32))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_small.npz"), meta=json.dumps(meta), init=init,
                    U=U, assign=assign, ind=ind, Zj=Zj, Zx=Zx, values=values, H=H)
print("wrote oracle_small.npz", meta)
