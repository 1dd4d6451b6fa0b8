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

# ---- second fixture: training objective, optimiser, SE grid, Nystrom driver (the round's widening rows) -------------
m = 120
idx = np.arange(m, dtype=np.int32)
obj_m, grad_m = O.regression_objective(V, values, Y[:m], idx, meta["K"], (6.0, 0.4), 1e-5, "marginal")
obj_p, grad_p = O.regression_objective(V, values, Y[:m], idx, meta["K"], (6.0, 0.4), 1e-5, "posterior")
obj_s, grad_s = O.regression_objective(V, values, Y[:8], idx[:8], meta["K"], (6.0, 0.4), 1e-5, "posterior")  # m <= K
pars, obj_t = O.train_regression(V, values, Y[:m], idx, meta["K"], 1e-5, "posterior")
a2s = np.array([0.3, 1.0, 3.0])
se = O.fit_se_regression(X[:m], Y[:m], X[m:], meta["s"], meta["r"], meta["K"], init, a2s, pars=(6.0, 0.4), iter_max=30)
ny = O.fit_nystrom_regression(X[:m], Y[:m], X[m:], meta["s"], meta["K"], init, a2s, pars=(6.0, 0.4), iter_max=30)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_train.npz"), m=m, a2s=a2s,
                    obj=np.array([obj_m, obj_p, obj_s]), grad=np.array([grad_m, grad_p, grad_s]), pars=pars, obj_t=obj_t,
                    se_a2=se["a2"], se_obj=se["obj"], se_test=se["test"][:200], se_cov=se["cov"][:200],
                    ny_a2=ny["a2"], ny_obj=ny["obj"], ny_test=ny["test"][:200], ny_cov=ny["cov"][:200])
print("wrote oracle_train.npz: pars", pars, "se a2", se["a2"], "nystrom a2", ny["a2"])

# ---- third fixture: the callers added late in round 2 (mini-batch subsample, noise = "different", the SE / Nystrom
# logit grids at a fixed diffusion time) ---------------------------------------------------------------------------
Umb, it_mb = O.minibatch_kmeans(X, meta["s"], init, max_iters=40, seed=9)
lab = (Y > np.median(Y)).astype(np.float64)
xd = np.concatenate([[6.0], 0.1 + 0.4 * (np.arange(m) % 7) / 7.0])
obj_d, grad_d = O.regression_objective_diff(V, values, Y[:m], idx, meta["K"], xd, 1e-5, "posterior")
pred_d = O.predict_regression_diff(V, values, Y[:m], idx, np.arange(m, meta["n"], dtype=np.int32), meta["K"], xd, 1e-5)
sl = O.fit_se_logit(X[:m], lab[:m], X[m:], meta["s"], meta["r"], meta["K"], init, a2s, iter_max=30, t=6.0)
nl = O.fit_nystrom_logit(X[:m], lab[:m], X[m:], meta["s"], meta["K"], init, a2s, iter_max=30, t=6.0)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_round2b.npz"), Umb=Umb, it_mb=it_mb, xd=xd,
                    obj_d=obj_d, grad_d=grad_d, pred_d=pred_d[:200],
                    sl_a2=sl["a2"], sl_obj=sl["obj"], sl_mean=sl["mean"][:200], sl_cov=sl["cov"][:200],
                    nl_a2=nl["a2"], nl_obj=nl["obj"], nl_mean=nl["mean"][:200], nl_cov=nl["cov"][:200])
print("wrote oracle_round2b.npz: minibatch batches", it_mb, "se logit a2", sl["a2"], sl["obj"], "nystrom logit a2",
      nl["a2"], nl["obj"])
