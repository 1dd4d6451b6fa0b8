"""Host logic of the hyper-parameter training (no GPU): the CCSA/MMA optimiser exported by the library against its
independent Python twin in the oracle and against known minima; the oracle's restatement of the reference's training
objective (src/train.cpp:333-436) against finite differences."""
import os

import numpy as np
import pytest

import flgp_b200 as F


def _rosen(x):
    v = 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2
    return v, np.array([-400 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]), 200 * (x[1] - x[0] ** 2)])


def _bowl(x):
    return (x[0] - 3) ** 2 + np.cosh(x[1] - 1), np.array([2 * (x[0] - 3), np.sinh(x[1] - 1)])


@pytest.mark.parametrize("f,x0,lb,ub,xmin", [
    (_bowl, [10.0, 1.0], [1e-3, 1e-4], [np.inf, np.inf], [3.0, 1.0]),          # the reference's start and bounds
    (_bowl, [10.0, 1.0], [4.0, 2.0], [np.inf, np.inf], [4.0, 2.0]),            # both bounds active
    (_rosen, [3.0, 3.0], [1.5, -np.inf], [np.inf, np.inf], [1.5, 2.25]),       # one bound active, curved valley
    (_bowl, [0.5, -2.0], [-5.0, -5.0], [5.0, 5.0], [3.0, 1.0]),                # finite box
])
def test_mma_matches_twin_and_minimum(oracle, f, x0, lb, ub, xmin):
    x, minf, nev = F.mma_minimize(f, x0, lb, ub, 1e-9, 5000)
    xo, mo, no = oracle.mma_minimize(f, x0, lb, ub, 1e-9, 5000)
    assert nev == no and np.array_equal(x, xo) and minf == mo      # same algorithm, same arithmetic
    np.testing.assert_allclose(x, xmin, rtol=1e-5, atol=1e-6)
    assert minf <= f(np.array(xmin))[0] + 1e-9


def test_mma_default_tolerance_and_eval_cap():
    x, minf, nev = F.mma_minimize(_bowl, [10.0, 1.0], [1e-3, 1e-4], [np.inf, np.inf])
    np.testing.assert_allclose(x, [3.0, 1.0], rtol=1e-3)
    x2, _, nev2 = F.mma_minimize(_rosen, [-1.2, 1.0], [-5, -5], [5, 5], 1e-12, 50)
    assert nev2 == 50


@pytest.mark.parametrize("m,K", [(40, 12), (10, 12)])
@pytest.mark.parametrize("approach", ["marginal", "posterior"])
def test_oracle_objective_gradient(oracle, m, K, approach):
    """Both branches (m > K Woodbury, m <= K direct) of the reference's objective: the analytic gradient the
    reference hands to nlopt equals central differences of its own objective (away from the clipping)."""
    rng = np.random.default_rng(m + K)
    n = 60
    V = rng.standard_normal((n, K)) * 1.2
    values = np.sort(rng.uniform(0.2, 1.0, K))[::-1].copy()
    idx = np.arange(m, dtype=np.int32)
    Y = V[:m, :3] @ np.array([1.0, -0.5, 0.3]) + 0.3 * rng.standard_normal(m)
    x = np.array([2.5, 3.0])
    f, g = oracle.regression_objective(V, values, Y, idx, K, x, 1e-5, approach)
    assert abs(g[1]) < 10.0
    for j in range(2):
        h = 1e-6 * x[j]
        xp, xm = x.copy(), x.copy()
        xp[j] += h
        xm[j] -= h
        fd = (oracle.regression_objective(V, values, Y, idx, K, xp, 1e-5, approach)[0] -
              oracle.regression_objective(V, values, Y, idx, K, xm, 1e-5, approach)[0]) / (2 * h)
        if m <= K or j == 1:
            np.testing.assert_allclose(g[j], fd, rtol=2e-5, atol=1e-7)
    # the Woodbury branch's value differs from the direct form by the reference's own "+1e-9" regularisers only:
    # its t-gradient is the exact gradient of the un-regularised likelihood
    if m > K:
        def direct(xx):
            ev = 1.0 - values[:K]
            Cm = (V[:m] * np.exp(-xx[0] * ev)) @ V[:m].T + (xx[1] + 1e-5) * np.eye(m)
            L = np.linalg.cholesky(Cm)
            a = np.linalg.solve(Cm, Y)
            v = 0.5 * Y @ a + np.log(np.diag(L)).sum()
            if approach == "posterior":
                v += np.log(xx[0] + 1e-9) + (xx[0] / 2.0) ** -10.0 + 1.1 * np.log(xx[1] + 1e-5) + 1e-3 / (xx[1] + 1e-5)
            return v
        for j in range(2):
            h = 1e-6 * x[j]
            xp, xm = x.copy(), x.copy()
            xp[j] += h
            xm[j] -= h
            np.testing.assert_allclose(g[j], (direct(xp) - direct(xm)) / (2 * h), rtol=2e-5, atol=1e-7)


def test_oracle_train_regression_improves_and_respects_bounds(oracle):
    """The reference clips grad[1] to +-10 (src/train.cpp:365-369, 409-413), so the gradient nlopt sees is not the
    objective's and a stationary point is not guaranteed; what must hold: the bounds, a decrease from the start
    x0 = (10, 1), and -obj = objective(pars)."""
    rng = np.random.default_rng(5)
    n, K, m = 300, 20, 120
    V = rng.standard_normal((n, K))
    values = np.linspace(1.0, 0.3, K)
    idx = np.arange(m, dtype=np.int32)
    Y = V[:m, :4] @ np.array([2.0, -1.0, 0.5, 0.25]) + 0.2 * rng.standard_normal(m)
    x, obj = oracle.train_regression(V, values, Y, idx, K, 1e-5, "posterior")
    assert x[0] >= 1e-3 and x[1] >= 1e-4
    f, _ = oracle.regression_objective(V, values, Y, idx, K, x, 1e-5, "posterior")
    f0, _ = oracle.regression_objective(V, values, Y, idx, K, (10.0, 1.0), 1e-5, "posterior")
    assert abs(f + obj) < 1e-9 and f < f0 - 50.0


# ------------------------------------------------------------------ logit training: COBYLA restatement (host only)
def _toy_logit_problem(seed=3, m=60, K=25):
    rng = np.random.default_rng(seed)
    n = 400
    V = np.linalg.qr(rng.standard_normal((n, K)))[0] * np.sqrt(n)
    values = np.sort(rng.uniform(0.2, 1.0, K))[::-1]
    values[0] = 1.0
    idx = np.arange(m, dtype=np.int32)
    f_true = V[:m, 1] * 1.5 + V[:m, 2]
    Y = (f_true + 0.3 * rng.standard_normal(m) > 0).astype(np.float64)
    return V, values, Y, idx


def test_cobyla_1d_library_matches_python_twin_and_scipy(oracle):
    """The one-variable COBYLA restatement: library (host code of libflgp_b200.so) == Python twin evaluation for
    evaluation, and both reach scipy's COBYLA optimum (an independent implementation of Powell's method) to the
    optimiser's tolerance on the Laplace objective of a small classifier and on two analytic functions."""
    import scipy.optimize as so
    import flgp_b200 as F

    V, values, Y, idx = _toy_logit_problem()
    K = V.shape[1]
    fobj = lambda t: oracle.logit_objective(V, values, Y, idx, K, t, 1e-3, "posterior")  # noqa: E731
    for f, x0 in ((fobj, 10.0), (lambda t: (np.log(t) - 1.3) ** 2 + 0.01 * t, 10.0), (lambda t: abs(t - 0.4) + 1.0, 10.0)):
        t_lib, f_lib, n_lib = F.cobyla_minimize_1d(f, x0)
        t_py, f_py, n_py = oracle.cobyla_minimize_1d(f, x0)
        assert n_lib == n_py and t_lib == t_py and f_lib == f_py
        ref = so.minimize(lambda x: f(float(x[0])), [x0], method="COBYLA", bounds=[(1e-3, None)],
                          options=dict(rhobeg=0.75 * (x0 - 1e-3), tol=1e-7, maxiter=2000))
        # xtol_rel = 1e-4 of the initial step (7.5e-4 here) is the resolution the reference asks of the optimiser
        assert f_lib <= ref.fun + 1e-3 * max(1.0, abs(ref.fun))
        assert abs(t_lib - float(ref.x[0])) <= 2e-3 * max(1.0, abs(float(ref.x[0])))


def test_laplace_mll_matches_direct_newton(oracle):
    """marginal_log_likelihood_logit_la_cpp restated: at the converged mode the value equals the textbook Laplace
    approximation log p(y|f) - f^T C^-1 f / 2 - log|I + W^1/2 C W^1/2| / 2 (GPML eq. 3.32)."""
    V, values, Y, idx = _toy_logit_problem(5, 40, 15)
    C = oracle.hk_from_spectrum(V, values, 15, 6.0, idx, idx)
    C[np.diag_indices(40)] += 1e-3
    got = oracle.laplace_mll(C, Y, tol=1e-12, max_iter=200)
    f = np.zeros(40)
    for _ in range(200):  # plain Newton on the log posterior
        pi = 1 / (1 + np.exp(-f))
        W = pi * (1 - pi)
        f = np.linalg.solve(np.eye(40) + C * W[None, :], C @ (W * f + Y - pi))
    pi = 1 / (1 + np.exp(-f))
    W = pi * (1 - pi)
    B = np.eye(40) + np.sqrt(W)[:, None] * C * np.sqrt(W)[None, :]
    want = (Y * np.log(pi) + (1 - Y) * np.log(1 - pi)).sum() - 0.5 * f @ np.linalg.solve(C, f) - 0.5 * np.linalg.slogdet(B)[1]
    assert abs(got - want) <= 1e-6 * abs(want)


def test_se_logit_grid_twins_select_the_largest_objective(oracle):
    """The oracle twins of fit_se_logit_gp_cpp / fit_se_logit_mult_gp_cpp (src/Fit.cpp:712-743, 839-867): the grid
    point with the largest (summed) objective wins, first one on ties; a one-point grid is that point's training;
    at a fixed t the objective is the plain logit objective of the winning spectrum."""
    rng = np.random.default_rng(4)
    n, m, s, r, K = 600, 50, 60, 3, 20
    ang = rng.uniform(0, 2 * np.pi, n)
    ring = rng.integers(0, 4, n)
    rad = 0.5 + 0.25 * ring + 0.05 * rng.standard_normal(n)
    X = np.asfortranarray(np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1))
    lab = np.where(rng.uniform(size=n) < 0.15, 1 - ring % 2, ring % 2).astype(np.float64)
    init = np.sort(rng.choice(n, s, replace=False)).astype(np.int32)
    a2s = np.array([0.2, 1.0, 5.0])
    idx0 = np.arange(m, dtype=np.int32)
    grid = list(oracle._se_grid(X[:m], X[m:], s, r, K, init, a2s, "cluster-normalized", True, 30, 2))
    objs = [-oracle.logit_objective(V, values, lab[:m], idx0, K, 6.0, 1e-3, "posterior") for _, values, V in grid]
    ref = oracle.fit_se_logit(X[:m], lab[:m], X[m:], s, r, K, init, a2s, iter_max=30, nthreads=2, t=6.0)
    assert ref["a2"] == a2s[int(np.argmax(objs))] and ref["obj"] == max(objs) and ref["t"] == 6.0
    assert ref["mean"].shape == (n - m,) and ref["C"].shape == (n, m) and np.all(ref["cov"] > 0)
    one = oracle.fit_se_logit(X[:m], lab[:m], X[m:], s, r, K, init, a2s[1:2], iter_max=30, nthreads=2)
    t1, o1, _ = oracle.train_lae_logit(grid[1][2], grid[1][1], lab[:m], idx0, K, 1e-3, "posterior")
    assert one["t"] == t1 and one["obj"] == o1
    lab3 = (ring % 3).astype(np.float64)
    lab3[:3] = [0, 1, 2]
    mult = oracle.fit_se_logit_mult(X[:m], lab3[:m], X[m:], s, r, K, init, a2s[:2], iter_max=30, nthreads=2)
    sums = [oracle.train_logit_mult(V, values, lab3[:m], idx0, K, 1e-3, "posterior")[1].sum() for _, values, V in grid[:2]]
    assert mult["a2"] == a2s[int(np.argmax(sums))] and abs(mult["obj"] - max(sums)) <= 1e-12 * abs(max(sums))
    assert len(mult["t"]) == 3


def test_nystrom_logit_twins_select_the_largest_objective(oracle):
    """The oracle twins of fit_nystrom_logit_gp_cpp / fit_nystrom_logit_mult_gp_cpp (src/Fit.cpp:942-990, 1087-1132)
    against their own pieces: grid selection by the largest (summed) objective, a fixed t evaluates the objective there,
    the regression twin's extension is the one the logit twins train on."""
    rng = np.random.default_rng(8)
    n, m, s, K = 500, 40, 50, 15
    ang = rng.uniform(0, 2 * np.pi, n)
    ring = rng.integers(0, 4, n)
    rad = 0.5 + 0.25 * ring + 0.05 * rng.standard_normal(n)
    X = np.asfortranarray(np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1))
    lab = np.where(rng.uniform(size=n) < 0.15, 1 - ring % 2, ring % 2).astype(np.float64)
    init = np.sort(rng.choice(n, s, replace=False)).astype(np.int32)
    a2s = np.array([0.3, 3.0])
    idx0 = np.arange(m, dtype=np.int32)
    grid = [(a2, values, extend(D_all[:m])) for a2, values, extend, D_all in
            oracle._nystrom_grid(X[:m], X[m:], s, K, init, a2s, 30, 2)]
    objs = [-oracle.logit_objective(Vm, values, lab[:m], idx0, K, 4.0, 1e-3, "marginal") for _, values, Vm in grid]
    ref = oracle.fit_nystrom_logit(X[:m], lab[:m], X[m:], s, K, init, a2s, approach="marginal", iter_max=30, nthreads=2,
                                   t=4.0)
    assert ref["a2"] == a2s[int(np.argmax(objs))] and ref["obj"] == max(objs)
    assert ref["V"].shape == (n, K) and ref["C"].shape == (n, m) and np.all(np.isfinite(ref["mean"]))
    lab3 = (ring % 3).astype(np.float64)
    lab3[:3] = [0, 1, 2]
    mult = oracle.fit_nystrom_logit_mult(X[:m], lab3[:m], X[m:], s, K, init, a2s, iter_max=30, nthreads=2)
    sums = [oracle.train_logit_mult(Vm, values, lab3[:m], idx0, K, 1e-3, "posterior")[1].sum() for _, values, Vm in grid]
    assert mult["a2"] == a2s[int(np.argmax(sums))] and abs(mult["obj"] - max(sums)) <= 1e-12 * abs(max(sums))


def test_small_host_exports_match_the_oracle(oracle):
    """The reference's small exported helpers behind the C ABI (host algebra, usable without a GPU):
    marginal_log_likelihood_logit_la_cpp against the oracle's laplace_mll, multi_train_split, negative_log_likelihood
    (type "regression", the reference's literal constant), test_regression_cpp against a dense solve."""
    V, values, Y, idx = _toy_logit_problem(seed=5, m=50, K=20)
    K = V.shape[1]
    Cm = oracle.hk_from_spectrum(V, values, K, 6.0, idx, idx)
    Cm[np.diag_indices(len(idx))] += 1e-3
    N = np.where(np.arange(len(idx)) % 3 == 0, 2.0, 1.0)
    Yn = np.minimum(Y, N)
    for nn, yy in ((None, Y), (N, Yn)):
        got = F.marginal_log_likelihood_logit_la_cpp(Cm, yy, nn)
        want = oracle.laplace_mll(Cm, yy, nn)
        assert abs(got - want) <= 1e-10 * max(1.0, abs(want))
    lab = np.array([2, 0, 1, 1, 3, 0], dtype=np.float64)
    aug = F.multi_train_split(lab)
    assert aug.shape == (6, 4) and np.array_equal(aug, (lab[:, None] == np.arange(4)[None, :]).astype(np.float64))
    with pytest.raises(F.FlgpError):
        F.multi_train_split([0.5, 1.0])
    rng = np.random.default_rng(1)
    mean, cov, tgt = rng.standard_normal(40), rng.uniform(0.1, 2.0, 40), rng.standard_normal(40)
    want = (np.mean((tgt - mean) ** 2 / cov + np.log(cov + 1e-9)) + np.log(2 * 3.1415926)) / 2
    assert abs(F.negative_log_likelihood(mean, cov, tgt, "regression") - want) <= 1e-13 * abs(want)
    with pytest.raises(F.FlgpError, match="RNG"):
        F.negative_log_likelihood(mean, cov, tgt, "binary")
    A = rng.standard_normal((30, 30))
    Cs = A @ A.T + 30 * np.eye(30)
    Cnv = rng.standard_normal((7, 30))
    y = rng.standard_normal(30)
    np.testing.assert_allclose(F.test_regression_cpp(Cs, y, Cnv), Cnv @ np.linalg.solve(Cs, y), rtol=1e-11, atol=1e-13)
    with pytest.raises(F.FlgpError, match="positive definite"):
        F.test_regression_cpp(-Cs, y, Cnv)


@pytest.mark.parametrize("m,K", [(40, 12), (10, 16)])
def test_diff_noise_objective_training_and_prediction_rows(oracle, m, K):
    """noise = "different" (src/train.cpp:438-556, src/Predict.cpp:76-113) on explicit training rows, library host code
    against the oracle's literal numpy restatement: objective 1e-12 and gradient 1e-10 in both branches (m > K with the
    reference's clipping, m <= K) and both approaches; the (m + 1)-variable MMA training to optimiser tolerance; the
    folded prediction coefficient against the literal prediction."""
    rng = np.random.default_rng(m)
    n = 200
    V = np.linalg.qr(rng.standard_normal((n, K)))[0] * np.sqrt(n)
    values = np.sort(rng.uniform(0.2, 1.0, K))[::-1]
    values[0] = 1.0
    idx = np.arange(m, dtype=np.int32)
    Y = V[:m, 1] * 1.2 + V[:m, 2] + 0.3 * rng.standard_normal(m)
    x = np.concatenate([[3.0], rng.uniform(0.05, 2.0, m)])
    for approach in ("marginal", "posterior"):
        fo, go = oracle.regression_objective_diff(V, values, Y, idx, K, x, 1e-5, approach)
        fl, gl = F.regression_objective_diff_rows(V[:m], values, Y, x, 1e-5, approach)
        assert abs(fo - fl) <= 1e-12 * max(1.0, abs(fo))
        np.testing.assert_allclose(gl, go, rtol=1e-10, atol=1e-12)
        if m > K:
            assert np.abs(go[1:]).max() <= 1.0 + (0.0 if approach == "marginal" else 1.2 / (m * 0.05))   # clipped
        h = 1e-6   # the gradient of the time is never clipped: central difference
        xp, xm = x.copy(), x.copy()
        xp[0] += h
        xm[0] -= h
        fd = (oracle.regression_objective_diff(V, values, Y, idx, K, xp, 1e-5, approach)[0]
              - oracle.regression_objective_diff(V, values, Y, idx, K, xm, 1e-5, approach)[0]) / (2 * h)
        assert abs(fd - go[0]) <= 1e-6 * max(1.0, abs(fd))
        # training: the optimiser is pinned against its twin on a common objective in test_mma_matches_twin_and_minimum
        # and the objective above; two runs on objectives that differ in the last bit may still take different paths over
        # a flat landscape, so the composed run is checked for what does not depend on the path
        x0 = np.concatenate([[10.0], np.ones(m)])
        xl, ol, nev = F.train_regression_diff_rows(V[:m], values, Y, 1e-5, approach)
        assert nev <= 1000 and xl[0] >= 1e-3 and xl[1:].min() >= 1e-4
        f_at = oracle.regression_objective_diff(V, values, Y, idx, K, xl, 1e-5, approach)[0]
        assert abs(-f_at - ol) <= 1e-10 * max(1.0, abs(ol))                       # obj = -(objective at the optimum)
        assert ol > -oracle.regression_objective_diff(V, values, Y, idx, K, x0, 1e-5, approach)[0]   # improved
        xo, oo = oracle.train_regression_diff(V, values, Y, idx, K, 1e-5, approach)
        if abs(ol - oo) <= 1e-6 * max(1.0, abs(oo)):                               # same basin: same point
            np.testing.assert_allclose(xl, xo, rtol=5e-3, atol=5e-3)
        coef = F.predict_coef_diff_rows(V[:m], values, Y, xl, 1e-5)
        pred = oracle.predict_regression_diff(V, values, Y, idx, np.arange(m, n), K, xl, 1e-5)
        np.testing.assert_allclose(V[m:, :K] @ coef, pred, rtol=1e-8, atol=1e-9 * np.abs(pred).max())
    with pytest.raises(F.FlgpError):
        F.regression_objective_diff_rows(V[:m], values, Y, x[:-1])


@pytest.mark.parametrize("m,K", [(60, 25), (20, 30)])
def test_classification_fold_rows_matches_the_literal_posterior(oracle, m, K):
    """posterior_distribution_classification (src/Utils.cpp:252-299) folded onto the eigenvector rows (library host code:
    Newton mode, coef = Lam V1^T (Y - pi), Mq = Lam - Lam V1^T beta V1 Lam through solves with the factor of B) against
    the oracle's literal restatement on explicit covariance blocks: mean and variance of the unlabelled rows to 1e-9."""
    V, values, Y, idx = _toy_logit_problem(seed=m, m=m, K=K)
    n, sigma = len(V), 1e-3
    idx1 = np.arange(m, n, dtype=np.int32)
    for t in (2.0, 9.0):
        coef, Mq = F.classification_fold_rows(V[:m], values, Y, t, sigma)
        mean = V[m:, :K] @ coef
        cov = ((V[m:, :K] @ Mq) * V[m:, :K]).sum(axis=1) + sigma
        C11 = oracle.hk_from_spectrum(V, values, K, t, idx, idx)
        C11[np.diag_indices(m)] += sigma
        C21 = oracle.hk_from_spectrum(V, values, K, t, idx1, idx)
        C22 = ((V[m:, :K] * np.exp(-t * (1.0 - values[:K]))) * V[m:, :K]).sum(axis=1) + sigma
        mo, co = oracle.posterior_distribution_classification(C11, C21, C22, Y)
        np.testing.assert_allclose(mean, mo, rtol=1e-9, atol=1e-10 * max(1.0, np.abs(mo).max()))
        np.testing.assert_allclose(cov, co, rtol=1e-8, atol=1e-9 * max(1.0, np.abs(co).max()))
        np.testing.assert_allclose(Mq, Mq.T, rtol=0, atol=1e-12 * np.abs(Mq).max())


def test_host_cholesky_is_the_textbook_form_bit_for_bit():
    """The library's blocked, column-oriented Cholesky / substitutions (csrc/capi.cu: chol_lower, chol_solve) give the
    bits of the textbook dot-product forms (every element updated in ascending k), across panel edges (n = 31, 32, 33,
    70): checked through test_regression_cpp = Cnv (L L^T)^{-1} Y against a plain Python transcription."""
    rng = np.random.default_rng(3)
    for n in (1, 5, 31, 32, 33, 70):
        A = rng.standard_normal((n, n))
        Cs = A @ A.T + n * np.eye(n)
        y = rng.standard_normal(n)
        Cnv = rng.standard_normal((3, n))
        L = [[0.0] * n for _ in range(n)]
        for j in range(n):
            d = float(Cs[j, j])
            for k in range(j):
                d -= L[j][k] * L[j][k]
            d = d ** 0.5
            L[j][j] = d
            for i in range(j + 1, n):
                v = float(Cs[i, j])
                for k in range(j):
                    v -= L[i][k] * L[j][k]
                L[i][j] = v / d
        b = [float(v) for v in y]
        for i in range(n):
            v = b[i]
            for k in range(i):
                v -= L[i][k] * b[k]
            b[i] = v / L[i][i]
        for i in range(n - 1, -1, -1):
            v = b[i]
            for k in range(i + 1, n):
                v -= L[k][i] * b[k]
            b[i] = v / L[i][i]
        want = []
        for i in range(3):
            acc = 0.0
            for j in range(n):
                acc += float(Cnv[i, j]) * b[j]
            want.append(acc)
        assert np.array_equal(F.test_regression_cpp(Cs, y, Cnv), np.array(want))


def test_logit_objective_and_training_rows_match_the_twin(oracle):
    """The host code of the logit training on explicit labelled rows (what flgp_logit_objective / flgp_train_logit run
    after fetching the rows from the handle): objective against the oracle's literal restatement (1e-10, both approaches,
    with trial counts N; m large enough for the threaded covariance build), the COBYLA training against the twin
    evaluation for evaluation, and the same objective for any number of host threads."""
    import subprocess
    import sys

    for m, K in ((60, 25), (300, 40)):
        V, values, Y, idx = _toy_logit_problem(seed=m, m=m, K=K)
        N = np.where(np.arange(m) % 4 == 0, 3.0, 1.0)
        for approach in ("posterior", "marginal"):
            for t in (1.5, 12.0):
                got = F.logit_objective_rows(V[:m], values, Y, t, 1e-3, approach)
                want = oracle.logit_objective(V, values, Y, idx, K, t, 1e-3, approach)
                assert abs(got - want) <= 1e-10 * max(1.0, abs(want))
            got = F.logit_objective_rows(V[:m], values, np.minimum(Y * 2, N), 5.0, 1e-3, approach, N=N)
            want = oracle.logit_objective(V, values, np.minimum(Y * 2, N), idx, K, 5.0, 1e-3, approach, N)
            assert abs(got - want) <= 1e-10 * max(1.0, abs(want))
        if m == 60:
            t_l, o_l, n_l = F.train_logit_rows(V[:m], values, Y, 1e-3, "posterior")
            f_lib = lambda t: F.logit_objective_rows(V[:m], values, Y, t, 1e-3, "posterior")  # noqa: E731
            t_p, f_p, n_p = oracle.cobyla_minimize_1d(f_lib, 10.0)      # the twin optimiser on the library's objective
            assert (t_l, -o_l, n_l) == (t_p, f_p, n_p)
            t_o, o_o, _ = oracle.train_lae_logit(V, values, Y, idx, K, 1e-3, "posterior")
            assert abs(t_l - t_o) <= 2e-3 * max(1.0, t_o) and abs(o_l - o_o) <= 1e-6 * max(1.0, abs(o_o))
    # thread-count independence of the bits (m = 300 uses the threaded covariance build and, from m >= 416, the Cholesky)
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, flgp_b200 as F\n"
            "rng = np.random.default_rng(1); m, K = 500, 30\n"
            "V = np.linalg.qr(rng.standard_normal((m, K)))[0] * np.sqrt(m); vals = np.sort(rng.uniform(0.2, 1, K))[::-1]\n"
            "Y = (V[:, 1] + 0.3 * rng.standard_normal(m) > 0).astype(float)\n"
            "print(repr(F.logit_objective_rows(V, vals, Y, 4.0)))" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for threads in ("1", "3", "8"):
        env = dict(os.environ, FLGP_HOST_THREADS=threads)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1] == outs[2] and "nan" not in outs[0]


@pytest.mark.parametrize("m,K", [(80, 20), (12, 20)])
def test_regression_objective_and_training_rows_match_the_twin(oracle, m, K):
    """noise = "same" on explicit training rows, the host code of flgp_regression_objective / flgp_train_regression:
    value 1e-11 and gradient 1e-9 against the oracle's literal restatement in both branches (m > K through V^T V, m <= K)
    and both approaches; the MMA training is the twin's run on the library's objective, step for step."""
    rng = np.random.default_rng(m + K)
    n = 300
    V = np.linalg.qr(rng.standard_normal((n, K)))[0] * np.sqrt(n)
    values = np.sort(rng.uniform(0.3, 1.0, K))[::-1]
    values[0] = 1.0
    idx = np.arange(m, dtype=np.int32)
    Y = V[:m, 1] + 0.7 * V[:m, 3] + 0.2 * rng.standard_normal(m)
    for approach in ("marginal", "posterior"):
        for x in ((6.0, 0.4), (1.2, 0.02), (30.0, 3.0)):
            fo, go = oracle.regression_objective(V, values, Y, idx, K, x, 1e-5, approach)
            fl, gl = F.regression_objective_rows(V[:m], values, Y, x, 1e-5, approach)
            assert abs(fo - fl) <= 1e-11 * max(1.0, abs(fo))
            np.testing.assert_allclose(gl, go, rtol=1e-9, atol=1e-10)
        xl, ol, nev = F.train_regression_rows(V[:m], values, Y, 1e-5, approach)
        f_lib = lambda z: F.regression_objective_rows(V[:m], values, Y, z, 1e-5, approach)  # noqa: E731
        xt, ft, nt = oracle.mma_minimize(f_lib, (10.0, 1.0), (1e-3, 1e-4), (np.inf, np.inf))
        assert nt == nev and np.array_equal(xt, xl) and -ft == ol
        xo, oo = oracle.train_regression(V, values, Y, idx, K, 1e-5, approach)
        assert abs(ol - oo) <= 1e-6 * max(1.0, abs(oo))


def test_posterior_distribution_multiclassification_against_the_literal_form(oracle):
    """posterior_distribution_multiclassification (src/Utils.cpp:339-370; C11 without sigma, C22 with): the Python mirror
    (library host fold on the labelled rows + two products per block of rows) against the oracle's literal restatement
    on explicit covariance blocks.  The eigenpair is a stand-in with the two members the mirror reads (.rows, .values),
    so the composition is checked on the CPU box; on the GPU the same members come from the spectrum handle."""
    V, values, _, idx = _toy_logit_problem(seed=9, m=70, K=22)
    n, m, K, sigma = len(V), 70, 22, 1e-3
    rng = np.random.default_rng(9)
    lab = rng.integers(0, 3, m).astype(np.float64)
    lab[:3] = [0, 1, 2]
    ts = np.array([3.0, 8.0, 15.0])

    class FakePair:
        n_local = n

        def __init__(self):
            self.values = values

        def rows(self, ii):
            return np.asfortranarray(V[np.asarray(ii)])

    idx_new = np.arange(m, n, dtype=np.int32)
    mean, cov = F.posterior_distribution_multiclassification(FakePair(), lab, m, K, ts, sigma, block=97)
    mo, co = oracle.posterior_distribution_multiclassification(V, values, lab, idx, idx_new, K, ts, sigma)
    np.testing.assert_allclose(mean, mo, rtol=1e-8, atol=1e-9 * np.abs(mo).max())
    np.testing.assert_allclose(cov, co, rtol=1e-8, atol=1e-9 * np.abs(co).max())
    sub = np.array([75, 80, 399], dtype=np.int32)
    m2, c2 = F.posterior_distribution_multiclassification(FakePair(), lab, m, K, ts, sigma, idx_new=sub)
    np.testing.assert_allclose(m2, mo[sub - m], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(c2, co[sub - m], rtol=1e-8, atol=1e-10)
