"""Host logic of the hyper-parameter training (no GPU): the CCSA/MMA optimiser exported by the library against its
independent Python twin in the oracle and against known minima; the oracle's restatement of the reference's training
objective (src/train.cpp:333-436) against finite differences."""
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
