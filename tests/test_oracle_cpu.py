"""CPU tests of the oracle itself: each C++ stage against an independent numpy twin, the properties the
domain offers, and the committed golden fixtures.  (The reference ships no tests or vectors: SURVEY §4.)"""
import json
import os

import numpy as np
import pytest

from conftest import spiral, swiss

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fixed_point_roundtrip_and_order_independence(oracle):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(5000) * 17.0
    hi, lo = oracle.fx_encode(x, np.abs(x).max(), x.size)
    back = oracle.fx_decode(hi, lo, np.abs(x).max(), x.size)
    assert np.max(np.abs(back - x)) <= 1e-18 * 32  # dropped bits are ~2^-70 of the scale
    perm = rng.permutation(x.size)
    assert hi.sum() == hi[perm].sum() and lo.sum() == lo[perm].sum()
    s = oracle.fx_decode(np.array([hi.sum()]), np.array([lo.sum()]), np.abs(x).max(), x.size)[0]
    import math

    assert abs(s - math.fsum(x)) <= 1e-12 * abs(math.fsum(x)) + 1e-12


def test_kmeans_matches_numpy_lloyd(oracle):
    X, _ = spiral(1500, 3)
    s = 40
    rng = np.random.default_rng(0)
    init = np.sort(rng.choice(len(X), s, replace=False))
    U, assign, iters = oracle.kmeans_lloyd(X, s, init, 100)
    # numpy twin (plain fp64 means): same partition unless a point sits on a bisector to 1e-12
    C = X[init].copy()
    a_prev = None
    for _ in range(100):
        D = (C ** 2).sum(1)[None, :] - 2 * X @ C.T
        a = D.argmin(1)
        if a_prev is not None and np.array_equal(a, a_prev):
            break
        a_prev = a
        for j in range(s):
            if (a == j).any():
                C[j] = X[a == j].mean(0)
    assert np.mean(a == assign) > 0.999
    assert U[:, 2].sum() == len(X)
    assert np.array_equal(np.bincount(assign, minlength=s), U[:, 2].astype(int))
    # centres are the means of their members
    for j in range(0, s, 7):
        np.testing.assert_allclose(U[j, :2], X[assign == j].mean(0), rtol=1e-12)
    assert 1 <= iters <= 100


def test_kmeans_thread_invariance(oracle):
    X, _ = swiss(4000, 5)
    init = np.arange(0, 4000, 80, dtype=np.int32)
    U1, a1, i1 = oracle.kmeans_lloyd(X, len(init), init, 100, nthreads=1)
    U4, a4, i4 = oracle.kmeans_lloyd(X, len(init), init, 100, nthreads=4)
    assert i1 == i4 and np.array_equal(a1, a4) and np.array_equal(U1, U4)  # bit-exact: integer limbs


def test_knn_matches_argsort_and_is_sorted(oracle):
    X, _ = swiss(2000, 7)
    U = X[::20].copy()
    ind, dist = oracle.knn(X, U, 4, want_dist=True, nthreads=3)
    D = ((-2 * (X @ U.T)) + (X ** 2).sum(1)[:, None]) + (U ** 2).sum(1)[None, :]
    ref = np.argsort(D, axis=1, kind="stable")[:, :4]
    assert np.array_equal(ind, ref)
    assert np.all(np.diff(dist, axis=1) >= 0)
    np.testing.assert_allclose(dist, np.take_along_axis(D, ref, 1), rtol=1e-12, atol=1e-12)


def test_knn_errors(oracle):
    X, _ = spiral(10)
    with pytest.raises(ValueError):
        oracle.knn(X, X[:3], 4)


def test_simplex_projection_properties(oracle):
    rng = np.random.default_rng(2)
    for r in (1, 2, 3, 5, 9):
        for _ in range(50):
            v = rng.standard_normal(r) * rng.choice([0.1, 1, 10])
            z = oracle.simplex_project(v)
            assert np.all(z >= 0) and abs(z.sum() - 1) < 1e-12
            # optimality: z = max(v - theta, 0) for a single theta
            th = (v - z)[z > 0]
            assert np.ptp(th) < 1e-12
            # idempotent on the simplex
            np.testing.assert_allclose(oracle.simplex_project(z), z, atol=1e-15)


def _lae_numpy(x, U, T=100, tol=1e-5):
    """Independent numpy restatement of src/lae.cpp:76-133 (vectorised ops, same control flow)."""
    r = U.shape[0]
    zp = np.full(r, 1.0 / r)
    zc = zp.copy()
    dp, dc, bc = 0.0, 1.0, 1.0
    UUt = U @ U.T

    def proj(v):
        vd = np.sort(v)[::-1]
        cs = np.cumsum(vd)
        vs = vd - (cs - 1) / np.arange(1, r + 1)
        rho = np.nonzero(vs > 0)[0].max() + 1
        th = (vd[:rho].sum() - 1.0) / rho
        return np.maximum(v - th, 0)

    for _ in range(T):
        al = (dp - 1) / dc
        v = zc + al * (zc - zp)
        gv = ((x - v @ U) ** 2).sum() / 2
        g = v @ UUt - x @ U.T
        j = 0
        while True:
            b = 2.0 ** j * bc
            z = proj(v - g / b)
            gz = ((x - z @ U) ** 2).sum() / 2
            gt = gv + g @ (z - v) + b * ((z - v) ** 2).sum() / 2
            if gz <= gt:
                bc = b
                zp, zc = zc, z
                break
            j += 1
        dp, dc = dc, (1 + np.sqrt(1 + 4 * dc * dc)) / 2
        if ((zc - zp) ** 2).sum() < tol:
            break
    return zc


def test_lae_point_matches_numpy_twin(oracle):
    rng = np.random.default_rng(4)
    worst = 0.0
    for _ in range(200):
        r, d = rng.choice([2, 3, 5]), rng.choice([2, 3, 6])
        U = rng.standard_normal((r, d)) * rng.choice([0.5, 3.0])
        x = U.mean(0) + 0.3 * rng.standard_normal(d)
        z = oracle.lae_point(x, U)
        zn = _lae_numpy(x, U)
        assert np.all(z >= 0) and abs(z.sum() - 1) < 1e-12
        worst = max(worst, np.abs(z - zn).max())
    # summation order differs (numpy pairwise/BLAS vs sequential) so borderline branch flips are possible;
    # they are rare and small
    assert worst < 1e-6


def test_lae_reconstruction_improves_on_uniform(oracle):
    X, _ = swiss(500, 8)
    U = X[::10].copy()
    Zj, Zx, W = oracle.lae(X, U, 3)
    assert np.allclose(Zx.sum(1), 1.0) and np.all(Zx >= 0)
    assert np.all(np.diff(Zj, axis=1) > 0)  # rows sorted by column, distinct
    rec = np.einsum("ij,ijk->ik", Zx, U[Zj])
    uni = U[Zj].mean(1)
    assert (np.linalg.norm(X - rec, axis=1) <= np.linalg.norm(X - uni, axis=1) + 1e-9).mean() > 0.95


def _dense(Zj, Zx, s):
    n, r = Zj.shape
    Z = np.zeros((n, s))
    np.put_along_axis(Z, Zj, Zx, 1)
    return Z


@pytest.mark.parametrize("mode", ["rw", "normalized", "cluster-normalized"])
def test_graph_laplacian_matches_dense_numpy(oracle, mode):
    X, _ = spiral(800, 2)
    U = X[::16].copy()
    s = len(U)
    Zj, Zx, _ = oracle.lae(X, U, 3)
    nc = np.arange(1, s + 1, dtype=float)
    Z = _dense(Zj, Zx, s)
    if mode != "rw":
        Z = Z * (1.0 / (Z.sum(0) + 1e-9))
    if mode == "cluster-normalized":
        Z = Z * nc
    Z = (1.0 / (Z.sum(1) + 1e-9))[:, None] * Z
    for exact in (0, 1):
        got = oracle.graph_laplacian(Zj, Zx, s, mode, nc, exact)
        np.testing.assert_allclose(_dense(Zj, got, s), Z, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-6)  # the +1e-9 regulariser is visible after scaling
    with pytest.raises(ValueError):
        oracle.graph_laplacian(Zj, Zx, s, "bogus")


def test_spectrum_against_dense_svd_and_properties(oracle):
    X, _ = spiral(1200, 6)
    rng = np.random.default_rng(0)
    s, r, K = 60, 3, 12
    init = np.sort(rng.choice(len(X), s, replace=False))
    U, _, _ = oracle.kmeans_lloyd(X, s, init)
    Zj, Zx = oracle.cross_similarity_lae(X, U, r, "cluster-normalized")
    values, V, I = oracle.spectrum_from_Z(Zj, Zx, s, K, root=True, want_internals=True)
    Z = _dense(Zj, Zx, s)
    A = Z * (1.0 / np.sqrt(np.abs(Z.sum(0)) + 1e-9))
    u, sv, _ = np.linalg.svd(A, full_matrices=False)
    np.testing.assert_allclose(values, sv[:K], rtol=1e-10)
    assert abs(values[0] - 1.0) < 1e-6  # row-stochastic Z: top singular value 1
    # lifted vectors: sqrt(n) * left singular vectors, up to sign
    for k in range(K):
        c = abs(np.dot(V[:, k], u[:, k])) / np.sqrt(len(X))
        assert abs(c - 1) < 1e-8
    np.testing.assert_allclose(V.T @ V / len(X), np.eye(K), atol=1e-9)
    # exact (fixed-point) and sequential sums agree far below the 1e-8 contract
    v0, V0 = oracle.spectrum_from_Z(Zj, Zx, s, K, root=True, exact=0)
    np.testing.assert_allclose(v0, values, rtol=1e-12)
    # heat kernel: symmetric PSD on a subset
    idx = np.arange(50, dtype=np.int32)
    H = oracle.hk_from_spectrum(V, values, K, 2.0, idx, idx)
    np.testing.assert_allclose(H, H.T, atol=1e-12)
    assert np.linalg.eigvalsh(H).min() > -1e-10
    Hn = (V[idx] * np.exp(-2.0 * (1 - values))) @ V[idx].T
    np.testing.assert_allclose(H, Hn, rtol=1e-12, atol=1e-12)


def test_regression_tail_branches_agree_with_direct_gp(oracle):
    """Both branches of predict_regression_cpp / posterior_covariance_regression equal the textbook GP."""
    X, Y = spiral(900, 9)
    rng = np.random.default_rng(1)
    s, r = 50, 3
    init = np.sort(rng.choice(len(X), s, replace=False))
    for m, K in ((30, 40), (120, 25)):  # m <= K and m > K
        values, V = oracle.heat_kernel_spectrum(X[:m], X[m:], s, r, K, init)
        idx0 = np.arange(m, dtype=np.int32)
        idx1 = np.arange(m, len(X), dtype=np.int32)
        pars, sigma = (3.0, 0.05), 1e-3
        mean = oracle.predict_regression(V, values, Y[:m], idx0, idx1, K, pars, sigma)
        cov = oracle.posterior_covariance_regression(V, values, idx0, idx1, K, pars, sigma)
        lam = np.exp(-pars[0] * (1 - values[:K]))
        C = (V[:, :K] * lam) @ V[:, :K].T
        Kvv = C[:m, :m] + (pars[1] + sigma) * np.eye(m)
        ref_mean = C[m:, :m] @ np.linalg.solve(Kvv, Y[:m])
        ref_cov = np.diag(C)[m:] + pars[1] + sigma - np.einsum("ij,ji->i", C[m:, :m], np.linalg.solve(Kvv, C[:m, m:]))
        np.testing.assert_allclose(mean, ref_mean, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(cov, ref_cov, rtol=1e-7, atol=1e-9)


def test_golden_fixture(oracle):
    """Oracle outputs pinned by the committed fixture (generated by tests/golden/make_golden.py from THIS oracle;
    the reference cannot run here, so this guards against drift, not against the reference)."""
    path = os.path.join(GOLD, "oracle_small.npz")
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    X, Y = spiral(meta["n"], meta["seed"])
    init = g["init"]
    U, assign, iters = oracle.kmeans_lloyd(X, meta["s"], init)
    assert iters == meta["iters"]
    assert np.array_equal(assign, g["assign"]) and np.array_equal(U, g["U"])
    ind = oracle.knn(X, U[:, :2], meta["r"])
    assert np.array_equal(ind, g["ind"])
    Zj, Zx = oracle.cross_similarity_lae(X, U, meta["r"], "cluster-normalized")
    assert np.array_equal(Zj, g["Zj"]) and np.array_equal(Zx, g["Zx"])
    values, V = oracle.spectrum_from_Z(Zj, Zx, meta["s"], meta["K"], True)
    np.testing.assert_allclose(values, g["values"], rtol=1e-11)


def test_nystrom_oracle_is_a_sane_regressor(oracle):
    """The dense restatement of fit_nystrom_regression_gp_cpp (oracle.fit_nystrom_regression) on the README spiral:
    extension rows are convex-like weights, predictions track the signal."""
    from conftest import spiral

    X, Y = spiral(1200, 9)
    m, s, K = 150, 100, 40
    init = np.sort(np.random.default_rng(1).choice(1200, s, replace=False)).astype(np.int32)
    res = oracle.fit_nystrom_regression(X[:m], Y[:m], X[m:], s, K, init, [0.5, 2.0], pars=(2.0, 1.0), iter_max=15)
    assert res["a2"] in (0.5, 2.0) and np.isfinite(res["obj"])
    assert np.all(np.isfinite(res["test"])) and np.all(res["cov"] > 0)
    assert np.sqrt(np.mean((res["test"] - Y[m:]) ** 2)) < 0.7 * np.std(Y[m:])  # labels carry N(0,1) noise


def test_golden_fixture_training_rows(oracle):
    """The round's widening rows (training objective, MMA optimiser, SE grid, Nystrom driver) pinned by the committed
    fixture tests/golden/oracle_train.npz (same generator script; guards the oracle against drift)."""
    g0 = np.load(os.path.join(GOLD, "oracle_small.npz"))
    g = np.load(os.path.join(GOLD, "oracle_train.npz"))
    meta = json.loads(str(g0["meta"]))
    X, Y = spiral(meta["n"], meta["seed"])
    init, m, K = g0["init"], int(g["m"]), meta["K"]
    U, _, _ = oracle.kmeans_lloyd(X, meta["s"], init)
    Zj, Zx = oracle.cross_similarity_lae(X, U, meta["r"], "cluster-normalized")
    values, V = oracle.spectrum_from_Z(Zj, Zx, meta["s"], K, True)
    idx = np.arange(m, dtype=np.int32)
    cases = [(Y[:m], idx, "marginal"), (Y[:m], idx, "posterior"), (Y[:8], idx[:8], "posterior")]
    for q, (yy, ii, ap) in enumerate(cases):
        f, gr = oracle.regression_objective(V, values, yy, ii, K, (6.0, 0.4), 1e-5, ap)
        np.testing.assert_allclose(f, g["obj"][q], rtol=1e-9)
        np.testing.assert_allclose(gr, g["grad"][q], rtol=1e-7, atol=1e-9)
    pars, obj_t = oracle.train_regression(V, values, Y[:m], idx, K, 1e-5, "posterior")
    np.testing.assert_allclose(pars, g["pars"], rtol=1e-6)
    np.testing.assert_allclose(obj_t, g["obj_t"], rtol=1e-8)
    se = oracle.fit_se_regression(X[:m], Y[:m], X[m:], meta["s"], meta["r"], K, init, g["a2s"], pars=(6.0, 0.4), iter_max=30)
    assert se["a2"] == float(g["se_a2"])
    np.testing.assert_allclose(se["test"][:200], g["se_test"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(se["cov"][:200], g["se_cov"], rtol=1e-8, atol=1e-10)
    ny = oracle.fit_nystrom_regression(X[:m], Y[:m], X[m:], meta["s"], K, init, g["a2s"], pars=(6.0, 0.4), iter_max=30)
    assert ny["a2"] == float(g["ny_a2"])
    np.testing.assert_allclose(ny["test"][:200], g["ny_test"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(ny["cov"][:200], g["ny_cov"], rtol=1e-7, atol=1e-9)


def test_laplace_posterior_oracle_against_direct_mode(oracle):
    """oracle.posterior_distribution_classification (src/Utils.cpp:252-299): its Newton mode satisfies the
    stationarity condition f = C11 (Y - pi(f)) of the Laplace approximation, and the predictive variance equals the
    textbook C22 - C21 (C11 + W^-1)^-1 C21^T."""
    rng = np.random.default_rng(4)
    m, q = 30, 12
    A = rng.standard_normal((m + q, 8))
    Kf = A @ A.T * 0.3 + 1e-3 * np.eye(m + q)
    C11, C21, C22 = Kf[:m, :m], Kf[m:, :m], np.diag(Kf)[m:]
    Y = (rng.uniform(size=m) > 0.5).astype(np.float64)
    mean, cov = oracle.posterior_distribution_classification(C11, C21, C22, Y, tol=1e-12, max_iter=200)
    # recover the mode from the mean relation on the training block itself
    f = np.zeros(m)
    for _ in range(500):
        pi = 1 / (1 + np.exp(-f))
        W = pi * (1 - pi)
        f = np.linalg.solve(np.eye(m) + C11 * W[None, :], C11 @ (W * f + Y - pi))
    pi = 1 / (1 + np.exp(-f))
    np.testing.assert_allclose(mean, C21 @ (Y - pi), rtol=1e-8, atol=1e-10)
    W = pi * (1 - pi)
    ref = C22 - np.einsum("ij,jk,ik->i", C21, np.linalg.inv(C11 + np.diag(1 / W)), C21)
    np.testing.assert_allclose(cov, ref, rtol=1e-8, atol=1e-10)


# ------------------------------------------------------------------ independent third-party cross-checks
def test_kmeans_against_sklearn_lloyd(oracle):
    """The Lloyd contract against scikit-learn's Lloyd (an implementation by other authors) from the same start rows:
    same partition, same centres, same sizes.  (stats::kmeans itself is Hartigan-Wong with an R-RNG start: unpinned.)"""
    sk = pytest.importorskip("sklearn.cluster")
    X, _ = swiss(4000, 7)
    s = 30
    init = np.sort(np.random.default_rng(5).choice(len(X), s, replace=False)).astype(np.int32)
    U, assign, iters = oracle.kmeans_lloyd(X, s, init, 300)
    assert iters < 300  # converged: no assignment changed in the last pass
    km = sk.KMeans(n_clusters=s, init=np.ascontiguousarray(X[init]), n_init=1, algorithm="lloyd", max_iter=300, tol=0.0)
    km.fit(np.ascontiguousarray(X))
    assert np.array_equal(km.labels_, assign)
    np.testing.assert_allclose(km.cluster_centers_, U[:, :3], rtol=1e-12, atol=1e-12)
    assert np.array_equal(np.bincount(km.labels_, minlength=s), U[:, 3].astype(np.int64))


def test_spectrum_against_scipy_svds(oracle):
    """spectrum_from_Z_cpp through the Gram route against ARPACK's svds on A = Z diag(w) (the method family of
    RSpectra::svds, src/TruncatedSVD.cpp:23-28): singular values to 1e-8, left singular subspace through H."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    X, _ = spiral(3000, 11)
    s, r, K = 120, 3, 20
    init = np.sort(np.random.default_rng(2).choice(len(X), s, replace=False)).astype(np.int32)
    U, _, _ = oracle.kmeans_lloyd(X, s, init, 50)
    Zj, Zx = oracle.cross_similarity_lae(X, U, r, "cluster-normalized")
    values, V, I = oracle.spectrum_from_Z(Zj, Zx, s, K, root=True, want_internals=True)
    n = len(X)
    A = sp.csr_matrix((Zx.reshape(-1) * I["w"][Zj.reshape(-1)], Zj.reshape(-1), np.arange(0, n * r + 1, r)), shape=(n, s))
    u, sv, _ = spl.svds(A, k=K, tol=1e-12, random_state=0)
    order = np.argsort(-sv)
    sv, u = sv[order], u[:, order]
    np.testing.assert_allclose(values, sv, rtol=1e-8)
    # heat kernel on a few rows: invariant under the sign / rotation freedom of the singular vectors
    idx = np.arange(0, n, 97)
    lam = np.exp(-5.0 * (1.0 - sv))
    H_arpack = n * (u[idx] * lam) @ u[idx].T
    H_oracle = (V[idx] * np.exp(-5.0 * (1.0 - values))) @ V[idx].T
    # the K cut must not split a cluster for the comparison to be meaningful
    gap = values[K - 1] - np.sqrt(max(np.linalg.eigvalsh(I["G"])[::-1][K], 0.0))
    if gap > 1e-4:
        assert np.abs(H_arpack - H_oracle).max() <= 1e-7 * np.abs(H_oracle).max()


# ---- independent cross-checks (other people's implementations of the same mathematics) ------------------------------
def test_knn_against_sklearn_brute_force(oracle):
    """KNN_cpp restated vs scikit-learn's brute-force neighbours on tie-free data: same anchors in the same order, squared
    distances to 1e-10 (sklearn expands the norm the same way but sums in BLAS order)."""
    from sklearn.neighbors import NearestNeighbors

    rng = np.random.default_rng(21)
    for n, d, s, r in [(400, 3, 60, 3), (300, 16, 50, 5), (200, 37, 33, 1)]:
        X = np.asfortranarray(rng.standard_normal((n, d)))
        U = np.asfortranarray(rng.standard_normal((s, d)))
        ind, dist = oracle.knn(X, U, r, want_dist=True)
        dd, ii = NearestNeighbors(n_neighbors=r, algorithm="brute").fit(U).kneighbors(X)
        assert np.array_equal(ind, ii)
        np.testing.assert_allclose(dist, dd ** 2, rtol=1e-9, atol=1e-10)


def test_lae_against_scipy_constrained_least_squares(oracle):
    """local_anchor_embedding_cpp restated vs an independent solver of the same problem (min 1/2 |x - z U|^2 over the
    simplex, scipy SLSQP): the accelerated projected gradient stops at |z - z_prev|^2 < 1e-5 (the reference's rule, not a
    tight one), so its objective is within 2e-3 of the constrained optimum and its weights within 5e-2; never below it."""
    import scipy.optimize as so

    rng = np.random.default_rng(22)
    for r, d in [(3, 2), (3, 3), (5, 16), (4, 8)]:
        for _ in range(6):
            U = rng.standard_normal((r, d))
            x = rng.dirichlet(np.ones(r)) @ U + 0.3 * rng.standard_normal(d)
            z = oracle.lae_point(x, np.asfortranarray(U))
            assert abs(z.sum() - 1.0) <= 1e-12 and z.min() >= 0.0
            f = lambda w: 0.5 * np.sum((x - w @ U) ** 2)  # noqa: E731
            ref = so.minimize(f, np.full(r, 1.0 / r), method="SLSQP", bounds=[(0, 1)] * r,
                              constraints=[dict(type="eq", fun=lambda w: w.sum() - 1.0)],
                              options=dict(ftol=1e-14, maxiter=500))
            assert ref.success
            assert ref.fun - 1e-9 <= f(z) <= ref.fun + 2e-3 * max(1.0, x @ x)
            assert np.abs(z - ref.x).max() <= 5e-2


def test_minibatch_contract_is_a_sane_minibatch_kmeans(oracle):
    """The mini-batch contract (ClusterR un-vendored) against scikit-learn's MiniBatchKMeans on separated blobs: another
    implementation of Sculley's algorithm with its own sampling and learning-rate details, so only the quality is
    comparable — the within-cluster sum of squares of the contract's centroids is within 25 % of scikit-learn's from the
    same start, and far below that of the start itself; sizes add up to n."""
    from sklearn.cluster import MiniBatchKMeans

    rng = np.random.default_rng(23)
    n, d, s = 6000, 2, 12
    centres = rng.uniform(-10, 10, (s, d))
    X = np.asfortranarray(centres[rng.integers(0, s, n)] + 0.4 * rng.standard_normal((n, d)))
    init = np.sort(rng.choice(n, s, replace=False)).astype(np.int32)
    U, iters = oracle.minibatch_kmeans(X, s, init, seed=4)

    def wss(C):
        return ((X[:, None, :] - C[None, :, :]) ** 2).sum(-1).min(1).sum()

    sk = MiniBatchKMeans(n_clusters=s, init=X[init], n_init=1, batch_size=10 * s, max_iter=100, random_state=0).fit(X)
    assert U[:, d].sum() == n and 1 <= iters <= 100
    assert wss(U[:, :d]) <= 1.25 * wss(sk.cluster_centers_)
    assert wss(U[:, :d]) <= 0.6 * wss(X[init])


def test_regression_training_reaches_the_lbfgsb_optimum(oracle):
    """The MMA restatement (NLopt un-vendored) against scipy's L-BFGS-B on the same bounded objective (marginal: smooth;
    the reference's clipping of the noise gradient is inactive near the optimum): same minimum to 1e-6 relative."""
    import scipy.optimize as so

    rng = np.random.default_rng(24)
    n, m, K = 300, 80, 20
    V = np.linalg.qr(rng.standard_normal((n, K)))[0] * np.sqrt(n)
    values = np.sort(rng.uniform(0.3, 1.0, K))[::-1]
    values[0] = 1.0
    idx = np.arange(m, dtype=np.int32)
    Y = V[:m, 1] + 0.7 * V[:m, 3] + 0.2 * rng.standard_normal(m)
    x, obj = oracle.train_regression(V, values, Y, idx, K, 1e-5, "marginal")
    f = lambda z: oracle.regression_objective(V, values, Y, idx, K, z, 1e-5, "marginal")  # noqa: E731
    ref = so.minimize(lambda z: f(z)[0], [10.0, 1.0], jac=lambda z: f(z)[1], method="L-BFGS-B",
                      bounds=[(1e-3, None), (1e-4, None)], options=dict(ftol=1e-15, gtol=1e-10))
    assert abs(-obj - ref.fun) <= 1e-6 * max(1.0, abs(ref.fun))
    np.testing.assert_allclose(x, ref.x, rtol=2e-2, atol=1e-3)


def test_golden_fixture_late_round2_rows(oracle):
    """The callers added late in round 2 (mini-batch subsample, noise = "different", SE / Nystrom logit grids at a fixed
    diffusion time) pinned by the committed fixture tests/golden/oracle_round2b.npz (same generator script)."""
    g0 = np.load(os.path.join(GOLD, "oracle_small.npz"))
    gt = np.load(os.path.join(GOLD, "oracle_train.npz"))
    g = np.load(os.path.join(GOLD, "oracle_round2b.npz"))
    meta = json.loads(str(g0["meta"]))
    X, Y = spiral(meta["n"], meta["seed"])
    init, m, K, s, r = g0["init"], int(gt["m"]), meta["K"], meta["s"], meta["r"]
    Umb, it_mb = oracle.minibatch_kmeans(X, s, init, max_iters=40, seed=9)
    assert it_mb == int(g["it_mb"]) and np.array_equal(Umb, g["Umb"])
    U, _, _ = oracle.kmeans_lloyd(X, s, init)
    Zj, Zx = oracle.cross_similarity_lae(X, U, r, "cluster-normalized")
    values, V = oracle.spectrum_from_Z(Zj, Zx, s, K, True)
    idx = np.arange(m, dtype=np.int32)
    f, gr = oracle.regression_objective_diff(V, values, Y[:m], idx, K, g["xd"], 1e-5, "posterior")
    np.testing.assert_allclose(f, g["obj_d"], rtol=1e-9)
    np.testing.assert_allclose(gr, g["grad_d"], rtol=1e-7, atol=1e-9)
    pred = oracle.predict_regression_diff(V, values, Y[:m], idx, np.arange(m, meta["n"], dtype=np.int32), K, g["xd"], 1e-5)
    np.testing.assert_allclose(pred[:200], g["pred_d"], rtol=1e-8, atol=1e-9)
    lab = (Y > np.median(Y)).astype(np.float64)
    sl = oracle.fit_se_logit(X[:m], lab[:m], X[m:], s, r, K, init, gt["a2s"], iter_max=30, t=6.0)
    assert sl["a2"] == float(g["sl_a2"])
    np.testing.assert_allclose(sl["obj"], g["sl_obj"], rtol=1e-9)
    np.testing.assert_allclose(sl["mean"][:200], g["sl_mean"], rtol=1e-7, atol=1e-8)
    nl = oracle.fit_nystrom_logit(X[:m], lab[:m], X[m:], s, K, init, gt["a2s"], iter_max=30, t=6.0)
    assert nl["a2"] == float(g["nl_a2"])
    np.testing.assert_allclose(nl["obj"], g["nl_obj"], rtol=1e-8)
    np.testing.assert_allclose(nl["mean"][:200], g["nl_mean"], rtol=1e-6, atol=1e-7)


def test_laplace_objective_against_sklearn_gpc(oracle):
    """marginal_log_likelihood_logit_la_cpp and posterior_distribution_classification restated vs scikit-learn's
    GaussianProcessClassifier (its own Newton iteration and Laplace formula, GPML algorithms 3.1 / 3.2, logistic link) on
    the same precomputed covariance: approximate log marginal likelihood to 1e-6 (the reference adds 1e-9 inside the log
    of the Cholesky diagonal and stops on |df|_1 < 1e-5), identical predicted labels on the held-out rows."""
    from sklearn.gaussian_process import GaussianProcessClassifier
    from sklearn.gaussian_process.kernels import Kernel

    class Precomputed(Kernel):
        """k(i, j) = C[i, j] on integer 'inputs' (no hyper-parameters)."""

        def __init__(self, C):
            self.C = C

        def __call__(self, X, Y=None, eval_gradient=False):
            i = np.asarray(X)[:, 0].astype(int)
            j = i if Y is None else np.asarray(Y)[:, 0].astype(int)
            Kij = self.C[np.ix_(i, j)]
            return (Kij, np.empty((len(i), len(i), 0))) if eval_gradient else Kij

        def diag(self, X):
            return np.diag(self.C)[np.asarray(X)[:, 0].astype(int)].copy()

        def is_stationary(self):
            return False

    rng = np.random.default_rng(0)
    n, K, m = 400, 25, 60
    V = np.linalg.qr(rng.standard_normal((n, K)))[0] * np.sqrt(n)
    values = np.sort(rng.uniform(0.2, 1.0, K))[::-1]
    values[0] = 1.0
    Y = (V[:m, 1] * 1.5 + V[:m, 2] + 0.3 * rng.standard_normal(m) > 0).astype(np.float64)
    every = np.arange(n, dtype=np.int32)
    for t in (2.0, 9.0):
        C = oracle.hk_from_spectrum(V, values, K, t, every, every)
        C[np.diag_indices(n)] += 1e-3
        gpc = GaussianProcessClassifier(kernel=Precomputed(C), optimizer=None, max_iter_predict=200)
        gpc.fit(np.arange(m, dtype=np.float64)[:, None], Y)
        ours = oracle.laplace_mll(C[:m, :m].copy(), Y)
        assert abs(ours - gpc.log_marginal_likelihood_value_) <= 1e-6 * abs(ours)
        mean, cov = oracle.posterior_distribution_classification(C[:m, :m].copy(), C[m:, :m].copy(), np.diag(C)[m:].copy(), Y)
        labels = gpc.predict(np.arange(m, n, dtype=np.float64)[:, None])
        assert np.array_equal(mean > 0, labels > 0.5) and np.all(cov > 0)


@pytest.mark.parametrize("m", [80, 12])
def test_regression_tail_and_objective_against_sklearn_gpr(oracle, m):
    """predict_regression_cpp / posterior_covariance_regression / negative_marginal_likelihood_regression_cpp restated vs
    scikit-learn's GaussianProcessRegressor on the same precomputed heat kernel (alpha = noise + sigma), in both branches
    (m > K: Woodbury; m <= K: direct): predictive mean 1e-10, predictive variance (+ noise + sigma, as the reference
    reports it) 1e-10, log marginal likelihood = -(objective) - m/2 log 2 pi to 1e-7 (the reference adds 1e-9 inside the
    log of the Cholesky diagonal)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Kernel

    class Precomputed(Kernel):
        def __init__(self, C):
            self.C = C

        def __call__(self, X, Y=None, eval_gradient=False):
            i = np.asarray(X)[:, 0].astype(int)
            j = i if Y is None else np.asarray(Y)[:, 0].astype(int)
            Kij = self.C[np.ix_(i, j)]
            return (Kij, np.empty((len(i), len(i), 0))) if eval_gradient else Kij

        def diag(self, X):
            return np.diag(self.C)[np.asarray(X)[:, 0].astype(int)].copy()

        def is_stationary(self):
            return False

    rng = np.random.default_rng(1)
    n, K = 300, 20
    V = np.linalg.qr(rng.standard_normal((n, K)))[0] * np.sqrt(n)
    values = np.sort(rng.uniform(0.3, 1.0, K))[::-1]
    values[0] = 1.0
    every = np.arange(n, dtype=np.int32)
    idx, idx1 = every[:m], every[m:]
    Y = V[:m, 1] + 0.7 * V[:m, 3] + 0.2 * rng.standard_normal(m)
    t, noise, sigma = 5.0, 0.3, 1e-5
    C = oracle.hk_from_spectrum(V, values, K, t, every, every)
    gpr = GaussianProcessRegressor(kernel=Precomputed(C), alpha=noise + sigma, optimizer=None)
    gpr.fit(np.arange(m, dtype=np.float64)[:, None], Y)
    mu, sd = gpr.predict(np.arange(m, n, dtype=np.float64)[:, None], return_std=True)
    pred = oracle.predict_regression(V, values, Y, idx, idx1, K, (t, noise), sigma)
    cov = oracle.posterior_covariance_regression(V, values, idx, idx1, K, (t, noise), sigma)
    np.testing.assert_allclose(pred, mu, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(cov, sd ** 2 + noise + sigma, rtol=1e-10, atol=1e-11)
    f, _ = oracle.regression_objective(V, values, Y, idx, K, (t, noise), sigma, "marginal")
    assert abs((-f - 0.5 * m * np.log(2 * np.pi)) - gpr.log_marginal_likelihood_value_) <= 1e-7


@pytest.mark.parametrize("m", [80, 12])
def test_diff_noise_against_sklearn_heteroscedastic_gpr(oracle, m):
    """noise = "different" restated (src/train.cpp:438-556, src/Predict.cpp:76-113) vs scikit-learn's
    GaussianProcessRegressor with one alpha per training row on the same precomputed heat kernel, in both branches:
    predictive mean 1e-10; log marginal likelihood = -(objective) - m/2 log 2 pi to 1e-6 (the reference adds 1e-9 inside
    its logarithms) — the value the reference's source calls "wrong" is the exact one up to those regularisers."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Kernel

    class Precomputed(Kernel):
        def __init__(self, C):
            self.C = C

        def __call__(self, X, Y=None, eval_gradient=False):
            i = np.asarray(X)[:, 0].astype(int)
            j = i if Y is None else np.asarray(Y)[:, 0].astype(int)
            Kij = self.C[np.ix_(i, j)]
            return (Kij, np.empty((len(i), len(i), 0))) if eval_gradient else Kij

        def diag(self, X):
            return np.diag(self.C)[np.asarray(X)[:, 0].astype(int)].copy()

        def is_stationary(self):
            return False

    rng = np.random.default_rng(2)
    n, K = 300, 20
    V = np.linalg.qr(rng.standard_normal((n, K)))[0] * np.sqrt(n)
    values = np.sort(rng.uniform(0.3, 1.0, K))[::-1]
    values[0] = 1.0
    every = np.arange(n, dtype=np.int32)
    idx, idx1 = every[:m], every[m:]
    Y = V[:m, 1] + 0.7 * V[:m, 3] + 0.2 * rng.standard_normal(m)
    sigma = 1e-5
    x = np.concatenate([[5.0], rng.uniform(0.05, 1.0, m)])
    C = oracle.hk_from_spectrum(V, values, K, x[0], every, every)
    gpr = GaussianProcessRegressor(kernel=Precomputed(C), alpha=x[1:] + sigma, optimizer=None)
    gpr.fit(np.arange(m, dtype=np.float64)[:, None], Y)
    mu = gpr.predict(np.arange(m, n, dtype=np.float64)[:, None])
    np.testing.assert_allclose(oracle.predict_regression_diff(V, values, Y, idx, idx1, K, x, sigma), mu, rtol=1e-10,
                               atol=1e-11)
    f, _ = oracle.regression_objective_diff(V, values, Y, idx, K, x, sigma, "marginal")
    assert abs((-f - 0.5 * m * np.log(2 * np.pi)) - gpr.log_marginal_likelihood_value_) <= 1e-6
