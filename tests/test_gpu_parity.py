"""GPU parity tests proper: every stage through the C ABI (ctypes mirror of the Rcpp exports) against the
oracle on the same seeded inputs.  Bit-exact: k-means assignments/centres, KNN indices+distances, LAE weights,
Z pattern and values, Gram-side sums.  1e-8 relative (fp64): eigenvalues, heat kernel, predictions."""
import os

import numpy as np
import pytest

from conftest import spiral, swiss

pytestmark = pytest.mark.gpu

NT = 8  # oracle threads


def _init(n, s, seed=0):
    return np.sort(np.random.default_rng(seed).choice(n, s, replace=False)).astype(np.int32)


def _csr_parts(Z):
    n = Z.shape[0]
    r = Z.indptr[1] - Z.indptr[0]
    return Z.indices.reshape(n, r), Z.data.reshape(n, r)


# ------------------------------------------------------------------------------------------- k-means
@pytest.mark.parametrize("n,d,s", [(4000, 2, 500), (20000, 3, 200), (1500, 1, 20), (3000, 4, 64),
                                    (3000, 5, 70), (2500, 16, 100), (700, 37, 33),
                                    # tile edges of the fused pruned pass: one point into a new warp / CTA, tiny inputs
                                    (257, 3, 7), (2049, 2, 30), (1025, 4, 12), (33, 1, 5), (2, 3, 2),
                                    # s >= 512: the first pass goes through pivot groups
                                    (40000, 3, 600), (30000, 2, 1024), (9000, 4, 513), (5000, 1, 700)])
def test_kmeans_bitexact(flgp, oracle, n, d, s):
    rng = np.random.default_rng(n + d)
    if d == 2:
        X, _ = spiral(n, 1)
    elif d == 3:
        X, _ = swiss(n, 1)
    else:
        X = np.asfortranarray(rng.standard_normal((n, d)) + 3 * rng.integers(0, 4, (n, 1)))
    init = _init(n, s, 2)
    U, assign, iters = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, return_info=True)
    Uo, ao, io = oracle.kmeans_lloyd(X, s, init, 100, NT)
    assert iters == io
    assert np.array_equal(assign, ao)
    assert np.array_equal(U, Uo)  # centres AND sizes, bit for bit
    assert U[:, d].sum() == n


def test_kmeans_pivot_first_pass_with_duplicates_and_lattice(flgp, oracle):
    """The pivot-pruned first pass on adversarial input: lattice points (exact score ties between centres) and
    duplicated start rows (identical centres, one of them left empty)."""
    rng = np.random.default_rng(12)
    n, s = 20000, 640
    X = np.asfortranarray(rng.integers(-20, 21, (n, 2)).astype(np.float64))
    init = _init(n, s, 4)
    X[init[7]] = X[init[3]]
    X[init[300]] = X[init[3]]
    for iter_max in (1, 2, 9):
        U, assign, iters = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, iter_max=iter_max, return_info=True)
        Uo, ao, io = oracle.kmeans_lloyd(X, s, init, iter_max, NT)
        assert iters == io and np.array_equal(assign, ao) and np.array_equal(U, Uo)


def test_kmeans_iter_cap_and_default_init(flgp, oracle):
    X, _ = swiss(6000, 3)
    init = flgp.default_init(6000, 300, seed=9)
    U, assign, iters = flgp.subsample_cpp(X, 300, "kmeans", seed=9, iter_max=5, return_info=True)
    Uo, ao, io = oracle.kmeans_lloyd(X, 300, init, 5, NT)
    assert iters == io == 5 and np.array_equal(U, Uo) and np.array_equal(assign, ao)


@pytest.mark.parametrize("n,d,s,iters,seed", [(3000, 2, 30, 100, 1), (2000, 3, 25, 40, 2), (900, 16, 12, 30, 3),
                                              (700, 37, 9, 25, 4), (150, 5, 20, 15, 5), (5000, 1, 40, 100, 6),
                                              (40000, 3, 600, 100, 7), (2500, 784, 64, 6, 8)])
def test_subsample_minibatchkmeans_bitexact(flgp, oracle, n, d, s, iters, seed):
    """subsample_cpp "minibatchkmeans" (src/Utils.cpp:49-62; ClusterR un-vendored: the mini-batch contract of
    csrc/minibatch.cu and the oracle): centroids, number of batches and the 1-NN sizes column bit for bit — lattice
    data with exact score ties (d = 2), tight clusters that end by the early stop (d = 3), batch = all rows (n < 10 s),
    d below / at / above the register chunk, the tensor-core 1-NN (d = 784)."""
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, d)) * (0.01 if d == 3 else 1.0) + 3 * rng.integers(0, 4, (n, 1)))
    if d == 2:
        X = np.asfortranarray(np.round(X))
    init = _init(n, s, seed)
    U, _, it = flgp.subsample_cpp(X, s, "minibatchkmeans", init_idx=init, seed=seed, iter_max=iters, return_info=True)
    Uo, it_o = oracle.minibatch_kmeans(X, s, init, max_iters=iters, seed=seed, nthreads=NT)
    assert it == it_o
    assert np.array_equal(U[:, :d], Uo[:, :d])
    assert np.array_equal(U[:, d], Uo[:, d]) and U[:, d].sum() == n


def test_pipeline_with_minibatch_anchors(flgp, oracle):
    """heat_kernel_spectrum_cpp with models$subsample = "minibatchkmeans": anchors from the mini-batch contract, the
    rest of the path unchanged (Z bit-exact, eigenvalues 1e-8 against the oracle fed the same anchors)."""
    X, _ = swiss(6000, 5)
    m, s, r, K = 300, 150, 3, 30
    init = _init(len(X), s, 3)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, models={"subsample": "minibatchkmeans"}, init_idx=init,
                                       seed=3)
    Uo, _ = oracle.minibatch_kmeans(X, s, init, seed=3, nthreads=NT)
    assert np.array_equal(ep.anchors(), Uo)
    Zj, Zx = oracle.cross_similarity_lae(X, Uo, r, "cluster-normalized", nthreads=NT)
    Z = ep.Z()
    assert np.array_equal(Z.indices.reshape(len(X), r), Zj) and np.array_equal(Z.data.reshape(len(X), r), Zx)
    vo, _ = oracle.spectrum_from_Z(Zj, Zx, s, K, True, nthreads=NT)
    np.testing.assert_allclose(ep.values, vo, rtol=1e-8, atol=1e-10)


def test_subsample_random_and_errors(flgp):
    X, _ = spiral(500, 4)
    init = _init(500, 30, 1)
    U = flgp.subsample_cpp(X, 30, "random", init_idx=init)
    assert U.shape == (30, 2) and np.array_equal(U, X[init])
    with pytest.raises(flgp.FlgpError, match="not supported"):
        flgp.subsample_cpp(X, 30, "kmedoids")
    with pytest.raises(flgp.FlgpError, match="nstart"):
        flgp.subsample_cpp(X, 30, "minibatchkmeans", nstart=2)
    with pytest.raises(flgp.FlgpError):
        flgp.subsample_cpp(X, 501, "kmeans")


# ------------------------------------------------------------------------------------------- KNN
@pytest.mark.parametrize("n,d,s,r", [(4000, 2, 500, 3), (5000, 3, 333, 3), (3000, 3, 64, 5), (2000, 1, 50, 1),
                                      (2000, 4, 100, 8), (1000, 3, 40, 32), (1500, 8, 130, 5), (900, 37, 65, 4),
                                      (600, 784, 100, 5)])
def test_knn_bitexact(flgp, oracle, n, d, s, r):
    rng = np.random.default_rng(n + s)
    X = np.asfortranarray(rng.standard_normal((n, d)))
    U = np.asfortranarray(X[_init(n, s, 3)] + 0.01 * rng.standard_normal((s, d)))
    ind, dist = flgp.knn_distances(X, U, r)
    io, do = oracle.knn(X, U, r, want_dist=True, nthreads=NT)
    assert np.array_equal(ind, io)
    assert np.array_equal(dist, do)
    assert np.all(np.diff(dist, axis=1) >= 0)
    res = flgp.KNN_cpp(X, U, r, "Euclidean", True)
    assert np.array_equal(res["ind_knn"], io)
    Zj, Zx = _csr_parts(res["distances_sp"])
    zj, zx = oracle.knn_csr(io, do)
    assert np.array_equal(Zj, zj) and np.array_equal(Zx, zx)


@pytest.mark.parametrize("d", [2, 3, 6])
def test_knn_exact_ties_follow_std_partial_sort(flgp, oracle, d):
    """Lattice points and duplicated anchors: the selected set and its order must match libstdc++."""
    rng = np.random.default_rng(d)
    X = np.asfortranarray(rng.integers(-3, 4, (3000, d)).astype(np.float64))
    U = np.asfortranarray(rng.integers(-3, 4, (90, d)).astype(np.float64))
    U[40:50] = U[10:20]  # exact duplicates
    for r in (1, 3, 4, 7):
        ind = flgp.KNN_cpp(X, U, r)["ind_knn"]
        assert np.array_equal(ind, oracle.knn(X, U, r, nthreads=NT))


@pytest.mark.parametrize("n,d,s,r,kernel", [(6000, 2, 300, 3, "lae"), (5000, 3, 200, 5, "lae"), (4000, 1, 40, 2, "se"),
                                             (5000, 4, 150, 4, "se"), (3000, 3, 30, 1, "lae"), (2500, 2, 64, 6, "se")])
def test_pipeline_knn_pruned_bitexact(flgp, oracle, n, d, s, r, kernel):
    """Inside heat_kernel_spectrum the KNN runs on the cluster-sorted rows with exact candidate pruning
    (r <= 5; r = 6 takes the full scan): the neighbour sets, their order (through the LAE weights) and the SE
    distances must still be the oracle's bit for bit."""
    rng = np.random.default_rng(n + 7 * s + r)
    X = np.asfortranarray(rng.standard_normal((n, d)) * rng.uniform(0.5, 3.0, d))
    init = _init(n, s, 11)
    K = min(s, 10)
    mo = dict(kernel=kernel, gl="rw")
    ep = flgp.heat_kernel_spectrum_cpp(X[:100], X[100:], s, r, K, models=mo, init_idx=init, iter_max=7)
    _, _, I = oracle.heat_kernel_spectrum(X[:100], X[100:], s, r, K, init, kernel=kernel, gl="rw", nthreads=NT,
                                          want_internals=True, iter_max=7)
    assert np.array_equal(ep.anchors(), I["U"])
    Zj, Zx = _csr_parts(ep.Z())
    assert np.array_equal(Zj, I["Zj"])
    if kernel == "lae":
        assert np.array_equal(Zx, I["Zx"])
    else:  # exp() of the device and of glibc differ in the last bit
        np.testing.assert_allclose(Zx, I["Zx"], rtol=1e-12, atol=1e-300)


def test_pipeline_knn_pruned_duplicate_anchors_fall_back(flgp, oracle):
    """Identical start rows tie for ever: the lowest index takes the points and moves, the others stay empty and
    keep their (identical) centre - one is peeled off per pass.  Six copies and four passes leave a duplicated pair
    of anchors: every point near it sees an exact tie inside its r + 1 smallest distances and must take the
    libstdc++ emulation path."""
    rng = np.random.default_rng(5)
    n, d, s, r = 4000, 2, 60, 3
    X = np.asfortranarray(rng.standard_normal((n, d)))
    init = _init(n, s, 2)
    for t in (5, 20, 41):  # six identical rows among the start rows, three times
        for q in range(1, 6):
            X[init[t + q]] = X[init[t]]
    ep = flgp.heat_kernel_spectrum_cpp(X[:50], X[50:], s, r, 8, models=dict(gl="rw"), init_idx=init, iter_max=4)
    _, _, I = oracle.heat_kernel_spectrum(X[:50], X[50:], s, r, 8, init, gl="rw", nthreads=NT, want_internals=True,
                                          iter_max=4)
    U = ep.anchors()
    assert np.array_equal(U, I["U"])
    assert len(np.unique(U[:, :d], axis=0)) < s  # the duplicated anchors really are there
    Zj, Zx = _csr_parts(ep.Z())
    assert np.array_equal(Zj, I["Zj"]) and np.array_equal(Zx, I["Zx"])
    # rows that hold a duplicated pair exist (so the tie path was taken)
    dup = [j for j in range(s) if np.sum(np.all(U[:, :d] == U[j, :d], axis=1)) > 1]
    assert np.isin(Zj, dup).any()


def test_knn_errors(flgp):
    X, _ = spiral(100, 1)
    with pytest.raises(flgp.FlgpError):
        flgp.KNN_cpp(X, X[:3], 4)
    with pytest.raises(flgp.FlgpError, match="not supported"):
        flgp.KNN_cpp(X, X[:30], 3, distance="cosine")


# ------------------------------------------------------------------------------------------- LAE
def test_v_to_z_and_single_point(flgp, oracle):
    rng = np.random.default_rng(0)
    for r in (1, 2, 3, 7, 40):
        v = rng.standard_normal(r) * 3
        assert np.array_equal(flgp.v_to_z_cpp(v), oracle.simplex_project(v))
    for r, d in ((3, 3), (5, 2), (4, 9), (16, 3)):
        U = rng.standard_normal((r, d))
        x = U.mean(0) + 0.2 * rng.standard_normal(d)
        assert np.array_equal(flgp.local_anchor_embedding_cpp(x, U), oracle.lae_point(x, U))


@pytest.mark.parametrize("n,d,s,r", [(4000, 2, 500, 3), (6000, 3, 300, 3), (3000, 3, 100, 5), (2000, 2, 80, 2),
                                      (2000, 3, 80, 4), (1500, 6, 60, 3), (1200, 3, 50, 7), (400, 784, 50, 5),
                                      (900, 33, 40, 8), (700, 100, 30, 1), (500, 257, 64, 16), (800, 16, 64, 5),
                                      (600, 16, 40, 2), (600, 40, 40, 3), (600, 12, 40, 4), (500, 64, 40, 4), (500, 9, 40, 3)])
def test_lae_bitexact(flgp, oracle, n, d, s, r):
    rng = np.random.default_rng(n + r)
    if d == 2:
        X, _ = spiral(n, 5)
    elif d == 3:
        X, _ = swiss(n, 5)
    else:
        X = np.asfortranarray(rng.standard_normal((n, d)))
    U = np.asfortranarray(X[_init(n, s, 4)] + 0.05 * rng.standard_normal((s, d)))
    Z, stats = flgp.LAE_cpp(X, U, r, return_stats=True)
    Zj, Zx = _csr_parts(Z)
    zj, zx, _, st = oracle.lae(X, U, r, nthreads=NT, want_stats=True)
    assert np.array_equal(Zj, zj)
    assert np.array_equal(Zx, zx)  # bit-exact weights: same operation order, no FMA contraction
    assert tuple(stats) == tuple(st)
    assert Z.nnz == n * r  # explicit zeros are stored


# ------------------------------------------------------------------------------------------- Z, GL
@pytest.mark.parametrize("gl", ["rw", "normalized", "cluster-normalized"])
def test_graph_laplacian_and_cross_similarity(flgp, oracle, gl):
    X, _ = swiss(5000, 6)
    s, r = 120, 3
    U, _, _ = oracle.kmeans_lloyd(X, s, _init(5000, s, 5), 100, NT)
    zj, zx, _ = oracle.lae(X, U[:, :3], r, nthreads=NT)
    Z0 = flgp.LAE_cpp(X, U[:, :3], r)
    want = oracle.graph_laplacian(zj, zx, s, gl, U[:, 3], 1)
    got = flgp.graphLaplacian_cpp(Z0, gl, U[:, 3])
    assert np.array_equal(_csr_parts(got)[1], want)  # fixed-point column sums: bit-exact
    Zc = flgp.cross_similarity_lae_cpp(X, U, r, gl)
    assert np.array_equal(_csr_parts(Zc)[0], zj) and np.array_equal(_csr_parts(Zc)[1], want)
    # the reference's sequential fp64 sums agree far inside the 1e-8 contract
    seq = oracle.graph_laplacian(zj, zx, s, gl, U[:, 3], 0)
    np.testing.assert_allclose(_csr_parts(got)[1], seq, rtol=1e-12, atol=1e-300)
    # SE weights: exp() differs by an ulp between libm and CUDA
    Zs = flgp.cross_similarity_se_cpp(X, U, r, gl, 0.7)
    sj, sx = oracle.cross_similarity_se(X, U, r, gl, 0.7, 1, NT)
    assert np.array_equal(_csr_parts(Zs)[0], sj)
    np.testing.assert_allclose(_csr_parts(Zs)[1], sx, rtol=1e-12, atol=1e-300)


def test_gl_errors(flgp):
    X, _ = spiral(300, 1)
    Z = flgp.LAE_cpp(X, X[:20], 3)
    with pytest.raises(flgp.FlgpError, match="graph Laplacian is not supported"):
        flgp.graphLaplacian_cpp(Z, "bogus")
    with pytest.raises(flgp.FlgpError):
        flgp.cross_similarity_lae_cpp(X, X[:20], 3, "cluster-normalized")  # no size column


# ------------------------------------------------------------------------------------------- spectrum
def _check_spectrum(flgp, oracle, ep, values_o, V_o, K, t=2.0, tol=1e-8):
    values = ep.values
    np.testing.assert_allclose(values, values_o, rtol=tol, atol=tol * 1e-3)
    n = ep.n_local
    idx0 = np.arange(0, n, max(1, n // 97), dtype=np.int32)
    idx1 = np.arange(0, n, max(1, n // 61), dtype=np.int32)
    H = flgp.HK_from_spectrum_cpp(ep, K, t, idx0, idx1)
    Ho = oracle.hk_from_spectrum(V_o, values_o, K, t, idx0, idx1, NT)
    assert np.abs(H - Ho).max() <= tol * np.abs(Ho).max()
    return H


@pytest.mark.parametrize("root", [True, False])
def test_spectrum_from_Z(flgp, oracle, root):
    X, _ = swiss(8000, 7)
    s, r, K = 150, 3, 40
    U, _, _ = oracle.kmeans_lloyd(X, s, _init(8000, s, 6), 100, NT)
    Zj, Zx = oracle.cross_similarity_lae(X, U, r, "cluster-normalized", 1, NT)
    import scipy.sparse as sp

    Z = sp.csr_matrix((Zx.reshape(-1), Zj.reshape(-1), np.arange(0, 8000 * r + 1, r)), shape=(8000, s))
    ep = flgp.spectrum_from_Z_cpp(Z, K, root)
    vo, Vo, I = oracle.spectrum_from_Z(Zj, Zx, s, K, root, 1, NT, True)
    _check_spectrum(flgp, oracle, ep, vo, Vo, K)
    V = ep.vectors
    np.testing.assert_allclose(V.T @ V / 8000, np.eye(K), atol=1e-9)
    # eigenvectors up to sign where the eigenvalue is isolated
    gaps = np.minimum(np.abs(np.diff(I["lam"], prepend=np.inf)), np.abs(np.diff(I["lam"], append=-np.inf)))
    for k in np.nonzero(gaps > 1e-4)[0]:
        c = np.dot(V[:, k], Vo[:, k]) / 8000
        assert abs(abs(c) - 1) < 1e-8
    np.testing.assert_allclose(ep.rows([5, 17, 4000]), V[[5, 17, 4000]], rtol=0, atol=0)


def test_spectrum_full_K_equals_s(flgp, oracle):
    X, _ = spiral(3000, 8)
    s, r = 40, 3
    U, _, _ = oracle.kmeans_lloyd(X, s, _init(3000, s, 7), 100, NT)
    Zj, Zx = oracle.cross_similarity_lae(X, U, r, "rw", 1, NT)
    ep = flgp.spectrum_from_Z_cpp((Zj, Zx, s), -1, True)
    vo, Vo = oracle.spectrum_from_Z(Zj, Zx, s, -1, True, 1, NT)
    assert ep.K == s
    np.testing.assert_allclose(ep.values, vo, rtol=1e-8, atol=1e-9)


def test_heat_kernel_spectrum_rings_degenerate(flgp, oracle):
    """C1: six disconnected rings => eigenvalue 1 with multiplicity 6; compare eigenvalues and H, never vectors."""
    from flgp_b200.datasets import make

    X, Y, cfg = make("C1")
    m, s, r, K = cfg["m"], cfg["s"], cfg["r"], cfg["K"]
    init = _init(len(X), s, 8)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init)
    vo, Vo, I = oracle.heat_kernel_spectrum(X[:m], X[m:], s, r, K, init, nthreads=NT, want_internals=True)
    assert ep.kmeans_iters == I["iters"]
    assert np.array_equal(ep.anchors(), I["U"])
    Zj, Zx = _csr_parts(ep.Z())
    assert np.array_equal(Zj, I["Zj"]) and np.array_equal(Zx, I["Zx"])
    assert np.sum(np.abs(vo - 1) < 1e-9) >= 2  # the spectrum really is degenerate
    # keep the K cut away from a cluster: compare on the largest K' <= K with a clear gap below it
    lam = I["lam"]
    Kc = max(k for k in range(10, K) if lam[k - 1] - lam[k] > 1e-6)
    np.testing.assert_allclose(ep.values, vo, rtol=1e-8, atol=1e-10)
    idx0 = np.arange(0, len(X), 37, dtype=np.int32)
    idx1 = np.arange(m, dtype=np.int32)
    H = flgp.HK_from_spectrum_cpp(ep, Kc, 5.0, idx0, idx1)
    Ho = oracle.hk_from_spectrum(Vo, vo, Kc, 5.0, idx0, idx1, NT)
    assert np.abs(H - Ho).max() <= 1e-8 * np.abs(Ho).max()


def test_exports_eigenmap_and_covariance(flgp, oracle):
    X, _ = spiral(2500, 9)
    s, r, ndim = 60, 3, 8
    init = _init(2500, s, 9)
    res = flgp.lae_eigenmap(X, s, r, ndim, "kmeans", "normalized", init_idx=init)
    evo, Vo = oracle.lae_eigenmap(X, s, r, ndim, init, "normalized", nthreads=NT)
    np.testing.assert_allclose(res["eigenvalues"], evo, rtol=1e-7, atol=1e-9)
    P = res["eigenvectors"] @ res["eigenvectors"].T[:, :50]
    Po = Vo @ Vo.T[:, :50]
    assert np.abs(P - Po).max() <= 1e-7 * np.abs(Po).max()
    m = 80
    H = flgp.heat_kernel_covariance_rcpp(X[:m], X[m:], s, r, 3.0, 20, init_idx=init)
    vo, Vv = oracle.heat_kernel_spectrum(X[:m], X[m:], s, r, 20, init, nthreads=NT)
    Ho = oracle.hk_from_spectrum(Vv, vo, 20, 3.0, np.arange(2500, dtype=np.int32), np.arange(m, dtype=np.int32), NT)
    assert H.shape == (2500, m)
    assert np.abs(H - Ho).max() <= 1e-8 * np.abs(Ho).max()


# ------------------------------------------------------------------------------------------- GPR tail
@pytest.mark.parametrize("m,K", [(60, 80), (400, 50)])
def test_fit_lae_regression_fixed_pars(flgp, oracle, m, K):
    X, Y = spiral(4000, 10)
    s, r = 120, 3
    init = _init(4000, s, 10)
    pars, sigma = (4.0, 0.05), 1e-3
    res = flgp.fit_lae_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, sigma, pars=pars, init_idx=init)
    ref = oracle.fit_lae_regression_fixed(X[:m], Y[:m], X[m:], s, r, K, pars, init, sigma, nthreads=NT)
    sc = np.abs(ref["test"]).max()
    assert np.abs(res["Y_pred"]["train"] - ref["train"]).max() <= 1e-8 * sc
    assert np.abs(res["Y_pred"]["test"] - ref["test"]).max() <= 1e-8 * sc
    assert np.abs(res["posterior"]["cov"] - ref["cov"]).max() <= 1e-8 * np.abs(ref["cov"]).max()
    assert np.all(res["posterior"]["cov"] > 0)


@pytest.mark.parametrize("m,K", [(150, 40), (30, 40)])
def test_fit_lae_regression_noise_different(flgp, oracle, m, K):
    """fit_lae_regression_gp_rcpp(noise = "different") (src/train.cpp:438-556, src/Predict.cpp:76-113): one noise variance
    per training row.  At fixed parameters, mean of every row against the oracle's literal prediction on the library's
    eigenvectors (1e-8), variance = the reference's posterior_covariance_regression at pars[1], objective 1e-9; trained:
    the result is the objective at the returned parameters, better than at the start, within the bounds."""
    X, Y = spiral(3000, 10)
    n, s, r, sigma = len(X), 120, 3, 1e-3
    init = _init(n, s, 10)
    rng = np.random.default_rng(m)
    pars = np.concatenate([[4.0], rng.uniform(0.02, 0.5, m)])
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init)
    V, values = ep.vectors, ep.values
    idx0, idx1 = np.arange(m, dtype=np.int32), np.arange(m, n, dtype=np.int32)
    for approach in ("marginal", "posterior"):
        res = flgp.fit_lae_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, sigma, approach=approach, noise="different",
                                              pars=pars, init_idx=init)
        want_tr = oracle.predict_regression_diff(V, values, Y[:m], idx0, idx0, K, pars, sigma)
        want_te = oracle.predict_regression_diff(V, values, Y[:m], idx0, idx1, K, pars, sigma)
        sc = np.abs(want_te).max()
        assert np.abs(res["Y_pred"]["train"] - want_tr).max() <= 1e-8 * sc
        assert np.abs(res["Y_pred"]["test"] - want_te).max() <= 1e-8 * sc
        want_cov = oracle.posterior_covariance_regression(V, values, idx0, idx1, K, pars[:2], sigma)
        assert np.abs(res["posterior"]["cov"] - want_cov).max() <= 1e-8 * np.abs(want_cov).max()
        fo = oracle.regression_objective_diff(V, values, Y[:m], idx0, K, pars, sigma, approach)[0]
        assert abs(res["obj"] + fo) <= 1e-9 * max(1.0, abs(fo))
    res = flgp.fit_lae_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, sigma, noise="different", init_idx=init)
    x = np.array(res["pars"])
    assert x.size == m + 1 and x[0] >= 1e-3 and x[1:].min() >= 1e-4
    f_at = oracle.regression_objective_diff(V, values, Y[:m], idx0, K, x, sigma, "posterior")[0]
    f_0 = oracle.regression_objective_diff(V, values, Y[:m], idx0, K, np.concatenate([[10.0], np.ones(m)]), sigma,
                                           "posterior")[0]
    assert abs(res["obj"] + f_at) <= 1e-8 * max(1.0, abs(f_at)) and f_at < f_0
    want_te = oracle.predict_regression_diff(V, values, Y[:m], idx0, idx1, K, x, sigma)
    assert np.abs(res["Y_pred"]["test"] - want_te).max() <= 1e-7 * np.abs(want_te).max()
    with pytest.raises(flgp.FlgpError, match="m \\+ 1"):
        flgp.fit_lae_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, sigma, noise="different", pars=(4.0, 0.1),
                                        init_idx=init)


def test_errors_mirror_rcpp_stop(flgp):
    X, Y = spiral(300, 1)
    with pytest.raises(flgp.FlgpError, match="kernel type is not supported"):
        flgp.heat_kernel_spectrum_cpp(X[:50], X[50:], 20, 3, 5, models=dict(kernel="bogus"))
    with pytest.raises(flgp.FlgpError, match="subsample method is not supported"):
        flgp.heat_kernel_spectrum_cpp(X[:50], X[50:], 20, 3, 5, models=dict(subsample="bogus"))
    with pytest.raises(flgp.FlgpError, match="cluster-normalized"):
        flgp.heat_kernel_spectrum_cpp(X[:50], X[50:], 20, 3, 5, models=dict(subsample="random"))
    with pytest.raises(flgp.FlgpError):
        flgp.heat_kernel_spectrum_cpp(X[:50], X[50:], 20, 3, 21)  # K > s
    with pytest.raises(flgp.FlgpError, match="noise"):
        flgp.fit_lae_regression_gp_rcpp(X[:50], Y[:50], X[50:], 20, 3, 5, noise="bogus", pars=(1, 1))


# ------------------------------------------------------------------------------------------- full-size properties
def test_c4_scale_properties(flgp):
    """BASELINE config 4 shape at n = 2e6 (full s, r, K): size-independent properties through the C ABI."""
    from flgp_b200.datasets import make

    n, m = 2_000_000, 5000
    X, Y, cfg = make("C4", n=n)
    s, r, K = cfg["s"], cfg["r"], cfg["K"]
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, seed=1, iter_max=10)
    U = ep.anchors()
    assert U[:, 3].sum() == n and ep.kmeans_iters == 10
    Z = ep.Z()
    Zj, Zx = _csr_parts(Z)
    assert np.all(np.diff(Zj, axis=1) > 0) and Zj.min() >= 0 and Zj.max() < s
    np.testing.assert_allclose(Zx.sum(1), 1.0, rtol=1e-6)
    vals = ep.values
    assert abs(vals[0] - 1) < 1e-6 and np.all(np.diff(vals) <= 1e-12) and vals[-1] > 0
    idx = np.arange(0, n, n // 300, dtype=np.int32)
    H = flgp.HK_from_spectrum_cpp(ep, K, 10.0, idx, idx)
    np.testing.assert_allclose(H, H.T, rtol=1e-10, atol=1e-12)
    assert np.linalg.eigvalsh(H).min() > -1e-8
    V = ep.rows(idx)
    assert np.abs(V[:, 0]).std() < 1e-6 * np.abs(V[:, 0]).mean()  # top eigenvector of a row-stochastic Z is constant
    y, cov = flgp.regression_fixed(ep, Y[:m], m, K, (10.0, 0.01), 1e-5)
    assert np.sqrt(np.mean((y[m:] - Y[m:]) ** 2)) < 0.5 and np.all(cov[m:] > 0)
    # the pipeline's pruned KNN against the exported full-scan KNN on the same anchors (two different kernels)
    ind = flgp.KNN_cpp(X, np.asfortranarray(U[:, :3]), r)["ind_knn"]
    assert np.array_equal(np.sort(ind, axis=1), Zj)


# ------------------------------------------------------------------------------------------- dense eigensolver
def _sym(rng, s, kind):
    Q, _ = np.linalg.qr(rng.standard_normal((s, s)))
    if kind == "decay":          # graph-like: eigenvalues 1, then slowly decaying towards 0
        lam = 1.0 / (1.0 + 0.05 * np.arange(s)) ** 2
    elif kind == "clustered":    # exact multiplicities and tight clusters at the top
        lam = np.sort(np.r_[np.ones(5), np.full(4, 0.9), 0.9 - 1e-9 * np.arange(3), rng.uniform(0, 0.8, s - 12)])[::-1]
    elif kind == "bigcluster":   # 50-fold eigenvalue 1 and a 30-member cluster 1e-10 apart (CGS2 re-orthogonalisation)
        lam = np.sort(np.r_[np.ones(50), 0.7 - 1e-10 * np.arange(30), rng.uniform(0, 0.6, s - 80)])[::-1]
    elif kind == "indefinite":
        lam = np.sort(rng.uniform(-1, 1, s))[::-1]
    else:                         # graded over 12 orders of magnitude
        lam = 10.0 ** (-12.0 * np.arange(s) / s)
    A = (Q * lam) @ Q.T
    return np.asfortranarray((A + A.T) / 2), lam


@pytest.mark.parametrize("s,K,kind", [(300, 40, "decay"), (257, 257, "decay"), (500, 60, "clustered"), (64, 5, "graded"),
                                       (400, 30, "indefinite"), (2, 1, "decay"), (1, 1, "decay"), (1000, 120, "decay"),
                                       (600, 100, "bigcluster"), (300, 65, "bigcluster")])
def test_eigs_sym_matches_lapack(flgp, s, K, kind):
    """The eigensolver behind spectrum_from_Z / the eigs_sym seam against LAPACK on dense symmetric matrices."""
    rng = np.random.default_rng(s + K)
    A, _ = _sym(rng, s, kind)
    res = flgp.eigs_sym(A, K)
    w, V = np.linalg.eigh(A)
    w, V = w[::-1][:K], V[:, ::-1][:, :K]
    scale = np.abs(w).max()
    assert np.abs(res["values"] - w).max() <= 1e-12 * max(scale, 1.0) * s
    Y = res["vectors"]
    assert np.abs(Y.T @ Y - np.eye(K)).max() < 1e-10
    assert np.abs(A @ Y - Y * res["values"]).max() <= 1e-10 * max(scale, 1.0)
    # invariant subspaces agree wherever the K cut does not split a cluster
    if K < s and w[-1] - np.linalg.eigvalsh(A)[::-1][K] > 1e-6:
        assert np.abs(Y @ Y.T - V @ V.T).max() < 1e-8


# ------------------------------------------------------------------------------------------- heat kernel GEMM
@pytest.mark.parametrize("n0,n1,K", [(64, 64, 16), (333, 257, 30), (1000, 700, 200), (70, 5, 9), (129, 640, 37)])
def test_hk_from_spectrum_tensor_core_gemm(flgp, oracle, n0, n1, K):
    """HK_from_spectrum_cpp through the DMMA + TMA GEMM (and its FMA fallback for small / odd shapes) against the
    oracle's plain triple loop on the same lifted eigenvectors."""
    X, _ = spiral(3000, 21)
    s, r, KK = 150, 3, 40 if K <= 40 else 200
    if KK > s:
        s = 260
    init = _init(3000, s, 4)
    ep = flgp.heat_kernel_spectrum_cpp(X[:100], X[100:], s, r, KK, init_idx=init, iter_max=15)
    rng = np.random.default_rng(n0 + n1)
    idx0 = rng.integers(0, 3000, n0).astype(np.int32)
    idx1 = rng.integers(0, 3000, n1).astype(np.int32)
    H = flgp.HK_from_spectrum_cpp(ep, K, 3.0, idx0, idx1)
    V = ep.vectors
    lam = np.exp(-3.0 * (1.0 - ep.values[:K]))
    Ho = (V[idx0, :K] * lam) @ V[idx1, :K].T
    assert H.shape == (n0, n1)
    assert np.abs(H - Ho).max() <= 1e-12 * max(1.0, np.abs(Ho).max())


# ------------------------------------------------------------------------------------------- large d: tensor-core path
@pytest.mark.parametrize("n,d,s", [(1200, 37, 70), (900, 784, 64), (3000, 6, 200)])
def test_kmeans_large_d_tensor_core_bitexact(flgp, oracle, n, d, s, monkeypatch):
    """d > 4, n >= 64, s >= 64: the assignment runs as DMMA inner products + certified arg-min (distsel.cu); the
    result must be the oracle's bit for bit, and identical to the oracle-order FMA kernel (FLGP_NO_DMMA_DIST)."""
    rng = np.random.default_rng(n + d + s)
    X = np.asfortranarray(rng.standard_normal((n, d)) * rng.uniform(0.2, 4.0, d) + rng.integers(0, 3, (n, 1)))
    init = _init(n, s, 8)
    U, assign, iters = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, iter_max=15, return_info=True)
    Uo, ao, io = oracle.kmeans_lloyd(X, s, init, 15, NT)
    assert iters == io and np.array_equal(assign, ao) and np.array_equal(U, Uo)
    monkeypatch.setenv("FLGP_NO_DMMA_DIST", "1")
    U2, a2, it2 = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, iter_max=15, return_info=True)
    assert it2 == iters and np.array_equal(a2, assign) and np.array_equal(U2, U)


def test_kmeans_large_d_bound_filter_many_passes(flgp, oracle, monkeypatch):
    """Config 5's data type (2-torus in 16 dimensions + noise) over many Lloyd passes: from pass 2 on most rows are
    skipped by their Hamerly bounds (neighbour-list-local decay) and only the survivors reach the tensor-core
    evaluation; centres, sizes, assignments and the iteration count must stay the brute-force oracle's bit for bit,
    and equal to the unfiltered run (FLGP_NO_HAMERLY)."""
    rng = np.random.default_rng(16)
    n, d, s = 20000, 16, 128
    a, b = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    Q, _ = np.linalg.qr(rng.standard_normal((d, 4)))
    X = np.asfortranarray(np.c_[np.cos(a), np.sin(a), np.cos(b), np.sin(b)] @ Q.T + 0.01 * rng.standard_normal((n, d)))
    init = _init(n, s, 12)
    U, assign, iters = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, iter_max=60, return_info=True)
    Uo, ao, io = oracle.kmeans_lloyd(X, s, init, 60, NT)
    assert iters == io and np.array_equal(assign, ao) and np.array_equal(U, Uo)
    monkeypatch.setenv("FLGP_NO_HAMERLY", "1")
    U2, a2, it2 = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, iter_max=60, return_info=True)
    assert it2 == iters and np.array_equal(a2, assign) and np.array_equal(U2, U)


def test_large_d_ties_take_the_exact_fallback(flgp, oracle):
    """Lattice data in d = 6 with duplicated centres / anchors: no row can be certified by the tensor-core
    selection (exact ties), every one is re-done in the oracle's order; results still bit-exact."""
    rng = np.random.default_rng(66)
    n, d, s = 2500, 6, 96
    X = np.asfortranarray(rng.integers(-3, 4, (n, d)).astype(np.float64))
    init = _init(n, s, 5)
    X[init[9]] = X[init[2]]
    X[init[70]] = X[init[2]]
    for iter_max in (1, 4):
        U, assign, iters = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, iter_max=iter_max, return_info=True)
        Uo, ao, io = oracle.kmeans_lloyd(X, s, init, iter_max, NT)
        assert iters == io and np.array_equal(assign, ao) and np.array_equal(U, Uo)
    A = np.asfortranarray(X[init])
    for r in (1, 2, 3, 5):
        ind, dist = flgp.knn_distances(X, A, r)
        io_, do_ = oracle.knn(X, A, r, want_dist=True, nthreads=NT)
        assert np.array_equal(ind, io_) and np.array_equal(dist, do_)


@pytest.mark.parametrize("n,d,s,r", [(5000, 16, 300, 5), (700, 101, 64, 2), (2000, 9, 257, 3), (333, 50, 70, 1)])
def test_knn_large_d_tensor_core_bitexact(flgp, oracle, n, d, s, r, monkeypatch):
    rng = np.random.default_rng(n + d)
    X = np.asfortranarray(rng.standard_normal((n, d)) * 5.0)
    U = np.asfortranarray(X[_init(n, s, 3)] + 0.3 * rng.standard_normal((s, d)))
    ind, dist = flgp.knn_distances(X, U, r)
    io_, do_ = oracle.knn(X, U, r, want_dist=True, nthreads=NT)
    assert np.array_equal(ind, io_) and np.array_equal(dist, do_)
    monkeypatch.setenv("FLGP_NO_DMMA_DIST", "1")
    ind2, dist2 = flgp.knn_distances(X, U, r)
    assert np.array_equal(ind2, ind) and np.array_equal(dist2, dist)


# ------------------------------------------------------------------------------------------- large d (config 3 shape)
def test_c3_shape_pipeline_large_d(flgp, oracle):
    """BASELINE config 3's shape at reduced n: d = 784, r = 5 through the tiled (any-d) k-means / KNN / LAE kernels."""
    rng = np.random.default_rng(33)
    n, d, s, r, K, m = 2500, 784, 80, 5, 20, 200
    means = 3.0 * rng.standard_normal((10, d))
    lab = rng.integers(0, 10, n)
    X = np.asfortranarray(means[lab] + rng.standard_normal((n, d)))
    init = _init(n, s, 6)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=12)
    vo, Vo, I = oracle.heat_kernel_spectrum(X[:m], X[m:], s, r, K, init, nthreads=NT, want_internals=True, iter_max=12)
    assert ep.kmeans_iters == I["iters"]
    assert np.array_equal(ep.anchors(), I["U"])
    Zj, Zx = _csr_parts(ep.Z())
    assert np.array_equal(Zj, I["Zj"]) and np.array_equal(Zx, I["Zx"])
    np.testing.assert_allclose(ep.values, vo, rtol=1e-8, atol=1e-10)


# ------------------------------------------------------------------------------------------- training (§8f rows 2, 3)
@pytest.mark.parametrize("m,K", [(300, 40), (30, 40)])
@pytest.mark.parametrize("approach", ["marginal", "posterior"])
def test_regression_objective_matches_oracle(flgp, oracle, m, K, approach):
    """The training objective and its gradient on a spectrum handle (K-sized statistics gathered on the device)
    against the oracle's literal restatement on the materialised eigenvectors; both branches of the reference."""
    X, Y = spiral(2500, 5)
    s, r = 120, 3
    init = _init(len(X), s, 2)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=20)
    V, values = ep.vectors, ep.values
    idx = np.arange(m, dtype=np.int32)
    for pars in [(10.0, 1.0), (2.0, 0.05), (40.0, 3.0)]:
        f, g = flgp.regression_objective(ep, Y[:m], m, K, pars, 1e-5, approach)
        fo, go = oracle.regression_objective(V, values, Y[:m], idx, K, pars, 1e-5, approach)
        np.testing.assert_allclose(f, fo, rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(g, go, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(go).max()))
    with pytest.raises(flgp.FlgpError, match="not supported"):
        flgp.regression_objective(ep, Y[:m], m, K, (10.0, 1.0), 1e-5, "bayes")


def test_train_regression_matches_oracle_optimiser(flgp, oracle):
    """train_regression_gp_cpp: the library's MMA run on the device-side statistics lands where the oracle's twin
    lands on the materialised eigenvectors (to optimiser tolerance), improving on the start x0 = (10, 1)."""
    X, Y = spiral(3000, 8)
    m, s, r, K = 200, 150, 3, 50
    init = _init(len(X), s, 3)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=20)
    idx = np.arange(m, dtype=np.int32)
    for approach in ("posterior", "marginal"):
        x, obj, nev = flgp.train_regression_gp(ep, Y[:m], m, K, 1e-5, approach)
        xo, objo = oracle.train_regression(ep.vectors, ep.values, Y[:m], idx, K, 1e-5, approach)
        assert x[0] >= 1e-3 and x[1] >= 1e-4 and nev >= 2
        f0, _ = flgp.regression_objective(ep, Y[:m], m, K, (10.0, 1.0), 1e-5, approach)
        assert -obj <= f0
        np.testing.assert_allclose(obj, objo, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(x, xo, rtol=1e-3)


def test_fit_lae_regression_trains_when_pars_missing(flgp, oracle):
    X, Y = swiss(3000, 4)
    m, s, r, K = 250, 100, 3, 30
    init = _init(len(X), s, 9)
    res = flgp.fit_lae_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, init_idx=init, iter_max=15)
    ref = oracle.fit_lae_regression_fixed(X[:m], Y[:m], X[m:], s, r, K, res["pars"], init, iter_max=15, nthreads=NT)
    np.testing.assert_allclose(res["Y_pred"]["test"], ref["test"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["posterior"]["cov"], ref["cov"], rtol=1e-7, atol=1e-9)
    xo, objo = oracle.train_regression(ref["vectors"], ref["values"], Y[:m], np.arange(m, dtype=np.int32), K)
    np.testing.assert_allclose(res["pars"], xo, rtol=1e-3)
    np.testing.assert_allclose(res["obj"], objo, rtol=1e-5, atol=1e-5)


def test_fit_se_regression_config2_grid(flgp, oracle):
    """BASELINE config 2 (README GPR spiral: n=4000, d=2, m=200, s=500, r=3, K=100, fit_se_regression_gp_rcpp):
    one k-means + KNN, ten bandwidths.  With fixed (t, noise) the whole path is deterministic: same winning a2,
    predictions and variances to 1e-8; with training, to optimiser tolerance."""
    X, Y = spiral(4000, 2)
    m, s, r, K = 200, 500, 3, 100
    init = _init(len(X), s, 1)
    a2s = np.exp(np.linspace(np.log(0.1), np.log(10.0), 10))
    pars = (8.0, 0.5)
    res = flgp.fit_se_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, pars=pars, init_idx=init)
    ref = oracle.fit_se_regression(X[:m], Y[:m], X[m:], s, r, K, init, a2s, pars=pars, nthreads=NT)
    assert res["a2"] == ref["a2"]
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-8)
    np.testing.assert_allclose(res["eigenpair"].values, ref["values"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(res["Y_pred"]["train"], ref["train"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["Y_pred"]["test"], ref["test"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["posterior"]["cov"], ref["cov"], rtol=1e-7, atol=1e-9)
    # trained
    res = flgp.fit_se_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, init_idx=init, output_cov=True)
    ref = oracle.fit_se_regression(X[:m], Y[:m], X[m:], s, r, K, init, a2s, nthreads=NT)
    assert res["a2"] == ref["a2"]
    np.testing.assert_allclose(res["pars"], ref["pars"], rtol=1e-3)
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-5)
    np.testing.assert_allclose(res["Y_pred"]["test"], ref["test"], rtol=1e-3, atol=1e-3)
    assert res["C"].shape == (4000, m)
    rmse = np.sqrt(np.mean((res["Y_pred"]["test"] - Y[m:]) ** 2))
    assert rmse < 2.0   # labels carry N(0,1) noise on a signal of amplitude ~10


# ------------------------------------------------------------------------------------------- Nystrom (§8f row 4)
@pytest.mark.parametrize("n,d,m,s,K", [(3000, 2, 200, 300, 60), (1500, 3, 40, 128, 50), (2000, 16, 150, 100, 100)])
def test_fit_nystrom_regression(flgp, oracle, n, d, m, s, K):
    """fit_nystrom_regression_gp_rcpp against the oracle's dense literal restatement: same winning bandwidth, objective
    to 1e-7, predictions and variances to 1e-6 at fixed (t, noise) (the dense s x s normalisation and the n x s x K
    extension accumulate in different orders); trained parameters to optimiser tolerance."""
    if d == 2:
        X, Y = spiral(n, 6)
    elif d == 3:
        X, Y = swiss(n, 6)
    else:
        rng = np.random.default_rng(6)
        a, b = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
        Q, _ = np.linalg.qr(rng.standard_normal((d, 4)))
        X = np.asfortranarray(np.c_[np.cos(a), np.sin(a), np.cos(b), np.sin(b)] @ Q.T + 0.01 * rng.standard_normal((n, d)))
        Y = np.sin(a) + np.cos(2 * b) + 0.1 * rng.standard_normal(n)
    init = _init(n, s, 4)
    a2s = np.exp(np.linspace(np.log(0.1), np.log(10.0), 5))
    pars = (6.0, 0.3)
    res = flgp.fit_nystrom_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, K, a2s=a2s, pars=pars, init_idx=init, iter_max=25)
    ref = oracle.fit_nystrom_regression(X[:m], Y[:m], X[m:], s, K, init, a2s, pars=pars, iter_max=25, nthreads=NT)
    assert res["a2"] == ref["a2"]
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-7, atol=1e-7)
    scale = max(1.0, np.abs(ref["test"]).max())
    np.testing.assert_allclose(res["Y_pred"]["train"], ref["train"], rtol=1e-6, atol=1e-6 * scale)
    np.testing.assert_allclose(res["Y_pred"]["test"], ref["test"], rtol=1e-6, atol=1e-6 * scale)
    np.testing.assert_allclose(res["posterior"]["cov"], ref["cov"], rtol=1e-6, atol=1e-8)
    if d == 2:  # with training
        res = flgp.fit_nystrom_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, K, a2s=a2s, init_idx=init, iter_max=25)
        ref = oracle.fit_nystrom_regression(X[:m], Y[:m], X[m:], s, K, init, a2s, iter_max=25, nthreads=NT)
        assert res["a2"] == ref["a2"]
        np.testing.assert_allclose(res["pars"], ref["pars"], rtol=1e-3)
        np.testing.assert_allclose(res["Y_pred"]["test"], ref["test"], rtol=1e-3, atol=1e-3 * scale)


# ------------------------------------------------------------------------------------------- committed golden fixtures
def test_cuda_path_against_golden_fixtures(flgp):
    """The CUDA path against the committed fixtures of tests/golden/ directly (no oracle call at run time): k-means,
    KNN, Z, eigenvalues (oracle_small.npz); training objective, trained pars, SE grid and Nystrom driver
    (oracle_train.npz)."""
    import json
    import os

    gold = os.path.join(os.path.dirname(__file__), "golden")
    g0 = np.load(os.path.join(gold, "oracle_small.npz"))
    g = np.load(os.path.join(gold, "oracle_train.npz"))
    meta = json.loads(str(g0["meta"]))
    X, Y = spiral(meta["n"], meta["seed"])
    init, s, r, K, m = g0["init"], meta["s"], meta["r"], meta["K"], int(g["m"])
    U, assign, iters = flgp.subsample_cpp(X, s, "kmeans", init_idx=init, return_info=True)
    assert iters == meta["iters"] and np.array_equal(assign, g0["assign"]) and np.array_equal(U, g0["U"])
    assert np.array_equal(flgp.KNN_cpp(X, np.asfortranarray(U[:, :2]), r)["ind_knn"], g0["ind"])
    Zj, Zx = _csr_parts(flgp.cross_similarity_lae_cpp(X, U, r, "cluster-normalized"))
    assert np.array_equal(Zj, g0["Zj"]) and np.array_equal(Zx, g0["Zx"])
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init)
    np.testing.assert_allclose(ep.values, g0["values"], rtol=1e-8)
    for q, (mm, ap) in enumerate([(m, "marginal"), (m, "posterior"), (8, "posterior")]):
        f, gr = flgp.regression_objective(ep, Y[:mm], mm, K, (6.0, 0.4), 1e-5, ap)
        np.testing.assert_allclose(f, g["obj"][q], rtol=1e-8)
        np.testing.assert_allclose(gr, g["grad"][q], rtol=1e-6, atol=1e-7)
    x, obj, _ = flgp.train_regression_gp(ep, Y[:m], m, K, 1e-5, "posterior")
    np.testing.assert_allclose(x, g["pars"], rtol=1e-3)
    np.testing.assert_allclose(obj, g["obj_t"], rtol=1e-6)
    se = flgp.fit_se_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, a2s=g["a2s"], pars=(6.0, 0.4), init_idx=init,
                                        iter_max=30)
    assert se["a2"] == float(g["se_a2"])
    np.testing.assert_allclose(se["Y_pred"]["test"][:200], g["se_test"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(se["posterior"]["cov"][:200], g["se_cov"], rtol=1e-7, atol=1e-9)
    ny = flgp.fit_nystrom_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, K, a2s=g["a2s"], pars=(6.0, 0.4), init_idx=init,
                                             iter_max=30)
    assert ny["a2"] == float(g["ny_a2"])
    np.testing.assert_allclose(ny["Y_pred"]["test"][:200], g["ny_test"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ny["posterior"]["cov"][:200], g["ny_cov"], rtol=1e-6, atol=1e-8)


def test_cuda_path_against_golden_fixtures_late_round2(flgp):
    """The CUDA path against tests/golden/oracle_round2b.npz directly (no oracle call at run time): mini-batch anchors
    bit for bit, noise = "different" objective and prediction, SE and Nystrom logit grids at a fixed diffusion time."""
    import json
    import os

    gold = os.path.join(os.path.dirname(__file__), "golden")
    g0 = np.load(os.path.join(gold, "oracle_small.npz"))
    gt = np.load(os.path.join(gold, "oracle_train.npz"))
    g = np.load(os.path.join(gold, "oracle_round2b.npz"))
    meta = json.loads(str(g0["meta"]))
    X, Y = spiral(meta["n"], meta["seed"])
    init, s, r, K, m = g0["init"], meta["s"], meta["r"], meta["K"], int(gt["m"])
    U, _, it = flgp.subsample_cpp(X, s, "minibatchkmeans", init_idx=init, seed=9, iter_max=40, return_info=True)
    assert it == int(g["it_mb"]) and np.array_equal(U, g["Umb"])
    res = flgp.fit_lae_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, 1e-5, noise="different", pars=g["xd"],
                                          init_idx=init)
    np.testing.assert_allclose(-res["obj"], g["obj_d"], rtol=1e-8)
    np.testing.assert_allclose(res["Y_pred"]["test"][:200], g["pred_d"], rtol=1e-7, atol=1e-8)
    lab = (Y > np.median(Y)).astype(np.float64)
    sl = flgp.fit_se_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, r, K, a2s=gt["a2s"], t=6.0, init_idx=init, iter_max=30)
    assert sl["a2"] == float(g["sl_a2"])
    np.testing.assert_allclose(sl["obj"], g["sl_obj"], rtol=1e-7)
    np.testing.assert_allclose(sl["posterior"]["mean"][:200], g["sl_mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sl["posterior"]["cov"][:200], g["sl_cov"], rtol=1e-6, atol=1e-7)
    nl = flgp.fit_nystrom_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, K, a2s=gt["a2s"], t=6.0, init_idx=init, iter_max=30)
    assert nl["a2"] == float(g["nl_a2"])
    np.testing.assert_allclose(nl["obj"], g["nl_obj"], rtol=1e-6)
    np.testing.assert_allclose(nl["posterior"]["mean"][:200], g["nl_mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(nl["posterior"]["cov"][:200], g["nl_cov"], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------- edge shapes of the new paths
@pytest.mark.parametrize("n,d,s,r,K", [(130, 16, 64, 5, 64), (64, 5, 64, 1, 10), (200, 9, 65, 5, 65), (97, 33, 64, 4, 7)])
def test_large_d_minimum_shapes(flgp, oracle, n, d, s, r, K):
    """Smallest shapes that still take the tensor-core path (n >= 64, s >= 64), K = s, r = 1, ragged last tiles."""
    rng = np.random.default_rng(n + d)
    X = np.asfortranarray(rng.standard_normal((n, d)) * 2.0 + rng.integers(0, 3, (n, 1)))
    init = _init(n, s, 2)
    m = 20
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=6)
    vo, Vo, I = oracle.heat_kernel_spectrum(X[:m], X[m:], s, r, K, init, nthreads=NT, want_internals=True, iter_max=6)
    assert ep.kmeans_iters == I["iters"] and np.array_equal(ep.anchors(), I["U"])
    Zj, Zx = _csr_parts(ep.Z())
    assert np.array_equal(Zj, I["Zj"]) and np.array_equal(Zx, I["Zx"])
    np.testing.assert_allclose(ep.values, vo, rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("gl", ["rw", "normalized"])
def test_fit_se_regression_other_laplacians_and_errors(flgp, oracle, gl):
    X, Y = spiral(1500, 13)
    m, s, r, K = 100, 80, 3, 20
    init = _init(len(X), s, 7)
    a2s = np.array([0.5, 2.0])
    res = flgp.fit_se_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, a2s=a2s, pars=(5.0, 0.3), models=dict(gl=gl),
                                         init_idx=init, iter_max=20)
    ref = oracle.fit_se_regression(X[:m], Y[:m], X[m:], s, r, K, init, a2s, gl=gl, pars=(5.0, 0.3), iter_max=20,
                                   nthreads=NT)
    assert res["a2"] == ref["a2"]
    np.testing.assert_allclose(res["Y_pred"]["test"], ref["test"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["posterior"]["cov"], ref["cov"], rtol=1e-7, atol=1e-9)
    with pytest.raises(flgp.FlgpError, match="not supported"):
        flgp.fit_se_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, approach="evidence", init_idx=init)
    with pytest.raises(flgp.FlgpError, match="illegal"):
        flgp.fit_se_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, r, K, noise="none", init_idx=init)
    with pytest.raises(flgp.FlgpError):
        flgp.fit_nystrom_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, s + 1, init_idx=init)   # K > s


def test_nystrom_random_anchors_and_K_equals_s(flgp, oracle):
    """subsample="random" anchors (no k-means) and K = s through the Nystrom driver: finite, positive variances."""
    X, Y = spiral(900, 3)
    m, s = 60, 40
    init = _init(len(X), s, 1)
    res = flgp.fit_nystrom_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, -1, a2s=[1.0], pars=(4.0, 0.5), subsample="random",
                                              init_idx=init)
    assert np.all(np.isfinite(res["Y_pred"]["test"])) and np.all(res["posterior"]["cov"] > 0)


# ------------------------------------------------------------------------------------------- Laplace posterior (logit)
@pytest.mark.parametrize("m,K", [(100, 100), (60, 25)])
def test_classification_posterior_config1_rings(flgp, oracle, m, K):
    """BASELINE config 1 (README GPC rings: n=4800, d=2, m=100, s=600, r=3, K=100, fit_lae_logit_gp_rcpp): the
    deterministic half of the driver at a fixed diffusion time — Laplace posterior mean / variance of every row,
    folded through the factored eigenvectors, against the oracle's literal posterior_distribution_classification on
    the dense covariance blocks (src/Fit.cpp:563-582)."""
    from flgp_b200.datasets import make

    X, lab, cfg = make("C1")  # rows already shuffled: training rows first
    s, r, t, sigma = cfg["s"], cfg["r"], 10.0, 1e-3
    init = _init(len(X), s, 1)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=50)
    mean, cov = flgp.posterior_distribution_classification(ep, lab[:m], m, K, t, sigma)
    V, values = ep.vectors, ep.values
    n = len(X)
    idx0, idx1 = np.arange(m, dtype=np.int32), np.arange(n, dtype=np.int32)
    C11 = oracle.hk_from_spectrum(V, values, K, t, idx0, idx0)
    C11[np.diag_indices(m)] += sigma
    C21 = oracle.hk_from_spectrum(V, values, K, t, idx1, idx0)
    C22 = ((V[:, :K] * np.exp(-t * (1.0 - values[:K]))) * V[:, :K]).sum(axis=1) + sigma
    mo, co = oracle.posterior_distribution_classification(C11, C21, C22, lab[:m])
    np.testing.assert_allclose(mean, mo, rtol=1e-7, atol=1e-8 * max(1.0, np.abs(mo).max()))
    np.testing.assert_allclose(cov, co, rtol=1e-7, atol=1e-8 * max(1.0, np.abs(co).max()))
    acc = np.mean((mean[m:] > 0) == (lab[m:] > 0.5))
    assert acc > 0.8  # six separated rings, untuned t: the Laplace mode classifies most held-out points


def test_posterior_distribution_classification_export(flgp, oracle):
    """The reference's own export on explicit covariance blocks (src/Utils.h:77-80), ragged sizes."""
    rng = np.random.default_rng(9)
    for m, q in [(37, 501), (64, 64), (5, 1)]:
        A = rng.standard_normal((m + q, 11))
        Kf = A @ A.T * 0.2 + 1e-3 * np.eye(m + q)
        C11, C21, C22 = np.asfortranarray(Kf[:m, :m]), np.asfortranarray(Kf[m:, :m]), np.diag(Kf)[m:].copy()
        Y = (rng.uniform(size=m) > 0.4).astype(np.float64)
        res = flgp.posterior_distribution_classification_rcpp(C11, C21, C22, Y)
        mo, co = oracle.posterior_distribution_classification(C11, C21, C22, Y)
        np.testing.assert_allclose(res["mean"], mo, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(res["cov"], co, rtol=1e-9, atol=1e-10)
    with pytest.raises(flgp.FlgpError, match="inconsistent"):
        flgp.posterior_distribution_classification_rcpp(C11, C21[:, :1], C22, Y)


# ------------------------------------------------------------------------------------------- round 2: iterative eigensolver
def _chfsi_matrix(rng, s, kind):
    Q, _ = np.linalg.qr(rng.standard_normal((s, s)))
    if kind == "graph":          # like the Gram of a row-stochastic Z: top eigenvalue 1, slow decay to a positive floor
        lam = 0.2 + 0.8 / (1.0 + 0.004 * np.arange(s)) ** 2
    elif kind == "degenerate":   # exact multiplicities at the top (disconnected anchor graph) and tight clusters
        lam = np.sort(np.r_[np.ones(6), np.full(4, 0.97), 0.9 - 1e-9 * np.arange(3), rng.uniform(0.0, 0.85, s - 13)])[::-1]
    elif kind == "indefinite":
        lam = np.sort(rng.uniform(-1.0, 1.0, s))[::-1]
    else:                        # "flat": the K-th gap is tiny compared with the spectrum: the iteration must either
        lam = np.sort(rng.uniform(0.5, 0.5001, s))[::-1]  # converge or hand over to the direct solver
        lam[:5] = [1.0, 0.9, 0.8, 0.7, 0.6]
    A = (Q * lam) @ Q.T
    return np.asfortranarray((A + A.T) / 2), lam


@pytest.mark.parametrize("s,K,kind", [(1200, 100, "graph"), (2000, 200, "graph"), (1024, 60, "indefinite"),
                                       (1500, 150, "degenerate"), (1100, 40, "flat")])
def test_eigs_sym_iterative_route(flgp, s, K, kind):
    """s >= 1024 and K <= s / 5 take the Chebyshev-filtered subspace iteration (chfsi.cu) — or fall back to the direct
    solver when it does not converge; either way the answer must be LAPACK's."""
    rng = np.random.default_rng(s + K)
    A, _ = _chfsi_matrix(rng, s, kind)
    res = flgp.eigs_sym(A, K)
    w_all, V_all = np.linalg.eigh(A)
    w, V = w_all[::-1][:K], V_all[:, ::-1][:, :K]
    scale = max(np.abs(w_all).max(), 1.0)
    assert np.abs(res["values"] - w).max() <= 1e-12 * scale * s
    Y = res["vectors"]
    assert np.abs(Y.T @ Y - np.eye(K)).max() < 1e-10
    assert np.abs(A @ Y - Y * res["values"]).max() <= 1e-10 * scale
    gap = w[-1] - w_all[::-1][K]
    if gap > 1e-4 * scale:
        assert np.abs(Y @ Y.T - V @ V.T).max() < 1e-8


def test_eigs_sym_iterative_equals_direct(flgp, monkeypatch):
    """The two routes behind one seam on the same matrix: eigenvalues to 1e-12, invariant subspace to 1e-9."""
    rng = np.random.default_rng(77)
    A, _ = _chfsi_matrix(rng, 1600, "graph")
    K = 160
    it = flgp.eigs_sym(A, K)
    monkeypatch.setenv("FLGP_EIGH_DIRECT", "1")
    import subprocess
    import sys as _sys
    # the switch is read once per process: the direct route runs in a child process
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import flgp_b200 as F; A = np.load(sys.argv[1]); "
            "r = F.eigs_sym(np.asfortranarray(A), %d); np.save(sys.argv[2], r['values']); np.save(sys.argv[3], r['vectors'])"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), K))
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "A.npy"), A)
        subprocess.run([_sys.executable, "-c", code, os.path.join(td, "A.npy"), os.path.join(td, "w.npy"),
                        os.path.join(td, "Y.npy")], check=True, env=dict(os.environ, FLGP_EIGH_DIRECT="1"))
        wd, Yd = np.load(os.path.join(td, "w.npy")), np.load(os.path.join(td, "Y.npy"))
    assert np.abs(it["values"] - wd).max() <= 1e-12
    assert np.abs(it["vectors"] @ it["vectors"].T - Yd @ Yd.T).max() < 1e-9


# ------------------------------------------------------------------------------------------- round 2: headline shapes vs the oracle
def test_c4_headline_shape_against_oracle(flgp, oracle):
    """BASELINE config 4 at n = 2e6 with the full s = 2000, r = 3, K = 200 and 5 Lloyd passes, against the oracle
    (about a minute of CPU time): anchors, sizes, Z pattern and Z values bit-exact, eigenvalues 1e-8."""
    from flgp_b200.datasets import make

    n, m = 2_000_000, 5000
    X, Y, cfg = make("C4", n=n)
    s, r, K = cfg["s"], cfg["r"], cfg["K"]
    init = _init(n, s, 1)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=5)
    vo, Vo, I = oracle.heat_kernel_spectrum(X[:m], X[m:], s, r, K, init, nthreads=NT, want_internals=True, iter_max=5)
    assert ep.kmeans_iters == I["iters"] == 5
    assert np.array_equal(ep.anchors(), I["U"])
    Zj, Zx = _csr_parts(ep.Z())
    assert np.array_equal(Zj, I["Zj"]) and np.array_equal(Zx, I["Zx"])
    np.testing.assert_allclose(ep.values, vo, rtol=1e-8, atol=1e-10)
    idx = np.arange(0, n, n // 200, dtype=np.int32)
    H = flgp.HK_from_spectrum_cpp(ep, K, 10.0, idx, idx)
    lam = np.exp(-10.0 * (1.0 - vo))
    Ho = (Vo[idx] * lam) @ Vo[idx].T
    assert np.abs(H - Ho).max() <= 1e-8 * np.abs(Ho).max()


def test_c3_headline_shape_against_oracle(flgp, oracle):
    """BASELINE config 3 at its full shape n = 70000, d = 784, s = 1000, r = 5, K = 200 with 2 Lloyd passes (the
    tensor-core distance path with certified selection): anchors, sizes, Z bit-exact, eigenvalues 1e-8."""
    from flgp_b200.datasets import make

    X, Y, cfg = make("C3")
    n, m, s, r, K = cfg["n"], cfg["m"], cfg["s"], cfg["r"], cfg["K"]
    init = _init(n, s, 3)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=2)
    vo, Vo, I = oracle.heat_kernel_spectrum(X[:m], X[m:], s, r, K, init, nthreads=NT, want_internals=True, iter_max=2)
    assert ep.kmeans_iters == I["iters"]
    assert np.array_equal(ep.anchors(), I["U"])
    Zj, Zx = _csr_parts(ep.Z())
    assert np.array_equal(Zj, I["Zj"]) and np.array_equal(Zx, I["Zx"])
    np.testing.assert_allclose(ep.values, vo, rtol=1e-8, atol=1e-10)


# ------------------------------------------------------------------------------------------- round 2: ADVICE items
def test_caller_supplied_csr_is_validated_and_scaled(flgp, oracle):
    """flgp_spectrum_from_z / flgp_graph_laplacian on a caller's sparse matrix: unsorted scipy rows are sorted by the
    binding, bad indices are rejected by the library, and values far outside [0, 1] keep full accuracy (the
    fixed-point scale follows max |Z| and max w instead of a constant)."""
    import scipy.sparse as sp

    rng = np.random.default_rng(8)
    n, s, r, K = 3000, 60, 4, 12
    cols = np.stack([rng.choice(s, r, replace=False) for _ in range(n)]).astype(np.int32)   # unsorted rows
    vals = rng.uniform(0.5, 40.0, (n, r)) * rng.choice([1.0, 1.0, 1.0, -0.2], (n, r))        # large, some negative
    Z = sp.csr_matrix((vals.reshape(-1), cols.reshape(-1), np.arange(0, n * r + 1, r)), shape=(n, s))
    res = flgp.spectrum_from_Z_cpp(Z, K, root=False)
    # dense reference: sigma^2 of A = Z diag(1/sqrt(|colsum| + 1e-9))
    Zd = Z.toarray()
    c = Zd.sum(0)
    A = Zd / np.sqrt(np.abs(c) + 1e-9)
    sv = np.linalg.svd(A, compute_uv=False)[:K] ** 2
    np.testing.assert_allclose(res.values if hasattr(res, "values") else res["values"], sv, rtol=1e-9)
    # a column index outside [0, s) and a duplicated column are rejected
    bad = cols.copy()
    bad.sort(axis=1)
    bad[5, r - 1] = s
    with pytest.raises(flgp.FlgpError, match="outside"):
        flgp.spectrum_from_Z_cpp((bad, np.abs(vals), s), K)
    dup = np.sort(cols, axis=1)
    dup[7, 1] = dup[7, 0]
    with pytest.raises(flgp.FlgpError, match="ascending"):
        flgp.graphLaplacian_cpp((dup, np.abs(vals), s), "rw")


def test_kmeans_nstart_keeps_the_best_run(flgp, oracle):
    """nstart > 1 = stats::kmeans's restarts: the returned run has the smallest total within-cluster sum of squares of
    the nstart runs (each run is the Lloyd contract from its own start rows)."""
    X, _ = swiss(5000, 4)
    s = 25
    U1 = flgp.subsample_cpp(X, s, "kmeans", nstart=1, seed=11)
    U4 = flgp.subsample_cpp(X, s, "kmeans", nstart=4, seed=11)

    def wss(U):
        D = ((X[:, None, :] - U[None, :, :3]) ** 2).sum(-1)
        return D.min(1).sum()

    assert wss(U4) <= wss(U1) * (1 + 1e-12)
    assert U4[:, 3].sum() == len(X)


@pytest.mark.parametrize("approach", ["posterior", "marginal"])
def test_fit_lae_logit_config1_end_to_end(flgp, oracle, approach):
    """BASELINE config 1 end to end (README GPC rings: n=4800, d=2, m=100, s=600, r=3, K=100, fit_lae_logit_gp_rcpp):
    spectrum -> training of the diffusion time t (COBYLA restatement on the Laplace objective) -> Laplace posterior of
    the test rows, against the oracle twin run on the library's own eigenvectors: the objective to 1e-8, the trained
    t to the optimiser's tolerance, posterior mean / variance at the trained t to 1e-7."""
    from flgp_b200.datasets import make

    X, lab, cfg = make("C1")
    m, s, r, K, sigma = cfg["m"], cfg["s"], cfg["r"], cfg["K"], 1e-3
    init = _init(len(X), s, 1)
    res = flgp.fit_lae_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, r, K, sigma=sigma, approach=approach, init_idx=init,
                                     iter_max=50, output_cov=True)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=50)
    V, values = ep.vectors, ep.values
    n = len(X)
    idx0 = np.arange(m, dtype=np.int32)
    for t in (2.0, 10.0, 40.0):
        got = flgp.logit_objective(ep, lab[:m], m, K, t, sigma, approach)
        want = oracle.logit_objective(V, values, lab[:m], idx0, K, t, sigma, approach)
        assert abs(got - want) <= 1e-8 * max(1.0, abs(want))
    t_o, obj_o, nev_o = oracle.train_lae_logit(V, values, lab[:m], idx0, K, sigma, approach)
    t_l, obj_l, nev_l = flgp.train_lae_logit_gp(ep, lab[:m], m, K, sigma, approach)
    assert abs(t_l - t_o) <= 1e-3 * max(1.0, t_o) and abs(obj_l - obj_o) <= 1e-6 * max(1.0, abs(obj_o))
    assert abs(res["pars"] - t_l) <= 1e-12 * t_l and abs(res["obj"] - obj_l) <= 1e-9 * max(1.0, abs(obj_l))
    # posterior of the test rows at the trained t (src/Fit.cpp:563-582)
    t = res["pars"]
    idx1 = np.arange(m, n, dtype=np.int32)
    C11 = oracle.hk_from_spectrum(V, values, K, t, idx0, idx0)
    C11[np.diag_indices(m)] += sigma
    C21 = oracle.hk_from_spectrum(V, values, K, t, idx1, idx0)
    C22 = ((V[m:, :K] * np.exp(-t * (1.0 - values[:K]))) * V[m:, :K]).sum(axis=1) + sigma
    mo, co = oracle.posterior_distribution_classification(C11, C21, C22, lab[:m])
    np.testing.assert_allclose(res["posterior"]["mean"], mo, rtol=1e-7, atol=1e-8 * max(1.0, np.abs(mo).max()))
    np.testing.assert_allclose(res["posterior"]["cov"], co, rtol=1e-7, atol=1e-8 * max(1.0, np.abs(co).max()))
    np.testing.assert_allclose(res["C"][:m], C11, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(res["C"][m:], C21, rtol=1e-8, atol=1e-10)
    assert np.mean((res["posterior"]["mean"] > 0) == (lab[m:] > 0.5)) > 0.9   # README: error rate 0.027 after training


@pytest.mark.gpu
def test_fit_lae_logit_mult_three_classes(flgp, oracle):
    """fit_lae_logit_mult_gp_rcpp's deterministic half (src/Fit.cpp:603-662, src/MultiClassification.cpp:30-53): three
    concentric rings as three classes, one-vs-rest training of the diffusion time per class against the oracle twin
    on the library's own eigenvectors (t to optimiser tolerance, objective to 1e-6), and the arg-max of the per-class
    Laplace posterior means classifies the held-out rows."""
    rng = np.random.default_rng(11)
    n, m, s, r, K, sigma = 3000, 150, 300, 3, 60, 1e-3
    lab = rng.integers(0, 3, size=n)
    lab[:3] = [0, 1, 2]
    ang = rng.uniform(0, 2 * np.pi, size=n)
    rad = 1.0 + lab + 0.08 * rng.standard_normal(n)
    X = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1)
    init = _init(n, s, 1)
    ep = flgp.heat_kernel_spectrum_cpp(X[:m], X[m:], s, r, K, init_idx=init, iter_max=50)
    t_l, obj_l = flgp.train_logit_mult_gp(ep, lab[:m].astype(np.float64), m, K, sigma, "posterior")
    t_o, obj_o = oracle.train_logit_mult(ep.vectors, ep.values, lab[:m], np.arange(m, dtype=np.int32), K, sigma, "posterior")
    assert len(t_l) == 3
    np.testing.assert_allclose(t_l, t_o, rtol=1e-3)
    np.testing.assert_allclose(obj_l, obj_o, rtol=1e-6, atol=1e-6)
    res = flgp.fit_lae_logit_mult_gp_rcpp(X[:m], lab[:m], X[m:], s, r, K, sigma=sigma, init_idx=init, iter_max=50)
    np.testing.assert_allclose(res["pars"], t_l, rtol=1e-12)
    assert np.mean(res["argmax_posterior_mean"] == lab[m:]) > 0.95
    with pytest.raises(flgp.FlgpError):
        flgp.train_logit_mult_gp(ep, np.array([0.5] * m), m, K, sigma, "posterior")


def _noisy_rings(n, n_rings, n_classes, seed):
    """Concentric rings, class = ring number modulo n_classes, 12 % of the labels moved to another class: the trained
    diffusion time is then an interior optimum (on cleanly separated classes t runs to the evaluation cap)."""
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0, 2 * np.pi, n)
    ring = rng.integers(0, n_rings, n)
    rad = 0.5 + 0.25 * ring + 0.05 * rng.standard_normal(n)
    X = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1)
    lab = ring % n_classes
    flip = rng.uniform(size=n) < 0.12
    lab = np.where(flip, (lab + rng.integers(1, n_classes, n)) % n_classes, lab)
    lab[:n_classes] = np.arange(n_classes)
    return np.asfortranarray(X), lab.astype(np.float64)


def test_fit_se_logit_config1_grid(flgp, oracle):
    """fit_se_logit_gp_rcpp (src/Fit.cpp:668-794; the call of the README's GPC example) against the oracle twin running
    its own pipeline (k-means, KNN, SE weights, graph Laplacian, LAPACK eigenvectors).  BASELINE config 1's rings
    (n=4800, d=2, m=100, s=600, r=3, K=100), default ten bandwidths, at a fixed t — the whole path is deterministic:
    same winning a2, objective to 1e-7, Laplace posterior to 1e-6.  Trained (COBYLA restatement per grid point) on
    rings with noisy labels: same a2, t to optimiser tolerance."""
    from flgp_b200.datasets import make

    X, lab, cfg = make("C1")
    m, s, r, K, sigma = cfg["m"], cfg["s"], cfg["r"], cfg["K"], 1e-3
    init = _init(len(X), s, 1)
    a2s = np.exp(np.linspace(np.log(0.1), np.log(10.0), 10))
    res = flgp.fit_se_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, r, K, sigma=sigma, t=8.0, init_idx=init, iter_max=50,
                                    output_cov=True)
    ref = oracle.fit_se_logit(X[:m], lab[:m], X[m:], s, r, K, init, a2s, sigma=sigma, iter_max=50, nthreads=NT, t=8.0)
    assert res["a2"] == ref["a2"] and res["pars"] == 8.0
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-7)
    np.testing.assert_allclose(res["eigenpair"].values, ref["values"], rtol=1e-8, atol=1e-10)
    sc = max(1.0, np.abs(ref["mean"]).max())
    np.testing.assert_allclose(res["posterior"]["mean"], ref["mean"], rtol=1e-6, atol=1e-7 * sc)
    np.testing.assert_allclose(res["posterior"]["cov"], ref["cov"], rtol=1e-6, atol=1e-7 * max(1.0, np.abs(ref["cov"]).max()))
    np.testing.assert_allclose(res["C"], ref["C"], rtol=1e-7, atol=1e-9)
    assert np.mean((res["posterior"]["mean"] > 0) == (lab[m:] > 0.5)) > 0.75   # t = 8 is not the trained optimum
    with pytest.raises(flgp.FlgpError, match="not supported"):
        flgp.fit_se_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, r, K, approach="evidence", init_idx=init)
    # trained
    X, lab = _noisy_rings(2400, 4, 2, 5)
    m, s, r, K = 120, 240, 3, 60
    init = _init(len(X), s, 1)
    a2s = np.array([0.1, 0.5, 2.0, 10.0])
    for approach in ("posterior", "marginal"):
        res = flgp.fit_se_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, r, K, sigma=sigma, a2s=a2s, approach=approach,
                                        init_idx=init, iter_max=50)
        ref = oracle.fit_se_logit(X[:m], lab[:m], X[m:], s, r, K, init, a2s, sigma=sigma, approach=approach, iter_max=50,
                                  nthreads=NT)
        assert res["a2"] == ref["a2"]
        assert abs(res["pars"] - ref["t"]) <= 2e-3 * max(1.0, ref["t"])
        np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-6, atol=1e-6)
        sc = max(1.0, np.abs(ref["mean"]).max())
        np.testing.assert_allclose(res["posterior"]["mean"], ref["mean"], rtol=1e-2, atol=1e-2 * sc)


def test_fit_se_logit_mult_three_classes_grid(flgp, oracle):
    """fit_se_logit_mult_gp_rcpp's deterministic half (src/Fit.cpp:797-895): bandwidth grid with the J one-vs-rest
    trainings per grid point, the summed objective selects a2; against the oracle twin's own pipeline."""
    X, lab = _noisy_rings(2400, 6, 3, 7)
    m, s, r, K, sigma = 150, 240, 3, 60, 1e-3
    init = _init(len(X), s, 1)
    a2s = np.array([0.1, 0.5, 2.0, 10.0])
    res = flgp.fit_se_logit_mult_gp_rcpp(X[:m], lab[:m], X[m:], s, r, K, sigma=sigma, a2s=a2s, init_idx=init,
                                         iter_max=50)
    ref = oracle.fit_se_logit_mult(X[:m], lab[:m], X[m:], s, r, K, init, a2s, sigma=sigma, iter_max=50, nthreads=NT)
    assert res["a2"] == ref["a2"] and len(res["pars"]) == 3
    np.testing.assert_allclose(res["pars"], ref["t"], rtol=2e-3)
    np.testing.assert_allclose(res["obj_classes"], ref["objs"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(res["eigenpair"].values, ref["values"], rtol=1e-8, atol=1e-10)
    assert res["posterior_mean"].shape == (len(X) - m, 3)
    assert np.mean(res["argmax_posterior_mean"] == lab[m:]) > 0.55   # 12 % of the held-out labels are noise (chance: 0.33)


def test_fit_nystrom_logit_grid(flgp, oracle):
    """fit_nystrom_logit_gp_rcpp (src/Fit.cpp:896-1038) against the oracle's dense literal restatement on rings with
    noisy labels: at a fixed t the same winning bandwidth, objective to 1e-6, Laplace posterior and the covariance block
    to 1e-5 (the dense s x s normalisation and the n x s x K extension accumulate in different orders, as for the
    regression driver); trained: same bandwidth, t to optimiser tolerance."""
    X, lab = _noisy_rings(2400, 4, 2, 5)
    m, s, K, sigma = 120, 240, 60, 1e-3
    init = _init(len(X), s, 1)
    a2s = np.array([0.1, 0.5, 2.0, 10.0])
    res = flgp.fit_nystrom_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, K, sigma=sigma, a2s=a2s, t=8.0, init_idx=init,
                                         iter_max=50, output_cov=True)
    ref = oracle.fit_nystrom_logit(X[:m], lab[:m], X[m:], s, K, init, a2s, sigma=sigma, iter_max=50, nthreads=NT, t=8.0)
    assert res["a2"] == ref["a2"] and res["pars"] == 8.0
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-6)
    sc = max(1.0, np.abs(ref["mean"]).max())
    np.testing.assert_allclose(res["posterior"]["mean"], ref["mean"], rtol=1e-5, atol=1e-6 * sc)
    np.testing.assert_allclose(res["posterior"]["cov"], ref["cov"], rtol=1e-5, atol=1e-6 * max(1.0, np.abs(ref["cov"]).max()))
    np.testing.assert_allclose(res["C"], ref["C"], rtol=1e-5, atol=1e-6 * max(1.0, np.abs(ref["C"]).max()))
    res = flgp.fit_nystrom_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, K, sigma=sigma, a2s=a2s, init_idx=init, iter_max=50)
    ref = oracle.fit_nystrom_logit(X[:m], lab[:m], X[m:], s, K, init, a2s, sigma=sigma, iter_max=50, nthreads=NT)
    assert res["a2"] == ref["a2"]
    assert abs(res["pars"] - ref["t"]) <= 2e-3 * max(1.0, ref["t"])
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-6, atol=1e-6)
    assert np.mean((res["posterior"]["mean"] > 0) == (lab[m:] > 0.5)) > 0.6
    with pytest.raises(flgp.FlgpError, match="not supported"):
        flgp.fit_nystrom_logit_gp_rcpp(X[:m], lab[:m], X[m:], s, K, approach="evidence", init_idx=init)


def test_fit_nystrom_logit_mult_grid(flgp, oracle):
    """The training half of fit_nystrom_logit_mult_gp_rcpp (src/Fit.cpp:1045-1162): per bandwidth the J one-vs-rest
    trainings on the extended labelled rows, the summed objective selects; the winning extended eigenpair is returned
    for the reference's sampler (eigenvalues 1e-8; the eigenvectors reproduce the oracle's heat kernel on the labelled
    rows)."""
    X, lab = _noisy_rings(2400, 6, 3, 7)
    m, s, K, sigma = 150, 240, 60, 1e-3
    init = _init(len(X), s, 1)
    a2s = np.array([0.1, 0.5, 2.0, 10.0])
    res = flgp.fit_nystrom_logit_mult_gp_rcpp(X[:m], lab[:m], X[m:], s, K, sigma=sigma, a2s=a2s, init_idx=init,
                                              iter_max=50, return_eigenpair=True)
    ref = oracle.fit_nystrom_logit_mult(X[:m], lab[:m], X[m:], s, K, init, a2s, sigma=sigma, iter_max=50, nthreads=NT)
    assert res["a2"] == ref["a2"] and len(res["pars"]) == 3
    np.testing.assert_allclose(res["pars"], ref["t"], rtol=2e-3)
    np.testing.assert_allclose(res["obj_classes"], ref["objs"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(res["obj"], ref["obj"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(res["values"], ref["values"], rtol=1e-8, atol=1e-12)
    idx0 = np.arange(m, dtype=np.int32)
    Hl = oracle.hk_from_spectrum(res["vectors"], res["values"], K, 5.0, idx0, idx0)
    Ho = oracle.hk_from_spectrum(ref["V"], ref["values"], K, 5.0, idx0, idx0)
    np.testing.assert_allclose(Hl, Ho, rtol=1e-5, atol=1e-6 * np.abs(Ho).max())


@pytest.mark.gpu
@pytest.mark.parametrize("nbytes", [(8 << 20) - 8, 8 << 20, (8 << 20) + 8, (37 << 20) + 4088, 200 << 20])
def test_staged_pageable_copies_roundtrip(flgp, nbytes):
    """csrc/hostcopy.cu: copies from / to pageable memory of 8 MB and more go through per-thread pinned bounce buffers
    (partial last chunk, fewer chunks than threads, many chunks per thread); pinned memory takes the plain path.
    Bit-exact round trip either way, and repeated calls reuse the bounce buffers."""
    import torch

    ctx = flgp.default_ctx()
    rng = np.random.default_rng(nbytes % 1000)
    src = rng.integers(0, 2 ** 63 - 1, size=nbytes // 8, dtype=np.int64)
    for _ in range(2):
        assert np.array_equal(ctx.copy_roundtrip(src), src)
    pinned = torch.from_numpy(src[: (9 << 20) // 8].copy()).pin_memory().numpy()
    assert np.array_equal(ctx.copy_roundtrip(pinned), pinned)
