"""ctypes binding + numpy glue for the CPU ORACLE (oracle/flgp_oracle.cpp).

TEST INFRASTRUCTURE ONLY — parity unpinned (see the header of flgp_oracle.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (flgp_b200) never does.

Stages that the reference delegates to un-vendored third-party code are restated here:
  * RSpectra::svds (src/TruncatedSVD.cpp:23-28)  -> scipy.linalg.eigh of the Gram A^T A
  * Eigen::LLT (src/Predict.cpp:57,66; src/Utils.cpp:234,242) -> scipy cho_factor/cho_solve
  * nloptr / NLopt (LD_MMA, LN_COBYLA; src/train.cpp:38-71, 557-671) -> mma_minimize, cobyla_minimize_1d below
  * ClusterR::MiniBatchKmeans (src/Utils.cpp:49-62) -> the mini-batch contract of flgp_oracle.cpp
All citations are relative to /root/reference.

What pins this oracle in the absence of reference-held vectors (tests/test_oracle_cpu.py): numpy twins of every stage,
the committed fixtures of tests/golden/, and other people's implementations of the same mathematics — scikit-learn's
Lloyd, brute-force neighbours, GaussianProcessRegressor / GaussianProcessClassifier on the precomputed heat kernel,
MiniBatchKMeans; scipy's ARPACK svds, SLSQP, L-BFGS-B, COBYLA; LAPACK.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GL_MODES = {"rw": 0, "normalized": 1, "cluster-normalized": 2}


def build(force: bool = False) -> str:
    """Compile the C++ restatement with the committed Makefile (g++ -O2)."""
    so = os.path.join(_HERE, "libflgp_oracle.so")
    src = os.path.join(_HERE, "flgp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libflgp_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_maxabs.restype = C.c_double
    return _LIB


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


I64 = C.c_int64


# ------------------------------------------------------------------ fixed point (tests)
def fx_encode(x, maxabs, count):
    x = np.ascontiguousarray(x, dtype=np.float64)
    hi = np.zeros(x.shape, np.int64)
    lo = np.zeros(x.shape, np.int64)
    rc = lib().orc_fx_encode(C.c_double(maxabs), I64(count), _p(x), I64(x.size), _p(hi, I64), _p(lo, I64))
    assert rc == 0
    return hi, lo


def fx_decode(hi, lo, maxabs, count):
    hi = np.ascontiguousarray(hi, dtype=np.int64)
    lo = np.ascontiguousarray(lo, dtype=np.int64)
    x = np.zeros(hi.shape, np.float64)
    rc = lib().orc_fx_decode(C.c_double(maxabs), I64(count), _p(hi, I64), _p(lo, I64), I64(hi.size), _p(x))
    assert rc == 0
    return x


# ------------------------------------------------------------------ k-means
def kmeans_acc_words(s, d):
    return 2 * s * d + s + 1


def kmeans_step(X, C_, maxabs, n_total, assign, acc, nthreads=1):
    """One Lloyd assign+accumulate pass over a shard; adds into acc (int64)."""
    X = _f(X)
    C_ = _f(C_)
    n, d = X.shape
    s = C_.shape[0]
    rc = lib().orc_kmeans_step(_p(X), I64(n), I64(n), d, _p(C_), s, C.c_double(maxabs), I64(n_total),
                               _p(assign, C.c_int32), _p(acc, I64), nthreads)
    assert rc == 0


def kmeans_update(acc, C_, maxabs, n_total):
    """Centroid update from (all-reduced) accumulators; returns (C_new, sizes)."""
    C_ = _f(C_).copy(order="F")
    s, d = C_.shape
    sizes = np.zeros(s)
    rc = lib().orc_kmeans_update(_p(acc, I64), s, d, C.c_double(maxabs), I64(n_total), _p(C_), _p(sizes))
    assert rc == 0
    return C_, sizes


def kmeans_lloyd(X, s, init_idx, iter_max=100, nthreads=1):
    """subsample_cpp(method="kmeans") contract (src/Utils.cpp:32-45): U = [centres, size]."""
    X = _f(X)
    n, d = X.shape
    init_idx = _i32(init_idx)
    U = np.zeros((s, d + 1), order="F")
    assign = np.zeros(n, np.int32)
    iters = C.c_int(0)
    rc = lib().orc_kmeans_lloyd(_p(X), I64(n), I64(n), d, s, _p(init_idx, C.c_int32), iter_max, nthreads,
                                _p(U), _p(assign, C.c_int32), C.byref(iters))
    if rc:
        raise ValueError("orc_kmeans_lloyd failed")
    return U, assign, iters.value


def mb_perm(k, n, key):
    """The batch sampler of the mini-batch contract: value k of the keyed bijection of [0, n)."""
    lib().orc_mb_perm.restype = C.c_int64
    return int(lib().orc_mb_perm(I64(k), I64(n), C.c_uint64(key)))


def mb_batch_key(seed, it):
    lib().orc_mb_batch_key.restype = C.c_uint64
    return int(lib().orc_mb_batch_key(C.c_uint64(seed), it))


def minibatch_kmeans(X, s, init_idx, max_iters=100, seed=0, nthreads=1, early_stop_iter=10, tol=1e-4,
                     want_batches=False):
    """subsample_cpp(method="minibatchkmeans") contract (src/Utils.cpp:49-62): centroids by Sculley's mini-batch
    k-means (ClusterR::MiniBatchKmeans is un-vendored: parity unpinned, contract in flgp_oracle.cpp), then the
    reference's own lines :57-62 — labels = KNN_cpp(X, centroids, 1), U[:, d] = rows per label."""
    X = _f(X)
    n, d = X.shape
    init_idx = _i32(init_idx)
    Cc = np.zeros((s, d), order="F")
    iters = C.c_int(0)
    b = min(10 * s, n)
    rows = np.zeros((max_iters, b), np.int64) if want_batches else None
    rc = lib().orc_minibatch_kmeans(_p(X), I64(n), I64(n), d, s, _p(init_idx, C.c_int32), max_iters, C.c_uint64(seed),
                                    early_stop_iter, C.c_double(tol), nthreads, _p(Cc), C.byref(iters),
                                    _p(rows, C.c_int64) if want_batches else None)
    if rc:
        raise ValueError("orc_minibatch_kmeans failed")
    labels = knn(X, Cc, 1, nthreads=nthreads)[:, 0]
    U = np.zeros((s, d + 1), order="F")
    U[:, :d] = Cc
    U[:, d] = np.bincount(labels, minlength=s)
    if want_batches:
        return U, iters.value, rows[:iters.value]
    return U, iters.value


# ------------------------------------------------------------------ KNN / LAE
def knn(X, U, r, want_dist=False, nthreads=1):
    """KNN_cpp (src/Utils.cpp:102-192): ind_knn n x r (0-based, ascending distance)."""
    X = _f(X)
    U = _f(U)
    n, d = X.shape
    s = U.shape[0]
    assert U.shape[1] == d
    ind = np.zeros((n, r), np.int32, order="F")
    dist = np.zeros((n, r), order="F") if want_dist else None
    rc = lib().orc_knn(_p(X), I64(n), I64(n), d, _p(U), s, I64(s), r, _p(ind, C.c_int32),
                       _p(dist) if want_dist else None, nthreads)
    if rc:
        raise ValueError("orc_knn failed (need 1 <= r <= s)")
    return (ind, dist) if want_dist else ind


def simplex_project(v):
    """v_to_z_cpp (src/lae.cpp:137-153)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    z = np.zeros_like(v)
    lib().orc_simplex_project(_p(v), v.size, _p(z))
    return z


def lae_point(x, Ur, want_stats=False):
    """local_anchor_embedding_cpp (src/lae.cpp:76-133); Ur is r x d."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    Ur = _f(Ur)
    r, d = Ur.shape
    z = np.zeros(r)
    it = C.c_int(0)
    bt = C.c_int(0)
    lib().orc_lae_point(_p(x), d, _p(Ur), I64(r), r, _p(z), C.byref(it), C.byref(bt))
    return (z, it.value, bt.value) if want_stats else z


def lae(X, U, r, ind=None, nthreads=1, want_stats=False):
    """LAE_cpp (src/lae.cpp:48-70): returns CSR (Zj, Zx) as (n, r) arrays, rows column-sorted,
    and the dense n x r weights in KNN order."""
    X = _f(X)
    U = _f(U)
    n, d = X.shape
    s = U.shape[0]
    if ind is None:
        ind = knn(X, U, r, nthreads=nthreads)
    ind = np.asfortranarray(ind, dtype=np.int32)
    Zj = np.zeros((n, r), np.int32)
    Zx = np.zeros((n, r))
    W = np.zeros((n, r), order="F")
    stats = np.zeros(2, np.int64)
    rc = lib().orc_lae(_p(X), I64(n), I64(n), d, _p(U), s, I64(s), r, _p(ind, C.c_int32), _p(Zj, C.c_int32),
                       _p(Zx), _p(W), _p(stats, I64), nthreads)
    assert rc == 0
    if want_stats:
        return Zj, Zx, W, stats
    return Zj, Zx, W


def knn_csr(ind, dist):
    """distances_sp of KNN_cpp(output=true) (src/Utils.cpp:145-189) as fixed-r CSR."""
    ind = np.asfortranarray(ind, dtype=np.int32)
    dist = _f(dist)
    n, r = ind.shape
    Zj = np.zeros((n, r), np.int32)
    Zx = np.zeros((n, r))
    lib().orc_knn_to_csr(I64(n), r, _p(ind, C.c_int32), _p(dist), _p(Zj, C.c_int32), _p(Zx))
    return Zj, Zx


def se_weights(dist, denom):
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    out = np.zeros_like(dist)
    lib().orc_se_weights(_p(dist), I64(dist.size), C.c_double(denom), _p(out))
    return out


# ------------------------------------------------------------------ GL, spectrum
def colsum(Zj, Zx, s, exact=1, n_total=None, limbs=None):
    Zj = _i32(Zj)
    Zx = np.ascontiguousarray(Zx, dtype=np.float64)
    n, r = Zj.shape
    c = np.zeros(s)
    hi = lo = None
    if limbs is not None:
        hi, lo = limbs
    rc = lib().orc_colsum(I64(n), s, r, _p(Zj, C.c_int32), _p(Zx), exact, I64(n_total or n), _p(c),
                          _p(hi, I64) if hi is not None else None, _p(lo, I64) if lo is not None else None)
    assert rc == 0
    return c


def graph_laplacian_apply(Zj, Zx, s, mode, colsum_, num_class=None):
    Zj = _i32(Zj)
    Zx = np.ascontiguousarray(Zx, dtype=np.float64).copy()
    n, r = Zj.shape
    m = GL_MODES[mode] if isinstance(mode, str) else mode
    nc = np.ascontiguousarray(num_class, dtype=np.float64) if num_class is not None else np.ones(s)
    cs = np.ascontiguousarray(colsum_, dtype=np.float64) if colsum_ is not None else np.zeros(s)
    rc = lib().orc_graph_laplacian_apply(I64(n), s, r, _p(Zj, C.c_int32), _p(Zx), m, _p(cs), _p(nc))
    if rc:
        raise ValueError("Error: the type of graph Laplacian is not supported!")
    return Zx


def graph_laplacian(Zj, Zx, s, mode, num_class=None, exact=1):
    """graphLaplacian_cpp (src/Utils.cpp:195-212); returns the scaled values."""
    if mode not in GL_MODES:
        raise ValueError("Error: the type of graph Laplacian is not supported!")
    if mode == "cluster-normalized" and num_class is None:
        raise ValueError("cluster-normalized needs cluster sizes")
    c = colsum(Zj, Zx, s, exact) if GL_MODES[mode] >= 1 else None
    return graph_laplacian_apply(Zj, Zx, s, mode, c, num_class)


def spectrum_scale(c):
    c = np.ascontiguousarray(c, dtype=np.float64)
    w = np.zeros_like(c)
    lib().orc_spectrum_scale(c.size, _p(c), _p(w))
    return w


def gram(Zj, Zx, w, s, exact=1, n_total=None, limbs=None):
    Zj = _i32(Zj)
    Zx = np.ascontiguousarray(Zx, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    n, r = Zj.shape
    G = np.zeros((s, s), order="F")
    hi = lo = None
    if limbs is not None:
        hi, lo = limbs
    rc = lib().orc_gram(I64(n), s, r, _p(Zj, C.c_int32), _p(Zx), _p(w), exact, I64(n_total or n), _p(G),
                        _p(hi, I64) if hi is not None else None, _p(lo, I64) if lo is not None else None)
    assert rc == 0
    return G


def gram_eigh(G, K):
    """Top-K eigenpairs of the symmetric Gram, descending (replaces RSpectra::svds on A)."""
    import scipy.linalg as sla

    s = G.shape[0]
    lam, Y = sla.eigh(G, subset_by_index=[s - K, s - 1])
    return lam[::-1].copy(), np.asfortranarray(Y[:, ::-1])


def lift(Zj, Zx, w, Y, sigma, n_total=None, nthreads=1):
    Zj = _i32(Zj)
    Zx = np.ascontiguousarray(Zx, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    Y = _f(Y)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    n, r = Zj.shape
    s, K = Y.shape
    V = np.zeros((n, K), order="F")
    lib().orc_lift(I64(n), s, r, _p(Zj, C.c_int32), _p(Zx), _p(w), _p(Y), _p(sigma), K, I64(n_total or n),
                   _p(V), nthreads)
    return V


def spectrum_from_Z(Zj, Zx, s, K, root=True, exact=1, nthreads=1, want_internals=False):
    """spectrum_from_Z_cpp (src/Spectrum.cpp:146-161) through the Gram route.
    values: sigma (root) or sigma^2, descending; vectors: sqrt(n) * U (n x K)."""
    if K < 0:
        K = s
    c2 = colsum(Zj, Zx, s, exact)
    w = spectrum_scale(c2)
    G = gram(Zj, Zx, w, s, exact)
    lam, Y = gram_eigh(G, K)
    lam = np.maximum(lam, 0.0)
    sigma = np.sqrt(lam)
    V = lift(Zj, Zx, w, Y, sigma, nthreads=nthreads)
    values = sigma if root else lam
    if want_internals:
        return values, V, dict(w=w, G=G, Y=Y, lam=lam, c2=c2)
    return values, V


def hk_from_spectrum(V, values, K, t, idx0, idx1, nthreads=1):
    """HK_from_spectrum_cpp (src/Spectrum.cpp:83-94)."""
    V = _f(V)
    values = np.ascontiguousarray(values, dtype=np.float64)
    idx0 = _i32(idx0)
    idx1 = _i32(idx1)
    H = np.zeros((idx0.size, idx1.size), order="F")
    lib().orc_hk_from_spectrum(_p(V), I64(V.shape[0]), _p(values), K, C.c_double(t), _p(idx0, C.c_int32),
                               I64(idx0.size), _p(idx1, C.c_int32), I64(idx1.size), _p(H), nthreads)
    return H


# ------------------------------------------------------------------ orchestrators
def cross_similarity_lae(X, U, r, gl, exact=1, nthreads=1):
    """cross_similarity_lae_cpp (src/Spectrum.cpp:101-117); U is s x (d+1) for cluster-normalized."""
    X = _f(X)
    d = X.shape[1]
    U = _f(U)
    if gl == "cluster-normalized" and U.shape[1] < d + 1:
        raise ValueError("cluster-normalized needs the cluster-size column of U")
    Zj, Zx, _ = lae(X, U[:, :d], r, nthreads=nthreads)
    nc = U[:, d] if gl == "cluster-normalized" else None
    return Zj, graph_laplacian(Zj, Zx, U.shape[0], gl, nc, exact)


def cross_similarity_se(X, U, r, gl, epsilon=0.1, exact=1, nthreads=1):
    """cross_similarity_se_cpp (src/Spectrum.cpp:120-142)."""
    X = _f(X)
    d = X.shape[1]
    U = _f(U)
    ind, dist = knn(X, U[:, :d], r, want_dist=True, nthreads=nthreads)
    Zj, Zd = knn_csr(ind, dist)
    Zx = se_weights(Zd, 4.0 * epsilon * epsilon)
    nc = U[:, d] if gl == "cluster-normalized" else None
    return Zj, graph_laplacian(Zj, Zx, U.shape[0], gl, nc, exact)


def heat_kernel_spectrum(X, X_new, s, r, K, init_idx, kernel="lae", gl="cluster-normalized", root=True,
                         epsilon=0.1, iter_max=100, exact=1, nthreads=1, want_internals=False):
    """heat_kernel_spectrum_cpp (src/Spectrum.cpp:48-76) with subsample="kmeans" (Lloyd contract)."""
    X_all = np.asfortranarray(np.vstack([X, X_new])) if X_new is not None and len(X_new) else _f(X)
    U, assign, iters = kmeans_lloyd(X_all, s, init_idx, iter_max, nthreads)
    if kernel == "lae":
        Zj, Zx = cross_similarity_lae(X_all, U, r, gl, exact, nthreads)
    elif kernel == "se":
        Zj, Zx = cross_similarity_se(X_all, U, r, gl, epsilon, exact, nthreads)
    else:
        raise ValueError("The kernel type is not supported!")
    if K < 0:
        K = s
    out = spectrum_from_Z(Zj, Zx, s, K, root, exact, nthreads, want_internals)
    if want_internals:
        out[2].update(U=U, assign=assign, iters=iters, Zj=Zj, Zx=Zx)
    return out


def lae_eigenmap(X, s, r, ndim, init_idx, norm="cluster-normalized", **kw):
    """lae_eigenmap (src/Spectrum.cpp:17-25): eigenvalues = 1 - values (root=true)."""
    values, V = heat_kernel_spectrum(X, None, s, r, ndim, init_idx, kernel="lae", gl=norm, root=True, **kw)
    return 1.0 - values, V


# ------------------------------------------------------------------ GPR tail (fixed hyper-parameters)
def predict_regression(V, values, Y, idx0, idx1, K, pars, sigma):
    """predict_regression_cpp, noise="same" (src/Predict.cpp:40-75)."""
    import scipy.linalg as sla

    t, noise = pars[0], pars[1]
    Y = np.asarray(Y, dtype=np.float64).reshape(-1)
    m = Y.size
    ev = 1.0 - values[:K]
    if m <= K:
        Cvv = hk_from_spectrum(V, values, K, t, idx0, idx0)
        Cn = Cvv.copy()
        Cn[np.diag_indices(m)] += sigma
        Cn[np.diag_indices(m)] += noise
        Cnv = hk_from_spectrum(V, values, K, t, idx1, idx0)
        alpha = sla.cho_solve(sla.cho_factor(Cn, lower=True), Y)
        return Cnv @ alpha
    V1 = V[np.asarray(idx0), :K]
    ls = np.exp(-0.5 * t * ev) + 0.0
    Q = (ls[:, None] * (V1.T @ V1)) * ls[None, :]
    Q[np.diag_indices(K)] += noise + sigma
    cf = sla.cho_factor(Q, lower=True)
    alpha = 1.0 / (noise + sigma) * (Y - V1 @ (ls * sla.cho_solve(cf, ls * (V1.T @ Y))))
    Vnv = V[np.asarray(idx1), :K]
    return Vnv @ (np.exp(-t * ev + 0.0) * (V1.T @ alpha))


def posterior_covariance_regression(V, values, idx0, idx1, K, pars, sigma):
    """posterior_covariance_regression (src/Utils.cpp:215-249)."""
    import scipy.linalg as sla

    m = len(idx0)
    t, var = pars[0], pars[1]
    ev = 1.0 - values[:K]
    V2 = V[np.asarray(idx1), :K]
    lam = np.exp(-t * ev)
    if m <= K:
        C11 = hk_from_spectrum(V, values, K, t, idx0, idx0)
        K11 = C11.copy()
        K11[np.diag_indices(m)] += var + sigma
        C21 = hk_from_spectrum(V, values, K, t, idx1, idx0)
        alpha = C21 @ sla.cho_solve(sla.cho_factor(K11, lower=True), np.eye(m))
        beta = (C21 * alpha).sum(axis=1)
    else:
        V1 = V[np.asarray(idx0), :K]
        ls = np.exp(-0.5 * t * ev) + 0.0
        G1 = V1.T @ V1
        Q = (ls[:, None] * G1) * ls[None, :]
        Q[np.diag_indices(K)] += var + sigma
        cf = sla.cho_factor(Q, lower=True)
        inner = V1 - V1 @ (ls[:, None] * sla.cho_solve(cf, ls[:, None] * G1))
        alpha = 1.0 / (var + sigma) * (lam[:, None] * (V1.T @ inner)) * lam[None, :]
        beta = (V2 * (V2 @ alpha)).sum(axis=1)
    return ((V2 * lam[None, :]) * V2).sum(axis=1) + var + sigma - beta


def fit_lae_regression_fixed(X, Y, X_new, s, r, K, pars, init_idx, sigma=1e-5, gl="cluster-normalized",
                             root=True, iter_max=100, nthreads=1):
    """fit_lae_regression_gp_cpp (src/Fit.cpp:20-99) with pars=(t, noise) supplied instead of optimised."""
    m = len(X)
    n = m + len(X_new)
    if K < 0:
        K = s
    values, V = heat_kernel_spectrum(X, X_new, s, r, K, init_idx, "lae", gl, root, iter_max=iter_max,
                                     nthreads=nthreads)
    idx0 = np.arange(m, dtype=np.int32)
    idx1 = np.arange(m, n, dtype=np.int32)
    train = predict_regression(V, values, Y, idx0, idx0, K, pars, sigma)
    test = predict_regression(V, values, Y, idx0, idx1, K, pars, sigma)
    cov = posterior_covariance_regression(V, values, idx0, idx1, K, pars, sigma)
    return dict(train=train, test=test, cov=cov, values=values, vectors=V)


# ------------------------------------------------------------------ hyper-parameter training (regression)
def regression_objective(V, values, Y, idx, K, x, sigma=1e-5, approach="marginal", prior=(1.0, 10.0, 2.0, 0.1, 1e-3)):
    """negative_marginal_likelihood_regression_cpp / negative_log_posterior_regression_cpp, noise="same"
    (src/train.cpp:333-436), on the materialised eigenvectors V.  Returns (objective, grad[2]) with the reference's
    gradient clipping of grad[1] to +-10.  prior = (p, q, tau, alpha, beta) of PostOFDataReg (src/train.h:144-156)."""
    import scipy.linalg as sla

    t, noise = float(x[0]), float(x[1])
    Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
    m, q = Y.shape
    ev = 1.0 - values[:K]
    Vm = V[np.asarray(idx), :K]
    grad = np.zeros(2)
    if m <= K:
        C = (Vm * np.exp(-t * ev)) @ Vm.T
        C[np.diag_indices(m)] += sigma
        C[np.diag_indices(m)] += noise
        L = sla.cholesky(C, lower=True)
        alpha = sla.cho_solve((L, True), Y)
        C_inv = sla.cho_solve((L, True), np.eye(m))
        U = alpha @ alpha.T / q - C_inv
        grad_t = (Vm * (-ev * np.exp(-t * ev))) @ Vm.T
        grad[0] = -0.5 * (U * grad_t.T).sum()
        grad[1] = -0.5 * np.trace(U)
        nmll = 0.5 * (Y * alpha).sum() / q + np.log(np.diag(L) + 1e-9).sum()
    else:
        ls = np.exp(-0.5 * t * ev) + 0.0
        VtV = Vm.T @ Vm
        Q = (ls[:, None] * VtV) * ls[None, :]
        Q[np.diag_indices(K)] += noise + sigma
        L = sla.cholesky(Q, lower=True)
        ns = noise + sigma
        alpha = 1.0 / ns * (Y - (Vm * ls) @ sla.cho_solve((L, True), ls[:, None] * (Vm.T @ Y)))
        Q_inv = sla.cho_solve((L, True), np.eye(K))
        A = -ev * (np.exp(-t * ev) + 0.0) + 0.0
        Vta = Vm.T @ alpha
        grad[0] = -0.5 * (Vta * (A[:, None] * Vta)).sum() / q
        grad[0] += 0.5 / ns * np.trace(A[:, None] * VtV)
        grad[0] += -0.5 / ns * ((Q_inv @ (ls[:, None] * VtV)) * ((A[:, None] * VtV) * ls[None, :]).T).sum()
        grad[1] = -0.5 * (alpha * alpha).sum() / q
        grad[1] += 0.5 / ns * (m - (Q_inv * ((ls[:, None] * VtV) * ls[None, :]).T).sum())
        nmll = 0.5 * (Y * alpha).sum() / q + np.log(np.diag(L) + 1e-9).sum() + 0.5 * (m - K) * np.log(ns)
    if abs(grad[1]) >= 10.0:
        grad[1] = grad[1] / abs(grad[1]) * 10.0
    if approach == "posterior":
        p, qq, tau, al, be = prior
        nmll += p * np.log(t + 1e-9) + (t / tau) ** (-qq)
        nmll += (al + 1) * np.log(noise + sigma) + be / (noise + sigma)
        grad[0] += p / (t + 1e-9) - (qq / tau) * (t / tau) ** (-qq - 1)
        grad[1] += (al + 1) / (noise + sigma) - be / (noise + sigma) ** 2
    elif approach != "marginal":
        raise ValueError("This model selection approach is not supported!")
    return float(nmll), grad


def regression_objective_diff(V, values, Y, idx, K, x, sigma=1e-5, approach="marginal",
                              prior=(1.0, 10.0, 2.0, 0.1, 1e-3)):
    """negative_marginal_likelihood_diff_noise_regression_cpp / negative_log_posterior_diff_noise_regression_cpp
    (src/train.cpp:438-556), literal: x = (t, noise_1 .. noise_m).  Returns (objective, grad[m + 1]); the noise
    gradients of the m > K branch are clipped to [-1, 1] as in the reference (:536-541)."""
    import scipy.linalg as sla

    x = np.asarray(x, dtype=np.float64)
    t = float(x[0])
    Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
    m, q = Y.shape
    ev = 1.0 - values[:K]
    Vm = V[np.asarray(idx), :K]
    grad = np.zeros(m + 1)
    if m <= K:
        C = (Vm * np.exp(-t * ev)) @ Vm.T
        C[np.diag_indices(m)] += sigma
        C[np.diag_indices(m)] += x[1:]
        L = sla.cholesky(C, lower=True)
        alpha = sla.cho_solve((L, True), Y)
        C_inv = sla.cho_solve((L, True), np.eye(m))
        U = alpha @ alpha.T / q - C_inv
        grad_t = (Vm * (-ev * np.exp(-t * ev))) @ Vm.T
        grad[0] = -0.5 * (U * grad_t.T).sum()
        grad[1:] = -0.5 * np.diag(U)
        nmll = 0.5 * (Y * alpha).sum() / q + np.log(np.diag(L) + 1e-9).sum()
    else:
        ls = np.exp(-0.5 * t * ev) + 0.0
        Z = x[1:] + sigma
        Zi = 1.0 / Z
        VtZiV = Vm.T @ (Zi[:, None] * Vm)
        Q = (ls[:, None] * VtZiV) * ls[None, :]
        Q[np.diag_indices(K)] += 1.0
        L = sla.cholesky(Q, lower=True)
        ZiY = Zi[:, None] * Y
        alpha = ZiY - Zi[:, None] * ((Vm * ls) @ sla.cho_solve((L, True), ls[:, None] * (Vm.T @ ZiY)))
        Q_inv = sla.cho_solve((L, True), np.eye(K))
        A = -ev * (np.exp(-t * ev) + 0.0) + 0.0
        Vta = Vm.T @ alpha
        grad[0] = -0.5 * (Vta * (A[:, None] * Vta)).sum() / q
        grad[0] += 0.5 * np.trace(A[:, None] * VtZiV)
        grad[0] += -0.5 * ((Q_inv @ (ls[:, None] * VtZiV)) * ((A[:, None] * VtZiV) * ls[None, :]).T).sum()
        for i in range(m):
            g = -0.5 * (alpha[i] * alpha[i]).sum() / q
            tmp = Zi[i] * Vm[i] * ls
            g += 0.5 * (Zi[i] - ((tmp @ Q_inv) * tmp).sum())
            if abs(g) >= 1.0:
                g = g / abs(g) * 1.0
            grad[i + 1] = g
        nmll = 0.5 * (Y * alpha).sum() / q + np.log(np.diag(L) + 1e-9).sum() + 0.5 * np.log(Z + 1e-9).sum()
    if approach == "posterior":
        p, qq, tau, al, be = prior
        nmll += p * np.log(t + 1e-9) + (t / tau) ** (-qq)
        grad[0] += p / (t + 1e-9) - (qq / tau) * (t / tau) ** (-qq - 1)
        ns = x[1:] + sigma
        nmll += (((al + 1) * np.log(ns) + be / ns) / m).sum()
        grad[1:] += ((al + 1) / ns - be / ns ** 2) / m
    elif approach != "marginal":
        raise ValueError("This model selection approach is not supported!")
    return float(nmll), grad


def predict_regression_diff(V, values, Y, idx0, idx1, K, pars, sigma):
    """predict_regression_cpp, noisepar = "different" (src/Predict.cpp:76-113), literal."""
    import scipy.linalg as sla

    pars = np.asarray(pars, dtype=np.float64)
    t = float(pars[0])
    Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
    m = Y.shape[0]
    ev = 1.0 - values[:K]
    V0 = V[np.asarray(idx0), :K]
    V1 = V[np.asarray(idx1), :K]
    if m <= K:
        C = (V0 * np.exp(-t * ev)) @ V0.T
        C[np.diag_indices(m)] += sigma
        C[np.diag_indices(m)] += pars[1:]
        Cnv = (V1 * np.exp(-t * ev)) @ V0.T
        alpha = sla.cho_solve(sla.cho_factor(C, lower=True), Y)
        return (Cnv @ alpha)[:, 0]
    ls = np.exp(-0.5 * t * ev) + 0.0
    Zi = 1.0 / (pars[1:] + sigma)
    VtZV = V0.T @ (Zi[:, None] * V0)
    Q = (ls[:, None] * VtZV) * ls[None, :]
    Q[np.diag_indices(K)] += 1.0
    ZiY = Zi[:, None] * Y
    alpha = ZiY - Zi[:, None] * ((V0 * ls) @ sla.cho_solve(sla.cho_factor(Q, lower=True), ls[:, None] * (V0.T @ ZiY)))
    return (V1 @ (np.exp(-t * ev + 0.0)[:, None] * (V0.T @ alpha)))[:, 0]


def train_regression_diff(V, values, Y, idx, K, sigma=1e-5, approach="posterior", x0=None):
    """train_regression_gp_cpp, noise = "different" (src/train.cpp:588-611, 614-671) with m = number of training rows
    (the reference reads it through a mistyped pointer): (x[m + 1], -minimum)."""
    m = len(idx)
    x0 = np.concatenate([[10.0], np.ones(m)]) if x0 is None else np.asarray(x0, dtype=np.float64)
    lb = np.concatenate([[1e-3], np.full(m, 1e-4)])
    ub = np.full(m + 1, np.inf)
    x, fmin, _ = mma_minimize(lambda z: regression_objective_diff(V, values, Y, idx, K, z, sigma, approach), x0, lb, ub)
    return x, -fmin


def mma_minimize(f, x0, lb, ub, xtol_rel=1e-5, maxeval=1000):
    """Svanberg's CCSA with MMA approximations for bound constraints only, as NLopt's NLOPT_LD_MMA runs it
    (nloptr is an un-vendored dependency of the reference, version unpinned: DESCRIPTION; call site
    src/train.cpp:621-650 with xtol_rel = 1e-5, x0 = (10, 1), lb = (1e-3, 1e-4), ub = +inf).  Restated from the
    published algorithm (K. Svanberg, SIAM J. Optim. 12, 2002) in NLopt's arrangement: per-coordinate asymptote
    widths sigma (1 when a bound is infinite, else half the box), conservative inner iterations that raise rho,
    0.7 / 1.2 sigma adaptation from the sign pattern of successive steps.  f(x) -> (value, grad).
    Returns (x, minf, nevals)."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    lb = np.asarray(lb, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    sigma = np.where(np.isinf(ub) | np.isinf(lb), 1.0, 0.5 * (ub - lb))
    rho = 1.0
    minf, dfdx = f(x)
    dfdx = np.array(dfdx, dtype=np.float64)
    nev = 1
    xcur = x.copy()
    xprev = x.copy()
    xprevprev = x.copy()
    k = 0
    while True:
        k += 1
        if k > 1:
            xprevprev = xprev.copy()
        xprev = xcur.copy()
        while True:
            gval, wval = minf, 0.0
            xcur = x.copy()
            for j in range(n):
                if sigma[j] == 0:
                    continue
                s2 = sigma[j] * sigma[j]
                v = abs(dfdx[j]) * sigma[j] + 0.5 * rho
                u = dfdx[j] * s2
                dx = (u / v) / (-1.0 - np.sqrt(abs(1.0 - (u / (v * sigma[j])) ** 2)))
                xj = x[j] + dx
                xj = min(max(xj, lb[j]), ub[j])
                xj = min(max(xj, x[j] - 0.9 * sigma[j]), x[j] + 0.9 * sigma[j])
                xcur[j] = xj
                dx = xj - x[j]
                dx2 = dx * dx
                den = 1.0 / (s2 - dx2)
                gval += (dfdx[j] * s2 * dx + v * dx2) * den
                wval += 0.5 * dx2 * den
            fcur, dcur = f(xcur)
            nev += 1
            inner_done = gval >= fcur
            if fcur < minf:
                minf = fcur
                x = xcur.copy()
                dfdx = np.array(dcur, dtype=np.float64)
            if nev >= maxeval:
                return x, minf, nev
            if inner_done:
                break
            if fcur > gval:
                rho = min(10.0 * rho, 1.1 * (rho + (fcur - gval) / wval))
        if np.abs(xcur - xprev).sum() <= xtol_rel * np.abs(xcur).sum():
            return x, minf, nev
        rho = max(0.1 * rho, 1e-5)
        if k > 1:
            for j in range(n):
                dx2 = (xcur[j] - xprev[j]) * (xprev[j] - xprevprev[j])
                sigma[j] *= 0.7 if dx2 < 0 else (1.2 if dx2 > 0 else 1.0)
                if not (np.isinf(ub[j]) or np.isinf(lb[j])):
                    sigma[j] = max(min(sigma[j], 10.0 * (ub[j] - lb[j])), 0.01 * (ub[j] - lb[j]))


def train_regression(V, values, Y, idx, K, sigma=1e-5, approach="posterior", x0=(10.0, 1.0), lb=(1e-3, 1e-4),
                     ub=(np.inf, np.inf)):
    """train_regression_gp_cpp, noise="same" (src/train.cpp:557-671): returns (pars, obj = -minimum)."""
    x, minf, _ = mma_minimize(lambda x: regression_objective(V, values, Y, idx, K, x, sigma, approach), x0, lb, ub)
    return x, -minf


def fit_se_regression(X, Y, X_new, s, r, K, init_idx, a2s, sigma=1e-5, approach="posterior",
                      gl="cluster-normalized", root=True, iter_max=100, nthreads=1, pars=None):
    """fit_se_regression_gp_cpp (src/Fit.cpp:102-219): one k-means + KNN, then for every a2 of the grid
    Z = exp(-dist / (a2 * mean dist)), graph Laplacian, spectrum, empirical-Bayes training; the a2 with the largest
    objective wins.  pars: fixed (t, noise) instead of training (then the objective is evaluated at pars)."""
    m = len(X)
    X_all = np.asfortranarray(np.vstack([X, X_new]))
    n = len(X_all)
    if K < 0:
        K = s
    U, assign, iters = kmeans_lloyd(X_all, s, init_idx, iter_max, nthreads)
    ind, dist = knn(X_all, np.asfortranarray(U[:, :-1]), r, want_dist=True, nthreads=nthreads)
    Zj, Dx = knn_csr(ind, dist)
    dmean = Dx.sum() / (n * r)
    idx0 = np.arange(m, dtype=np.int32)
    idx1 = np.arange(m, n, dtype=np.int32)
    best = None
    for a2 in a2s:
        Zx = se_weights(Dx, a2 * dmean)
        nc = U[:, -1].copy() if gl == "cluster-normalized" else None
        Zx = graph_laplacian(Zj, Zx, s, gl, nc)
        values, V = spectrum_from_Z(Zj, Zx, s, K, root, nthreads=nthreads)
        if pars is None:
            x, obj = train_regression(V, values, Y, idx0, K, sigma, approach)
        else:
            x = np.asarray(pars, dtype=np.float64)
            obj = -regression_objective(V, values, Y, idx0, K, x, sigma, approach)[0]
        if best is None or obj > best["obj"]:
            best = dict(obj=obj, pars=x, a2=a2, values=values, V=V)
    V, values, x = best["V"], best["values"], best["pars"]
    best["train"] = predict_regression(V, values, Y, idx0, idx0, K, x, sigma)
    best["test"] = predict_regression(V, values, Y, idx0, idx1, K, x, sigma)
    best["cov"] = posterior_covariance_regression(V, values, idx0, idx1, K, x, sigma)
    return best


# ------------------------------------------------------------------ Nystrom extension
def _sqdist(A, B):
    """((-2 A B^T) + |a|^2) + |b|^2 with sequential, separately rounded sums (the contract for every distance)."""
    A = _f(A)
    B = _f(B)
    na, d = A.shape
    dot = np.zeros((na, B.shape[0]))
    an = np.zeros(na)
    bn = np.zeros(B.shape[0])
    for k in range(d):
        dot = dot + np.outer(A[:, k], B[:, k])
        an = an + A[:, k] * A[:, k]
        bn = bn + B[:, k] * B[:, k]
    return ((-2.0 * dot) + an[:, None]) + bn[None, :]


def _nystrom_grid(X, X_new, s, K, init_idx, a2s, iter_max, nthreads):
    """The part the fit_nystrom_* drivers share (src/Fit.cpp:242-292, 919-968), dense and literal: anchors, the doubly
    normalised SE anchor kernel per bandwidth, its top-K eigenpairs (RSpectra::eigs_sym -> LAPACK), the rescaled anchor
    eigenvectors and the extension formula.  Yields (a2, values, extend) with extend(D rows x s) -> rows x K."""
    X_all = np.asfortranarray(np.vstack([X, X_new]))
    U = kmeans_lloyd(X_all, s, init_idx, iter_max, nthreads)[0][:, :-1]
    D_UU = _sqdist(U, U)
    D_all = _sqdist(X_all, U)
    dmean = D_UU.sum() / (s * s)
    for a2 in a2s:
        Z_UU = np.exp(-D_UU / (a2 * dmean))
        rs = Z_UU.sum(axis=1) + 1e-9
        A_UU = (1.0 / rs)[:, None] * Z_UU * (1.0 / rs)[None, :]
        sdi = 1.0 / np.sqrt(A_UU.sum(axis=1) + 1e-9)
        W_UU = sdi[:, None] * A_UU * sdi[None, :]
        w, Q = np.linalg.eigh((W_UU + W_UU.T) / 2)
        values, vecs = w[::-1][:K].copy(), Q[:, ::-1][:, :K].copy()
        vecs = sdi[:, None] * vecs
        vecs = np.sqrt(s) * vecs * (1.0 / (np.linalg.norm(vecs, axis=0) + 1e-9))[None, :]

        def extend(Dx, a2=a2, rs=rs, vecs=vecs, values=values):  # bound now: the closure outlives the loop
            Zx = np.exp(-Dx / (a2 * dmean))
            Ax = (1.0 / (Zx.sum(axis=1) + 1e-9))[:, None] * Zx * (1.0 / rs)[None, :]
            Wx = (1.0 / (Ax.sum(axis=1) + 1e-9))[:, None] * Ax
            return Wx @ vecs * (1.0 / (np.abs(values) + 1e-9))[None, :]

        yield a2, values, extend, D_all


def fit_nystrom_logit(X, Y, X_new, s, K, init_idx, a2s, sigma=1e-3, approach="posterior", iter_max=100, nthreads=1,
                      N=None, t=None):
    """fit_nystrom_logit_gp_cpp (src/Fit.cpp:896-1038) without the label sampler: per bandwidth the diffusion time
    trained on the extended labelled rows (t given: the objective at t); Laplace posterior of the test rows from the
    winning extension (:1003-1022)."""
    m = len(X)
    n = m + len(X_new)
    if K < 0:
        K = s
    idx0 = np.arange(m, dtype=np.int32)
    idx1 = np.arange(m, n, dtype=np.int32)
    best = None
    for a2, values, extend, D_all in _nystrom_grid(X, X_new, s, K, init_idx, a2s, iter_max, nthreads):
        Vm = extend(D_all[:m])
        if t is None:
            tq, obj, _ = train_lae_logit(Vm, values, Y, idx0, K, sigma, approach, N=N)
        else:
            tq, obj = t, -logit_objective(Vm, values, Y, idx0, K, t, sigma, approach, N)
        if best is None or obj > best["obj"]:
            best = dict(obj=obj, t=tq, a2=a2, values=values, extend=extend)
    V = best["extend"](D_all)
    values, tq = best["values"], best["t"]
    C11 = hk_from_spectrum(V, values, K, tq, idx0, idx0)
    C11[np.diag_indices(m)] += sigma
    C21 = hk_from_spectrum(V, values, K, tq, idx1, idx0)
    C22 = ((V[m:, :K] * np.exp(-tq * (1.0 - values[:K]))) * V[m:, :K]).sum(axis=1) + sigma
    best["mean"], best["cov"] = posterior_distribution_classification(C11, C21, C22, Y)
    best["C"] = np.vstack([C11, C21])
    best["V"] = V
    del best["extend"]
    return best


def fit_nystrom_logit_mult(X, Y, X_new, s, K, init_idx, a2s, sigma=1e-3, approach="posterior", iter_max=100,
                           nthreads=1):
    """The training half of fit_nystrom_logit_mult_gp_cpp (src/Fit.cpp:1045-1162): per bandwidth the J one-vs-rest
    trainings on the extended labelled rows; the summed objective selects."""
    m = len(X)
    if K < 0:
        K = s
    idx0 = np.arange(m, dtype=np.int32)
    best = None
    for a2, values, extend, D_all in _nystrom_grid(X, X_new, s, K, init_idx, a2s, iter_max, nthreads):
        Vm = extend(D_all[:m])
        ts, objs = train_logit_mult(Vm, values, Y, idx0, K, sigma, approach)
        obj = 0.0
        for o in objs:
            obj += o
        if best is None or obj > best["obj"]:
            best = dict(obj=obj, t=ts, objs=objs, a2=a2, values=values, extend=extend)
    best["V"] = best["extend"](D_all)
    del best["extend"]
    return best


def fit_nystrom_regression(X, Y, X_new, s, K, init_idx, a2s, sigma=1e-5, approach="posterior", iter_max=100,
                           nthreads=1, pars=None):
    """fit_nystrom_regression_gp_cpp (src/Fit.cpp:222-357), dense and literal; RSpectra::eigs_sym replaced by LAPACK
    (top-K algebraic: W_UU is positive semi-definite, so largest magnitude = largest algebraic)."""
    m = len(X)
    X_all = np.asfortranarray(np.vstack([X, X_new]))
    n = len(X_all)
    if K < 0:
        K = s
    U = kmeans_lloyd(X_all, s, init_idx, iter_max, nthreads)[0][:, :-1]
    D_UU = _sqdist(U, U)
    D_all = _sqdist(X_all, U)
    dmean = D_UU.sum() / (s * s)
    idx0 = np.arange(m, dtype=np.int32)
    idx1 = np.arange(m, n, dtype=np.int32)
    best = None
    for a2 in a2s:
        Z_UU = np.exp(-D_UU / (a2 * dmean))
        rs = Z_UU.sum(axis=1) + 1e-9
        A_UU = (1.0 / rs)[:, None] * Z_UU * (1.0 / rs)[None, :]
        sdi = 1.0 / np.sqrt(A_UU.sum(axis=1) + 1e-9)
        W_UU = sdi[:, None] * A_UU * sdi[None, :]
        w, Q = np.linalg.eigh((W_UU + W_UU.T) / 2)
        values, vecs = w[::-1][:K].copy(), Q[:, ::-1][:, :K].copy()
        vecs = sdi[:, None] * vecs
        vecs = np.sqrt(s) * vecs * (1.0 / (np.linalg.norm(vecs, axis=0) + 1e-9))[None, :]

        def extend(Dx, a2=a2, rs=rs, vecs=vecs, values=values):  # bound now: the closure outlives the loop
            Zx = np.exp(-Dx / (a2 * dmean))
            Ax = (1.0 / (Zx.sum(axis=1) + 1e-9))[:, None] * Zx * (1.0 / rs)[None, :]
            Wx = (1.0 / (Ax.sum(axis=1) + 1e-9))[:, None] * Ax
            return Wx @ vecs * (1.0 / (np.abs(values) + 1e-9))[None, :]

        Vm = extend(D_all[:m])
        if pars is None:
            x, obj = train_regression(Vm, values, Y, idx0, K, sigma, approach)
        else:
            x = np.asarray(pars, dtype=np.float64)
            obj = -regression_objective(Vm, values, Y, idx0, K, x, sigma, approach)[0]
        if best is None or obj > best["obj"]:
            best = dict(obj=obj, pars=x, a2=a2, values=values, extend=extend)
    V = best["extend"](D_all)
    values, x = best["values"], best["pars"]
    best["train"] = predict_regression(V, values, Y, idx0, idx0, K, x, sigma)
    best["test"] = predict_regression(V, values, Y, idx0, idx1, K, x, sigma)
    best["cov"] = posterior_covariance_regression(V, values, idx0, idx1, K, x, sigma)
    del best["extend"]
    return best


# ------------------------------------------------------------------ Laplace posterior of the binary classifier
def posterior_distribution_classification(C11, C21, C22, Y, tol=1e-5, max_iter=100):
    """posterior_distribution_classification (src/Utils.cpp:252-299), literal: Newton iterations for the posterior
    mode (GPML algorithm 3.1) from f = 0, then mean = C21 (Y - pi), cov = C22 - rowsum((C21 beta) o C21)."""
    import scipy.linalg as sla

    C11 = np.asarray(C11, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(-1)
    m = Y.size
    f = np.zeros(m)

    def factor(f):
        pi = 1.0 / (1.0 + np.exp(-f))
        W = pi * (1.0 - pi)
        sw = np.sqrt(W)
        B = sw[:, None] * C11 * sw[None, :]
        B[np.diag_indices(m)] += 1.0
        return pi, W, sw, sla.cho_factor(B, lower=True)

    for _ in range(max_iter):
        pi, W, sw, cf = factor(f)
        b = W * f + (Y - pi)
        a = b - sw * sla.cho_solve(cf, sw * (C11 @ b))
        f_new = C11 @ a
        done = np.abs(f - f_new).sum() < tol
        f = f_new
        if done:
            break
    pi, W, sw, cf = factor(f)
    mean = C21 @ (Y - pi)
    beta = sw[:, None] * sla.cho_solve(cf, np.eye(m)) * sw[None, :]
    cov = np.asarray(C22) - ((C21 @ beta) * C21).sum(axis=1)
    return mean, cov


# ------------------------------------------------------------------ binary GP classifier: training of t (TEST ORACLE)
def posterior_distribution_multiclassification(V, values, Y, idx, idx_new, K, ts, sigma):
    """posterior_distribution_multiclassification (src/Utils.cpp:339-370), literal: per class j the Laplace posterior of
    the one-vs-rest labels at the class's own diffusion time t_j.  Note the reference's blocks here: C11 WITHOUT sigma
    on its diagonal (unlike the binary drivers, src/Fit.cpp:566), C22 with + sigma.  Returns (mean, cov), m_new x J."""
    Y = np.asarray(Y)
    J = len(ts)
    mean = np.zeros((len(idx_new), J))
    cov = np.zeros((len(idx_new), J))
    V2 = V[np.asarray(idx_new), :K]
    for j in range(J):
        C11 = hk_from_spectrum(V, values, K, ts[j], idx, idx)
        C21 = hk_from_spectrum(V, values, K, ts[j], idx_new, idx)
        C22 = ((V2 * np.exp(-ts[j] * (1.0 - values[:K]))) * V2).sum(axis=1) + sigma
        mean[:, j], cov[:, j] = posterior_distribution_classification(C11, C21, C22, (Y == j).astype(np.float64))
    return mean, cov


def laplace_mll(Cm, Y, N=None, tol=1e-5, max_iter=100):
    """marginal_log_likelihood_logit_la_cpp (src/train.cpp:716-760), literal: Newton from f = 0, value from the LAST
    Newton step's a and chol(B)."""
    import scipy.linalg as sla

    Cm = np.asarray(Cm, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(-1)
    m = Y.size
    N = np.ones(m) if N is None else np.asarray(N, dtype=np.float64).reshape(-1)
    f = np.zeros(m)
    a = np.zeros(m)
    Lb = np.eye(m)
    for _ in range(max_iter):
        pi = 1.0 / (1.0 + np.exp(-f))
        W = N * pi * (1.0 - pi)
        sw = np.sqrt(W)
        B = (sw[:, None] * Cm) * sw[None, :]
        B[np.diag_indices(m)] += 1.0
        Lb = sla.cholesky(B, lower=True)
        b = W * f + Y * (1.0 - pi) + (N - Y) * (-pi)
        a = b - sw * sla.cho_solve((Lb, True), sw * (Cm @ b))
        f_new = Cm @ a
        done = np.abs(f - f_new).sum() < tol
        f = f_new
        if done:
            break
    pi = 1.0 / (1.0 + np.exp(-f))
    amll = -0.5 * float(a @ f)
    amll += float((Y * np.log(pi)).sum() + ((N - Y) * np.log(1.0 - pi)).sum())
    amll -= float(np.log(np.diag(Lb) + 1e-9).sum())
    return amll


def logit_objective(V, values, Y, idx, K, t, sigma=1e-3, approach="posterior", N=None, prior=(1e-2, 10.0, 2.0)):
    """negative_marginal_likelihood_logit_cpp / negative_log_posterior_logit_cpp (src/train.cpp:14-36)."""
    Cm = hk_from_spectrum(V, values, K, t, idx, idx)
    Cm[np.diag_indices(len(idx))] += sigma
    mll = laplace_mll(Cm, Y, N)
    if approach == "marginal":
        return -mll
    p, q, tau = prior
    return -mll + p * np.log(t + 1e-9) + (t / tau) ** (-q)


def cobyla_minimize_1d(f, x0, lb=1e-3, ub=np.inf, xtol_rel=1e-4, maxeval=1000):
    """Twin of the library's one-variable restatement of NLOPT_LN_COBYLA (csrc/train.inl: cobyla_minimize_1d): two-point
    simplex, linear model, trust-region step rho from the best vertex clipped to the bounds, rho / 10 when a step does
    not pay off, from NLopt's default initial step down to xtol_rel times it.  Returns (x, fmin, evaluations)."""
    step = np.inf
    if np.isfinite(ub) and np.isfinite(lb) and (ub - lb) * 0.25 < step and ub > lb:
        step = (ub - lb) * 0.25
    if np.isfinite(ub) and ub - x0 < step and ub > x0:
        step = (ub - x0) * 0.75
    if np.isfinite(lb) and x0 - lb < step and x0 > lb:
        step = (x0 - lb) * 0.75
    if not np.isfinite(step):
        step = abs(x0)
    if not step > 0.0:
        step = 1.0
    rhobeg, rhoend = step, xtol_rel * step
    clip = lambda x: min(ub, max(lb, x))  # noqa: E731
    xa = clip(x0)
    fa = f(xa)
    xb = clip(xa + rhobeg if xa + rhobeg <= ub else xa - rhobeg)
    fb = f(xb)
    nev = 2
    if fb < fa:
        xa, xb, fa, fb = xb, xa, fb, fa
    rho = rhobeg
    while nev < maxeval:
        if xb != xa and fb != fa:
            direction = -1.0 if (fb - fa) / (xb - xa) > 0.0 else 1.0
        else:
            direction = -1.0 if xb > xa else 1.0
        xt = clip(xa + direction * rho)
        if xt == xa:
            xt = clip(xa - direction * rho)
        improved = False
        if xt != xa:
            ft = f(xt)
            nev += 1
            if ft < fa:
                predicted = abs((fb - fa) / (xb - xa)) * abs(xt - xa) if xb != xa else 0.0
                improved = (fa - ft) >= 0.1 * predicted
                xb, fb, xa, fa = xa, fa, xt, ft
            else:
                xb, fb = xt, ft
        if not improved:
            if rho <= rhoend:
                break
            rho *= 0.1
            if rho <= 1.5 * rhoend:
                rho = rhoend
    return xa, fa, nev


def train_lae_logit(V, values, Y, idx, K, sigma=1e-3, approach="posterior", t0=10.0, N=None):
    """train_lae_logit_gp_cpp (src/train.cpp:38-71): (t, -minimum, evaluations)."""
    t, fmin, nev = cobyla_minimize_1d(lambda t: logit_objective(V, values, Y, idx, K, t, sigma, approach, N), t0)
    return t, -fmin, nev


def train_logit_mult(V, values, Y, idx, K, sigma=1e-3, approach="posterior"):
    """train_logit_mult_gp_cpp (src/MultiClassification.cpp:30-53): J = max(Y) + 1 one-vs-rest binary trainings
    (multi_train_split, :14-27), each train_lae_logit_gp_cpp with N = 1 from t0 = 10.  Returns (t[J], obj[J])."""
    Y = np.asarray(Y)
    J = int(Y.max()) + 1
    ts, objs = np.zeros(J), np.zeros(J)
    for j in range(J):
        ts[j], objs[j], _ = train_lae_logit(V, values, (Y == j).astype(np.float64), idx, K, sigma, approach)
    return ts, objs


def _se_grid(X, X_new, s, r, K, init_idx, a2s, gl, root, iter_max, nthreads):
    """The part the fit_se_* drivers share (src/Fit.cpp:121-156, 690-733, 817-850): one k-means + KNN, then per a2
    the SE weights on the KNN graph, the graph Laplacian and the spectrum.  Yields (a2, values, V)."""
    X_all = np.asfortranarray(np.vstack([X, X_new]))
    n = len(X_all)
    U, assign, iters = kmeans_lloyd(X_all, s, init_idx, iter_max, nthreads)
    ind, dist = knn(X_all, np.asfortranarray(U[:, :-1]), r, want_dist=True, nthreads=nthreads)
    Zj, Dx = knn_csr(ind, dist)
    dmean = Dx.sum() / (n * r)
    for a2 in a2s:
        Zx = se_weights(Dx, a2 * dmean)
        nc = U[:, -1].copy() if gl == "cluster-normalized" else None
        Zx = graph_laplacian(Zj, Zx, s, gl, nc)
        values, V = spectrum_from_Z(Zj, Zx, s, K, root, nthreads=nthreads)
        yield a2, values, V


def fit_se_logit(X, Y, X_new, s, r, K, init_idx, a2s, sigma=1e-3, approach="posterior", gl="cluster-normalized",
                 root=True, iter_max=100, nthreads=1, N=None, t=None):
    """fit_se_logit_gp_cpp (src/Fit.cpp:668-794) without the label sampler: the bandwidth grid with the diffusion time
    trained per grid point (t given: the objective at t), the largest objective wins (:737-742); Laplace posterior of
    the test rows at the winner (:752-773)."""
    m = len(X)
    n = m + len(X_new)
    if K < 0:
        K = s
    idx0 = np.arange(m, dtype=np.int32)
    idx1 = np.arange(m, n, dtype=np.int32)
    best = None
    for a2, values, V in _se_grid(X, X_new, s, r, K, init_idx, a2s, gl, root, iter_max, nthreads):
        if t is None:
            tq, obj, _ = train_lae_logit(V, values, Y, idx0, K, sigma, approach, N=N)
        else:
            tq, obj = t, -logit_objective(V, values, Y, idx0, K, t, sigma, approach, N)
        if best is None or obj > best["obj"]:
            best = dict(obj=obj, t=tq, a2=a2, values=values, V=V)
    V, values, tq = best["V"], best["values"], best["t"]
    C11 = hk_from_spectrum(V, values, K, tq, idx0, idx0)
    C11[np.diag_indices(m)] += sigma
    C21 = hk_from_spectrum(V, values, K, tq, idx1, idx0)
    C22 = ((V[m:, :K] * np.exp(-tq * (1.0 - values[:K]))) * V[m:, :K]).sum(axis=1) + sigma
    best["mean"], best["cov"] = posterior_distribution_classification(C11, C21, C22, Y)
    best["C"] = np.vstack([C11, C21])
    return best


def fit_se_logit_mult(X, Y, X_new, s, r, K, init_idx, a2s, sigma=1e-3, approach="posterior",
                      gl="cluster-normalized", root=True, iter_max=100, nthreads=1):
    """fit_se_logit_mult_gp_cpp (src/Fit.cpp:797-895) without the label sampler: per a2 the J one-vs-rest trainings;
    the grid point's objective is the sum of the class objectives (:855-859)."""
    m = len(X)
    if K < 0:
        K = s
    idx0 = np.arange(m, dtype=np.int32)
    best = None
    for a2, values, V in _se_grid(X, X_new, s, r, K, init_idx, a2s, gl, root, iter_max, nthreads):
        ts, objs = train_logit_mult(V, values, Y, idx0, K, sigma, approach)
        obj = 0.0
        for o in objs:
            obj += o
        if best is None or obj > best["obj"]:
            best = dict(obj=obj, t=ts, objs=objs, a2=a2, values=values, V=V)
    return best
