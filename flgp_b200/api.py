"""Host-side mirror of FLGP's Rcpp export surface for the spectral core, bound to libflgp_b200.so.

Function names, argument meaning and error behaviour follow the reference's exports
(/root/reference/R/RcppExports.R, src/RcppExports.cpp:471-504) so that tests read like calls into
the R package: KNN_cpp, LAE_cpp, subsample_cpp, v_to_z_cpp, local_anchor_embedding_cpp,
cross_similarity_lae_cpp, lae_eigenmap, heat_kernel_covariance_rcpp, fit_lae_regression_gp_rcpp ...
Matrices go in as anything numpy can view; they are handed to the C ABI column-major (R's layout).
Sparse results come back as scipy CSR with exactly r stored entries per row (the dgRMatrix of the
reference, explicit zeros kept).  Every numeric operation runs in the CUDA library; nothing here
computes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FlgpError, check

GL = {"rw": 0, "normalized": 1, "cluster-normalized": 2}
DEFAULT_MODELS = dict(subsample="kmeans", kernel="lae", gl="cluster-normalized", root=True)


def _gl(name) -> int:
    if isinstance(name, str):
        if name not in GL:
            raise FlgpError("Error: the type of graph Laplacian is not supported!")
        return GL[name]
    return int(name)


def _f64(a):
    return np.asfortranarray(a, dtype=np.float64)


def _pf(a):
    return a.ctypes.data_as(_lib.p_f64) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_lib.p_i32) if a is not None else None


def _idx(a):
    return np.ascontiguousarray(a, dtype=np.int32) if a is not None else None


def _b(s: Optional[str]):
    return s.encode() if s is not None else None


class Context:
    """One per process / GPU: device, stream, (optional) NCCL communicator."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.flgp_ctx_create(device, C.byref(self._h)))
        self.device = device
        self.rank, self.nranks = 0, 1

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.flgp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- streams / accounting
    def set_stream(self, cuda_stream: Optional[int]):
        check(self._lib.flgp_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(self._lib.flgp_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.flgp_ctx_launch_count(self._h))

    def set_timing(self, on: bool):
        check(self._lib.flgp_ctx_set_timing(self._h, int(on)))

    def stage_reset(self):
        check(self._lib.flgp_ctx_stage_reset(self._h))

    def stages(self):
        out = []
        for i in range(self._lib.flgp_ctx_stage_count(self._h)):
            name = C.create_string_buffer(64)
            ms, fl, by = C.c_double(), C.c_double(), C.c_double()
            ln = C.c_uint64()
            check(self._lib.flgp_ctx_stage_get(self._h, i, name, 64, C.byref(ms), C.byref(ln), C.byref(fl), C.byref(by)))
            out.append(dict(name=name.value.decode(), ms=ms.value, launches=int(ln.value), flops=fl.value,
                            bytes=by.value))
        return out

    def dfma_peak_tflops(self, iters: int = 1 << 15) -> float:
        v = C.c_double()
        check(self._lib.flgp_dfma_peak(self._h, iters, C.byref(v)))
        return v.value

    def copy_roundtrip(self, arr: np.ndarray) -> np.ndarray:
        """Host -> device -> host through the library's copy path (pageable arrays of 8 MB and more are staged by
        several host threads, csrc/hostcopy.cu)."""
        src = np.ascontiguousarray(arr)
        out = np.empty_like(src)
        check(self._lib.flgp_copy_roundtrip(self._h, src.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                            src.nbytes))
        return out

    # -- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().flgp_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: Optional[bytes], rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(self._lib.flgp_ctx_comm_init(self._h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks


_default_ctx: Optional[Context] = None


def default_ctx() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def default_init(n: int, s: int, seed: int = 0) -> np.ndarray:
    """The s distinct sorted row indices the library starts Lloyd from when none are given."""
    out = np.zeros(s, np.int32)
    check(_lib.load().flgp_default_init(n, s, seed, _pi(out)))
    return out


def _csr(n, s, r, Zj, Zx):
    import scipy.sparse as sp

    return sp.csr_matrix((Zx.reshape(-1), Zj.reshape(-1), np.arange(0, n * r + 1, r)), shape=(n, s))


def _from_csr(Z, r=None):
    """(Zj, Zx, n, s, r) from a scipy CSR / (Zj, Zx, s) triple with exactly r entries per row."""
    if isinstance(Z, tuple):
        Zj, Zx, s = Z
        Zj = np.ascontiguousarray(Zj, dtype=np.int32)
        Zx = np.ascontiguousarray(Zx, dtype=np.float64)
        n, r = Zj.shape
        if r > 1 and np.any(np.diff(Zj, axis=1) < 0):  # the library wants column-sorted rows (it rejects others)
            order = np.argsort(Zj, axis=1, kind="stable")
            Zj = np.ascontiguousarray(np.take_along_axis(Zj, order, axis=1))
            Zx = np.ascontiguousarray(np.take_along_axis(Zx.reshape(n, r), order, axis=1))
        return Zj, Zx, n, s, r
    n, s = Z.shape
    cnt = np.diff(Z.indptr)
    if n == 0 or not np.all(cnt == cnt[0]):
        raise FlgpError("the sparse matrix must store the same number of entries in every row")
    r = int(cnt[0])
    if not Z.has_sorted_indices:  # scipy does not keep rows column-sorted; dgRMatrix (the R side) does
        Z = Z.copy()
        Z.sort_indices()
    return (np.ascontiguousarray(Z.indices, dtype=np.int32).reshape(n, r),
            np.ascontiguousarray(Z.data, dtype=np.float64).reshape(n, r), n, s, r)


# ------------------------------------------------------------------------------------------------
# Rcpp exports
# ------------------------------------------------------------------------------------------------
def subsample_cpp(X, s: int, method: str = "kmeans", nstart: int = 1, *, init_idx=None, seed: int = 0,
                  iter_max: int = 100, return_info: bool = False, ctx: Optional[Context] = None):
    """subsample_cpp (src/Utils.cpp:32-68).  "kmeans" / "minibatchkmeans": U = [centres, size] (s x (d+1)); "random":
    s x d.  "minibatchkmeans": mini-batch k-means from the start rows (at most iter_max batches of 10 s rows keyed by
    seed), sizes = rows per nearest centre; return_info then gives the number of batches run (no assignments)."""
    ctx = ctx or default_ctx()
    X = _f64(X)
    n, d = X.shape
    ucols = d if method == "random" else d + 1
    U = np.zeros((s, ucols), order="F")
    assign = np.zeros(n, np.int32)
    iters = C.c_int(0)
    check(ctx._lib.flgp_subsample(ctx._h, _pf(X), n, d, s, _b(method), iter_max, nstart, _pi(_idx(init_idx)), seed,
                                  _pf(U), _pi(assign), C.byref(iters)))
    if return_info:
        return U, assign, iters.value
    return U


def KNN_cpp(X, U, r: int, distance: str = "Euclidean", output: bool = False, batch: int = 100, *,
            ctx: Optional[Context] = None):
    """KNN_cpp (src/Utils.cpp:102-192) -> {"ind_knn": n x r int32 (0-based), ["distances_sp": CSR n x s]}.
    `batch` only bounded the reference's working set; it never changed the result and is ignored."""
    if distance != "Euclidean":
        raise FlgpError("The distance method of KNN is not supported!")
    ctx = ctx or default_ctx()
    X = _f64(X)
    U = _f64(U)
    n, d = X.shape
    s = U.shape[0]
    if U.shape[1] != d:
        raise FlgpError("X and U must have the same number of columns")
    ind = np.zeros((n, r), np.int32, order="F")
    if output:
        Zj = np.zeros((n, r), np.int32)
        Zx = np.zeros((n, r))
        check(ctx._lib.flgp_knn(ctx._h, _pf(X), n, d, _pf(U), s, r, _pi(ind), None, _pi(Zj), _pf(Zx)))
        return {"ind_knn": ind, "distances_sp": _csr(n, s, r, Zj, Zx)}
    check(ctx._lib.flgp_knn(ctx._h, _pf(X), n, d, _pf(U), s, r, _pi(ind), None, None, None))
    return {"ind_knn": ind}


def knn_distances(X, U, r: int, *, ctx: Optional[Context] = None):
    """ind_knn together with the n x r squared distances in the same (ascending) order."""
    ctx = ctx or default_ctx()
    X = _f64(X)
    U = _f64(U)
    n, d = X.shape
    s = U.shape[0]
    ind = np.zeros((n, r), np.int32, order="F")
    dist = np.zeros((n, r), order="F")
    check(ctx._lib.flgp_knn(ctx._h, _pf(X), n, d, _pf(U), s, r, _pi(ind), _pf(dist), None, None))
    return ind, dist


def v_to_z_cpp(v, *, ctx: Optional[Context] = None):
    """v_to_z_cpp (src/lae.cpp:137-153)."""
    ctx = ctx or default_ctx()
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
    z = np.zeros_like(v)
    check(ctx._lib.flgp_simplex_project(ctx._h, _pf(v), v.size, _pf(z)))
    return z


def local_anchor_embedding_cpp(x, U, *, ctx: Optional[Context] = None):
    """local_anchor_embedding_cpp (src/lae.cpp:76-133); U is r x d."""
    ctx = ctx or default_ctx()
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    U = _f64(U)
    r, d = U.shape
    if x.size != d:
        raise FlgpError("x and U must have the same number of columns")
    z = np.zeros(r)
    check(ctx._lib.flgp_lae_point(ctx._h, _pf(x), d, _pf(U), r, _pf(z)))
    return z


def LAE_cpp(X, U, r: int, *, return_stats: bool = False, ctx: Optional[Context] = None):
    """LAE_cpp (src/lae.cpp:48-70): sparse n x s matrix of simplex weights on the r nearest anchors."""
    ctx = ctx or default_ctx()
    X = _f64(X)
    U = _f64(U)
    n, d = X.shape
    s = U.shape[0]
    Zj = np.zeros((n, r), np.int32)
    Zx = np.zeros((n, r))
    stats = np.zeros(2, np.int64)
    check(ctx._lib.flgp_lae(ctx._h, _pf(X), n, d, _pf(U), s, r, _pi(Zj), _pf(Zx), stats.ctypes.data_as(_lib.p_i64)))
    Z = _csr(n, s, r, Zj, Zx)
    return (Z, stats) if return_stats else Z


def graphLaplacian_cpp(Z, gl: str, num_class=None, *, ctx: Optional[Context] = None):
    """graphLaplacian_cpp (src/Utils.cpp:195-212); returns the scaled matrix (the reference works in place)."""
    ctx = ctx or default_ctx()
    Zj, Zx, n, s, r = _from_csr(Z)
    Zx = Zx.copy()
    nc = np.ascontiguousarray(num_class, dtype=np.float64) if num_class is not None else None
    check(ctx._lib.flgp_graph_laplacian(ctx._h, n, s, r, _pi(Zj), _pf(Zx), _gl(gl), _pf(nc)))
    return _csr(n, s, r, Zj, Zx)


def cross_similarity_lae_cpp(X, U, r: int, gl: str = "rw", *, ctx: Optional[Context] = None):
    """cross_similarity_lae_cpp (src/Spectrum.cpp:101-117); U is s x (d+1) for cluster-normalized."""
    ctx = ctx or default_ctx()
    X = _f64(X)
    U = _f64(U)
    n, d = X.shape
    s, ucols = U.shape
    Zj = np.zeros((n, r), np.int32)
    Zx = np.zeros((n, r))
    check(ctx._lib.flgp_cross_similarity_lae(ctx._h, _pf(X), n, d, _pf(U), s, ucols, r, _gl(gl), _pi(Zj), _pf(Zx)))
    return _csr(n, s, r, Zj, Zx)


def cross_similarity_se_cpp(X, U, r: int, gl: str = "rw", epsilon: float = 0.1, *, ctx: Optional[Context] = None):
    """cross_similarity_se_cpp (src/Spectrum.cpp:120-142)."""
    ctx = ctx or default_ctx()
    X = _f64(X)
    U = _f64(U)
    n, d = X.shape
    s, ucols = U.shape
    Zj = np.zeros((n, r), np.int32)
    Zx = np.zeros((n, r))
    check(ctx._lib.flgp_cross_similarity_se(ctx._h, _pf(X), n, d, _pf(U), s, ucols, r, _gl(gl), epsilon, _pi(Zj),
                                            _pf(Zx)))
    return _csr(n, s, r, Zj, Zx)


class EigenPair:
    """Device-resident EigenPair (src/Spectrum.h:117-124): values (K) and the n x K vectors, which are
    kept factored (sparse A times an s x K lift operator) and only materialised on request."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle
        info = np.zeros(10, np.int64)
        check(ctx._lib.flgp_spectrum_info(self._h, info.ctypes.data_as(_lib.p_i64)))
        (self.n_local, self.n_total, self.row_offset, self.d, self.s, self.r, self.K, self.kmeans_iters,
         self.lae_iters, self.lae_backtracks) = (int(v) for v in info)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.ctx._lib.flgp_spectrum_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def values(self) -> np.ndarray:
        v = np.zeros(self.K)
        check(self.ctx._lib.flgp_spectrum_values(self._h, _pf(v)))
        return v

    @property
    def vectors(self) -> np.ndarray:
        V = np.zeros((self.n_local, self.K), order="F")
        check(self.ctx._lib.flgp_spectrum_vectors(self._h, _pf(V)))
        return V

    def anchors(self, ucols: Optional[int] = None) -> np.ndarray:
        U = np.zeros((self.s, ucols or self.d + 1), order="F")
        check(self.ctx._lib.flgp_spectrum_anchors(self._h, _pf(U)))
        return U

    def Z(self):
        Zj = np.zeros((self.n_local, self.r), np.int32)
        Zx = np.zeros((self.n_local, self.r))
        check(self.ctx._lib.flgp_spectrum_z(self._h, _pi(Zj), _pf(Zx)))
        return _csr(self.n_local, self.s, self.r, Zj, Zx)

    def rows(self, idx) -> np.ndarray:
        """mat_indexing(eigenvectors, idx, 0..K-1) (src/Utils.h:130-137)."""
        idx = _idx(idx)
        V = np.zeros((idx.size, self.K), order="F")
        check(self.ctx._lib.flgp_spectrum_gather_rows(self._h, _pi(idx), idx.size, _pf(V)))
        return V


def spectrum_from_Z_cpp(Z, K: int, root: bool = False, *, ctx: Optional[Context] = None) -> EigenPair:
    """spectrum_from_Z_cpp (src/Spectrum.cpp:146-161)."""
    ctx = ctx or default_ctx()
    Zj, Zx, n, s, r = _from_csr(Z)
    h = C.c_void_p()
    check(ctx._lib.flgp_spectrum_from_z(ctx._h, n, s, r, _pi(Zj), _pf(Zx), K, int(root), None, None, C.byref(h)))
    return EigenPair(ctx, h)


def eigs_sym(A, k: int, *, ctx: Optional[Context] = None):
    """RSpectra::eigs_sym(A, k) as the Nystrom / GLGP drivers call it (src/Fit.cpp:262-278): the k algebraically
    largest eigenpairs of the dense symmetric matrix A, values descending, orthonormal vectors (signs arbitrary)."""
    ctx = ctx or default_ctx()
    A = _f64(A)
    s = A.shape[0]
    if A.ndim != 2 or A.shape[1] != s:
        raise FlgpError("eigs_sym: A must be square")
    values = np.zeros(k)
    vectors = np.zeros((s, k), order="F")
    check(ctx._lib.flgp_eigs_sym(ctx._h, _pf(A), s, k, _pf(values), _pf(vectors)))
    return {"values": values, "vectors": vectors}


def heat_kernel_spectrum_cpp(X, X_new, s: int, r: int, K: int, models=None, nstart: int = 1, epsilon: float = 0.1, *,
                             init_idx=None, seed: int = 0, iter_max: int = 100,
                             ctx: Optional[Context] = None) -> EigenPair:
    """heat_kernel_spectrum_cpp (src/Spectrum.cpp:48-76)."""
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    X = _f64(X)
    m, d = X.shape
    if X_new is not None and len(X_new):
        X_new = _f64(X_new)
        m_new = X_new.shape[0]
    else:
        X_new, m_new = None, 0
    h = C.c_void_p()
    check(ctx._lib.flgp_heat_kernel_spectrum(ctx._h, _pf(X), m, _pf(X_new), m_new, d, s, r, K, _b(mo["subsample"]),
                                             _b(mo["kernel"]), _gl(mo["gl"]), int(bool(mo["root"])), nstart, epsilon,
                                             iter_max, _pi(_idx(init_idx)), seed, C.byref(h)))
    return EigenPair(ctx, h)


def heat_kernel_spectrum_sharded(X_local, n_total: int, row_offset: int, s: int, r: int, K: int, models=None,
                                 nstart: int = 1, epsilon: float = 0.1, *, init_idx=None, seed: int = 0,
                                 iter_max: int = 100, ctx: Optional[Context] = None, device_ptr: Optional[int] = None,
                                 n_local: Optional[int] = None, d: Optional[int] = None) -> EigenPair:
    """The same on one contiguous block of rows of X_all (multi-GPU); with `device_ptr` the block already
    lives in HBM (column-major, leading dimension n_local)."""
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    h = C.c_void_p()
    args = (s, r, K, _b(mo["subsample"]), _b(mo["kernel"]), _gl(mo["gl"]), int(bool(mo["root"])), nstart, epsilon,
            iter_max, _pi(_idx(init_idx)), seed, C.byref(h))
    if device_ptr is not None:
        check(ctx._lib.flgp_heat_kernel_spectrum_dev(ctx._h, C.c_void_p(device_ptr), n_local, n_total, row_offset, d,
                                                     *args))
    else:
        X_local = _f64(X_local)
        n_local, d = X_local.shape
        check(ctx._lib.flgp_heat_kernel_spectrum_sharded(ctx._h, _pf(X_local), n_local, n_total, row_offset, d, *args))
    return EigenPair(ctx, h)


def HK_from_spectrum_cpp(eigenpair: EigenPair, K: int, t: float, idx0, idx1) -> np.ndarray:
    """HK_from_spectrum_cpp (src/Spectrum.cpp:83-94)."""
    idx0 = _idx(idx0)
    idx1 = _idx(idx1)
    H = np.zeros((idx0.size, idx1.size), order="F")
    check(eigenpair.ctx._lib.flgp_hk_from_spectrum(eigenpair._h, K, t, _pi(idx0), idx0.size, _pi(idx1), idx1.size,
                                                   _pf(H)))
    return H


def lae_eigenmap(X, s: int, r: int, ndim: int, subsample: str = "kmeans", norm: str = "cluster-normalized",
                 nstart: int = 1, *, init_idx=None, seed: int = 0, iter_max: int = 100,
                 ctx: Optional[Context] = None):
    """lae_eigenmap (src/Spectrum.cpp:17-25) -> {"eigenvalues", "eigenvectors"}."""
    ctx = ctx or default_ctx()
    X = _f64(X)
    n, d = X.shape
    ev = np.zeros(ndim)
    V = np.zeros((n, ndim), order="F")
    check(ctx._lib.flgp_lae_eigenmap(ctx._h, _pf(X), n, d, s, r, ndim, _b(subsample), _gl(norm), nstart, iter_max,
                                     _pi(_idx(init_idx)), seed, _pf(ev), _pf(V)))
    return {"eigenvalues": ev, "eigenvectors": V}


def heat_kernel_covariance_rcpp(X, X_new, s: int, r: int, t: float, K: int = -1, models=None, nstart: int = 1,
                                epsilon: float = 0.1, *, init_idx=None, seed: int = 0, iter_max: int = 100,
                                ctx: Optional[Context] = None) -> np.ndarray:
    """heat_kernel_covariance_rcpp (R/Fit.R:760-770 -> src/Spectrum.cpp:28-43): H is (m+m_new) x m."""
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    X = _f64(X)
    m, d = X.shape
    if X_new is not None and len(X_new):
        X_new = _f64(X_new)
        m_new = X_new.shape[0]
    else:
        X_new, m_new = None, 0
    H = np.zeros((m + m_new, m), order="F")
    check(ctx._lib.flgp_heat_kernel_covariance(ctx._h, _pf(X), m, _pf(X_new), m_new, d, s, r, t, K,
                                               _b(mo["subsample"]), _b(mo["kernel"]), _gl(mo["gl"]),
                                               int(bool(mo["root"])), nstart, epsilon, iter_max, _pi(_idx(init_idx)),
                                               seed, _pf(H)))
    return H


def regression_fixed(eigenpair: EigenPair, Y_local, m_total: int, K: int, pars: Sequence[float], sigma: float = 1e-5,
                     want_cov: bool = True):
    """predict_regression_cpp + posterior_covariance_regression at fixed pars = (t, noise) on a handle."""
    Y_local = np.ascontiguousarray(Y_local, dtype=np.float64).reshape(-1)
    y = np.zeros(eigenpair.n_local)
    cov = np.zeros(eigenpair.n_local) if want_cov else None
    check(eigenpair.ctx._lib.flgp_regression_fixed(eigenpair._h, _pf(Y_local), m_total, K, pars[0], pars[1], sigma,
                                                   _pf(y), _pf(cov)))
    return y, cov


def regression_objective(eigenpair: EigenPair, Y_local, m_total: int, K: int, pars: Sequence[float],
                         sigma: float = 1e-5, approach: str = "posterior"):
    """negative_marginal_likelihood_regression_cpp / negative_log_posterior_regression_cpp at pars = (t, noise)
    (src/train.cpp:333-436, noise="same").  Returns (objective, grad[2])."""
    Y_local = np.ascontiguousarray(Y_local, dtype=np.float64).reshape(-1)
    x = np.ascontiguousarray(pars, dtype=np.float64)
    obj = C.c_double()
    grad = np.zeros(2)
    check(eigenpair.ctx._lib.flgp_regression_objective(eigenpair._h, _pf(Y_local), m_total, K, sigma, _b(approach),
                                                       _pf(x), C.byref(obj), _pf(grad)))
    return obj.value, grad


def train_regression_gp(eigenpair: EigenPair, Y_local, m_total: int, K: int, sigma: float = 1e-5,
                        approach: str = "posterior", x0: Optional[Sequence[float]] = None):
    """train_regression_gp_cpp, noise="same" (src/train.cpp:557-671).  Returns (pars, obj, evaluations)."""
    Y_local = np.ascontiguousarray(Y_local, dtype=np.float64).reshape(-1)
    x = np.array(x0 if x0 is not None else (np.nan, np.nan), dtype=np.float64)
    obj = C.c_double()
    nev = C.c_int()
    check(eigenpair.ctx._lib.flgp_train_regression(eigenpair._h, _pf(Y_local), m_total, K, sigma, _b(approach), _pf(x),
                                                   C.byref(obj), C.byref(nev)))
    return x, obj.value, nev.value


def regression_objective_rows(V1, values, Y, pars, sigma: float = 1e-5, approach: str = "marginal"):
    """negative_marginal_likelihood_regression_cpp / negative_log_posterior_regression_cpp, noise = "same"
    (src/train.cpp:333-436), on explicit training rows V1 (m x K): (objective, grad[2]).  Host only."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    x = np.ascontiguousarray(pars, dtype=np.float64)
    obj = C.c_double()
    grad = np.zeros(2)
    check(_lib.load().flgp_regression_objective_rows(_pf(V1), _pf(values), _pf(Y), m, K, sigma, _b(approach), _pf(x),
                                                     C.byref(obj), _pf(grad)))
    return obj.value, grad


def train_regression_rows(V1, values, Y, sigma: float = 1e-5, approach: str = "posterior", x0=None):
    """train_regression_gp_cpp, noise = "same" (src/train.cpp:557-671), on explicit training rows:
    (pars[2], objective, evaluations).  Host only."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    x = np.array(x0 if x0 is not None else (np.nan, np.nan), dtype=np.float64)
    obj = C.c_double()
    nev = C.c_int()
    check(_lib.load().flgp_train_regression_rows(_pf(V1), _pf(values), _pf(Y), m, K, sigma, _b(approach), _pf(x),
                                                 C.byref(obj), C.byref(nev)))
    return x, obj.value, nev.value


def regression_objective_diff_rows(V1, values, Y, x, sigma: float = 1e-5, approach: str = "marginal"):
    """negative_marginal_likelihood_diff_noise_regression_cpp / negative_log_posterior_diff_noise_regression_cpp
    (src/train.cpp:438-556) on explicit training rows V1 (m x K) of the eigenvectors: (objective, grad[m + 1]).  Host only."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    x = np.ascontiguousarray(x, dtype=np.float64)
    if Y.size != m or x.size != m + 1 or values.size != K:
        raise FlgpError("noise=\"different\": need m labels, m + 1 parameters and K eigenvalues")
    obj = C.c_double()
    grad = np.zeros(m + 1)
    check(_lib.load().flgp_regression_objective_diff_rows(_pf(V1), _pf(values), _pf(Y), m, K, sigma, _b(approach),
                                                          _pf(x), C.byref(obj), _pf(grad)))
    return obj.value, grad


def train_regression_diff_rows(V1, values, Y, sigma: float = 1e-5, approach: str = "posterior", x0=None):
    """train_regression_gp_cpp, noise = "different", on explicit training rows: (x[m + 1], objective, evaluations)."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    x = np.full(m + 1, np.nan) if x0 is None else np.array(x0, dtype=np.float64)
    obj = C.c_double()
    nev = C.c_int()
    check(_lib.load().flgp_train_regression_diff_rows(_pf(V1), _pf(values), _pf(Y), m, K, sigma, _b(approach), _pf(x),
                                                      C.byref(obj), C.byref(nev)))
    return x, obj.value, nev.value


def predict_coef_diff_rows(V1, values, Y, x, sigma: float = 1e-5) -> np.ndarray:
    """predict_regression_cpp, noisepar = "different": coef (K) with Y_pred = V_new @ coef.  Host only."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    x = np.ascontiguousarray(x, dtype=np.float64)
    coef = np.zeros(K)
    check(_lib.load().flgp_predict_coef_diff_rows(_pf(V1), _pf(values), _pf(Y), m, K, sigma, _pf(x), _pf(coef)))
    return coef


def mma_minimize(f, x0, lb, ub, xtol_rel: float = 1e-5, maxeval: int = 1000):
    """The library's NLOPT_LD_MMA restatement on a Python objective f(x) -> (value, grad).  Host only."""
    from ._lib import OBJECTIVE_FN, load

    x = np.array(x0, dtype=np.float64)
    n = x.size

    def cb(nn, xp, gp, _):
        xx = np.array([xp[i] for i in range(nn)])
        v, g = f(xx)
        if gp:
            for i in range(nn):
                gp[i] = float(g[i])
        return float(v)

    cfn = OBJECTIVE_FN(cb)
    lbv = np.ascontiguousarray(lb, dtype=np.float64)
    ubv = np.ascontiguousarray(ub, dtype=np.float64)
    minf = C.c_double()
    nev = C.c_int()
    check(load().flgp_mma_minimize(n, cfn, None, _pf(lbv), _pf(ubv), _pf(x), C.byref(minf), xtol_rel, maxeval,
                                   C.byref(nev)))
    return x, minf.value, nev.value


def cobyla_minimize_1d(f, x0: float, lb: float = 1e-3, ub: float = float("inf"), xtol_rel: float = 1e-4,
                       maxeval: int = 1000):
    """The library's one-variable restatement of NLOPT_LN_COBYLA (src/train.cpp:38-71) on a Python objective f(t).
    Host only.  Returns (t, minimum, evaluations)."""
    from ._lib import OBJECTIVE_FN, load

    def cb(nn, xp, gp, _):
        return float(f(float(xp[0])))

    cfn = OBJECTIVE_FN(cb)
    x = np.array([x0], dtype=np.float64)
    minf = C.c_double()
    nev = C.c_int()
    check(load().flgp_cobyla_minimize_1d(cfn, None, lb, ub, _pf(x), C.byref(minf), xtol_rel, maxeval, C.byref(nev)))
    return float(x[0]), minf.value, nev.value


def marginal_log_likelihood_logit_la_cpp(Cm, Y, N=None, tol: float = 1e-5, max_iter: int = 100) -> float:
    """marginal_log_likelihood_logit_la_cpp (src/train.cpp:716-760; exported): Laplace-approximate marginal
    log-likelihood of the labelled rows for the m x m covariance Cm.  Host only."""
    Cm = _f64(Cm)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    out = C.c_double()
    check(_lib.load().flgp_marginal_log_likelihood_logit_la(_pf(Cm), _pf(Y), _pf(Nv), Y.size, tol, max_iter,
                                                            C.byref(out)))
    return out.value


def logit_objective_rows(V1, values, Y, t: float, sigma: float = 1e-3, approach: str = "posterior", N=None) -> float:
    """negative_marginal_likelihood_logit_cpp / negative_log_posterior_logit_cpp (src/train.cpp:14-36) on explicit
    labelled rows V1 (m x K) of the eigenvectors.  Host only."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    obj = C.c_double()
    check(_lib.load().flgp_logit_objective_rows(_pf(V1), _pf(values), _pf(Y), _pf(Nv), m, K, sigma, _b(approach), t,
                                                C.byref(obj)))
    return obj.value


def train_logit_rows(V1, values, Y, sigma: float = 1e-3, approach: str = "posterior", t0: Optional[float] = None, N=None):
    """train_lae_logit_gp_cpp (src/train.cpp:38-71) on explicit labelled rows: (t, objective, evaluations).  Host only."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    tt = np.array([np.nan if t0 is None else t0], dtype=np.float64)
    obj = C.c_double()
    nev = C.c_int()
    check(_lib.load().flgp_train_logit_rows(_pf(V1), _pf(values), _pf(Y), _pf(Nv), m, K, sigma, _b(approach), _pf(tt),
                                            C.byref(obj), C.byref(nev)))
    return float(tt[0]), obj.value, nev.value


def classification_fold_rows(V1, values, Y, t: float, sigma: float = 1e-3, tol: float = 1e-5, max_iter: int = 100):
    """The m-sized half of posterior_distribution_classification as the logit drivers call it, folded onto the
    eigenvector rows: (coef[K], Mq[K, K]) with mean = V_new @ coef, cov = rowsum((V_new @ Mq) * V_new) + sigma.  Host only."""
    V1 = np.ascontiguousarray(V1, dtype=np.float64)
    m, K = V1.shape
    values = np.ascontiguousarray(values, dtype=np.float64)[:K].copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    coef = np.zeros(K)
    Mq = np.zeros((K, K), order="F")
    check(_lib.load().flgp_classification_fold_rows(_pf(V1), _pf(values), _pf(Y), m, K, t, sigma, tol, max_iter,
                                                    _pf(coef), _pf(Mq)))
    return coef, Mq


def multi_train_split(Y) -> np.ndarray:
    """multi_train_split (src/MultiClassification.cpp:14-27): m x J one-vs-rest indicator matrix, J = max(Y) + 1."""
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    J = C.c_int()
    check(_lib.load().flgp_multi_train_split(_pf(Y), Y.size, 0, C.byref(J), None))
    aug = np.zeros((Y.size, J.value), order="F")
    check(_lib.load().flgp_multi_train_split(_pf(Y), Y.size, J.value, C.byref(J), _pf(aug)))
    return aug


def negative_log_likelihood(mean, cov, target, type: str = "regression") -> float:
    """negative_log_likelihood (src/Utils.cpp:302-318), type "regression" (the other types sample from R's RNG)."""
    mean = np.ascontiguousarray(mean, dtype=np.float64).reshape(-1)
    cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(-1)
    target = np.ascontiguousarray(target, dtype=np.float64).reshape(-1)
    if not (mean.size == cov.size == target.size):
        raise FlgpError("negative_log_likelihood: mean, cov and target must have the same length")
    out = C.c_double()
    check(_lib.load().flgp_negative_log_likelihood(_pf(mean), _pf(cov), _pf(target), mean.size, _b(type), C.byref(out)))
    return out.value


def test_regression_cpp(Cm, Y, Cnv) -> np.ndarray:
    """test_regression_cpp (src/Predict.cpp:29-37): Cnv C^{-1} Y by Cholesky.  Host only."""
    Cm = _f64(Cm)
    Cnv = _f64(Cnv)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    out = np.zeros(Cnv.shape[0])
    check(_lib.load().flgp_test_regression(_pf(Cm), _pf(Y), _pf(Cnv), Y.size, Cnv.shape[0], _pf(out)))
    return out


test_regression_cpp.__test__ = False  # the reference's export name; not a pytest case


def logit_objective(eigenpair: EigenPair, Y, m_total: int, K: int, t: float, sigma: float = 1e-3,
                    approach: str = "posterior", N=None) -> float:
    """negative_marginal_likelihood_logit_cpp / negative_log_posterior_logit_cpp at diffusion time t
    (src/train.cpp:14-36) on the m labelled rows (labels 0/1, N trials per row)."""
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    obj = C.c_double()
    check(eigenpair.ctx._lib.flgp_logit_objective(eigenpair._h, _pf(Y), _pf(Nv), m_total, K, sigma, _b(approach), t,
                                                  C.byref(obj)))
    return obj.value


def train_lae_logit_gp(eigenpair: EigenPair, Y, m_total: int, K: int, sigma: float = 1e-3,
                       approach: str = "posterior", t0: Optional[float] = None, N=None):
    """train_lae_logit_gp_cpp (src/train.cpp:38-71).  Returns (t, obj, evaluations)."""
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    t = np.array([np.nan if t0 is None else t0], dtype=np.float64)
    obj = C.c_double()
    nev = C.c_int()
    check(eigenpair.ctx._lib.flgp_train_logit(eigenpair._h, _pf(Y), _pf(Nv), m_total, K, sigma, _b(approach), _pf(t),
                                              C.byref(obj), C.byref(nev)))
    return float(t[0]), obj.value, nev.value


def train_logit_mult_gp(eigenpair: EigenPair, Y, m_total: int, K: int, sigma: float = 1e-3,
                        approach: str = "posterior", max_classes: int = 256):
    """train_logit_mult_gp_cpp (src/MultiClassification.cpp:30-53): labels 0 .. J-1, one binary training per class
    against the rest.  Returns (t[J], obj[J])."""
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    t = np.zeros(max_classes)
    obj = np.zeros(max_classes)
    J = C.c_int()
    check(eigenpair.ctx._lib.flgp_train_logit_mult(eigenpair._h, _pf(Y), m_total, K, sigma, _b(approach), max_classes,
                                                   C.byref(J), _pf(t), _pf(obj)))
    return t[:J.value].copy(), obj[:J.value].copy()


def posterior_distribution_multiclassification(eigenpair, Y, m_total: int, K: int, ts, sigma: float = 1e-3, idx_new=None,
                                               block: int = 1 << 16):
    """posterior_distribution_multiclassification (src/Utils.cpp:339-370): per class j the Laplace posterior of the
    one-vs-rest labels at that class's diffusion time ts[j], on the rows idx_new (default: every row after the m_total
    labelled ones).  As in the reference the labelled block C11 carries NO sigma on its diagonal here (the binary
    drivers add it), C22 does.  The m-sized half is the library's host fold (flgp_classification_fold_rows) on the
    labelled rows of the eigenvectors, the rest two products per block of rows.  Returns (mean, cov), len(idx_new) x J."""
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    ts = np.ascontiguousarray(ts, dtype=np.float64).reshape(-1)
    if idx_new is None:
        idx_new = np.arange(m_total, eigenpair.n_local, dtype=np.int32)
    idx_new = np.ascontiguousarray(idx_new, dtype=np.int32)
    values = np.asarray(eigenpair.values, dtype=np.float64)[:K]
    V1 = np.ascontiguousarray(eigenpair.rows(np.arange(m_total, dtype=np.int32))[:, :K])
    folds = [classification_fold_rows(V1, values, (Y == j).astype(np.float64), float(ts[j]), 0.0) for j in range(ts.size)]
    mean = np.zeros((idx_new.size, ts.size))
    cov = np.zeros((idx_new.size, ts.size))
    for b0 in range(0, idx_new.size, block):
        V2 = eigenpair.rows(idx_new[b0:b0 + block])[:, :K]
        for j, (coef, Mq) in enumerate(folds):
            mean[b0:b0 + block, j] = V2 @ coef
            cov[b0:b0 + block, j] = ((V2 @ Mq) * V2).sum(axis=1) + sigma
    return mean, cov


def fit_lae_logit_mult_gp_rcpp(X, Y, X_new, s: int, r: int, K: int = -1, sigma: float = 1e-3, approach="posterior",
                               models=None, nstart: int = 1, *, init_idx=None, seed: int = 0, iter_max: int = 100,
                               ctx: Optional[Context] = None):
    """fit_lae_logit_mult_gp_rcpp (R/Fit.R -> src/Fit.cpp:603-662; BASELINE config 3): spectrum, then J one-vs-rest
    trainings of the diffusion time.  The reference's Y_pred is the arg-max of Polya-Gamma-sampled class probabilities
    (R RNG, stochastic) and is not produced; the deterministic part is returned: the per-class (t_j, objective_j) and
    the per-class Laplace posterior means of the latent function on the test rows, with their arg-max."""
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    X = np.asarray(X, dtype=np.float64)
    X_new = np.asarray(X_new, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    m, m_new = X.shape[0], X_new.shape[0]
    if K < 0:
        K = s
    ep = heat_kernel_spectrum_cpp(X, X_new, s, r, K, models=models, nstart=nstart, init_idx=init_idx, seed=seed,
                                  iter_max=iter_max, ctx=ctx)
    try:
        t, obj = train_logit_mult_gp(ep, Y, m, K, sigma, approach)
        means, covs = posterior_distribution_multiclassification(ep, Y, m, K, t, sigma)   # src/Fit.cpp:655-656
    finally:
        ep.close()
    return {"pars": t, "obj": obj, "posterior": {"mean": means, "cov": covs}, "posterior_mean": means,
            "argmax_posterior_mean": means.argmax(axis=1)}


def fit_lae_logit_gp_rcpp(X, Y, X_new, s: int, r: int, K: int = -1, N=None, sigma: float = 1e-3, approach="posterior",
                          models=None, output_cov: bool = False, nstart: int = 1, *, t: Optional[float] = None,
                          init_idx=None, seed: int = 0, iter_max: int = 100, ctx: Optional[Context] = None):
    """fit_lae_logit_gp_rcpp (R/Fit.R:528-545 -> src/Fit.cpp:521-600): spectrum, empirical-Bayes training of the diffusion
    time t (COBYLA restatement; t given: used as is), Laplace posterior of the test rows.  The reference's Y_pred comes
    from a Polya-Gamma Gibbs sampler on R's RNG (stochastic) and is not produced; the
    deterministic part of the result list is: posterior$mean, posterior$cov, pars (= t), optional C."""
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    m, d = X.shape
    m_new = X_new.shape[0]
    mean = np.zeros(m_new)
    cov = np.zeros(m_new)
    Cm = np.zeros((m + m_new, m), order="F") if output_cov else None
    tt = np.array([np.nan if t is None else t], dtype=np.float64)
    obj = C.c_double()
    check(ctx._lib.flgp_fit_lae_logit(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, r, K, _pf(Nv), sigma,
                                      _b(approach), _b(mo["subsample"]), _b(mo["kernel"]), _gl(mo["gl"]),
                                      int(bool(mo["root"])), nstart, iter_max, _pi(_idx(init_idx)), seed, _pf(tt),
                                      _pf(mean), _pf(cov), _pf(Cm), C.byref(obj)))
    res = {"posterior": {"mean": mean, "cov": cov}, "pars": float(tt[0]), "obj": obj.value}
    if output_cov:
        res["C"] = Cm
    return res


def _default_a2s(a2s):
    if a2s is None:  # R/Fit.R: exp(seq(log(0.1), log(10), length.out = 10))
        a2s = np.exp(np.linspace(np.log(0.1), np.log(10.0), 10))
    return np.ascontiguousarray(a2s, dtype=np.float64)


def fit_se_logit_gp_rcpp(X, Y, X_new, s: int, r: int, K: int = -1, N=None, sigma: float = 1e-3, a2s=None,
                         approach="posterior", models=None, output_cov: bool = False, nstart: int = 1, *,
                         t: Optional[float] = None, init_idx=None, seed: int = 0, iter_max: int = 100,
                         ctx: Optional[Context] = None):
    """fit_se_logit_gp_rcpp (R/Fit.R:598-626 -> src/Fit.cpp:668-794; the README's GPC call): squared-exponential weights on the
    KNN graph, grid search over the bandwidth a2 with the diffusion time t trained per grid point (COBYLA restatement;
    t given: the objective is evaluated there), Laplace posterior of the test rows at the winner.  models["kernel"] is
    ignored, as in the reference.  Y_pred (Polya-Gamma sampler on R's RNG) is not produced;
    returned: posterior$mean, posterior$cov, pars (= t), a2, obj, the winning eigenpair, optional C."""
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    a2s = _default_a2s(a2s)
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    m, d = X.shape
    m_new = X_new.shape[0]
    mean = np.zeros(m_new)
    cov = np.zeros(m_new)
    Cm = np.zeros((m + m_new, m), order="F") if output_cov else None
    tt = np.array([np.nan if t is None else t], dtype=np.float64)
    a2 = C.c_double()
    obj = C.c_double()
    h = C.c_void_p()
    check(ctx._lib.flgp_fit_se_logit(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, r, K, _pf(Nv), sigma,
                                     _pf(a2s), a2s.size, _b(approach), _b(mo["subsample"]), _gl(mo["gl"]),
                                     int(bool(mo["root"])), nstart, iter_max, _pi(_idx(init_idx)), seed, _pf(tt),
                                     _pf(mean), _pf(cov), _pf(Cm), C.byref(a2), C.byref(obj), C.byref(h)))
    res = {"posterior": {"mean": mean, "cov": cov}, "pars": float(tt[0]), "a2": a2.value, "obj": obj.value,
           "eigenpair": EigenPair(ctx, h)}
    if output_cov:
        res["C"] = Cm
    return res


def fit_se_logit_mult_gp_rcpp(X, Y, X_new, s: int, r: int, K: int = -1, sigma: float = 1e-3, a2s=None,
                              approach="posterior", models=None, nstart: int = 1, *, init_idx=None, seed: int = 0,
                              iter_max: int = 100, ctx: Optional[Context] = None):
    """fit_se_logit_mult_gp_rcpp (R/Fit.R:365-380 -> src/Fit.cpp:797-895): the bandwidth grid with the J one-vs-rest trainings
    per grid point; the a2 with the largest summed objective wins.  As for fit_lae_logit_mult_gp_rcpp the sampled labels
    are not produced; returned: per-class (t_j, objective_j) of the winner, a2, the summed objective, the per-class
    Laplace posterior means on the test rows with their arg-max, and the winning eigenpair."""
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    a2s = _default_a2s(a2s)
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    m, d = X.shape
    m_new = X_new.shape[0]
    cap = 1024
    tj = np.zeros(cap)
    oj = np.zeros(cap)
    J = C.c_int()
    a2 = C.c_double()
    obj = C.c_double()
    h = C.c_void_p()
    check(ctx._lib.flgp_fit_se_logit_mult(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, r, K, sigma, _pf(a2s),
                                          a2s.size, _b(approach), _b(mo["subsample"]), _gl(mo["gl"]),
                                          int(bool(mo["root"])), nstart, iter_max, _pi(_idx(init_idx)), seed, cap,
                                          C.byref(J), _pf(tj), _pf(oj), C.byref(a2), C.byref(obj), C.byref(h)))
    ep = EigenPair(ctx, h)
    tj, oj = tj[:J.value].copy(), oj[:J.value].copy()
    Kk = s if K < 0 else K
    means, covs = posterior_distribution_multiclassification(ep, Y, m, Kk, tj, sigma)   # src/Fit.cpp:888-889
    return {"pars": tj, "obj_classes": oj, "obj": obj.value, "a2": a2.value, "posterior": {"mean": means, "cov": covs},
            "posterior_mean": means, "argmax_posterior_mean": means.argmax(axis=1), "eigenpair": ep}


def fit_lae_regression_gp_rcpp(X, Y, X_new, s: int, r: int, K: int = -1, sigma: float = 1e-5, approach="posterior",
                               noise="same", models=None, output_cov: bool = False, nstart: int = 1, *,
                               pars: Optional[Sequence[float]] = None, init_idx=None, seed: int = 0,
                               iter_max: int = 100, ctx: Optional[Context] = None):
    """fit_lae_regression_gp_rcpp (R/Fit.R:56-69 -> src/Fit.cpp:20-99).  pars = (t, noise variance) given: used as
    is; pars = None: trained by empirical Bayes as the reference does (src/train.cpp:557-671; the optimiser is a
    restatement of NLopt's MMA, so trained pars agree with the reference to optimiser tolerance)."""
    if noise not in ("same", "different"):
        raise FlgpError("The noise setting is illegal!")
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    m, d = X.shape
    m_new = X_new.shape[0]
    train = np.zeros(m)
    test = np.zeros(m_new)
    cov = np.zeros(m_new)
    obj = C.c_double()
    if noise == "different":  # pars = (t, noise_1 .. noise_m); src/train.cpp:438-556, src/Predict.cpp:76-113
        x = np.array(pars if pars is not None else np.full(m + 1, np.nan), dtype=np.float64)
        if x.size != m + 1:
            raise FlgpError("noise=\"different\" takes m + 1 parameters (t, one noise variance per training row)")
        check(ctx._lib.flgp_fit_lae_regression_diff_noise(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, r, K, sigma,
                                                          _b(approach), _b(mo["subsample"]), _b(mo["kernel"]),
                                                          _gl(mo["gl"]), int(bool(mo["root"])), nstart, iter_max,
                                                          _pi(_idx(init_idx)), seed, _pf(x), _pf(train), _pf(test),
                                                          _pf(cov), C.byref(obj)))
        return {"Y_pred": {"train": train, "test": test}, "posterior": {"mean": test, "cov": cov}, "pars": list(x),
                "obj": obj.value}
    x = np.array(pars if pars is not None else (np.nan, np.nan), dtype=np.float64)
    check(ctx._lib.flgp_fit_lae_regression(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, r, K, sigma,
                                           _b(approach), _b(mo["subsample"]), _b(mo["kernel"]), _gl(mo["gl"]),
                                           int(bool(mo["root"])), nstart, iter_max, _pi(_idx(init_idx)), seed, _pf(x),
                                           _pf(train), _pf(test), _pf(cov), C.byref(obj)))
    res = {"Y_pred": {"train": train, "test": test}, "posterior": {"mean": test, "cov": cov}, "pars": list(x),
           "obj": obj.value}
    if output_cov:
        res["C"] = heat_kernel_covariance_rcpp(X, X_new, s, r, x[0], K, mo, nstart, init_idx=init_idx, seed=seed,
                                               iter_max=iter_max, ctx=ctx)
    return res


def fit_se_regression_gp_rcpp(X, Y, X_new, s: int, r: int, K: int = -1, sigma: float = 1e-5, a2s=None,
                              approach="posterior", noise="same", models=None, output_cov: bool = False,
                              nstart: int = 1, *, pars: Optional[Sequence[float]] = None, init_idx=None, seed: int = 0,
                              iter_max: int = 100, ctx: Optional[Context] = None):
    """fit_se_regression_gp_rcpp (R/Fit.R:119-136 -> src/Fit.cpp:102-219): squared-exponential weights on the KNN
    graph with a grid search over the bandwidth a2 (default exp(seq(log 0.1, log 10, length 10))); one k-means and one
    KNN serve the whole grid.  models["kernel"] is ignored, as in the reference."""
    if noise != "same":
        raise FlgpError("The noise setting is illegal!" if noise != "different"
                        else "noise=\"different\" (one variance per training point) is not part of this path")
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    ctx = ctx or default_ctx()
    mo = dict(DEFAULT_MODELS)
    mo.update(models or {})
    if a2s is None:
        a2s = np.exp(np.linspace(np.log(0.1), np.log(10.0), 10))
    a2s = np.ascontiguousarray(a2s, dtype=np.float64)
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    m, d = X.shape
    m_new = X_new.shape[0]
    train = np.zeros(m)
    test = np.zeros(m_new)
    cov = np.zeros(m_new)
    xo = np.zeros(2)
    fixed = np.ascontiguousarray(pars, dtype=np.float64) if pars is not None else None
    a2 = C.c_double()
    obj = C.c_double()
    h = C.c_void_p()
    check(ctx._lib.flgp_fit_se_regression(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, r, K, sigma, _pf(a2s),
                                          a2s.size, _b(approach), _b(mo["subsample"]), _gl(mo["gl"]),
                                          int(bool(mo["root"])), nstart, iter_max, _pi(_idx(init_idx)), seed,
                                          _pf(fixed), _pf(train), _pf(test), _pf(cov), _pf(xo), C.byref(a2),
                                          C.byref(obj), C.byref(h)))
    ep = EigenPair(ctx, h)
    res = {"Y_pred": {"train": train, "test": test}, "posterior": {"mean": test, "cov": cov}, "pars": list(xo),
           "a2": a2.value, "obj": obj.value, "eigenpair": ep}
    if output_cov:
        n = m + m_new
        Kk = s if K < 0 else K
        res["C"] = HK_from_spectrum_cpp(ep, Kk, xo[0], np.arange(n, dtype=np.int32), np.arange(m, dtype=np.int32))
    return res


def fit_nystrom_regression_gp_rcpp(X, Y, X_new, s: int, K: int = -1, sigma: float = 1e-5, a2s=None,
                                   approach="posterior", noise="same", subsample="kmeans", output_cov: bool = False,
                                   nstart: int = 1, *, pars: Optional[Sequence[float]] = None, init_idx=None,
                                   seed: int = 0, iter_max: int = 100, ctx: Optional[Context] = None):
    """fit_nystrom_regression_gp_rcpp (R/Fit.R:177-195 -> src/Fit.cpp:222-357): the Nystrom-extension baseline —
    dense SE kernel on the s anchors, top-K eigenpairs, extension of the eigenvectors to every row (an n x s x K
    product on the FP64 tensor cores, in row blocks), bandwidth grid, training, GPR tail."""
    if noise != "same":
        raise FlgpError("The noise setting is illegal!" if noise != "different"
                        else "noise=\"different\" (one variance per training point) is not part of this path")
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    if output_cov:
        raise FlgpError("output_cov is not offloaded for the Nystrom driver")
    ctx = ctx or default_ctx()
    if a2s is None:
        a2s = np.exp(np.linspace(np.log(0.1), np.log(10.0), 10))
    a2s = np.ascontiguousarray(a2s, dtype=np.float64)
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    m, d = X.shape
    m_new = X_new.shape[0]
    train = np.zeros(m)
    test = np.zeros(m_new)
    cov = np.zeros(m_new)
    xo = np.zeros(2)
    fixed = np.ascontiguousarray(pars, dtype=np.float64) if pars is not None else None
    a2 = C.c_double()
    obj = C.c_double()
    check(ctx._lib.flgp_fit_nystrom_regression(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, K, sigma, _pf(a2s),
                                               a2s.size, _b(approach), _b(subsample), nstart, iter_max,
                                               _pi(_idx(init_idx)), seed, _pf(fixed), _pf(train), _pf(test), _pf(cov),
                                               _pf(xo), C.byref(a2), C.byref(obj)))
    return {"Y_pred": {"train": train, "test": test}, "posterior": {"mean": test, "cov": cov}, "pars": list(xo),
            "a2": a2.value, "obj": obj.value}


def fit_nystrom_logit_gp_rcpp(X, Y, X_new, s: int, K: int = -1, N=None, sigma: float = 1e-3, a2s=None,
                              approach="posterior", subsample="kmeans", output_cov: bool = False, nstart: int = 1, *,
                              t: Optional[float] = None, init_idx=None, seed: int = 0, iter_max: int = 100,
                              ctx: Optional[Context] = None):
    """fit_nystrom_logit_gp_rcpp (R/Fit.R:661-690 -> src/Fit.cpp:896-1038) without the Polya-Gamma labels: Nystrom grid with
    the diffusion time trained per bandwidth (t given: the objective there), Laplace posterior of the test rows at the
    winner.  Returned: posterior$mean, posterior$cov, pars (= t), a2, obj, optional C."""
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    ctx = ctx or default_ctx()
    a2s = _default_a2s(a2s)
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    Nv = np.ascontiguousarray(N, dtype=np.float64).reshape(-1) if N is not None else None
    m, d = X.shape
    m_new = X_new.shape[0]
    mean = np.zeros(m_new)
    cov = np.zeros(m_new)
    Cm = np.zeros((m + m_new, m), order="F") if output_cov else None
    tt = np.array([np.nan if t is None else t], dtype=np.float64)
    a2 = C.c_double()
    obj = C.c_double()
    check(ctx._lib.flgp_fit_nystrom_logit(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, K, _pf(Nv), sigma,
                                          _pf(a2s), a2s.size, _b(approach), _b(subsample), nstart, iter_max,
                                          _pi(_idx(init_idx)), seed, _pf(tt), _pf(mean), _pf(cov), _pf(Cm),
                                          C.byref(a2), C.byref(obj)))
    res = {"posterior": {"mean": mean, "cov": cov}, "pars": float(tt[0]), "a2": a2.value, "obj": obj.value}
    if output_cov:
        res["C"] = Cm
    return res


def fit_nystrom_logit_mult_gp_rcpp(X, Y, X_new, s: int, K: int = -1, sigma: float = 1e-3, a2s=None,
                                   approach="posterior", subsample="kmeans", nstart: int = 1, *,
                                   return_eigenpair: bool = False, init_idx=None, seed: int = 0, iter_max: int = 100,
                                   ctx: Optional[Context] = None):
    """The training half of fit_nystrom_logit_mult_gp_rcpp (R/Fit.R:424-435 -> src/Fit.cpp:1045-1162): per bandwidth the J one-vs-rest
    trainings on the extended labelled rows, summed objective selects.  Returned: pars (t_j), obj_classes, obj, a2 and,
    with return_eigenpair, the winning extended eigenpair (values K, vectors n x K) the reference's sampler consumes."""
    if approach not in ("posterior", "marginal"):
        raise FlgpError("This model selection approach is not supported!")
    ctx = ctx or default_ctx()
    a2s = _default_a2s(a2s)
    X = _f64(X)
    X_new = _f64(X_new)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    m, d = X.shape
    m_new = X_new.shape[0]
    Kk = s if K < 0 else K
    cap = 1024
    tj = np.zeros(cap)
    oj = np.zeros(cap)
    J = C.c_int()
    a2 = C.c_double()
    obj = C.c_double()
    values = np.zeros(Kk) if return_eigenpair else None
    vectors = np.zeros((m + m_new, Kk), order="F") if return_eigenpair else None
    check(ctx._lib.flgp_fit_nystrom_logit_mult(ctx._h, _pf(X), _pf(Y), _pf(X_new), m, m_new, d, s, K, sigma, _pf(a2s),
                                               a2s.size, _b(approach), _b(subsample), nstart, iter_max,
                                               _pi(_idx(init_idx)), seed, cap, C.byref(J), _pf(tj), _pf(oj),
                                               _pf(values), _pf(vectors), C.byref(a2), C.byref(obj)))
    res = {"pars": tj[:J.value].copy(), "obj_classes": oj[:J.value].copy(), "obj": obj.value, "a2": a2.value}
    if return_eigenpair:
        res["values"], res["vectors"] = values, vectors
    return res


def fit_nystrom_regression_sharded(X_local, n_total: int, row_offset: int, Y_local, m_total: int, s: int, K: int = -1,
                                   sigma: float = 1e-5, a2s=None, approach="posterior", subsample="kmeans",
                                   nstart: int = 1, *, pars: Optional[Sequence[float]] = None, init_idx=None,
                                   seed: int = 0, iter_max: int = 100, ctx: Optional[Context] = None):
    """fit_nystrom_regression_gp_cpp on one contiguous block of rows per rank (multi-GPU; BASELINE config 5's Nystrom
    variant): returns the posterior mean / variance of the local rows, pars, a2, obj."""
    ctx = ctx or default_ctx()
    if a2s is None:
        a2s = np.exp(np.linspace(np.log(0.1), np.log(10.0), 10))
    a2s = np.ascontiguousarray(a2s, dtype=np.float64)
    X_local = _f64(X_local)
    n_local, d = X_local.shape
    Y_local = np.ascontiguousarray(Y_local, dtype=np.float64).reshape(-1)
    mean = np.zeros(max(n_local, 1))
    cov = np.zeros(max(n_local, 1))
    xo = np.zeros(2)
    fixed = np.ascontiguousarray(pars, dtype=np.float64) if pars is not None else None
    a2 = C.c_double()
    obj = C.c_double()
    check(ctx._lib.flgp_fit_nystrom_regression_sharded(ctx._h, _pf(X_local), n_local, n_total, row_offset, d,
                                                       _pf(Y_local) if Y_local.size else None, m_total, s, K, sigma,
                                                       _pf(a2s), a2s.size, _b(approach), _b(subsample), nstart, iter_max,
                                                       _pi(_idx(init_idx)), seed, _pf(fixed), _pf(mean), _pf(cov),
                                                       _pf(xo), C.byref(a2), C.byref(obj)))
    return {"mean": mean[:n_local], "cov": cov[:n_local], "pars": list(xo), "a2": a2.value, "obj": obj.value}


def posterior_distribution_classification(eigenpair: EigenPair, Y_local, m_total: int, K: int, t: float,
                                          sigma: float = 1e-3, tol: float = 1e-5, max_iter: int = 100):
    """posterior_distribution_classification (src/Utils.cpp:252-299) on a spectrum handle at a fixed diffusion time
    t, as fit_lae_logit_gp_cpp calls it (src/Fit.cpp:563-582): Laplace posterior mean and variance of the latent
    function at every local row (training rows first).  Returns (mean, cov)."""
    Y_local = np.ascontiguousarray(Y_local, dtype=np.float64).reshape(-1)
    mean = np.zeros(eigenpair.n_local)
    cov = np.zeros(eigenpair.n_local)
    check(eigenpair.ctx._lib.flgp_classification_posterior_fixed(eigenpair._h, _pf(Y_local), m_total, K, t, sigma, tol,
                                                                 max_iter, _pf(mean), _pf(cov)))
    return mean, cov


def posterior_distribution_classification_rcpp(C11, C21, C22, Y, tol: float = 1e-5, max_iter: int = 100, *,
                                               ctx: Optional[Context] = None):
    """posterior_distribution_classification(C11, C21, C22, Y, tol, max_iter) — the reference's export on explicit
    covariance blocks (src/Utils.h:77-80, src/Utils.cpp:252-299) -> {"mean", "cov"}."""
    ctx = ctx or default_ctx()
    C11 = _f64(C11)
    C21 = _f64(C21)
    C22 = np.ascontiguousarray(C22, dtype=np.float64).reshape(-1)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    m, m_new = C11.shape[0], C21.shape[0]
    if C11.shape != (m, m) or C21.shape != (m_new, m) or C22.size != m_new or Y.size != m:
        raise FlgpError("posterior_distribution_classification: inconsistent shapes")
    mean = np.zeros(m_new)
    cov = np.zeros(m_new)
    check(ctx._lib.flgp_posterior_distribution_classification(ctx._h, _pf(C11), _pf(C21), _pf(C22), _pf(Y), m, m_new,
                                                              tol, max_iter, _pf(mean), _pf(cov)))
    return {"mean": mean, "cov": cov}
