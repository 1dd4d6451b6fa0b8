// capi.cu — the C ABI (include/flgp.h) and the host-side orchestration that mirrors the reference's
// C++ seams: subsample_cpp, KNN_cpp, LAE_cpp, graphLaplacian_cpp, cross_similarity_{lae,se}_cpp,
// spectrum_from_Z_cpp, heat_kernel_spectrum_cpp, HK_from_spectrum_cpp, lae_eigenmap,
// heat_kernel_covariance_cpp and the fixed-hyper-parameter GPR tail (file:line in include/flgp.h).
// Error behaviour mirrors Rcpp::stop: C++ exception inside, status + message at the boundary.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <atomic>
#include <thread>
#include <unordered_set>

#include "../../include/flgp.h"
#include "kernels.cuh"

using namespace flgp;

namespace flgp {
void comm_unique_id(void* out128);
void comm_init(Ctx* c, const void* id128, int rank, int nranks);
}  // namespace flgp

struct flgp_ctx {
  Ctx c;
};

struct flgp_spectrum {
  Ctx* c = nullptr;
  int64_t n_local = 0, n_total = 0, row_offset = 0;
  int d = 0, s = 0, r = 0, K = 0, ucols = 0;
  bool root = true;
  DevBuf<int32_t> Zj;
  DevBuf<double> Zx;   // final Z (after graph-Laplacian scaling), n_local * r
  DevBuf<double> w;    // s: A = Z diag(w)
  DevBuf<double> Wm;   // s x K row-major: Y(:,k) * sqrt(n_total) / sigma_k   (lift operator)
  DevBuf<double> U;    // s x ucols anchors (+ sizes)
  std::vector<double> values;  // K, as exported (sigma or sigma^2)
  int kmeans_iters = 0;
  long long lae_iters = 0, lae_bts = 0;
  KMeansSorted sorted;  // cluster-sorted rows left by k-means; consumed (and released) by the KNN stage
};

namespace {

thread_local std::string g_err;

template <class F>
int guard(F f) {
  try {
    f();
    return 0;
  } catch (const Error& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// every entry point works on its context's device, whatever device the calling thread had current
Ctx* on_device(Ctx* c) {
  FLGP_CUDA(cudaSetDevice(c->device));
  return c;
}

void need(bool ok, const char* msg) {
  if (!ok) fail(2, "%s", msg);
}

int parse_gl(int gl) {
  if (gl < 0 || gl > 2) fail(2, "Error: the type of graph Laplacian is not supported!");
  return gl;
}

std::vector<int32_t> default_init(int64_t n, int s, uint64_t seed) {
  if (s < 1 || s > n) fail(2, "subsample: need 1 <= s <= n");
  std::unordered_set<int64_t> seen;
  std::vector<int32_t> idx;
  idx.reserve(s);
  uint64_t state = seed;
  while ((int)idx.size() < s) {
    state += 0x9E3779B97F4A7C15ull;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    int64_t i = (int64_t)(z % (uint64_t)n);
    if (seen.insert(i).second) idx.push_back((int32_t)i);
  }
  std::sort(idx.begin(), idx.end());
  return idx;
}

// ---- small dense host algebra for the m- and K-sized GP back-end (reference: Eigen LLT) ----------
// column-major n x n, lower Cholesky in place; returns false if not positive definite
// Right-looking in panels of 32 columns, every inner loop down a contiguous column, four panel columns per sweep of a
// trailing column; each element still receives its updates in ascending k, so the factor is the same bit pattern as
// from the textbook dot-product form (the training objectives run hundreds of m x m factorisations: m = 1000 at
// config 3).  For large trailing blocks the columns are dealt round-robin to host threads — a column is always updated
// by one thread in the same order, so the bits do not depend on the thread count.
thread_local int g_outer_workers = 1;  // how many sibling host threads run dense algebra at the same time

// host threads this call may use: the machine's cores (or FLGP_HOST_THREADS) divided among the sibling workers, <= 16
static int host_threads() {
  static const int hw = [] {
    const char* e = std::getenv("FLGP_HOST_THREADS");  // 1 = no threading inside the host algebra
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : (int)std::max(1u, std::thread::hardware_concurrency());
  }();
  return std::max(1, std::min(16, hw / std::max(1, g_outer_workers)));
}
// fn(begin, step): the caller's loop over items begin, begin + step, ... — one item is always done by one thread
template <class Fn>
static void host_parallel(int T, Fn fn) {
  if (T <= 1) {
    fn(0, 1);
    return;
  }
  std::vector<std::thread> th;
  for (int q = 1; q < T; ++q) th.emplace_back(fn, q, T);
  fn(0, T);
  for (auto& t : th) t.join();
}

static void chol_trailing_columns(double* A, int n, int p0, int p1, int j_begin, int j_step) {
  for (int j = j_begin; j < n; j += j_step) {
    double* aj = A + (size_t)n * j;
    int k = p0;
    for (; k + 4 <= p1; k += 4) {
      const double* a0 = A + (size_t)n * k;
      const double* a1 = a0 + n;
      const double* a2 = a1 + n;
      const double* a3 = a2 + n;
      const double l0 = a0[j], l1 = a1[j], l2 = a2[j], l3 = a3[j];
      for (int i = j; i < n; ++i) {
        double v = aj[i];
        v -= a0[i] * l0;
        v -= a1[i] * l1;
        v -= a2[i] * l2;
        v -= a3[i] * l3;
        aj[i] = v;
      }
    }
    for (; k < p1; ++k) {
      const double* ak = A + (size_t)n * k;
      const double ljk = ak[j];
      for (int i = j; i < n; ++i) aj[i] -= ak[i] * ljk;
    }
  }
}

bool chol_lower(std::vector<double>& A, int n) {
  constexpr int NB = 32;
  const int tmax = host_threads();
  for (int p0 = 0; p0 < n; p0 += NB) {
    const int p1 = std::min(n, p0 + NB);
    for (int j = p0; j < p1; ++j) {  // the panel's own columns (earlier panels are already applied)
      double* aj = &A[(size_t)n * j];
      for (int k = p0; k < j; ++k) {
        const double* ak = &A[(size_t)n * k];
        const double ljk = ak[j];
        for (int i = j; i < n; ++i) aj[i] -= ak[i] * ljk;
      }
      double dj = aj[j];
      if (!(dj > 0.0)) return false;
      dj = std::sqrt(dj);
      aj[j] = dj;
      for (int i = j + 1; i < n; ++i) aj[i] = aj[i] / dj;
    }
    const int rest = n - p1;
    const int T = rest >= 384 ? std::min(tmax, rest / 96) : 1;
    host_parallel(T, [&](int q, int step) { chol_trailing_columns(A.data(), n, p0, p1, p1 + q, step); });
  }
  return true;
}
// solve L L^T x = b in place for nrhs columns (B column-major n x nrhs)
void chol_solve(const std::vector<double>& L, int n, double* B, int nrhs) {
  for (int c = 0; c < nrhs; ++c) {
    double* b = B + (size_t)n * c;
    for (int k = 0; k < n; ++k) {  // forward: column k of L leaves the rows below it
      const double* lk = &L[(size_t)n * k];
      const double bk = b[k] / lk[k];
      b[k] = bk;
      for (int i = k + 1; i < n; ++i) b[i] -= lk[i] * bk;
    }
    for (int i = n - 1; i >= 0; --i) {
      double v = b[i];
      for (int k = i + 1; k < n; ++k) v -= L[k + (size_t)n * i] * b[k];
      b[i] = v / L[i + (size_t)n * i];
    }
  }
}

#include "train.inl"

// deterministic sum of len doubles: fixed grid, per-thread sequential partial sums, fixed-order tree
__global__ void __launch_bounds__(256) sum_partial_kernel(const double* __restrict__ v, int64_t len, double* part) {
  __shared__ double sh[256];
  double a = 0.0;
  for (int64_t e = blockIdx.x * 256 + threadIdx.x; e < len; e += (int64_t)gridDim.x * 256) a += v[e];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void sum_final_kernel(const double* __restrict__ part, int np, double* out) {
  if (threadIdx.x || blockIdx.x) return;
  double a = 0.0;
  for (int q = 0; q < np; ++q) a += part[q];
  *out = a;
}

__global__ void build_lift_kernel(const double* __restrict__ Y, int s, int K, const double* __restrict__ scale,
                                  double* __restrict__ Wm) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s * K) return;
  int k = e % K, cidx = e / K;
  Wm[e] = Y[cidx + (size_t)s * k] * scale[k];
}

__global__ void gather_rows_kernel(const double* __restrict__ src, int64_t ld, int cols, const int32_t* __restrict__ idx,
                                   int s, double* __restrict__ dst) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s * cols) return;
  int j = e % s, k = e / s;
  dst[j + (size_t)s * k] = src[idx[j] + ld * k];
}

// ---- stages on device buffers ----------------------------------------------------------------------
struct Models {
  std::string subsample, kernel;
  int gl;
  bool root;
  double epsilon;
  int nstart, iter_max;
};

// anchors: fills sp->U (s x ucols)
void stage_subsample(Ctx* c, flgp_spectrum* sp, const double* X, const Models& mo, const int32_t* init_idx,
                     uint64_t seed, int32_t* assign_out) {
  const int s = sp->s, d = sp->d;
  std::vector<int32_t> own;
  if (!init_idx) {
    own = default_init(sp->n_total, s, seed);
    init_idx = own.data();
  }
  if (mo.subsample == "kmeans") {
    need(mo.nstart >= 1, "nstart must be >= 1");
    sp->ucols = d + 1;
    sp->U.alloc((size_t)s * (d + 1));
    DevBuf<int32_t> assign;
    int32_t* ap = assign_out;
    if (!ap) {
      assign.alloc(std::max<int64_t>(sp->n_local, 1));
      ap = assign.p;
    }
    StageScope st(c, "kmeans");
    kmeans_run(c, X, sp->n_local, sp->n_local, d, s, sp->n_total, sp->row_offset, init_idx, mo.iter_max, sp->U.p, ap,
               &sp->kmeans_iters, &sp->sorted);
    if (mo.nstart > 1) {
      // stats::kmeans(nstart = k) (src/Utils.cpp:37-42): k runs from different random starts, the one with the
      // smallest total within-cluster sum of squares is returned.  Start 0 = init_idx (or the seed's default
      // rows), start q > 0 = default rows of a seed derived from (seed, q); the comparison is on exact
      // fixed-point sums, so every rank keeps the same run.
      double best = kmeans_withinss_run(c, X, sp->n_local, sp->n_local, d, sp->U.p, s, ap, sp->n_total);
      DevBuf<double> U2((size_t)s * (d + 1));
      DevBuf<int32_t> a2(std::max<int64_t>(sp->n_local, 1));
      for (int q = 1; q < mo.nstart; ++q) {
        const std::vector<int32_t> init_q = default_init(sp->n_total, s, seed * 0x9E3779B97F4A7C15ull + (uint64_t)q);
        KMeansSorted srt;
        int iters = 0;
        kmeans_run(c, X, sp->n_local, sp->n_local, d, s, sp->n_total, sp->row_offset, init_q.data(), mo.iter_max, U2.p,
                   a2.p, &iters, &srt);
        const double w = kmeans_withinss_run(c, X, sp->n_local, sp->n_local, d, U2.p, s, a2.p, sp->n_total);
        if (w < best) {
          best = w;
          FLGP_CUDA(cudaMemcpyAsync(sp->U.p, U2.p, sizeof(double) * s * (d + 1), cudaMemcpyDeviceToDevice, c->stream));
          FLGP_CUDA(cudaMemcpyAsync(ap, a2.p, sizeof(int32_t) * sp->n_local, cudaMemcpyDeviceToDevice, c->stream));
          sync(c);
          sp->kmeans_iters = iters;
          sp->sorted = std::move(srt);
        }
      }
    }
    if (st.idx >= 0) {
      c->stages[st.idx].flops = 2.0 * s * d * (double)sp->n_local * sp->kmeans_iters;
      c->stages[st.idx].bytes = (8.0 * d + 4.0) * (double)sp->n_local * sp->kmeans_iters;
    }
  } else if (mo.subsample == "random") {
    if (mo.gl == FLGP_GL_CLUSTER_NORMALIZED)
      fail(2, "subsample=\"random\" returns no cluster sizes; it cannot be combined with gl=\"cluster-normalized\"");
    need(c->nranks == 1, "subsample=\"random\" is single-GPU only");
    sp->ucols = d;
    sp->U.alloc((size_t)s * d);
    for (int j = 0; j < s; ++j) need(init_idx[j] >= 0 && init_idx[j] < sp->n_local, "initial index out of range");
    DevBuf<int32_t> idx(s);
    idx.upload(init_idx, s, c->stream);
    FLGP_LAUNCH(c, gather_rows_kernel, ceil_div(s * d, 256), 256, 0, X, sp->n_local, d, idx.p, s, sp->U.p);
    sync(c);
  } else if (mo.subsample == "minibatchkmeans") {
    // src/Utils.cpp:49-62: centroids by mini-batch k-means (ClusterR un-vendored: the contract of minibatch.cu), sizes
    // by the reference's own 1-NN count.  ClusterR's num_init picks among kmeans++ starts; this path takes explicit
    // start rows, so there is nothing for nstart > 1 to select from.
    need(mo.nstart == 1, "subsample=\"minibatchkmeans\" takes one explicit start (nstart must be 1)");
    need(c->nranks == 1, "subsample=\"minibatchkmeans\" is single-GPU only");
    sp->ucols = d + 1;
    sp->U.alloc((size_t)s * (d + 1));
    StageScope st(c, "minibatchkmeans");
    minibatch_kmeans_run(c, X, sp->n_local, sp->n_local, d, s, init_idx, mo.iter_max, seed, sp->U.p, &sp->kmeans_iters);
  } else {
    fail(2, "The subsample method is not supported!");
  }
}

// Z for the local rows (before graph-Laplacian scaling) into sp->Zj / sp->Zx
void stage_cross_similarity(Ctx* c, flgp_spectrum* sp, const double* X, const double* U, const std::string& kernel,
                            double epsilon) {
  const int64_t n = sp->n_local;
  const int s = sp->s, d = sp->d, r = sp->r;
  sp->Zj.alloc(std::max<int64_t>(n * r, 1));
  sp->Zx.alloc(std::max<int64_t>(n * r, 1));
  DevBuf<int32_t> ind(std::max<int64_t>(n * r, 1));
  bool srt = false;  // are ind / dist in cluster-sorted order?
  if (kernel == "lae") {
    {
      StageScope st(c, "knn", 2.0 * s * d * (double)n, (8.0 * d + 4.0 * r) * (double)n);
      knn_run(c, X, n, n, d, U, s, s, r, ind.p, nullptr, &sp->sorted, &srt);
    }
    DevBuf<long long> stats(2);
    stats.zero(c->stream);
    {
      // after a pruned KNN the neighbours are in cluster-sorted order: the solver runs in that order too (coalesced
      // reads, warps whose points share anchors and converge alike) and writes each CSR row at its original place
      StageScope st(c, "lae", 0.0, (8.0 * d + 4.0 * r + 12.0 * r) * (double)n);
      if (srt) lae_run(c, sp->sorted.Xs.p, n, n, d, U, s, s, r, ind.p, sp->Zj.p, sp->Zx.p, nullptr, stats.p,
                       sp->sorted.perm.p);
      else lae_run(c, X, n, n, d, U, s, s, r, ind.p, sp->Zj.p, sp->Zx.p, nullptr, stats.p);
    }
    sync(c);
    sp->sorted = KMeansSorted();
    long long h[2];
    stats.download(h, 2, c->stream);
    sync(c);
    sp->lae_iters = h[0];
    sp->lae_bts = h[1];
  } else if (kernel == "se") {
    DevBuf<double> dist(std::max<int64_t>(n * r, 1));
    {
      StageScope st(c, "knn", 2.0 * s * d * (double)n, (8.0 * d + 12.0 * r) * (double)n);
      knn_run(c, X, n, n, d, U, s, s, r, ind.p, dist.p, &sp->sorted, &srt);
    }
    StageScope st(c, "se_weights", 0.0, 36.0 * r * (double)n);
    knn_to_csr_run(c, n, r, ind.p, dist.p, sp->Zj.p, sp->Zx.p, srt ? sp->sorted.perm.p : nullptr);
    sync(c);
    sp->sorted = KMeansSorted();
    se_weights_run(c, sp->Zx.p, n * r, 4.0 * epsilon * epsilon, sp->Zx.p);
    sync(c);
  } else {
    fail(2, "The kernel type is not supported!");
  }
}

void stage_graph_laplacian(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, double* Zx, int gl,
                           const double* num_class, int64_t n_total, double vmax = 1.0) {
  StageScope st(c, "graph_laplacian", 0.0, (gl >= 1 ? 36.0 : 24.0) * r * (double)n);
  DevBuf<double> cs(s);
  if (gl >= 1) colsum_run(c, n, s, r, Zj, Zx, n_total, cs.p, vmax);
  gl_apply_run(c, n, s, r, Zj, Zx, gl, cs.p, num_class);
  sync(c);
}

// spectrum_from_Z_cpp on sp->Zj/Zx: fills w, Wm, values
// zmax > 0: Z came from the caller (any finite values): the fixed-point scales follow max |Z| and max w
void stage_spectrum(Ctx* c, flgp_spectrum* sp, int K, bool root, double zmax = 0.0) {
  const int s = sp->s, r = sp->r;
  const int64_t n = sp->n_local;
  if (K < 0) K = s;
  need(K >= 1 && K <= s, "need 1 <= K <= s");
  sp->K = K;
  sp->root = root;
  sp->w.alloc(s);
  DevBuf<double> G((size_t)s * s);
  {
    StageScope st(c, "gram", 2.0 * r * r * (double)n, 24.0 * r * (double)n + 8.0 * s * s);
    DevBuf<double> cs(s);
    colsum_run(c, n, s, r, sp->Zj.p, sp->Zx.p, sp->n_total, cs.p, zmax > 0.0 ? zmax : 1.0);
    spectrum_scale_run(c, s, cs.p, sp->w.p);
    double pmax = 1.0;
    if (zmax > 0.0) {
      std::vector<double> wh(s);
      sp->w.download(wh.data(), s, c->stream);
      sync(c);
      double wmax = 0.0;
      for (double v : wh) wmax = std::max(wmax, std::fabs(v));
      pmax = zmax * wmax * zmax * wmax * (1.0 + 1e-12) + 1e-300;
    }
    gram_run(c, n, s, r, sp->Zj.p, sp->Zx.p, sp->w.p, sp->n_total, G.p, pmax);
  }
  DevBuf<double> lam(K), Y((size_t)s * K);
  {
    StageScope st(c, "eigh", (4.0 / 3.0) * s * (double)s * s + 2.0 * s * (double)s * K, 8.0 * s * (double)s * s);
    eigh_topk_run(c, G.p, s, K, lam.p, Y.p, /*psd=*/true);
  }
  std::vector<double> lam_h(K), scale(K);
  lam.download(lam_h.data(), K, c->stream);
  sync(c);
  sp->values.resize(K);
  const double sq = std::sqrt((double)sp->n_total);
  for (int k = 0; k < K; ++k) {
    double l = lam_h[k] > 0.0 ? lam_h[k] : 0.0;
    double sg = std::sqrt(l);
    sp->values[k] = root ? sg : l;           // src/Spectrum.cpp:153-155
    scale[k] = sg > 0.0 ? sq / sg : 0.0;     // src/Spectrum.cpp:157-158 (sqrt(n)) and u = A v / sigma
  }
  DevBuf<double> dscale(K);
  dscale.upload(scale.data(), K, c->stream);
  sp->Wm.alloc((size_t)s * K);
  FLGP_LAUNCH(c, build_lift_kernel, ceil_div(s * K, 256), 256, 0, Y.p, s, K, dscale.p, sp->Wm.p);
  sync(c);
}

std::unique_ptr<flgp_spectrum> spectrum_pipeline(Ctx* c, const double* Xdev, int64_t n_local, int64_t n_total,
                                                 int64_t row_offset, int d, int s, int r, int K, const Models& mo,
                                                 const int32_t* init_idx, uint64_t seed) {
  need(n_local >= 0 && n_total >= 1 && d >= 1, "bad matrix shape");
  need(row_offset >= 0 && row_offset + n_local <= n_total, "shard outside the matrix");
  need(s >= 1 && s <= n_total, "need 1 <= s <= n");
  need(r >= 1 && r <= s, "need 1 <= r <= s");
  need(n_total < ((int64_t)1 << 31), "n must fit in int32 indices per process API");
  parse_gl(mo.gl);
  std::unique_ptr<flgp_spectrum> sp(new flgp_spectrum);
  sp->c = c;
  sp->n_local = n_local;
  sp->n_total = n_total;
  sp->row_offset = row_offset;
  sp->d = d;
  sp->s = s;
  sp->r = r;
  stage_subsample(c, sp.get(), Xdev, mo, init_idx, seed, nullptr);
  stage_cross_similarity(c, sp.get(), Xdev, sp->U.p, mo.kernel, mo.epsilon);
  const double* nc = (mo.gl == FLGP_GL_CLUSTER_NORMALIZED) ? sp->U.p + (size_t)s * d : nullptr;
  stage_graph_laplacian(c, n_local, s, r, sp->Zj.p, sp->Zx.p, mo.gl, nc, n_total);
  stage_spectrum(c, sp.get(), K, mo.root);
  return sp;
}

// upload [X; X_new] as one column-major n x d device matrix (the concat of src/Spectrum.cpp:50-53)
DevBuf<double> upload_concat(Ctx* c, const double* X, int64_t m, const double* X_new, int64_t m_new, int d) {
  const int64_t n = m + m_new;
  DevBuf<double> dX((size_t)std::max<int64_t>(n * d, 1));
  if (m > 0)
    FLGP_CUDA(cudaMemcpy2DAsync(dX.p, n * sizeof(double), X, m * sizeof(double), m * sizeof(double), d,
                                cudaMemcpyHostToDevice, c->stream));
  if (m_new > 0)
    FLGP_CUDA(cudaMemcpy2DAsync(dX.p + m, n * sizeof(double), X_new, m_new * sizeof(double), m_new * sizeof(double), d,
                                cudaMemcpyHostToDevice, c->stream));
  return dX;
}

Models make_models(const char* subsample, const char* kernel, int gl, int root, int nstart, double epsilon,
                   int iter_max) {
  Models mo;
  mo.subsample = subsample ? subsample : "kmeans";
  mo.kernel = kernel ? kernel : "lae";
  mo.gl = gl;
  mo.root = root != 0;
  mo.epsilon = epsilon;
  mo.nstart = nstart;
  mo.iter_max = iter_max > 0 ? iter_max : 100;
  return mo;
}

// ---- GPR tail -------------------------------------------------------------------------------------
// Everything n-sized is folded through the r non-zeros of each row of A:
//   mean_i = a_i . (Wm coef),   var_i = add + a_i^T (Wm M Wm^T) a_i,
// coef (K) and M (K x K) come from the K x K (m > K, Woodbury) or m x m (m <= K) system on the host.
// coef (KK, zero padded beyond K) and M (KK x KK) of the GPR tail from the training rows V1 (row-major
// m_local x KK on the device; this rank's share, starting at global row row_offset):
//   mean_i = V_i . coef,   var_i = (noise + sigma) + V_i^T M V_i      (src/Predict.cpp:40-75, src/Utils.cpp:215-249)
void gpr_tail_system(Ctx* c, const double* V1, int KK, int64_t m_local, int64_t row_offset, const double* Ydev,
                     int64_t m_total, int K, const std::vector<double>& values, double t, double noise, double sigma,
                     DevBuf<double>& dcoef, DevBuf<double>& dM) {
  std::vector<double> lam(K), ls(K);
  for (int k = 0; k < K; ++k) {
    double ev = 1.0 - values[k];
    lam[k] = std::exp(-t * ev);
    ls[k] = std::exp(-0.5 * t * ev) + 0.0;
  }
  const double ns = noise + sigma;
  dcoef.alloc(KK);
  dM.alloc((size_t)KK * KK);
  if (m_total > K) {
    // Woodbury branch (src/Predict.cpp:61-74, src/Utils.cpp:237-244): the K x K algebra stays on the device (tail.cu)
    DevBuf<double> Gg((size_t)KK * KK + KK), dls(K), dlam(K);
    DevBuf<int> flag(1);
    gram_small_run(c, V1, Ydev, m_local, KK, Gg.p, Gg.p + (size_t)KK * KK);
    comm_allreduce_f64(c, Gg.p, (size_t)KK * KK + KK);
    dls.upload(ls.data(), K, c->stream);
    dlam.upload(lam.data(), K, c->stream);
    tail_woodbury_run(c, Gg.p, KK, K, dls.p, dlam.p, ns, dcoef.p, dM.p, flag.p);
    int bad = 0;
    flag.download(&bad, 1, c->stream);
    sync(c);
    if (bad) fail(2, "regression: K x K system is not positive definite");
  } else {
    // direct branch (src/Predict.cpp:47-59, src/Utils.cpp:228-236): every rank needs all m rows; m <= K is small,
    // the m x m system is solved on the host
    const int m = (int)m_total;
    std::vector<double> coef(KK, 0.0), M((size_t)KK * KK, 0.0);  // padded to the handle's K with zeros
    DevBuf<double> Vall((size_t)m * KK + m);
    Vall.zero(c->stream);
    if (m_local > 0) {
      FLGP_CUDA(cudaMemcpyAsync(Vall.p + (size_t)row_offset * KK, V1, sizeof(double) * m_local * KK,
                                cudaMemcpyDeviceToDevice, c->stream));
      FLGP_CUDA(cudaMemcpyAsync(Vall.p + (size_t)m * KK + row_offset, Ydev, sizeof(double) * m_local,
                                cudaMemcpyDeviceToDevice, c->stream));
    }
    comm_allreduce_f64(c, Vall.p, (size_t)m * KK + m);
    std::vector<double> Vh((size_t)m * KK + m);
    Vall.download(Vh.data(), Vh.size(), c->stream);
    sync(c);
    const double* Yh = Vh.data() + (size_t)m * KK;
    auto V = [&](int i, int k) { return Vh[(size_t)i * KK + k]; };
    std::vector<double> Cn((size_t)m * m);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc += (V(i, k) * lam[k]) * V(j, k);
        Cn[i + (size_t)m * j] = acc + (i == j ? ns : 0.0);
      }
    if (!chol_lower(Cn, m)) fail(2, "regression: m x m covariance is not positive definite");
    std::vector<double> alpha(Yh, Yh + m);
    chol_solve(Cn, m, alpha.data(), 1);
    for (int k = 0; k < K; ++k) {
      double acc = 0.0;
      for (int i = 0; i < m; ++i) acc += V(i, k) * alpha[i];
      coef[k] = lam[k] * acc;
    }
    // M = Lam - Lam V1^T K11^-1 V1 Lam
    std::vector<double> S((size_t)m * K);  // V1 Lam, column-major m x K
    for (int k = 0; k < K; ++k)
      for (int i = 0; i < m; ++i) S[i + (size_t)m * k] = V(i, k) * lam[k];
    std::vector<double> S2 = S;
    chol_solve(Cn, m, S2.data(), K);  // K11^-1 V1 Lam
    for (int b = 0; b < K; ++b)
      for (int a = 0; a < K; ++a) {
        double acc = 0.0;
        for (int i = 0; i < m; ++i) acc += S[i + (size_t)m * a] * S2[i + (size_t)m * b];
        M[a + (size_t)KK * b] = (a == b ? lam[a] : 0.0) - acc;
      }
    dcoef.upload(coef.data(), KK, c->stream);
    dM.upload(M.data(), (size_t)KK * KK, c->stream);
    sync(c);  // coef / M are host temporaries of this branch
  }
}

void regression_fixed_dev(flgp_spectrum* sp, const double* Ydev, int64_t m_total, int K, double t, double noise,
                          double sigma, double* y_pred, double* cov) {
  Ctx* c = sp->c;
  need(K >= 1 && K <= sp->K, "K exceeds the number of computed eigenpairs");
  need(m_total >= 1 && m_total <= sp->n_total, "bad number of training rows");
  const int s = sp->s, r = sp->r, KK = sp->K;
  const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(sp->n_local, m_total - sp->row_offset));
  StageScope st(c, "gpr_tail", 2.0 * r * (double)sp->n_local * (1 + r), 24.0 * r * (double)sp->n_local);
  const double ns = noise + sigma;
  // training rows of the lifted eigenvectors, row-major m_local x KK
  DevBuf<double> V1((size_t)std::max<int64_t>(m_local * KK, 1));
  lift_rows_run(c, r, sp->Zj.p, sp->Zx.p, sp->w.p, sp->Wm.p, KK, nullptr, m_local, V1.p, KK, false);
  DevBuf<double> dcoef, dM;
  gpr_tail_system(c, V1.p, KK, m_local, sp->row_offset, Ydev, m_total, K, sp->values, t, noise, sigma, dcoef, dM);
  // fold through the lift operator Wm (s x KK row-major)
  DevBuf<double> wv(s), T((size_t)s * KK), B((size_t)s * s);
  gemv_run(c, sp->Wm.p, dcoef.p, s, KK, wv.p);
  sparse_rowdot_run(c, sp->n_local, r, sp->Zj.p, sp->Zx.p, sp->w.p, wv.p, y_pred);
  if (cov) {
    // M is symmetric, so its column-major image is also its row-major image
    gemm_nn_run(c, sp->Wm.p, dM.p, s, KK, KK, T.p);               // T = Wm M        (s x KK row-major)
    gemm_nt_run(c, T.p, sp->Wm.p, nullptr, s, s, KK, B.p, s);     // B = T Wm^T      (s x s)
    sparse_quadform_run(c, sp->n_local, s, r, sp->Zj.p, sp->Zx.p, sp->w.p, B.p, ns, cov);
  }
  sync(c);
}

// ---- fit_se_regression_gp_cpp's grid search (src/Fit.cpp:126-178) ------------------------------------------
// One k-means and one KNN; per bandwidth a2: Z = exp(-dist / (a2 * mean dist)) -> graph Laplacian -> spectrum ->
// training objective.  Returns the handle of the best a2 (largest objective).
template <class T>
void dev_copy(Ctx* c, DevBuf<T>& dst, const DevBuf<T>& src) {
  dst.alloc(src.n);
  if (src.n) FLGP_CUDA(cudaMemcpyAsync(dst.p, src.p, sizeof(T) * src.n, cudaMemcpyDeviceToDevice, c->stream));
}

// prepare(q, handle): device-side preparation of grid point q's training (runs on the caller's thread, in grid order);
// train(q) -> objective to be maximised: host work, the grid points run concurrently when `concurrent`.
template <class Prepare, class Train>
std::unique_ptr<flgp_spectrum> se_grid_search(Ctx* c, const double* Xdev, int64_t n_local, int64_t n_total,
                                              int64_t row_offset, int d, int s, int r, int K, const Models& mo,
                                              const int32_t* init_idx, uint64_t seed, const double* a2s, int n_a2,
                                              bool concurrent, Prepare prepare, Train train, int* best_q,
                                              double* best_a2, double* best_obj) {
  need(n_local >= 0 && n_total >= 1 && d >= 1, "bad matrix shape");
  need(s >= 1 && s <= n_total, "need 1 <= s <= n");
  need(r >= 1 && r <= s, "need 1 <= r <= s");
  need(n_total < ((int64_t)1 << 31), "n must fit in int32 indices per process API");
  need(n_a2 >= 1 && a2s, "empty bandwidth grid");
  parse_gl(mo.gl);
  if (K < 0) K = s;
  flgp_spectrum base;
  base.c = c;
  base.n_local = n_local;
  base.n_total = n_total;
  base.row_offset = row_offset;
  base.d = d;
  base.s = s;
  base.r = r;
  stage_subsample(c, &base, Xdev, mo, init_idx, seed, nullptr);
  const int64_t nr = std::max<int64_t>(n_local * r, 1);
  DevBuf<int32_t> ind(nr), Zj(nr);
  DevBuf<double> dist(nr), D(nr);
  bool srt = false;
  {
    StageScope st(c, "knn", 2.0 * s * d * (double)n_local, (8.0 * d + 12.0 * r) * (double)n_local);
    knn_run(c, Xdev, n_local, n_local, d, base.U.p, s, s, r, ind.p, dist.p, &base.sorted, &srt);
  }
  knn_to_csr_run(c, n_local, r, ind.p, dist.p, Zj.p, D.p, srt ? base.sorted.perm.p : nullptr);
  sync(c);
  base.sorted = KMeansSorted();
  // distances_mean = distances_sp.coeffs().sum() / (n r)   (src/Fit.cpp:131)
  const int np = 256;
  DevBuf<double> part(np + 1);
  part.zero(c->stream);
  if (n_local > 0) FLGP_LAUNCH(c, sum_partial_kernel, np, 256, 0, D.p, n_local * r, part.p);
  FLGP_LAUNCH(c, sum_final_kernel, 1, 32, 0, part.p, np, part.p + np);
  comm_allreduce_f64(c, part.p + np, 1);
  double dsum = 0.0;
  FLGP_CUDA(cudaMemcpyAsync(&dsum, part.p + np, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  sync(c);
  const double dmean = dsum / ((double)n_total * r);
  // The grid points are independent.  Their device work (weights, graph Laplacian, Gram, eigensolve, training
  // statistics) runs back to back on the stream; the empirical-Bayes trainings are pure host work on K x K statistics
  // (hundreds of objective evaluations each: 70 ms at config 2, against 3 ms of device work per grid point) and run
  // concurrently, one thread per grid point, so the grid costs about one training instead of n_a2 of them.
  std::vector<std::unique_ptr<flgp_spectrum>> sps(n_a2);
  for (int q = 0; q < n_a2; ++q) {
    std::unique_ptr<flgp_spectrum> sp(new flgp_spectrum);
    sp->c = c;
    sp->n_local = n_local;
    sp->n_total = n_total;
    sp->row_offset = row_offset;
    sp->d = d;
    sp->s = s;
    sp->r = r;
    sp->ucols = base.ucols;
    sp->kmeans_iters = base.kmeans_iters;
    dev_copy(c, sp->U, base.U);
    dev_copy(c, sp->Zj, Zj);
    sp->Zx.alloc(nr);
    se_weights_run(c, D.p, n_local * r, a2s[q] * dmean, sp->Zx.p);  // src/Fit.cpp:150
    const double* nc = (mo.gl == FLGP_GL_CLUSTER_NORMALIZED) ? sp->U.p + (size_t)s * d : nullptr;
    stage_graph_laplacian(c, n_local, s, r, sp->Zj.p, sp->Zx.p, mo.gl, nc, n_total);
    stage_spectrum(c, sp.get(), K, mo.root);
    prepare(q, sp.get());
    sps[q] = std::move(sp);
  }
  std::vector<double> objs(n_a2);
  std::vector<std::string> errs(n_a2);
  std::vector<int> codes(n_a2, 0);
  auto train_one = [&](int q) {
    try {
      objs[q] = train(q);
    } catch (const Error& e) {
      codes[q] = e.code;
      errs[q] = e.what();
    } catch (const std::exception& e) {
      codes[q] = 3;
      errs[q] = e.what();
    }
  };
  if (n_a2 > 1 && concurrent) {
    std::vector<std::thread> th;
    for (int q = 0; q < n_a2; ++q)
      th.emplace_back([&, q] {
        g_outer_workers = n_a2;
        train_one(q);
      });
    for (auto& t : th) t.join();
  } else {
    for (int q = 0; q < n_a2; ++q) train_one(q);
  }
  for (int q = 0; q < n_a2; ++q)
    if (codes[q]) fail(codes[q], "%s", errs[q].c_str());
  std::unique_ptr<flgp_spectrum> best;
  double max_obj = -std::numeric_limits<double>::infinity();
  for (int q = 0; q < n_a2; ++q) {
    // src/Fit.cpp:172-177, 737-742, 861-866 (the first candidate is kept even when every objective is -inf)
    if (objs[q] > max_obj || !best) {
      max_obj = objs[q];
      *best_q = q;
      if (best_a2) *best_a2 = a2s[q];
      best = std::move(sps[q]);
    }
  }
  if (best_obj) *best_obj = max_obj;
  return best;
}

// fit_se_regression_gp_cpp's trainings (src/Fit.cpp:158-168): (t, noise) by MMA, or the objective at fixed_pars
std::unique_ptr<flgp_spectrum> se_grid_pipeline(Ctx* c, const double* Xdev, int64_t n_local, int64_t n_total,
                                                int64_t row_offset, int d, int s, int r, int K, const Models& mo,
                                                const int32_t* init_idx, uint64_t seed, const double* Ydev,
                                                int64_t m_total, double sigma, bool posterior, const double* a2s,
                                                int n_a2, const double* fixed_pars, double* pars_out, double* best_a2,
                                                double* best_obj) {
  need(n_a2 >= 1 && a2s, "empty bandwidth grid");
  std::vector<RegTrain> Ts(n_a2);
  std::vector<double> xs((size_t)2 * n_a2);
  int bq = 0;
  std::unique_ptr<flgp_spectrum> best = se_grid_search(
      c, Xdev, n_local, n_total, row_offset, d, s, r, K, mo, init_idx, seed, a2s, n_a2, !fixed_pars,
      [&](int q, flgp_spectrum* sp) { Ts[q] = reg_train_prepare(sp, Ydev, m_total, K < 0 ? s : K, sigma); },
      [&](int q) {
        double x[2] = {std::nan(""), std::nan("")};
        double obj;
        if (fixed_pars) {
          x[0] = fixed_pars[0];
          x[1] = fixed_pars[1];
          obj = -reg_objective(Ts[q], x, nullptr, posterior);
        } else {
          obj = train_regression(Ts[q], posterior, x, nullptr);
        }
        xs[2 * q] = x[0];
        xs[2 * q + 1] = x[1];
        return obj;
      },
      &bq, best_a2, best_obj);
  pars_out[0] = xs[2 * bq];
  pars_out[1] = xs[2 * bq + 1];
  return best;
}

// out[i] = base[i] - sum_k A(k, i) B(k, i) for row-major m x n operands (columns are the test rows)
__global__ void coldot_sub_kernel(const double* __restrict__ A, const double* __restrict__ B, int m, int64_t n,
                                  const double* __restrict__ base, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0;
  for (int k = 0; k < m; ++k) a = fma(A[(size_t)k * n + i], B[(size_t)k * n + i], a);
  out[i] = base[i] - a;
}

// ---- Laplace posterior of the binary GP classifier (posterior_distribution_classification,
//      /root/reference/src/Utils.cpp:252-299; GPML algorithms 3.1 / 3.2) -------------------------------------------
// Newton iterations on the m training points (host, m x m: m is the number of labelled points), from f = 0:
//   pi = 1 / (1 + exp(-f)), W = pi (1 - pi), B = I + sqrt(W) C11 sqrt(W), b = W f + (Y - pi),
//   a = b - sqrt(W) B^-1 sqrt(W) (C11 b), f <- C11 a, until |f - f_new|_1 < tol.
// Returns pi (m) and beta = sqrt(W) B^-1 sqrt(W) (m x m, column-major) at the mode.
// beta == nullptr: the caller takes the factors instead (sw_out = sqrt(W), Bchol_out = lower Cholesky factor of B) and
// applies beta to what it needs by solves — m x K right-hand sides instead of the m x m inverse.
void laplace_mode(const std::vector<double>& C11, const double* Y, int m, double tol, int max_iter,
                  std::vector<double>& pi, std::vector<double>* beta, std::vector<double>* sw_out = nullptr,
                  std::vector<double>* Bchol_out = nullptr) {
  std::vector<double> f(m, 0.0), W(m), sw(m), b(m), cb(m), a(m), fn(m), B((size_t)m * m);
  pi.assign(m, 0.5);
  auto refresh = [&]() {
    for (int i = 0; i < m; ++i) {
      pi[i] = 1.0 / (1.0 + std::exp(-f[i]));
      W[i] = pi[i] * (1.0 - pi[i]);
      sw[i] = std::sqrt(W[i]);
    }
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) B[i + (size_t)m * j] = (sw[i] * C11[i + (size_t)m * j]) * sw[j] + (i == j ? 1.0 : 0.0);
    if (!chol_lower(B, m)) fail(2, "classification: the Newton system is not positive definite");
  };
  for (int iter = 0; iter < max_iter; ++iter) {
    refresh();
    for (int i = 0; i < m; ++i) b[i] = W[i] * f[i] + (Y[i] - pi[i]);
    // C11 b and C11 a as column sweeps (contiguous; every row still adds its terms in ascending j: the same bits)
    std::fill(fn.begin(), fn.end(), 0.0);
    for (int j = 0; j < m; ++j) {
      const double bj = b[j];
      const double* cj = &C11[(size_t)m * j];
      for (int i = 0; i < m; ++i) fn[i] += cj[i] * bj;
    }
    for (int i = 0; i < m; ++i) cb[i] = sw[i] * fn[i];
    chol_solve(B, m, cb.data(), 1);
    for (int i = 0; i < m; ++i) a[i] = b[i] - sw[i] * cb[i];
    double diff = 0.0;
    std::fill(fn.begin(), fn.end(), 0.0);
    for (int j = 0; j < m; ++j) {
      const double aj = a[j];
      const double* cj = &C11[(size_t)m * j];
      for (int i = 0; i < m; ++i) fn[i] += cj[i] * aj;
    }
    for (int i = 0; i < m; ++i) diff += std::fabs(f[i] - fn[i]);
    f = fn;
    if (diff < tol) break;
  }
  refresh();
  if (beta) {
    beta->assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i) (*beta)[i + (size_t)m * i] = 1.0;
    chol_solve(B, m, beta->data(), m);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) (*beta)[i + (size_t)m * j] = (sw[i] * (*beta)[i + (size_t)m * j]) * sw[j];
  }
  if (sw_out) *sw_out = sw;
  if (Bchol_out) *Bchol_out = std::move(B);
}

// The m-sized half of posterior_distribution_classification as the logit drivers call it (src/Fit.cpp:563-582), from the
// labelled rows Vh of the eigenvectors (m x KK row-major, the first K columns used) and ev = 1 - values:
// C11 = V1 Lam V1^T + sigma I, Newton mode, then the two folded operators
//   coef = Lam V1^T (Y - pi)   (KK, zero padded)        Mq = Lam - Lam V1^T beta V1 Lam   (KK x KK, zero padded)
// so that for any row v of the eigenvectors  mean = v . coef  and  cov = v Mq v^T + sigma.
void laplace_fold(const double* Vh, int KK, int K, const std::vector<double>& ev, const double* Yh, int m, double t,
                  double sigma, double tol, int max_iter, std::vector<double>& coef, std::vector<double>& Mq) {
  auto V = [&](int i, int k) { return Vh[(size_t)i * KK + k]; };
  std::vector<double> lam(K);
  for (int k = 0; k < K; ++k) lam[k] = std::exp(-t * ev[k]);
  std::vector<double> C11((size_t)m * m);
  const int Tn = m >= 256 ? std::min(host_threads(), m / 64) : 1;
  host_parallel(Tn, [&](int q, int step) {  // a column by one thread: the same sums for any thread count
    for (int j = q; j < m; j += step)
      for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc += (V(i, k) * lam[k]) * V(j, k);
        C11[i + (size_t)m * j] = acc + (i == j ? sigma : 0.0);  // Cvv.diagonal() += sigma (src/Fit.cpp:566)
      }
  });
  std::vector<double> pi, sw, Bc;
  laplace_mode(C11, Yh, m, tol, max_iter, pi, nullptr, &sw, &Bc);
  std::vector<double> T1((size_t)m * K);
  coef.assign(KK, 0.0);
  Mq.assign((size_t)KK * KK, 0.0);
  for (int k = 0; k < K; ++k) {
    double acc = 0.0;
    for (int i = 0; i < m; ++i) acc += V(i, k) * (Yh[i] - pi[i]);
    coef[k] = lam[k] * acc;
  }
  // T1 = beta V1 = sqrt(W) B^-1 (sqrt(W) V1)  (m x K, column-major): K solves with the factor of B, no m x m inverse
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < m; ++i) T1[i + (size_t)m * k] = sw[i] * V(i, k);
  chol_solve(Bc, m, T1.data(), K);
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < m; ++i) T1[i + (size_t)m * k] *= sw[i];
  for (int b2 = 0; b2 < K; ++b2)
    for (int a2 = 0; a2 < K; ++a2) {
      double acc = 0.0;
      for (int i = 0; i < m; ++i) acc += V(i, a2) * T1[i + (size_t)m * b2];
      Mq[a2 + (size_t)KK * b2] = (a2 == b2 ? lam[a2] : 0.0) - (lam[a2] * acc) * lam[b2];
    }
}

// The n-sized half on a spectrum handle, folded like the GPR tail (C21 = V2 Lam V1^T is never formed):
//   mean_i = V_i . (Lam V1^T (Y - pi)),   cov_i = V_i (Lam - Lam V1^T beta V1 Lam) V_i^T + sigma
void classification_posterior_dev(flgp_spectrum* sp, const double* Ydev, int64_t m_total, int K, double t, double sigma,
                                  double tol, int max_iter, double* mean, double* cov) {
  Ctx* c = sp->c;
  need(K >= 1 && K <= sp->K, "K exceeds the number of computed eigenpairs");
  need(m_total >= 1 && m_total <= sp->n_total && m_total <= 8192, "classification: need 1 <= m <= 8192 labelled rows");
  const int s = sp->s, r = sp->r, KK = sp->K, m = (int)m_total;
  const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(sp->n_local, m_total - sp->row_offset));
  DevBuf<double> V1((size_t)std::max<int64_t>(m_local * KK, 1));
  lift_rows_run(c, r, sp->Zj.p, sp->Zx.p, sp->w.p, sp->Wm.p, KK, nullptr, m_local, V1.p, KK, false);
  DevBuf<double> Vall((size_t)m * KK + m);
  Vall.zero(c->stream);
  if (m_local > 0) {
    FLGP_CUDA(cudaMemcpyAsync(Vall.p + (size_t)sp->row_offset * KK, V1.p, sizeof(double) * m_local * KK,
                              cudaMemcpyDeviceToDevice, c->stream));
    FLGP_CUDA(cudaMemcpyAsync(Vall.p + (size_t)m * KK + sp->row_offset, Ydev, sizeof(double) * m_local,
                              cudaMemcpyDeviceToDevice, c->stream));
  }
  comm_allreduce_f64(c, Vall.p, (size_t)m * KK + m);
  std::vector<double> Vh((size_t)m * KK + m);
  Vall.download(Vh.data(), Vh.size(), c->stream);
  sync(c);
  const double* Yh = Vh.data() + (size_t)m * KK;
  std::vector<double> ev(K), coef, Mq;
  for (int k = 0; k < K; ++k) ev[k] = 1.0 - sp->values[k];
  laplace_fold(Vh.data(), KK, K, ev, Yh, m, t, sigma, tol, max_iter, coef, Mq);
  DevBuf<double> dcoef(KK), dM((size_t)KK * KK);
  dcoef.upload(coef.data(), KK, c->stream);
  dM.upload(Mq.data(), (size_t)KK * KK, c->stream);
  DevBuf<double> wv(s), T((size_t)s * KK), Bq((size_t)s * s);
  gemv_run(c, sp->Wm.p, dcoef.p, s, KK, wv.p);
  sparse_rowdot_run(c, sp->n_local, r, sp->Zj.p, sp->Zx.p, sp->w.p, wv.p, mean);
  if (cov) {
    gemm_nn_run(c, sp->Wm.p, dM.p, s, KK, KK, T.p);            // Mq is symmetric up to rounding: beta is
    gemm_nt_run(c, T.p, sp->Wm.p, nullptr, s, s, KK, Bq.p, s);
    sparse_quadform_run(c, sp->n_local, s, r, sp->Zj.p, sp->Zx.p, sp->w.p, Bq.p, sigma, cov);
  }
  sync(c);
}

}  // namespace

extern "C" {

int flgp_version(void) { return 100; }
const char* flgp_last_error(void) { return g_err.c_str(); }

int flgp_ctx_create(int device, flgp_ctx** out) {
  return guard([&] {
    need(out != nullptr, "null output");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      fail(3, "no CUDA device available (%s): libflgp_b200 has no CPU fallback", cudaGetErrorString(e));
    need(device >= 0 && device < count, "device index out of range");
    FLGP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FLGP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) fail(3, "libflgp_b200 is built for sm_100a only (found sm_%d%d)", prop.major, prop.minor);
    std::unique_ptr<flgp_ctx> h(new flgp_ctx);
    h->c.device = device;
    h->c.sm_count = prop.multiProcessorCount;
    FLGP_CUDA(cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking));
    h->c.own_stream = true;
    FLGP_CUDA(cudaMallocHost(&h->c.pinned, 64 * sizeof(int64_t)));
    pool_ctx_count(+1);
    *out = h.release();
  });
}

void flgp_ctx_destroy(flgp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->c.device);
  cudaDeviceSynchronize();
  pool_ctx_count(-1);
  pool_trim();
  comm_destroy(&ctx->c);
  for (auto& s : ctx->c.stages) {
    cudaEventDestroy(s.beg);
    cudaEventDestroy(s.end);
  }
  if (ctx->c.pinned) cudaFreeHost(ctx->c.pinned);
  if (ctx->c.own_stream && ctx->c.stream) cudaStreamDestroy(ctx->c.stream);
  delete ctx;
}

int flgp_ctx_set_stream(flgp_ctx* ctx, void* cuda_stream) {
  return guard([&] {
    need(ctx != nullptr, "null context");
    FLGP_CUDA(cudaSetDevice(ctx->c.device));
    FLGP_CUDA(cudaDeviceSynchronize());  // pooled blocks are ordered by the stream: drain before switching
    if (ctx->c.own_stream && ctx->c.stream) {
      FLGP_CUDA(cudaStreamSynchronize(ctx->c.stream));
      cudaStreamDestroy(ctx->c.stream);
    }
    if (cuda_stream) {
      ctx->c.stream = (cudaStream_t)cuda_stream;
      ctx->c.own_stream = false;
    } else {
      FLGP_CUDA(cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking));
      ctx->c.own_stream = true;
    }
  });
}

int flgp_ctx_synchronize(flgp_ctx* ctx) {
  return guard([&] {
    need(ctx != nullptr, "null context");
    sync(&ctx->c);
  });
}

uint64_t flgp_ctx_launch_count(const flgp_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int flgp_ctx_set_timing(flgp_ctx* ctx, int on) {
  if (!ctx) return 2;
  ctx->c.timing = on != 0;
  return 0;
}
int flgp_ctx_stage_reset(flgp_ctx* ctx) {
  if (!ctx) return 2;
  for (auto& s : ctx->c.stages) {
    cudaEventDestroy(s.beg);
    cudaEventDestroy(s.end);
  }
  ctx->c.stages.clear();
  return 0;
}
int flgp_ctx_stage_count(flgp_ctx* ctx) { return ctx ? (int)ctx->c.stages.size() : 0; }
int flgp_ctx_stage_get(flgp_ctx* ctx, int i, char* name, int name_len, double* ms, uint64_t* launches, double* flops,
                       double* bytes) {
  return guard([&] {
    need(ctx && i >= 0 && i < (int)ctx->c.stages.size(), "stage index out of range");
    StageRec& s = ctx->c.stages[i];
    FLGP_CUDA(cudaEventSynchronize(s.end));
    float t = 0.f;
    FLGP_CUDA(cudaEventElapsedTime(&t, s.beg, s.end));
    if (name && name_len > 0) {
      std::strncpy(name, s.name.c_str(), name_len - 1);
      name[name_len - 1] = 0;
    }
    if (ms) *ms = t;
    if (launches) *launches = s.launches;
    if (flops) *flops = s.flops;
    if (bytes) *bytes = s.bytes;
  });
}

int flgp_dfma_peak(flgp_ctx* ctx, int iters, double* tflops) {
  return guard([&] {
    need(ctx && tflops, "null argument");
    *tflops = dfma_peak_run(&ctx->c, iters);
  });
}

int flgp_copy_roundtrip(flgp_ctx* ctx, const void* in, void* out, size_t bytes) {
  return guard([&] {
    need(ctx && (bytes == 0 || (in && out)), "null argument");
    Ctx* c = on_device(&ctx->c);
    DevBuf<unsigned char> buf(std::max<size_t>(bytes, 1));
    if (bytes) {
      buf.upload(static_cast<const unsigned char*>(in), bytes, c->stream);
      buf.download(static_cast<unsigned char*>(out), bytes, c->stream);
    }
    sync(c);
  });
}

int flgp_comm_unique_id(void* out128) {
  return guard([&] {
    need(out128 != nullptr, "null output");
    comm_unique_id(out128);
  });
}
int flgp_ctx_comm_init(flgp_ctx* ctx, const void* id128, int rank, int nranks) {
  return guard([&] {
    need(ctx && (id128 || nranks == 1), "null argument");
    comm_init(&ctx->c, id128, rank, nranks);
  });
}

int flgp_default_init(int64_t n, int s, uint64_t seed, int32_t* init_idx) {
  return guard([&] {
    need(init_idx != nullptr, "null output");
    auto v = default_init(n, s, seed);
    std::memcpy(init_idx, v.data(), sizeof(int32_t) * s);
  });
}

int flgp_subsample(flgp_ctx* ctx, const double* X, int64_t n, int d, int s, const char* method, int iter_max,
                   int nstart, const int32_t* init_idx, uint64_t seed, double* U, int32_t* assign, int* iters) {
  return guard([&] {
    need(ctx && X && U, "null argument");
    need(n >= 1 && d >= 1, "bad matrix shape");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "flgp_subsample is single-process; use the sharded spectrum entry for multi-GPU");
    DevBuf<double> dX((size_t)n * d);
    dX.upload(X, (size_t)n * d, c->stream);
    flgp_spectrum sp;
    sp.c = c;
    sp.n_local = sp.n_total = n;
    sp.d = d;
    sp.s = s;
    Models mo = make_models(method, "lae", FLGP_GL_RW, 1, nstart, 0.1, iter_max);
    DevBuf<int32_t> dassign((size_t)n);
    stage_subsample(c, &sp, dX.p, mo, init_idx, seed, dassign.p);
    sp.U.download(U, (size_t)s * sp.ucols, c->stream);
    if (assign && mo.subsample == "kmeans") dassign.download(assign, n, c->stream);
    sync(c);
    if (iters) *iters = sp.kmeans_iters;
  });
}

int flgp_knn(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int r, int32_t* ind,
             double* dist, int32_t* Zj, double* Zx) {
  return guard([&] {
    need(ctx && X && U && ind, "null argument");
    need(n >= 0 && d >= 1 && s >= 1, "bad matrix shape");
    need((Zj == nullptr) == (Zx == nullptr), "Zj and Zx go together");
    Ctx* c = on_device(&ctx->c);
    const bool want_dist = dist || Zj;
    DevBuf<double> dX((size_t)std::max<int64_t>(n * d, 1)), dU((size_t)s * d), ddist;
    DevBuf<int32_t> dind((size_t)std::max<int64_t>(n * r, 1));
    if (n) dX.upload(X, (size_t)n * d, c->stream);
    dU.upload(U, (size_t)s * d, c->stream);
    if (want_dist) ddist.alloc((size_t)std::max<int64_t>(n * r, 1));
    knn_run(c, dX.p, n, n, d, dU.p, s, s, r, dind.p, want_dist ? ddist.p : nullptr);
    if (n) {
      dind.download(ind, (size_t)n * r, c->stream);
      if (dist) ddist.download(dist, (size_t)n * r, c->stream);
      if (Zj) {
        DevBuf<int32_t> dZj((size_t)n * r);
        DevBuf<double> dZx((size_t)n * r);
        knn_to_csr_run(c, n, r, dind.p, ddist.p, dZj.p, dZx.p);
        dZj.download(Zj, (size_t)n * r, c->stream);
        dZx.download(Zx, (size_t)n * r, c->stream);
        sync(c);
      }
    }
    sync(c);
  });
}

int flgp_simplex_project(flgp_ctx* ctx, const double* v, int r, double* z) {
  return guard([&] {
    need(ctx && v && z && r >= 1, "bad argument");
    Ctx* c = on_device(&ctx->c);
    DevBuf<double> dv(r), dz(r);
    dv.upload(v, r, c->stream);
    simplex_project_run(c, dv.p, r, dz.p);
    dz.download(z, r, c->stream);
    sync(c);
  });
}

int flgp_lae_point(flgp_ctx* ctx, const double* x, int d, const double* Ur, int r, double* z) {
  return guard([&] {
    need(ctx && x && Ur && z && d >= 1, "bad argument");
    Ctx* c = on_device(&ctx->c);
    DevBuf<double> dx(d), dU((size_t)r * d), dz(r);
    dx.upload(x, d, c->stream);
    dU.upload(Ur, (size_t)r * d, c->stream);
    lae_point_run(c, dx.p, d, dU.p, r, dz.p);
    dz.download(z, r, c->stream);
    sync(c);
  });
}

int flgp_lae(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int r, int32_t* Zj, double* Zx,
             int64_t* stats) {
  return guard([&] {
    need(ctx && X && U && Zj && Zx, "null argument");
    need(n >= 1 && d >= 1 && s >= 1, "bad matrix shape");
    Ctx* c = on_device(&ctx->c);
    DevBuf<double> dX((size_t)n * d), dU((size_t)s * d), dZx((size_t)n * r);
    DevBuf<int32_t> dind((size_t)n * r), dZj((size_t)n * r);
    DevBuf<long long> dst(2);
    dst.zero(c->stream);
    dX.upload(X, (size_t)n * d, c->stream);
    dU.upload(U, (size_t)s * d, c->stream);
    knn_run(c, dX.p, n, n, d, dU.p, s, s, r, dind.p, nullptr);
    lae_run(c, dX.p, n, n, d, dU.p, s, s, r, dind.p, dZj.p, dZx.p, nullptr, dst.p);
    dZj.download(Zj, (size_t)n * r, c->stream);
    dZx.download(Zx, (size_t)n * r, c->stream);
    long long h[2] = {0, 0};
    dst.download(h, 2, c->stream);
    sync(c);
    if (stats) {
      stats[0] = h[0];
      stats[1] = h[1];
    }
  });
}

int flgp_graph_laplacian(flgp_ctx* ctx, int64_t n, int s, int r, const int32_t* Zj, double* Zx, int gl,
                         const double* num_class) {
  return guard([&] {
    need(ctx && Zj && Zx, "null argument");
    need(n >= 1 && s >= 1 && r >= 1, "bad shape");
    parse_gl(gl);
    need(gl != FLGP_GL_CLUSTER_NORMALIZED || num_class, "cluster-normalized needs the cluster sizes");
    Ctx* c = on_device(&ctx->c);
    DevBuf<int32_t> dZj((size_t)n * r);
    DevBuf<double> dZx((size_t)n * r), dnc(s);
    dZj.upload(Zj, (size_t)n * r, c->stream);
    dZx.upload(Zx, (size_t)n * r, c->stream);
    if (num_class) dnc.upload(num_class, s, c->stream);
    const double zmax = csr_validate_run(c, n, s, r, dZj.p, dZx.p);
    stage_graph_laplacian(c, n, s, r, dZj.p, dZx.p, gl, dnc.p, n, std::max(zmax, 1e-300));
    dZx.download(Zx, (size_t)n * r, c->stream);
    sync(c);
  });
}

static int cross_similarity(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int ucols, int r,
                            int gl, const char* kernel, double epsilon, int32_t* Zj, double* Zx) {
  return guard([&] {
    need(ctx && X && U && Zj && Zx, "null argument");
    need(n >= 1 && d >= 1 && s >= 1, "bad matrix shape");
    need(ucols == d || ucols == d + 1, "U must have d or d+1 columns");
    parse_gl(gl);
    if (gl == FLGP_GL_CLUSTER_NORMALIZED && ucols != d + 1)
      fail(2, "gl=\"cluster-normalized\" needs the cluster-size column of U (SURVEY Appendix A.1)");
    Ctx* c = on_device(&ctx->c);
    flgp_spectrum sp;
    sp.c = c;
    sp.n_local = sp.n_total = n;
    sp.d = d;
    sp.s = s;
    sp.r = r;
    DevBuf<double> dX((size_t)n * d), dU((size_t)s * ucols);
    dX.upload(X, (size_t)n * d, c->stream);
    dU.upload(U, (size_t)s * ucols, c->stream);
    stage_cross_similarity(c, &sp, dX.p, dU.p, kernel, epsilon);
    stage_graph_laplacian(c, n, s, r, sp.Zj.p, sp.Zx.p, gl, ucols == d + 1 ? dU.p + (size_t)s * d : nullptr, n);
    sp.Zj.download(Zj, (size_t)n * r, c->stream);
    sp.Zx.download(Zx, (size_t)n * r, c->stream);
    sync(c);
  });
}

int flgp_cross_similarity_lae(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int ucols, int r,
                              int gl, int32_t* Zj, double* Zx) {
  return cross_similarity(ctx, X, n, d, U, s, ucols, r, gl, "lae", 0.1, Zj, Zx);
}
int flgp_cross_similarity_se(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int ucols, int r,
                             int gl, double epsilon, int32_t* Zj, double* Zx) {
  return cross_similarity(ctx, X, n, d, U, s, ucols, r, gl, "se", epsilon, Zj, Zx);
}

int flgp_spectrum_from_z(flgp_ctx* ctx, int64_t n, int s, int r, const int32_t* Zj, const double* Zx, int K, int root,
                         double* values, double* vectors, flgp_spectrum** handle) {
  return guard([&] {
    need(ctx && Zj && Zx, "null argument");
    need(n >= 1 && s >= 1 && r >= 1 && r <= s, "bad shape");
    Ctx* c = on_device(&ctx->c);
    std::unique_ptr<flgp_spectrum> sp(new flgp_spectrum);
    sp->c = c;
    sp->n_local = sp->n_total = n;
    sp->s = s;
    sp->r = r;
    sp->Zj.alloc((size_t)n * r);
    sp->Zx.alloc((size_t)n * r);
    sp->Zj.upload(Zj, (size_t)n * r, c->stream);
    sp->Zx.upload(Zx, (size_t)n * r, c->stream);
    const double zmax = csr_validate_run(c, n, s, r, sp->Zj.p, sp->Zx.p);
    stage_spectrum(c, sp.get(), K, root != 0, std::max(zmax, 1e-300));
    if (values) std::memcpy(values, sp->values.data(), sizeof(double) * sp->K);
    if (vectors) {
      DevBuf<double> V((size_t)n * sp->K);
      lift_rows_run(c, r, sp->Zj.p, sp->Zx.p, sp->w.p, sp->Wm.p, sp->K, nullptr, n, V.p, n, true);
      V.download(vectors, (size_t)n * sp->K, c->stream);
      sync(c);
    }
    if (handle) *handle = sp.release();
  });
}

int flgp_eigs_sym(flgp_ctx* ctx, const double* A, int s, int K, double* values, double* vectors) {
  return guard([&] {
    need(ctx && A && values, "null argument");
    need(s >= 1 && K >= 1 && K <= s, "need 1 <= K <= s");
    Ctx* c = on_device(&ctx->c);
    DevBuf<double> G((size_t)s * s), lam(K), Y((size_t)s * K);
    G.upload(A, (size_t)s * s, c->stream);
    {
      StageScope st(c, "eigh", (4.0 / 3.0) * s * (double)s * s + 2.0 * s * (double)s * K, 8.0 * s * (double)s * s);
      eigh_topk_run(c, G.p, s, K, lam.p, Y.p);
    }
    lam.download(values, K, c->stream);
    if (vectors) Y.download(vectors, (size_t)s * K, c->stream);
    sync(c);
  });
}

int flgp_heat_kernel_spectrum_dev(flgp_ctx* ctx, const double* X_local_dev, int64_t n_local, int64_t n_total,
                                  int64_t row_offset, int d, int s, int r, int K, const char* subsample,
                                  const char* kernel, int gl, int root, int nstart, double epsilon, int iter_max,
                                  const int32_t* init_idx, uint64_t seed, flgp_spectrum** out) {
  return guard([&] {
    need(ctx && out && (X_local_dev || n_local == 0), "null argument");
    Models mo = make_models(subsample, kernel, gl, root, nstart, epsilon, iter_max);
    *out = spectrum_pipeline(&ctx->c, X_local_dev, n_local, n_total, row_offset, d, s, r, K, mo, init_idx, seed)
               .release();
  });
}

int flgp_heat_kernel_spectrum_sharded(flgp_ctx* ctx, const double* X_local, int64_t n_local, int64_t n_total,
                                      int64_t row_offset, int d, int s, int r, int K, const char* subsample,
                                      const char* kernel, int gl, int root, int nstart, double epsilon, int iter_max,
                                      const int32_t* init_idx, uint64_t seed, flgp_spectrum** out) {
  return guard([&] {
    need(ctx && out && (X_local || n_local == 0), "null argument");
    need(n_local >= 0 && d >= 1, "bad matrix shape");
    Ctx* c = on_device(&ctx->c);
    DevBuf<double> dX((size_t)std::max<int64_t>(n_local * d, 1));
    if (n_local) dX.upload(X_local, (size_t)n_local * d, c->stream);
    Models mo = make_models(subsample, kernel, gl, root, nstart, epsilon, iter_max);
    *out = spectrum_pipeline(c, dX.p, n_local, n_total, row_offset, d, s, r, K, mo, init_idx, seed).release();
  });
}

int flgp_heat_kernel_spectrum(flgp_ctx* ctx, const double* X, int64_t m, const double* X_new, int64_t m_new, int d,
                              int s, int r, int K, const char* subsample, const char* kernel, int gl, int root,
                              int nstart, double epsilon, int iter_max, const int32_t* init_idx, uint64_t seed,
                              flgp_spectrum** out) {
  return guard([&] {
    need(ctx && out && X, "null argument");
    need(m >= 1 && m_new >= 0 && d >= 1 && (X_new || m_new == 0), "bad matrix shape");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "use flgp_heat_kernel_spectrum_sharded for multi-GPU runs");
    DevBuf<double> dX = upload_concat(c, X, m, X_new, m_new, d);
    Models mo = make_models(subsample, kernel, gl, root, nstart, epsilon, iter_max);
    *out = spectrum_pipeline(c, dX.p, m + m_new, m + m_new, 0, d, s, r, K, mo, init_idx, seed).release();
  });
}

void flgp_spectrum_free(flgp_spectrum* h) { delete h; }

int flgp_spectrum_info(const flgp_spectrum* h, int64_t* info10) {
  if (!h || !info10) return 2;
  int64_t v[10] = {h->n_local, h->n_total, h->row_offset, h->d, h->s, h->r, h->K, h->kmeans_iters, h->lae_iters,
                   h->lae_bts};
  std::memcpy(info10, v, sizeof v);
  return 0;
}

int flgp_spectrum_values(const flgp_spectrum* h, double* values) {
  if (!h || !values) return 2;
  std::memcpy(values, h->values.data(), sizeof(double) * h->K);
  return 0;
}

int flgp_spectrum_anchors(const flgp_spectrum* h, double* U) {
  return guard([&] {
    need(h && U && h->U.p, "no anchors in this handle");
    h->U.download(U, (size_t)h->s * h->ucols, h->c->stream);
    sync(h->c);
  });
}

int flgp_spectrum_z(const flgp_spectrum* h, int32_t* Zj, double* Zx) {
  return guard([&] {
    need(h && Zj && Zx, "null argument");
    h->Zj.download(Zj, (size_t)h->n_local * h->r, h->c->stream);
    h->Zx.download(Zx, (size_t)h->n_local * h->r, h->c->stream);
    sync(h->c);
  });
}

int flgp_spectrum_vectors(flgp_spectrum* h, double* vectors) {
  return guard([&] {
    need(h && vectors, "null argument");
    Ctx* c = on_device(h->c);
    DevBuf<double> V((size_t)std::max<int64_t>(h->n_local * h->K, 1));
    lift_rows_run(c, h->r, h->Zj.p, h->Zx.p, h->w.p, h->Wm.p, h->K, nullptr, h->n_local, V.p, h->n_local, true);
    if (h->n_local) V.download(vectors, (size_t)h->n_local * h->K, c->stream);
    sync(c);
  });
}

int flgp_spectrum_gather_rows(flgp_spectrum* h, const int32_t* idx, int64_t n_idx, double* V) {
  return guard([&] {
    need(h && idx && V && n_idx >= 1, "bad argument");
    for (int64_t a = 0; a < n_idx; ++a) need(idx[a] >= 0 && idx[a] < h->n_local, "row index out of range");
    Ctx* c = on_device(h->c);
    DevBuf<int32_t> di((size_t)n_idx);
    DevBuf<double> dV((size_t)n_idx * h->K);
    di.upload(idx, n_idx, c->stream);
    lift_rows_run(c, h->r, h->Zj.p, h->Zx.p, h->w.p, h->Wm.p, h->K, di.p, n_idx, dV.p, n_idx, true);
    dV.download(V, (size_t)n_idx * h->K, c->stream);
    sync(c);
  });
}

int flgp_hk_from_spectrum(flgp_spectrum* h, int K, double t, const int32_t* idx0, int64_t n0, const int32_t* idx1,
                          int64_t n1, double* H) {
  return guard([&] {
    need(h && idx0 && idx1 && H && n0 >= 1 && n1 >= 1, "bad argument");
    if (K < 0) K = h->K;
    need(K >= 1 && K <= h->K, "K exceeds the number of computed eigenpairs");
    for (int64_t a = 0; a < n0; ++a) need(idx0[a] >= 0 && idx0[a] < h->n_local, "row index out of range");
    for (int64_t a = 0; a < n1; ++a) need(idx1[a] >= 0 && idx1[a] < h->n_local, "row index out of range");
    Ctx* c = on_device(h->c);
    const int KK = h->K;
    StageScope st(c, "hk_from_spectrum", 2.0 * K * (double)n0 * n1, 8.0 * (double)n0 * n1);
    DevBuf<int32_t> d0((size_t)n0), d1((size_t)n1);
    DevBuf<double> V0((size_t)n0 * KK), V1((size_t)n1 * KK), dl(K), dH((size_t)n0 * n1);
    d0.upload(idx0, n0, c->stream);
    d1.upload(idx1, n1, c->stream);
    std::vector<double> lam(K);
    for (int k = 0; k < K; ++k) lam[k] = std::exp(-t * (1.0 - h->values[k]));  // src/Spectrum.cpp:86,90
    dl.upload(lam.data(), K, c->stream);
    lift_rows_run(c, h->r, h->Zj.p, h->Zx.p, h->w.p, h->Wm.p, KK, d0.p, n0, V0.p, KK, false);
    lift_rows_run(c, h->r, h->Zj.p, h->Zx.p, h->w.p, h->Wm.p, KK, d1.p, n1, V1.p, KK, false);
    // rows are KK long but only the first K columns enter: the tensor maps carry the pitch KK, no compaction
    {
      StageScope sg(c, "hk_gemm", 2.0 * K * (double)n0 * n1, 8.0 * (double)n0 * n1);
      gemm_nt_ld_run(c, V0.p, KK, V1.p, KK, dl.p, n0, n1, K, dH.p, n0);
    }
    dH.download(H, (size_t)n0 * n1, c->stream);
    sync(c);
  });
}

int flgp_lae_eigenmap(flgp_ctx* ctx, const double* X, int64_t n, int d, int s, int r, int ndim, const char* subsample,
                      int gl, int nstart, int iter_max, const int32_t* init_idx, uint64_t seed, double* eigenvalues,
                      double* eigenvectors) {
  flgp_spectrum* h = nullptr;
  int rc = flgp_heat_kernel_spectrum(ctx, X, n, nullptr, 0, d, s, r, ndim, subsample, "lae", gl, 1, nstart, 0.1, iter_max,
                                     init_idx, seed, &h);
  if (rc) return rc;
  std::unique_ptr<flgp_spectrum> own(h);
  if (eigenvalues)
    for (int k = 0; k < h->K; ++k) eigenvalues[k] = 1.0 - h->values[k];  // src/Spectrum.cpp:23
  if (eigenvectors) return flgp_spectrum_vectors(h, eigenvectors);
  return 0;
}

int flgp_heat_kernel_covariance(flgp_ctx* ctx, const double* X, int64_t m, const double* X_new, int64_t m_new, int d,
                                int s, int r, double t, int K, const char* subsample, const char* kernel, int gl,
                                int root, int nstart, double epsilon, int iter_max, const int32_t* init_idx,
                                uint64_t seed, double* H) {
  if (K < 0) K = s;  // src/Spectrum.cpp:31-33
  flgp_spectrum* h = nullptr;
  int rc = flgp_heat_kernel_spectrum(ctx, X, m, X_new, m_new, d, s, r, K, subsample, kernel, gl, root, nstart, epsilon,
                                     iter_max, init_idx, seed, &h);
  if (rc) return rc;
  std::unique_ptr<flgp_spectrum> own(h);
  const int64_t n = m + m_new;
  std::vector<int32_t> idx0(n), idx1(m);
  for (int64_t i = 0; i < n; ++i) idx0[i] = (int32_t)i;
  for (int64_t i = 0; i < m; ++i) idx1[i] = (int32_t)i;
  return flgp_hk_from_spectrum(h, K, t, idx0.data(), n, idx1.data(), m, H);
}

int flgp_regression_fixed_dev(flgp_spectrum* h, const double* Y_local_dev, int64_t m_total, int K, double t, double noise,
                              double sigma, double* y_pred_dev, double* cov_dev) {
  return guard([&] {
    need(h && y_pred_dev, "null argument");
    if (K < 0) K = h->K;
    regression_fixed_dev(h, Y_local_dev, m_total, K, t, noise, sigma, y_pred_dev, cov_dev);
  });
}

int flgp_regression_fixed(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double t, double noise,
                          double sigma, double* y_pred, double* cov) {
  return guard([&] {
    need(h && y_pred, "null argument");
    if (K < 0) K = h->K;
    Ctx* c = on_device(h->c);
    const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(h->n_local, m_total - h->row_offset));
    need(Y_local || m_local == 0, "labels missing");
    DevBuf<double> dY((size_t)std::max<int64_t>(m_local, 1)), dy((size_t)std::max<int64_t>(h->n_local, 1)),
        dc((size_t)std::max<int64_t>(h->n_local, 1));
    if (m_local) dY.upload(Y_local, m_local, c->stream);
    regression_fixed_dev(h, dY.p, m_total, K, t, noise, sigma, dy.p, cov ? dc.p : nullptr);
    if (h->n_local) {
      dy.download(y_pred, h->n_local, c->stream);
      if (cov) dc.download(cov, h->n_local, c->stream);
    }
    sync(c);
  });
}

int flgp_fit_lae_regression_fixed(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                  int64_t m_new, int d, int s, int r, int K, double sigma, double t, double noise,
                                  const char* subsample, const char* kernel, int gl, int root, int nstart, int iter_max,
                                  const int32_t* init_idx, uint64_t seed, double* train, double* test, double* cov) {
  if (K < 0) K = s;  // src/Fit.cpp:37-39
  flgp_spectrum* h = nullptr;
  int rc = flgp_heat_kernel_spectrum(ctx, X, m, X_new, m_new, d, s, r, K, subsample, kernel, gl, root, nstart, 0.1,
                                     iter_max, init_idx, seed, &h);
  if (rc) return rc;
  std::unique_ptr<flgp_spectrum> own(h);
  return guard([&] {
    need(Y && train, "null argument");
    const int64_t n = m + m_new;
    std::vector<double> y(n), cv(n);
    int rc2 = flgp_regression_fixed(h, Y, m, K, t, noise, sigma, y.data(), cv.data());
    if (rc2) fail(rc2, "%s", g_err.c_str());
    std::memcpy(train, y.data(), sizeof(double) * m);
    if (test && m_new) std::memcpy(test, y.data() + m, sizeof(double) * m_new);
    if (cov && m_new) std::memcpy(cov, cv.data() + m, sizeof(double) * m_new);
  });
}

static int approach_flag(const char* approach, bool* posterior) {
  const std::string a = approach ? approach : "posterior";
  if (a == "posterior") *posterior = true;
  else if (a == "marginal") *posterior = false;
  else return 1;
  return 0;
}

int flgp_regression_objective(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double sigma,
                              const char* approach, const double* pars, double* obj, double* grad) {
  return guard([&] {
    need(h && Y_local && pars && obj, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(h->c);
    const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(h->n_local, m_total - h->row_offset));
    DevBuf<double> dY(std::max<int64_t>(m_local, 1));
    if (m_local > 0) dY.upload(Y_local, m_local, c->stream);
    const RegTrain T = reg_train_prepare(h, dY.p, m_total, K, sigma);
    *obj = reg_objective(T, pars, grad, post);
  });
}

int flgp_train_regression(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double sigma,
                          const char* approach, double* pars_io, double* obj, int* nevals) {
  return guard([&] {
    need(h && Y_local && pars_io, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(h->c);
    const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(h->n_local, m_total - h->row_offset));
    DevBuf<double> dY(std::max<int64_t>(m_local, 1));
    if (m_local > 0) dY.upload(Y_local, m_local, c->stream);
    const RegTrain T = reg_train_prepare(h, dY.p, m_total, K, sigma);
    const double o = train_regression(T, post, pars_io, nevals);
    if (obj) *obj = o;
  });
}

int flgp_mma_minimize(int n, flgp_objective_fn f, void* data, const double* lb, const double* ub, double* x,
                      double* minf, double xtol_rel, int maxeval, int* nevals) {
  return guard([&] {
    need(n >= 1 && f && lb && ub && x && minf, "null argument");
    const int nev = mma_minimize(n, [&](const double* xx, double* g) { return f((unsigned)n, xx, g, data); }, lb, ub, x,
                                 minf, xtol_rel, maxeval > 0 ? maxeval : 1000);
    if (nevals) *nevals = nev;
  });
}

int flgp_fit_lae_regression(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                            int64_t m_new, int d, int s, int r, int K, double sigma, const char* approach,
                            const char* subsample, const char* kernel, int gl, int root, int nstart, int iter_max,
                            const int32_t* init_idx, uint64_t seed, double* pars_io, double* train, double* test,
                            double* cov, double* obj) {
  if (K < 0) K = s;  // src/Fit.cpp:37-39
  bool post = true;
  if (approach_flag(approach, &post)) {
    g_err = "This model selection approach is not supported!";
    return 2;
  }
  flgp_spectrum* h = nullptr;
  int rc = flgp_heat_kernel_spectrum(ctx, X, m, X_new, m_new, d, s, r, K, subsample, kernel, gl, root, nstart, 0.1,
                                     iter_max, init_idx, seed, &h);
  if (rc) return rc;
  std::unique_ptr<flgp_spectrum> own(h);
  return guard([&] {
    need(Y && train && pars_io, "null argument");
    if (!(pars_io[0] == pars_io[0] && pars_io[1] == pars_io[1])) {  // NaN: train (src/Fit.cpp:45-62)
      int rc1 = flgp_train_regression(h, Y, m, K, sigma, approach, pars_io, obj, nullptr);
      if (rc1) fail(rc1, "%s", g_err.c_str());
    } else if (obj) {
      int rc1 = flgp_regression_objective(h, Y, m, K, sigma, approach, pars_io, obj, nullptr);
      if (rc1) fail(rc1, "%s", g_err.c_str());
      *obj = -*obj;
    }
    const int64_t n = m + m_new;
    std::vector<double> y(n), cv(n);
    int rc2 = flgp_regression_fixed(h, Y, m, K, pars_io[0], pars_io[1], sigma, y.data(), cv.data());
    if (rc2) fail(rc2, "%s", g_err.c_str());
    std::memcpy(train, y.data(), sizeof(double) * m);
    if (test && m_new) std::memcpy(test, y.data() + m, sizeof(double) * m_new);
    if (cov && m_new) std::memcpy(cov, cv.data() + m, sizeof(double) * m_new);
  });
}

// ---- noise = "same" on explicit training rows (host only): the statistics the device path reduces on the GPU are
// formed here by plain loops; the objective / optimiser code is the one the handle-based entries run --------------------
static RegTrain make_reg_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma) {
  need(V1 && values && Y, "null argument");
  need(m >= 1 && K >= 1, "bad matrix shape");
  RegTrain T;
  T.m = m;
  T.K = K;
  T.sigma = sigma;
  T.ev.resize(K);
  for (int k = 0; k < K; ++k) T.ev[k] = 1.0 - values[k];
  if (m > K) {
    T.VtV.assign((size_t)K * K, 0.0);
    T.b.assign(K, 0.0);
    for (int i = 0; i < m; ++i) {
      const double* v = V1 + (size_t)i * K;
      for (int b = 0; b < K; ++b) {
        T.b[b] += v[b] * Y[i];
        for (int a = 0; a < K; ++a) T.VtV[a + (size_t)K * b] += v[a] * v[b];
      }
      T.yty += Y[i] * Y[i];
    }
  } else {
    T.V.assign(V1, V1 + (size_t)m * K);
    T.Y.assign(Y, Y + m);
  }
  return T;
}

int flgp_regression_objective_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                                   const char* approach, const double* pars, double* obj, double* grad) {
  return guard([&] {
    need(pars && obj, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const RegTrain T = make_reg_rows(V1, values, Y, m, K, sigma);
    double g[2] = {0.0, 0.0};
    *obj = reg_objective(T, pars, g, post);
    if (grad) {
      grad[0] = g[0];
      grad[1] = g[1];
    }
  });
}

int flgp_train_regression_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                               const char* approach, double* pars_io, double* obj, int* nevals) {
  return guard([&] {
    need(pars_io != nullptr, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const RegTrain T = make_reg_rows(V1, values, Y, m, K, sigma);
    const double o = train_regression(T, post, pars_io, nevals);
    if (obj) *obj = o;
  });
}

// ---- noise = "different" (src/train.cpp:438-556, src/Predict.cpp:76-113): host algebra on the m training rows -----------
static RegTrainDiff make_diff_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma) {
  need(V1 && values && Y, "null argument");
  need(m >= 1 && K >= 1 && m <= 8192, "noise=\"different\": need 1 <= m <= 8192 training rows");
  RegTrainDiff T;
  T.m = m;
  T.K = K;
  T.sigma = sigma;
  T.V.assign(V1, V1 + (size_t)m * K);
  T.Y.assign(Y, Y + m);
  T.ev.resize(K);
  for (int k = 0; k < K; ++k) T.ev[k] = 1.0 - values[k];
  return T;
}

int flgp_regression_objective_diff_rows(const double* V1, const double* values, const double* Y, int m, int K,
                                        double sigma, const char* approach, const double* x, double* obj,
                                        double* grad) {
  return guard([&] {
    need(x && obj, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const RegTrainDiff T = make_diff_rows(V1, values, Y, m, K, sigma);
    std::vector<double> g((size_t)m + 1);
    *obj = reg_objective_diff(T, x, g.data(), post);
    if (grad) std::memcpy(grad, g.data(), sizeof(double) * (m + 1));
  });
}

int flgp_train_regression_diff_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                                    const char* approach, double* x_io, double* obj, int* nevals) {
  return guard([&] {
    need(x_io, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const RegTrainDiff T = make_diff_rows(V1, values, Y, m, K, sigma);
    const double o = train_regression_diff(T, post, x_io, nevals);
    if (obj) *obj = o;
  });
}

int flgp_predict_coef_diff_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                                const double* x, double* coef) {
  return guard([&] {
    need(x && coef, "null argument");
    const RegTrainDiff T = make_diff_rows(V1, values, Y, m, K, sigma);
    std::vector<double> cf;
    if (!reg_diff_coef(T, x, cf)) fail(2, "regression: the covariance of the training rows is not positive definite");
    std::memcpy(coef, cf.data(), sizeof(double) * K);
  });
}

// fit_lae_regression_gp_cpp with noise = "different" (src/Fit.cpp:20-99): spectrum, training of (t, noise_1 .. noise_m)
// on the m training rows, mean folded through coef = Lam V1^T alpha; the posterior variance is the reference's
// posterior_covariance_regression(eigenpair, idx0, idx1, K, res.x, sigma), which reads pars[1] — the FIRST row's noise —
// as the common variance (src/Utils.cpp:218-220).
int flgp_fit_lae_regression_diff_noise(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                       int64_t m_new, int d, int s, int r, int K, double sigma, const char* approach,
                                       const char* subsample, const char* kernel, int gl, int root, int nstart,
                                       int iter_max, const int32_t* init_idx, uint64_t seed, double* pars_io,
                                       double* train, double* test, double* cov, double* obj) {
  if (K < 0) K = s;  // src/Fit.cpp:37-39
  bool post = true;
  if (approach_flag(approach, &post)) {
    g_err = "This model selection approach is not supported!";
    return 2;
  }
  flgp_spectrum* h = nullptr;
  int rc = flgp_heat_kernel_spectrum(ctx, X, m, X_new, m_new, d, s, r, K, subsample, kernel, gl, root, nstart, 0.1,
                                     iter_max, init_idx, seed, &h);
  if (rc) return rc;
  std::unique_ptr<flgp_spectrum> own(h);
  return guard([&] {
    need(Y && train && pars_io, "null argument");
    need(m >= 1 && m <= 8192, "noise=\"different\": need 1 <= m <= 8192 training rows");
    Ctx* c = on_device(h->c);
    const int KK = h->K;
    // the m training rows of the eigenvectors on the host (row-major m x K)
    DevBuf<double> V1((size_t)m * KK);
    lift_rows_run(c, h->r, h->Zj.p, h->Zx.p, h->w.p, h->Wm.p, KK, nullptr, m, V1.p, KK, false);
    std::vector<double> Vh((size_t)m * KK), Vk((size_t)m * K);
    V1.download(Vh.data(), Vh.size(), c->stream);
    sync(c);
    for (int64_t i = 0; i < m; ++i)
      for (int k = 0; k < K; ++k) Vk[(size_t)i * K + k] = Vh[(size_t)i * KK + k];
    const RegTrainDiff T = make_diff_rows(Vk.data(), h->values.data(), Y, (int)m, K, sigma);
    bool train_it = false;
    for (int64_t i = 0; i <= m; ++i) train_it = train_it || !(pars_io[i] == pars_io[i]);
    if (train_it) {
      const double o = train_regression_diff(T, post, pars_io, nullptr);
      if (obj) *obj = o;
    } else if (obj) {
      std::vector<double> g((size_t)m + 1);
      *obj = -reg_objective_diff(T, pars_io, g.data(), post);
    }
    std::vector<double> coef;
    if (!reg_diff_coef(T, pars_io, coef)) fail(2, "regression: the covariance of the training rows is not positive definite");
    coef.resize(KK, 0.0);
    const int64_t n = m + m_new;
    DevBuf<double> dY(m), dcoef(KK), wv(h->s), dy(n), dtmp(n), dc(n);
    dY.upload(Y, m, c->stream);
    dcoef.upload(coef.data(), KK, c->stream);
    gemv_run(c, h->Wm.p, dcoef.p, h->s, KK, wv.p);
    sparse_rowdot_run(c, n, h->r, h->Zj.p, h->Zx.p, h->w.p, wv.p, dy.p);
    regression_fixed_dev(h, dY.p, m, K, pars_io[0], pars_io[1], sigma, dtmp.p, dc.p);
    std::vector<double> y(n), cv(n);
    dy.download(y.data(), n, c->stream);
    dc.download(cv.data(), n, c->stream);
    sync(c);
    std::memcpy(train, y.data(), sizeof(double) * m);
    if (test && m_new) std::memcpy(test, y.data() + m, sizeof(double) * m_new);
    if (cov && m_new) std::memcpy(cov, cv.data() + m, sizeof(double) * m_new);
  });
}

int flgp_fit_se_regression(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                           int64_t m_new, int d, int s, int r, int K, double sigma, const double* a2s, int n_a2,
                           const char* approach, const char* subsample, int gl, int root, int nstart, int iter_max,
                           const int32_t* init_idx, uint64_t seed, const double* fixed_pars, double* train,
                           double* test, double* cov, double* pars_out, double* best_a2, double* best_obj,
                           flgp_spectrum** out) {
  return guard([&] {
    need(ctx && X && Y && train && pars_out, "null argument");
    need(m >= 1 && m_new >= 0, "bad matrix shape");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "flgp_fit_se_regression is the single-process entry point");
    const int64_t n = m + m_new;
    DevBuf<double> dX = upload_concat(c, X, m, X_new, m_new, d);
    DevBuf<double> dY(m);
    dY.upload(Y, m, c->stream);
    const Models mo = make_models(subsample, "se", gl, root, nstart, 0.1, iter_max);
    std::unique_ptr<flgp_spectrum> sp = se_grid_pipeline(c, dX.p, n, n, 0, d, s, r, K, mo, init_idx, seed, dY.p, m,
                                                         sigma, post, a2s, n_a2, fixed_pars, pars_out, best_a2,
                                                         best_obj);
    const int Kk = K < 0 ? s : K;
    DevBuf<double> dy(n), dc(n);
    regression_fixed_dev(sp.get(), dY.p, m, Kk, pars_out[0], pars_out[1], sigma, dy.p, dc.p);
    std::vector<double> y(n), cv(n);
    dy.download(y.data(), n, c->stream);
    dc.download(cv.data(), n, c->stream);
    sync(c);
    std::memcpy(train, y.data(), sizeof(double) * m);
    if (test && m_new) std::memcpy(test, y.data() + m, sizeof(double) * m_new);
    if (cov && m_new) std::memcpy(cov, cv.data() + m, sizeof(double) * m_new);
    if (out) *out = sp.release();
  });
}

// fit_nystrom_regression_gp_cpp (src/Fit.cpp:222-357) on one contiguous block of rows of X_all (dX: n_local x d on the
// device, rows row_offset ..; the first m_total rows of X_all are the training rows, dY their labels on this rank).
// The anchors and the s x s anchor operator are replicated (k-means sums all-reduced as everywhere); the extension is
// row-parallel; the K x K training statistics and the tail system are all-reduced.  y / cv: n_local on the host.
static void nystrom_fit(Ctx* c, const double* dX, int64_t n_local, int64_t n_total, int64_t row_offset, int d,
                        const double* dY, int64_t m_total, int s, int K, double sigma, const double* a2s, int n_a2,
                        bool post, const char* subsample, int nstart, int iter_max, const int32_t* init_idx,
                        uint64_t seed, const double* fixed_pars, double* y, double* cv, double* pars_out,
                        double* best_a2, double* best_obj) {
  const int64_t n = n_total, m = m_total;
  const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(n_local, m_total - row_offset));
  // anchors (src/Fit.cpp:242): subsample_cpp(...).leftCols(d)
  flgp_spectrum base;
  base.c = c;
  base.n_local = n_local;
  base.n_total = n;
  base.row_offset = row_offset;
  base.d = d;
  base.s = s;
  base.r = 1;
  const Models mo = make_models(subsample, "se", FLGP_GL_RW, 1, nstart, 0.1, iter_max);
  stage_subsample(c, &base, dX, mo, init_idx, seed, nullptr);
  base.sorted = KMeansSorted();
  const double* U = base.U.p;
  DevBuf<double> un(s), D((size_t)s * s);
  double dmean = 0.0;
  nys_anchor_distances_run(c, U, s, s, d, un.p, D.p, &dmean);
  // row blocks: the n_b x s weight block stays below ~512 MB
  const int64_t nb_max = std::max<int64_t>(256, std::min<int64_t>(std::max<int64_t>(n_local, 1), ((int64_t)64 << 20) / s));
  DevBuf<double> Wx((size_t)std::min<int64_t>(nb_max, std::max<int64_t>(n_local, 1)) * s);
  DevBuf<double> V1((size_t)std::max<int64_t>(m_local, 1) * K);
  struct Cand {
    DevBuf<double> rs, Bt;
    std::vector<double> lam;
    double denom = 0.0;
  } best;
  double max_obj = -std::numeric_limits<double>::infinity();
  bool have = false;
  for (int q = 0; q < n_a2; ++q) {  // grid search (src/Fit.cpp:262-312)
    Cand cd;
    cd.rs.alloc(s);
    cd.Bt.alloc((size_t)K * s);
    cd.denom = a2s[q] * dmean;
    DevBuf<double> lam(K);
    nys_anchor_operator_run(c, D.p, s, K, cd.denom, cd.rs.p, lam.p, cd.Bt.p);
    cd.lam.resize(K);
    lam.download(cd.lam.data(), K, c->stream);
    sync(c);
    for (int64_t r0 = 0; r0 < m_local; r0 += nb_max) {
      const int64_t nb = std::min<int64_t>(nb_max, m_local - r0);
      nys_extend_rows_run(c, dX, n_local, r0, nb, d, U, s, s, un.p, cd.rs.p, cd.denom, cd.Bt.p, K, Wx.p,
                          V1.p + (size_t)r0 * K);
    }
    const RegTrain T = reg_train_from_rows(c, V1.p, K, m_local, row_offset, dY, m, K, sigma, cd.lam);
    double x[2] = {std::nan(""), std::nan("")};
    double obj;
    if (fixed_pars) {
      x[0] = fixed_pars[0];
      x[1] = fixed_pars[1];
      obj = -reg_objective(T, x, nullptr, post);
    } else {
      obj = train_regression(T, post, x, nullptr);
    }
    if (obj > max_obj || !have) {
      max_obj = obj;
      pars_out[0] = x[0];
      pars_out[1] = x[1];
      if (best_a2) *best_a2 = a2s[q];
      best = std::move(cd);
      have = true;
    }
  }
  if (best_obj) *best_obj = max_obj;
  // predictions with the winning bandwidth (src/Fit.cpp:318-340)
  for (int64_t r0 = 0; r0 < m_local; r0 += nb_max) {
    const int64_t nb = std::min<int64_t>(nb_max, m_local - r0);
    nys_extend_rows_run(c, dX, n_local, r0, nb, d, U, s, s, un.p, best.rs.p, best.denom, best.Bt.p, K, Wx.p,
                        V1.p + (size_t)r0 * K);
  }
  DevBuf<double> dcoef, dM;
  gpr_tail_system(c, V1.p, K, m_local, row_offset, dY, m, K, best.lam, pars_out[0], pars_out[1], sigma, dcoef, dM);
  const double ns = pars_out[1] + sigma;
  const int64_t nbv = std::min<int64_t>(nb_max, std::max<int64_t>(n_local, 1));
  DevBuf<double> Vb((size_t)nbv * K), Tb((size_t)nbv * K), dy(std::max<int64_t>(n_local, 1)), dc(std::max<int64_t>(n_local, 1));
  for (int64_t r0 = 0; r0 < n_local; r0 += nb_max) {
    const int64_t nb = std::min<int64_t>(nb_max, n_local - r0);
    nys_extend_rows_run(c, dX, n_local, r0, nb, d, U, s, s, un.p, best.rs.p, best.denom, best.Bt.p, K, Wx.p, Vb.p);
    gemv_run(c, Vb.p, dcoef.p, nb, K, dy.p + r0);
    gemm_nn_run(c, Vb.p, dM.p, nb, K, K, Tb.p);  // M symmetric: row-major image = column-major image
    nys_rowdot_run(c, Tb.p, Vb.p, nb, K, ns, dc.p + r0);
  }
  if (n_local > 0) {
    dy.download(y, n_local, c->stream);
    dc.download(cv, n_local, c->stream);
  }
  sync(c);
}

int flgp_fit_nystrom_regression(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                int64_t m_new, int d, int s, int K, double sigma, const double* a2s, int n_a2,
                                const char* approach, const char* subsample, int nstart, int iter_max,
                                const int32_t* init_idx, uint64_t seed, const double* fixed_pars, double* train,
                                double* test, double* cov, double* pars_out, double* best_a2, double* best_obj) {
  return guard([&] {
    need(ctx && X && Y && train && pars_out && a2s, "null argument");
    need(m >= 1 && m_new >= 0 && d >= 1 && n_a2 >= 1, "bad matrix shape");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "flgp_fit_nystrom_regression is the single-process entry point (see ..._sharded)");
    const int64_t n = m + m_new;
    need(s >= 1 && s <= n && n < ((int64_t)1 << 31), "need 1 <= s <= n");
    if (K < 0) K = s;  // R/Fit.R:183-185
    need(K >= 1 && K <= s, "need 1 <= K <= s");
    DevBuf<double> dX = upload_concat(c, X, m, X_new, m_new, d);
    DevBuf<double> dY(m);
    dY.upload(Y, m, c->stream);
    std::vector<double> y(n), cv(n);
    nystrom_fit(c, dX.p, n, n, 0, d, dY.p, m, s, K, sigma, a2s, n_a2, post, subsample, nstart, iter_max, init_idx, seed,
                fixed_pars, y.data(), cv.data(), pars_out, best_a2, best_obj);
    std::memcpy(train, y.data(), sizeof(double) * m);
    if (test && m_new) std::memcpy(test, y.data() + m, sizeof(double) * m_new);
    if (cov && m_new) std::memcpy(cov, cv.data() + m, sizeof(double) * m_new);
  });
}

int flgp_fit_nystrom_regression_sharded(flgp_ctx* ctx, const double* X_local, int64_t n_local, int64_t n_total,
                                        int64_t row_offset, int d, const double* Y_local, int64_t m_total, int s, int K,
                                        double sigma, const double* a2s, int n_a2, const char* approach,
                                        const char* subsample, int nstart, int iter_max, const int32_t* init_idx,
                                        uint64_t seed, const double* fixed_pars, double* mean_local, double* cov_local,
                                        double* pars_out, double* best_a2, double* best_obj) {
  return guard([&] {
    need(ctx && mean_local && cov_local && pars_out && a2s, "null argument");
    need(n_local >= 0 && n_total >= 1 && d >= 1 && n_a2 >= 1 && m_total >= 1 && m_total <= n_total, "bad matrix shape");
    need(row_offset >= 0 && row_offset + n_local <= n_total, "shard outside the matrix");
    need(n_local == 0 || X_local, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(&ctx->c);
    need(s >= 1 && s <= n_total && n_total < ((int64_t)1 << 31), "need 1 <= s <= n");
    if (K < 0) K = s;
    need(K >= 1 && K <= s, "need 1 <= K <= s");
    const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(n_local, m_total - row_offset));
    need(m_local == 0 || Y_local, "null argument");
    DevBuf<double> dX((size_t)std::max<int64_t>(n_local * d, 1)), dY(std::max<int64_t>(m_local, 1));
    if (n_local > 0) dX.upload(X_local, (size_t)n_local * d, c->stream);
    if (m_local > 0) dY.upload(Y_local, m_local, c->stream);
    nystrom_fit(c, dX.p, n_local, n_total, row_offset, d, dY.p, m_total, s, K, sigma, a2s, n_a2, post, subsample, nstart,
                iter_max, init_idx, seed, fixed_pars, mean_local, cov_local, pars_out, best_a2, best_obj);
  });
}

// the m labelled rows of the lifted eigenvectors on the host (m x K row-major), gathered over the ranks
static LogitTrain logit_train_prepare(flgp_spectrum* sp, const double* Y, const double* N, int64_t m_total, int K,
                                      double sigma, bool posterior) {
  Ctx* c = sp->c;
  need(K >= 1 && K <= sp->K, "K exceeds the number of computed eigenpairs");
  need(m_total >= 1 && m_total <= sp->n_total && m_total <= 8192, "classification: need 1 <= m <= 8192 labelled rows");
  const int KK = sp->K, m = (int)m_total;
  const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(sp->n_local, m_total - sp->row_offset));
  DevBuf<double> V1((size_t)std::max<int64_t>(m_local * KK, 1)), Vall((size_t)m * KK);
  lift_rows_run(c, sp->r, sp->Zj.p, sp->Zx.p, sp->w.p, sp->Wm.p, KK, nullptr, m_local, V1.p, KK, false);
  Vall.zero(c->stream);
  if (m_local > 0)
    FLGP_CUDA(cudaMemcpyAsync(Vall.p + (size_t)sp->row_offset * KK, V1.p, sizeof(double) * m_local * KK,
                              cudaMemcpyDeviceToDevice, c->stream));
  comm_allreduce_f64(c, Vall.p, (size_t)m * KK);
  std::vector<double> Vh((size_t)m * KK);
  Vall.download(Vh.data(), Vh.size(), c->stream);
  sync(c);
  LogitTrain T;
  T.m = m;
  T.K = K;
  T.sigma = sigma;
  T.posterior = posterior;
  T.V.resize((size_t)m * K);
  for (int i = 0; i < m; ++i)
    for (int k = 0; k < K; ++k) T.V[(size_t)i * K + k] = Vh[(size_t)i * KK + k];
  T.ev.resize(K);
  for (int k = 0; k < K; ++k) T.ev[k] = 1.0 - sp->values[k];
  T.Y.assign(Y, Y + m);
  T.N.assign(m, 1.0);
  if (N) T.N.assign(N, N + m);
  return T;
}

int flgp_logit_objective(flgp_spectrum* h, const double* Y, const double* N, int64_t m_total, int K, double sigma,
                         const char* approach, double t, double* obj) {
  return guard([&] {
    need(h && Y && obj, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const LogitTrain T = logit_train_prepare(h, Y, N, m_total, K, sigma, post);
    *obj = logit_objective(T, t);
  });
}

int flgp_cobyla_minimize_1d(flgp_objective_fn f, void* data, double lb, double ub, double* x, double* minf,
                            double xtol_rel, int maxeval, int* nevals) {
  return guard([&] {
    need(f && x, "null argument");
    need(lb <= *x && *x <= ub, "the start lies outside the bounds");
    auto fn = [&](double t) { return f(1, &t, nullptr, data); };
    *x = cobyla_minimize_1d(fn, *x, lb, ub, xtol_rel > 0 ? xtol_rel : 1e-4, maxeval > 0 ? maxeval : 1000, minf, nevals);
  });
}

int flgp_train_logit(flgp_spectrum* h, const double* Y, const double* N, int64_t m_total, int K, double sigma,
                     const char* approach, double* t_io, double* obj, int* nevals) {
  return guard([&] {
    need(h && Y && t_io, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const LogitTrain T = logit_train_prepare(h, Y, N, m_total, K, sigma, post);
    double t0 = *t_io;
    if (!(t0 == t0) || t0 < 0.0) t0 = 10.0;  // src/train.cpp:41-43
    double fmin = 0.0;
    auto fn = [&](double t) { return logit_objective(T, t); };
    *t_io = cobyla_minimize_1d(fn, std::max(t0, 1e-3), 1e-3, HUGE_VAL, 1e-4, 1000, &fmin, nevals);
    if (obj) *obj = -fmin;  // ReturnValue(t, -obj), src/train.cpp:70
  });
}

// multi_train_split (src/MultiClassification.cpp:14-27): J = max(Y) + 1, class j against the rest
static int multi_class_count(const double* Y, int64_t m_total) {
  double ymax = 0.0;
  for (int64_t i = 0; i < m_total; ++i) {
    need(Y[i] >= 0.0 && Y[i] == std::floor(Y[i]), "multi-class labels must be the integers 0 .. J-1");
    ymax = std::max(ymax, Y[i]);
  }
  return (int)ymax + 1;
}

// the J binary trainings (src/MultiClassification.cpp:41-50) are independent host-side Newton / COBYLA loops
static void train_logit_classes(const LogitTrain& base, const double* Y, int64_t m_total, int J, double* t_out,
                                double* obj_out) {
  std::vector<std::string> errs(J);
  auto one = [&](int j) {
    try {
      LogitTrain T = base;
      for (int64_t i = 0; i < m_total; ++i) T.Y[i] = (Y[i] == (double)j) ? 1.0 : 0.0;
      double fmin = 0.0;
      auto fn = [&](double t) { return logit_objective(T, t); };
      t_out[j] = cobyla_minimize_1d(fn, 10.0, 1e-3, HUGE_VAL, 1e-4, 1000, &fmin, nullptr);
      if (obj_out) obj_out[j] = -fmin;
    } catch (const std::exception& e) {
      errs[j] = e.what();
    }
  };
  const int nthr = std::max(1u, std::min<unsigned>(J, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  std::atomic<int> next{0};
  for (int q = 0; q < nthr; ++q)
    pool.emplace_back([&] {
      g_outer_workers = nthr;  // the dense algebra inside shares the cores with its siblings
      for (int j = next++; j < J; j = next++) one(j);
    });
  for (auto& th : pool) th.join();
  for (int j = 0; j < J; ++j)
    if (!errs[j].empty()) fail(3, "class %d: %s", j, errs[j].c_str());
}

int flgp_train_logit_mult(flgp_spectrum* h, const double* Y, int64_t m_total, int K, double sigma,
                          const char* approach, int J_cap, int* J_out, double* t_out, double* obj_out) {
  return guard([&] {
    need(h && Y && J_out && t_out, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const int J = multi_class_count(Y, m_total);
    *J_out = J;
    need(J <= J_cap, "more classes than the output arrays hold");
    std::vector<double> y0((size_t)m_total, 0.0);  // placeholder labels: every class fills in its own below
    LogitTrain base = logit_train_prepare(h, y0.data(), nullptr, m_total, K, sigma, post);  // V, ev: shared by all classes
    train_logit_classes(base, Y, m_total, J, t_out, obj_out);
  });
}

// the tail of the binary logit drivers: posterior_distribution_classification on the test rows and the optional
// covariance block C = [Cvv + sigma I; Cnv] (src/Fit.cpp:566-582, 752-773)
static void logit_fit_tail(flgp_spectrum* h, const double* Y, int64_t m, int64_t m_new, int K, double t, double sigma,
                           double* post_mean, double* post_cov, double* C_out) {
  const int64_t n = m + m_new;
  if (post_mean) {
    std::vector<double> mean(n), cov(n);
    int rc2 = flgp_classification_posterior_fixed(h, Y, m, K, t, sigma, 1e-5, 100, mean.data(),
                                                  post_cov ? cov.data() : nullptr);
    if (rc2) fail(rc2, "%s", g_err.c_str());
    if (m_new) std::memcpy(post_mean, mean.data() + m, sizeof(double) * m_new);
    if (post_cov && m_new) std::memcpy(post_cov, cov.data() + m, sizeof(double) * m_new);
  }
  if (C_out) {
    std::vector<int32_t> i0(n), i1(m);
    for (int64_t i = 0; i < n; ++i) i0[i] = (int32_t)i;
    for (int64_t i = 0; i < m; ++i) i1[i] = (int32_t)i;
    int rc3 = flgp_hk_from_spectrum(h, K, t, i0.data(), n, i1.data(), m, C_out);
    if (rc3) fail(rc3, "%s", g_err.c_str());
    for (int64_t i = 0; i < m; ++i) C_out[i + n * i] += sigma;
  }
}

int flgp_fit_lae_logit(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m, int64_t m_new,
                       int d, int s, int r, int K, const double* N, double sigma, const char* approach,
                       const char* subsample, const char* kernel, int gl, int root, int nstart, int iter_max,
                       const int32_t* init_idx, uint64_t seed, double* t_io, double* post_mean, double* post_cov,
                       double* C_out, double* obj) {
  if (K < 0) K = s;  // src/Fit.cpp:538-540
  bool post = true;
  if (approach_flag(approach, &post)) {
    g_err = "This model selection approach is not supported!";
    return 2;
  }
  flgp_spectrum* h = nullptr;
  int rc = flgp_heat_kernel_spectrum(ctx, X, m, X_new, m_new, d, s, r, K, subsample, kernel, gl, root, nstart, 0.1,
                                     iter_max, init_idx, seed, &h);
  if (rc) return rc;
  std::unique_ptr<flgp_spectrum> own(h);
  return guard([&] {
    need(Y && t_io, "null argument");
    if (!(*t_io == *t_io)) {  // NaN: train t (src/Fit.cpp:548-563)
      int rc1 = flgp_train_logit(h, Y, N, m, K, sigma, approach, t_io, obj, nullptr);
      if (rc1) fail(rc1, "%s", g_err.c_str());
    } else if (obj) {
      int rc1 = flgp_logit_objective(h, Y, N, m, K, sigma, approach, *t_io, obj);
      if (rc1) fail(rc1, "%s", g_err.c_str());
      *obj = -*obj;
    }
    logit_fit_tail(h, Y, m, m_new, K, *t_io, sigma, post_mean, post_cov, C_out);
  });
}

// fit_se_logit_gp_cpp (src/Fit.cpp:668-794): one k-means + KNN, per bandwidth a2 the SE weights, graph Laplacian,
// spectrum and the COBYLA training of t (or the objective at the given t); the a2 with the largest objective wins.
int flgp_fit_se_logit(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m, int64_t m_new,
                      int d, int s, int r, int K, const double* N, double sigma, const double* a2s, int n_a2,
                      const char* approach, const char* subsample, int gl, int root, int nstart, int iter_max,
                      const int32_t* init_idx, uint64_t seed, double* t_io, double* post_mean, double* post_cov,
                      double* C_out, double* best_a2, double* best_obj, flgp_spectrum** out) {
  return guard([&] {
    need(ctx && X && Y && t_io && a2s, "null argument");
    need(m >= 1 && m_new >= 0 && n_a2 >= 1, "bad matrix shape");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "flgp_fit_se_logit is the single-process entry point");
    const int64_t n = m + m_new;
    if (K < 0) K = s;  // src/Fit.cpp:686-688
    DevBuf<double> dX = upload_concat(c, X, m, X_new, m_new, d);
    const Models mo = make_models(subsample, "se", gl, root, nstart, 0.1, iter_max);
    const bool fixed = (*t_io == *t_io);
    const double t_fixed = *t_io;
    std::vector<LogitTrain> Ts(n_a2);
    std::vector<double> ts(n_a2);
    int bq = 0;
    std::unique_ptr<flgp_spectrum> sp = se_grid_search(
        c, dX.p, n, n, 0, d, s, r, K, mo, init_idx, seed, a2s, n_a2, !fixed,
        [&](int q, flgp_spectrum* h) { Ts[q] = logit_train_prepare(h, Y, N, m, K, sigma, post); },
        [&](int q) {
          if (fixed) {
            ts[q] = t_fixed;
            return -logit_objective(Ts[q], t_fixed);
          }
          double fmin = 0.0;
          auto fn = [&](double t) { return logit_objective(Ts[q], t); };
          ts[q] = cobyla_minimize_1d(fn, 10.0, 1e-3, HUGE_VAL, 1e-4, 1000, &fmin, nullptr);  // src/train.cpp:38-71
          return -fmin;
        },
        &bq, best_a2, best_obj);
    *t_io = ts[bq];
    logit_fit_tail(sp.get(), Y, m, m_new, K, *t_io, sigma, post_mean, post_cov, C_out);
    if (out) *out = sp.release();
  });
}

// fit_se_logit_mult_gp_cpp (src/Fit.cpp:797-895): as above with the J one-vs-rest trainings of train_logit_mult_gp_cpp
// per bandwidth; a grid point's objective is the sum of its J class objectives (:855-859).
int flgp_fit_se_logit_mult(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                           int64_t m_new, int d, int s, int r, int K, double sigma, const double* a2s, int n_a2,
                           const char* approach, const char* subsample, int gl, int root, int nstart, int iter_max,
                           const int32_t* init_idx, uint64_t seed, int J_cap, int* J_out, double* t_out,
                           double* obj_out, double* best_a2, double* best_obj, flgp_spectrum** out) {
  return guard([&] {
    need(ctx && X && Y && a2s && J_out && t_out, "null argument");
    need(m >= 1 && m_new >= 0 && n_a2 >= 1, "bad matrix shape");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "flgp_fit_se_logit_mult is the single-process entry point");
    const int64_t n = m + m_new;
    if (K < 0) K = s;  // src/Fit.cpp:813-815
    const int J = multi_class_count(Y, m);
    *J_out = J;
    need(J <= J_cap, "more classes than the output arrays hold");
    DevBuf<double> dX = upload_concat(c, X, m, X_new, m_new, d);
    const Models mo = make_models(subsample, "se", gl, root, nstart, 0.1, iter_max);
    std::vector<double> y0((size_t)m, 0.0);  // placeholder labels: every class fills in its own
    std::vector<LogitTrain> Ts(n_a2);
    std::vector<double> ts((size_t)n_a2 * J), os((size_t)n_a2 * J);
    int bq = 0;
    std::unique_ptr<flgp_spectrum> sp = se_grid_search(
        c, dX.p, n, n, 0, d, s, r, K, mo, init_idx, seed, a2s, n_a2, false,
        [&](int q, flgp_spectrum* h) { Ts[q] = logit_train_prepare(h, y0.data(), nullptr, m, K, sigma, post); },
        [&](int q) {  // the J trainings of a grid point run on J threads; the grid points one after the other
          train_logit_classes(Ts[q], Y, m, J, &ts[(size_t)q * J], &os[(size_t)q * J]);
          double sum = 0.0;
          for (int j = 0; j < J; ++j) sum += os[(size_t)q * J + j];
          return sum;
        },
        &bq, best_a2, best_obj);
    for (int j = 0; j < J; ++j) {
      t_out[j] = ts[(size_t)bq * J + j];
      if (obj_out) obj_out[j] = os[(size_t)bq * J + j];
    }
    if (out) *out = sp.release();
  });
}

// fit_nystrom_logit_gp_cpp (src/Fit.cpp:896-1038) and the training half of fit_nystrom_logit_mult_gp_cpp
// (src/Fit.cpp:1045-1162): the Nystrom grid of nystrom_fit with the logit trainings in place of MMA.  Single process.
// J = 0: binary — t_io (NaN trains), Laplace posterior of the test rows, optional C = [Cvv + sigma I; Cnv] (n x m).
// J > 0: the J one-vs-rest trainings per bandwidth, summed objective selects — t_out / obj_out of the winner.
// values_out (K) / vectors_out (n x K column-major), optional: the winning extended eigenpair (what the reference's
// prediction code consumes).
static void nystrom_logit_run(Ctx* c, const double* X, const double* Y, const double* X_new, int64_t m, int64_t m_new,
                              int d, int s, int K, const double* N, double sigma, const double* a2s, int n_a2,
                              bool post, const char* subsample, int nstart, int iter_max, const int32_t* init_idx,
                              uint64_t seed, int J, double* t_io, double* post_mean, double* post_cov, double* C_out,
                              double* t_out, double* obj_out, double* values_out, double* vectors_out, double* best_a2,
                              double* best_obj) {
  const int64_t n = m + m_new;
  need(m >= 1 && m <= 8192, "classification: need 1 <= m <= 8192 labelled rows");
  DevBuf<double> dX = upload_concat(c, X, m, X_new, m_new, d);
  flgp_spectrum base;  // anchors (src/Fit.cpp:919): subsample_cpp(...).leftCols(d)
  base.c = c;
  base.n_local = base.n_total = n;
  base.d = d;
  base.s = s;
  base.r = 1;
  const Models mo = make_models(subsample, "se", FLGP_GL_RW, 1, nstart, 0.1, iter_max);
  stage_subsample(c, &base, dX.p, mo, init_idx, seed, nullptr);
  base.sorted = KMeansSorted();
  const double* U = base.U.p;
  DevBuf<double> un(s), D((size_t)s * s);
  double dmean = 0.0;
  nys_anchor_distances_run(c, U, s, s, d, un.p, D.p, &dmean);
  const int64_t nb_max = std::max<int64_t>(256, std::min<int64_t>(n, ((int64_t)64 << 20) / s));
  DevBuf<double> Wx((size_t)std::min<int64_t>(nb_max, n) * s), V1((size_t)m * K);
  struct Cand {
    DevBuf<double> rs, Bt;
    std::vector<double> lam;
    double denom = 0.0;
    LogitTrain T;
  };
  std::vector<Cand> cds(n_a2);
  for (int q = 0; q < n_a2; ++q) {  // device side of the grid (src/Fit.cpp:942-968): anchor operator, labelled rows
    Cand& cd = cds[q];
    cd.rs.alloc(s);
    cd.Bt.alloc((size_t)K * s);
    cd.denom = a2s[q] * dmean;
    DevBuf<double> lam(K);
    nys_anchor_operator_run(c, D.p, s, K, cd.denom, cd.rs.p, lam.p, cd.Bt.p);
    cd.lam.resize(K);
    lam.download(cd.lam.data(), K, c->stream);
    for (int64_t r0 = 0; r0 < m; r0 += nb_max) {
      const int64_t nb = std::min<int64_t>(nb_max, m - r0);
      nys_extend_rows_run(c, dX.p, n, r0, nb, d, U, s, s, un.p, cd.rs.p, cd.denom, cd.Bt.p, K, Wx.p,
                          V1.p + (size_t)r0 * K);
    }
    LogitTrain& T = cd.T;
    T.m = (int)m;
    T.K = K;
    T.sigma = sigma;
    T.posterior = post;
    T.V.resize((size_t)m * K);
    V1.download(T.V.data(), (size_t)m * K, c->stream);
    sync(c);
    T.ev.resize(K);
    for (int k = 0; k < K; ++k) T.ev[k] = 1.0 - cd.lam[k];
    T.Y.assign(m, 0.0);
    if (J == 0) T.Y.assign(Y, Y + m);
    T.N.assign(m, 1.0);
    if (N && J == 0) T.N.assign(N, N + m);
  }
  // the trainings: host work on m x K rows (src/Fit.cpp:970-990)
  const bool fixed = J == 0 && (*t_io == *t_io);
  const double t_fixed = J == 0 ? *t_io : 0.0;
  const int Jn = std::max(J, 1);
  std::vector<double> objs(n_a2), ts((size_t)n_a2 * Jn), os((size_t)n_a2 * Jn);
  std::vector<std::string> errs(n_a2);
  auto train_one = [&](int q) {
    try {
      const LogitTrain& T = cds[q].T;
      if (J > 0) {
        train_logit_classes(T, Y, m, J, &ts[(size_t)q * J], &os[(size_t)q * J]);
        double sum = 0.0;
        for (int j = 0; j < J; ++j) sum += os[(size_t)q * J + j];
        objs[q] = sum;
      } else if (fixed) {
        ts[q] = t_fixed;
        objs[q] = -logit_objective(T, t_fixed);
      } else {
        double fmin = 0.0;
        auto fn = [&](double t) { return logit_objective(T, t); };
        ts[q] = cobyla_minimize_1d(fn, 10.0, 1e-3, HUGE_VAL, 1e-4, 1000, &fmin, nullptr);
        objs[q] = -fmin;
      }
    } catch (const std::exception& e) {
      errs[q] = e.what();
      if (errs[q].empty()) errs[q] = "training failed";
    }
  };
  if (J == 0 && !fixed && n_a2 > 1) {
    std::vector<std::thread> th;
    for (int q = 0; q < n_a2; ++q)
      th.emplace_back([&, q] {
        g_outer_workers = n_a2;
        train_one(q);
      });
    for (auto& t : th) t.join();
  } else {
    for (int q = 0; q < n_a2; ++q) train_one(q);
  }
  for (int q = 0; q < n_a2; ++q)
    if (!errs[q].empty()) fail(3, "%s", errs[q].c_str());
  int bq = 0;
  double max_obj = -std::numeric_limits<double>::infinity();
  for (int q = 0; q < n_a2; ++q)
    if (objs[q] > max_obj || q == 0) {  // src/Fit.cpp:992-998
      max_obj = objs[q];
      bq = q;
    }
  if (best_a2) *best_a2 = a2s[bq];
  if (best_obj) *best_obj = max_obj;
  Cand& best = cds[bq];
  if (J > 0) {
    for (int j = 0; j < J; ++j) {
      t_out[j] = ts[(size_t)bq * J + j];
      if (obj_out) obj_out[j] = os[(size_t)bq * J + j];
    }
  } else {
    *t_io = ts[bq];
  }
  if (values_out) std::memcpy(values_out, best.lam.data(), sizeof(double) * K);
  const bool want_post = J == 0 && post_mean;
  if (!want_post && !vectors_out && !(J == 0 && C_out)) return;
  // the winning bandwidth's extension of every row (src/Fit.cpp:1003-1010), in blocks; binary: Laplace posterior folded
  // through coef / Mq (classification_posterior_dev), covariance block C = V Lam V1^T
  DevBuf<double> dcoef, dM, dlamt, V1b, Cdev;
  const double t = J == 0 ? *t_io : 0.0;
  if (J == 0) {
    std::vector<double> coef, Mq, lamt(K);
    if (want_post) {
      laplace_fold(best.T.V.data(), K, K, best.T.ev, Y, (int)m, t, sigma, 1e-5, 100, coef, Mq);
      dcoef.alloc(K);
      dM.alloc((size_t)K * K);
      dcoef.upload(coef.data(), K, c->stream);
      dM.upload(Mq.data(), (size_t)K * K, c->stream);
    }
    if (C_out) {
      for (int k = 0; k < K; ++k) lamt[k] = std::exp(-t * best.T.ev[k]);
      dlamt.alloc(K);
      dlamt.upload(lamt.data(), K, c->stream);
      V1b.alloc((size_t)m * K);
      V1b.upload(best.T.V.data(), (size_t)m * K, c->stream);
      Cdev.alloc((size_t)n * m);
    }
  }
  const int64_t nbv = std::min<int64_t>(nb_max, n);
  DevBuf<double> Vb((size_t)nbv * K), Tb((size_t)nbv * K), dmean_all(n), dcov_all(n);
  std::vector<double> blk;
  for (int64_t r0 = 0; r0 < n; r0 += nb_max) {
    const int64_t nb = std::min<int64_t>(nb_max, n - r0);
    nys_extend_rows_run(c, dX.p, n, r0, nb, d, U, s, s, un.p, best.rs.p, best.denom, best.Bt.p, K, Wx.p, Vb.p);
    if (want_post) {
      gemv_run(c, Vb.p, dcoef.p, nb, K, dmean_all.p + r0);
      gemm_nn_run(c, Vb.p, dM.p, nb, K, K, Tb.p);  // Mq symmetric up to rounding, as in classification_posterior_dev
      nys_rowdot_run(c, Tb.p, Vb.p, nb, K, sigma, dcov_all.p + r0);
    }
    if (J == 0 && C_out) gemm_nt_run(c, Vb.p, V1b.p, dlamt.p, nb, m, K, Cdev.p + r0, n);
    if (vectors_out) {
      blk.resize((size_t)nb * K);
      Vb.download(blk.data(), (size_t)nb * K, c->stream);
      sync(c);
      for (int64_t i = 0; i < nb; ++i)
        for (int k = 0; k < K; ++k) vectors_out[(r0 + i) + n * (int64_t)k] = blk[(size_t)i * K + k];
    }
  }
  if (want_post) {
    std::vector<double> mean(n), cov(n);
    dmean_all.download(mean.data(), n, c->stream);
    dcov_all.download(cov.data(), n, c->stream);
    sync(c);
    if (m_new) std::memcpy(post_mean, mean.data() + m, sizeof(double) * m_new);
    if (post_cov && m_new) std::memcpy(post_cov, cov.data() + m, sizeof(double) * m_new);
  }
  if (J == 0 && C_out) {
    Cdev.download(C_out, (size_t)n * m, c->stream);
    sync(c);
    for (int64_t i = 0; i < m; ++i) C_out[i + n * i] += sigma;  // Cvv.diagonal() += sigma (src/Fit.cpp:1015)
  }
  sync(c);
}

int flgp_fit_nystrom_logit(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                           int64_t m_new, int d, int s, int K, const double* N, double sigma, const double* a2s,
                           int n_a2, const char* approach, const char* subsample, int nstart, int iter_max,
                           const int32_t* init_idx, uint64_t seed, double* t_io, double* post_mean, double* post_cov,
                           double* C_out, double* best_a2, double* best_obj) {
  return guard([&] {
    need(ctx && X && Y && t_io && a2s, "null argument");
    need(m >= 1 && m_new >= 0 && d >= 1 && n_a2 >= 1, "bad matrix shape");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "flgp_fit_nystrom_logit is the single-process entry point");
    const int64_t n = m + m_new;
    need(s >= 1 && s <= n && n < ((int64_t)1 << 31), "need 1 <= s <= n");
    if (K < 0) K = s;
    need(K >= 1 && K <= s, "need 1 <= K <= s");
    nystrom_logit_run(c, X, Y, X_new, m, m_new, d, s, K, N, sigma, a2s, n_a2, post, subsample, nstart, iter_max,
                      init_idx, seed, 0, t_io, post_mean, post_cov, C_out, nullptr, nullptr, nullptr, nullptr, best_a2,
                      best_obj);
  });
}

int flgp_fit_nystrom_logit_mult(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                int64_t m_new, int d, int s, int K, double sigma, const double* a2s, int n_a2,
                                const char* approach, const char* subsample, int nstart, int iter_max,
                                const int32_t* init_idx, uint64_t seed, int J_cap, int* J_out, double* t_out,
                                double* obj_out, double* values_out, double* vectors_out, double* best_a2,
                                double* best_obj) {
  return guard([&] {
    need(ctx && X && Y && a2s && J_out && t_out, "null argument");
    need(m >= 1 && m_new >= 0 && d >= 1 && n_a2 >= 1, "bad matrix shape");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    Ctx* c = on_device(&ctx->c);
    need(c->nranks == 1, "flgp_fit_nystrom_logit_mult is the single-process entry point");
    const int64_t n = m + m_new;
    need(s >= 1 && s <= n && n < ((int64_t)1 << 31), "need 1 <= s <= n");
    if (K < 0) K = s;
    need(K >= 1 && K <= s, "need 1 <= K <= s");
    const int J = multi_class_count(Y, m);
    *J_out = J;
    need(J <= J_cap, "more classes than the output arrays hold");
    nystrom_logit_run(c, X, Y, X_new, m, m_new, d, s, K, nullptr, sigma, a2s, n_a2, post, subsample, nstart, iter_max,
                      init_idx, seed, J, nullptr, nullptr, nullptr, nullptr, t_out, obj_out, values_out, vectors_out,
                      best_a2, best_obj);
  });
}

// ---- the small host-side exports of the reference (no device work: m-sized dense algebra, as in the reference) -------
int flgp_marginal_log_likelihood_logit_la(const double* Cm, const double* Y, const double* N, int m, double tol,
                                          int max_iter, double* out) {
  return guard([&] {
    need(Cm && Y && out, "null argument");
    need(m >= 1 && m <= 8192, "classification: need 1 <= m <= 8192 labelled rows");
    std::vector<double> Cv(Cm, Cm + (size_t)m * m), ones;
    if (!N) {
      ones.assign(m, 1.0);
      N = ones.data();
    }
    *out = laplace_mll(Cv, Y, N, m, tol > 0.0 ? tol : 1e-5, max_iter > 0 ? max_iter : 100);
  });
}

static LogitTrain make_logit_rows(const double* V1, const double* values, const double* Y, const double* N, int m, int K,
                                  double sigma, bool posterior) {
  need(V1 && values && Y, "null argument");
  need(m >= 1 && m <= 8192 && K >= 1, "classification: need 1 <= m <= 8192 labelled rows");
  LogitTrain T;
  T.m = m;
  T.K = K;
  T.sigma = sigma;
  T.posterior = posterior;
  T.V.assign(V1, V1 + (size_t)m * K);
  T.ev.resize(K);
  for (int k = 0; k < K; ++k) T.ev[k] = 1.0 - values[k];
  T.Y.assign(Y, Y + m);
  T.N.assign(m, 1.0);
  if (N) T.N.assign(N, N + m);
  return T;
}

int flgp_logit_objective_rows(const double* V1, const double* values, const double* Y, const double* N, int m, int K,
                              double sigma, const char* approach, double t, double* obj) {
  return guard([&] {
    need(obj != nullptr, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    *obj = logit_objective(make_logit_rows(V1, values, Y, N, m, K, sigma, post), t);
  });
}

int flgp_train_logit_rows(const double* V1, const double* values, const double* Y, const double* N, int m, int K,
                          double sigma, const char* approach, double* t_io, double* obj, int* nevals) {
  return guard([&] {
    need(t_io != nullptr, "null argument");
    bool post = true;
    if (approach_flag(approach, &post)) fail(2, "This model selection approach is not supported!");
    const LogitTrain T = make_logit_rows(V1, values, Y, N, m, K, sigma, post);
    double t0 = *t_io;
    if (!(t0 == t0) || t0 < 0.0) t0 = 10.0;  // src/train.cpp:41-43
    double fmin = 0.0;
    auto fn = [&](double t) { return logit_objective(T, t); };
    *t_io = cobyla_minimize_1d(fn, std::max(t0, 1e-3), 1e-3, HUGE_VAL, 1e-4, 1000, &fmin, nevals);
    if (obj) *obj = -fmin;
  });
}

int flgp_classification_fold_rows(const double* V1, const double* values, const double* Y, int m, int K, double t,
                                  double sigma, double tol, int max_iter, double* coef, double* Mq) {
  return guard([&] {
    need(V1 && values && Y && coef && Mq, "null argument");
    need(m >= 1 && m <= 8192 && K >= 1, "classification: need 1 <= m <= 8192 labelled rows");
    std::vector<double> ev(K), cf, mq;
    for (int k = 0; k < K; ++k) ev[k] = 1.0 - values[k];
    laplace_fold(V1, K, K, ev, Y, m, t, sigma, tol > 0.0 ? tol : 1e-5, max_iter > 0 ? max_iter : 100, cf, mq);
    std::memcpy(coef, cf.data(), sizeof(double) * K);
    std::memcpy(Mq, mq.data(), sizeof(double) * K * K);
  });
}

int flgp_multi_train_split(const double* Y, int64_t m, int J_cap, int* J_out, double* aug_y) {
  return guard([&] {
    need(Y && J_out, "null argument");
    need(m >= 1, "bad matrix shape");
    const int J = multi_class_count(Y, m);
    *J_out = J;
    if (!aug_y) return;
    need(J <= J_cap, "more classes than the output array holds");
    for (int j = 0; j < J; ++j)
      for (int64_t i = 0; i < m; ++i) aug_y[i + m * (int64_t)j] = (Y[i] == (double)j) ? 1.0 : 0.0;
  });
}

int flgp_negative_log_likelihood(const double* mean, const double* cov, const double* target, int64_t n,
                                 const char* type, double* out) {
  return guard([&] {
    need(mean && cov && target && out && type, "null argument");
    need(n >= 1, "bad matrix shape");
    if (std::string(type) != "regression")
      fail(2, "negative_log_likelihood: only type = \"regression\" is deterministic; \"binary\" / \"multinomial\" draw "
              "rnorm samples from R's RNG and stay in R");
    double acc = 0.0;  // ((target - mean)^2 / cov + log(cov + 1e-9)).mean(), sequential
    for (int64_t i = 0; i < n; ++i) {
      const double df = target[i] - mean[i];
      acc += df * df / cov[i] + std::log(cov[i] + 1e-9);
    }
    *out = (acc / (double)n + std::log(2 * 3.1415926)) / 2;  // the reference's literal constant (src/Utils.cpp:306)
  });
}

int flgp_test_regression(const double* Cm, const double* Y, const double* Cnv, int m, int64_t m_new, double* Y_pred) {
  return guard([&] {
    need(Cm && Y && Y_pred && (Cnv || m_new == 0), "null argument");
    need(m >= 1 && m <= 8192 && m_new >= 0, "test_regression: need 1 <= m <= 8192 training rows");
    std::vector<double> L(Cm, Cm + (size_t)m * m), alpha(Y, Y + m);
    if (!chol_lower(L, m)) fail(2, "regression: m x m covariance is not positive definite");
    chol_solve(L, m, alpha.data(), 1);
    for (int64_t i = 0; i < m_new; ++i) {
      double acc = 0.0;
      for (int j = 0; j < m; ++j) acc += Cnv[i + m_new * (int64_t)j] * alpha[j];
      Y_pred[i] = acc;
    }
  });
}

int flgp_classification_posterior_fixed(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double t,
                                        double sigma, double tol, int max_iter, double* mean, double* cov) {
  return guard([&] {
    need(h && Y_local && mean, "null argument");
    Ctx* c = on_device(h->c);
    const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(h->n_local, m_total - h->row_offset));
    DevBuf<double> dY(std::max<int64_t>(m_local, 1)), dm(std::max<int64_t>(h->n_local, 1)),
        dc(std::max<int64_t>(h->n_local, 1));
    if (m_local > 0) dY.upload(Y_local, m_local, c->stream);
    classification_posterior_dev(h, dY.p, m_total, K, t, sigma, tol, max_iter > 0 ? max_iter : 100, dm.p,
                                 cov ? dc.p : nullptr);
    if (h->n_local > 0) {
      dm.download(mean, h->n_local, c->stream);
      if (cov) dc.download(cov, h->n_local, c->stream);
    }
    sync(c);
  });
}

int flgp_posterior_distribution_classification(flgp_ctx* ctx, const double* C11, const double* C21, const double* C22,
                                               const double* Y, int m, int64_t m_new, double tol, int max_iter,
                                               double* mean, double* cov) {
  return guard([&] {
    need(ctx && C11 && C21 && C22 && Y && mean && cov, "null argument");
    need(m >= 1 && m <= 8192 && m_new >= 0, "classification: need 1 <= m <= 8192 labelled rows");
    Ctx* c = on_device(&ctx->c);
    std::vector<double> c11(C11, C11 + (size_t)m * m), pi, beta;
    laplace_mode(c11, Y, m, tol, max_iter > 0 ? max_iter : 100, pi, &beta);
    if (m_new == 0) return;
    // C21 (m_new x m, column-major) is the row-major m x m_new matrix R = C21^T:
    //   mean = (Y - pi)^T R  (1 x m_new),   T^T = beta R  (m x m_new),   cov = C22 - colsum(T^T o R)
    std::vector<double> ypi(m);
    for (int i = 0; i < m; ++i) ypi[i] = Y[i] - pi[i];
    DevBuf<double> dR((size_t)m * m_new), dB((size_t)m * m), dT((size_t)m * m_new), dv(m), d22(m_new), dmean(m_new),
        dcov(m_new);
    dR.upload(C21, (size_t)m * m_new, c->stream);
    dB.upload(beta.data(), (size_t)m * m, c->stream);  // symmetric: row-major image = column-major image
    dv.upload(ypi.data(), m, c->stream);
    d22.upload(C22, m_new, c->stream);
    gemm_nn_run(c, dv.p, dR.p, 1, m_new, m, dmean.p);
    gemm_nn_run(c, dB.p, dR.p, m, m_new, m, dT.p);
    FLGP_LAUNCH(c, coldot_sub_kernel, ceil_div(m_new, 256), 256, 0, dT.p, dR.p, m, m_new, d22.p, dcov.p);
    dmean.download(mean, m_new, c->stream);
    dcov.download(cov, m_new, c->stream);
    sync(c);
  });
}

}  // extern "C"
