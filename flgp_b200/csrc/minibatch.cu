// minibatch.cu — subsample_cpp(method = "minibatchkmeans"), /root/reference/src/Utils.cpp:49-62.
//
// The reference gets its centroids from ClusterR::MiniBatchKmeans(data, clusters = s, batch_size = 10 s,
// init_fraction = 20 s / n, num_init = nstart) — an un-vendored, unpinned R package whose start is a kmeans++ draw
// on R's RNG: parity unpinned.  What runs here is the contract the oracle states (oracle/flgp_oracle.cpp,
// DESIGN.md §2): Sculley's mini-batch k-means with ClusterR's defaults (max_iters = 100, early_stop_iter = 10,
// tol = 1e-4), explicit start rows, batches of b = min(10 s, n) DISTINCT rows drawn by a keyed bijection of [0, n)
// (core_math.cuh: mb_perm), nearest centre under the Lloyd score rule, per-sample updates in batch order with
// eta = 1 / count.  The sizes column is the reference's own code (:57-62): 1-NN labels by KNN_cpp, counted.
//
// Per batch: mb_gather (b rows -> column-major b x d block), mb_assign (thread per batch row, centres staged in
// shared memory in groups of 8 held as 8 running scores per thread), mb_update (one warp per centre walks the batch
// in order; the lanes own the coordinates, which are independent), mb_delta (the stopping criterion in the
// contract's summation order).  A batch is 10 s rows — small against the n-sized stages, so these kernels are plain.
#include "kernels.cuh"

namespace flgp {

namespace {

constexpr int MB_THREADS = 128;
constexpr int MB_GROUP = 8;    // centres scored together per thread (running scores in registers)
constexpr int MB_QCHUNK = 16;  // coordinates of a batch row held in registers at a time

struct BatchRow {  // coordinate q of batch row k in the column-major b x d block
  const double* Xb;
  int64_t k, b;
  __device__ double operator()(int q) const { return Xb[k + b * q]; }
};

__global__ void __launch_bounds__(256)
mb_gather_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, int64_t b, uint64_t key,
                 double* __restrict__ Xb) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= b) return;
  const int64_t row = mb_perm(k, n, key);
  for (int q = 0; q < d; ++q) Xb[k + b * q] = X[row + ldx * q];
}

// score_j(x) = fma chain over q of x_q * (-2 c_jq), started from |c_j|^2 (fma chain) + 2 d maxabs^2; arg-min with the
// lowest index on ties — the Lloyd score rule of kmeans.cu / the oracle's orc_kmeans_step.
// Shared memory: a tile of ng groups: rec[ng * MB_GROUP][dp] (-2 c, zero padded to dp = multiple of MB_QCHUNK) and
// cn[ng * MB_GROUP]; one barrier pair per tile.
__global__ void __launch_bounds__(MB_THREADS)
mb_assign_kernel(const double* __restrict__ Xb, int64_t b, int d, int dp, int ng, const double* __restrict__ C, int s,
                 int64_t ldc, double m2, int32_t* __restrict__ assign) {
  extern __shared__ double sm[];
  const int tile = ng * MB_GROUP;     // centres per tile
  double* rec = sm;                   // tile x dp
  double* cn = sm + (size_t)tile * dp;  // tile
  const int64_t k = blockIdx.x * (int64_t)MB_THREADS + threadIdx.x;
  const bool live = k < b;
  const BatchRow xk{Xb, k, b};
  double best = 0.0;
  int bj = 0;
  for (int j0 = 0; j0 < s; j0 += tile) {
    __syncthreads();
    for (int t = threadIdx.x; t < tile * dp; t += MB_THREADS) {
      const int jj = t / dp, q = t - jj * dp;
      const int j = j0 + jj;
      rec[t] = (j < s && q < d) ? -2.0 * C[j + ldc * q] : 0.0;
    }
    for (int jj = threadIdx.x; jj < tile; jj += MB_THREADS) {
      const int j = j0 + jj;
      cn[jj] = (j < s) ? mb_centre_norm(C, ldc, j, d, m2) : 0.0;
    }
    __syncthreads();
    if (!live) continue;
    for (int g = 0; g < ng; ++g) {
      if (j0 + g * MB_GROUP >= s) break;
      double e[MB_GROUP];
      mb_score_group<MB_GROUP, MB_QCHUNK>(xk, d, dp, rec + (size_t)g * MB_GROUP * dp, cn + g * MB_GROUP, e);
      mb_argmin_group<MB_GROUP>(e, j0 + g * MB_GROUP, s, &best, &bj);
    }
  }
  if (live) assign[k] = bj;
}

// One warp per centre: walk the batch in order, 32 entries per step; for every member cnt += 1, eta = 1 / cnt and the
// lanes update their coordinates (q = lane, lane + 32, ...).  dsq[j] = sum_q (c_jq - old_jq)^2, sequential in q.
__global__ void __launch_bounds__(256)
mb_update_kernel(const double* __restrict__ Xb, int64_t b, int d, const int32_t* __restrict__ assign,
                 double* __restrict__ C, int s, int64_t ldc, const double* __restrict__ Cold,
                 long long* __restrict__ cnt, double* __restrict__ dsq) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= s) return;  // whole warps leave together
  long long n_j = cnt[j];
  bool touched = false;
  for (int64_t k0 = 0; k0 < b; k0 += 128) {  // four independent loads in flight; members still taken in batch order
    unsigned bal[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t k = k0 + 32 * u + lane;
      const int a = (k < b) ? assign[k] : -1;
      bal[u] = __ballot_sync(0xffffffffu, a == j);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      while (bal[u]) {
        const int src = __ffs(bal[u]) - 1;
        bal[u] &= bal[u] - 1;
        const int64_t km = k0 + 32 * u + src;
        n_j += 1;
        const double eta = 1.0 / (double)n_j;
        for (int q = lane; q < d; q += 32) C[j + ldc * q] = mb_update_coord(C[j + ldc * q], Xb[km + b * q], eta);
        touched = true;
      }
  }
  __syncwarp();
  if (lane == 0) {
    cnt[j] = n_j;
    double dj = 0.0;
    if (touched)
      for (int q = 0; q < d; ++q) {
        const double df = C[j + ldc * q] - Cold[j + (int64_t)s * q];
        dj = dj + df * df;
      }
    dsq[j] = dj;
  }
}

__global__ void mb_delta_kernel(const double* __restrict__ dsq, int s, double* __restrict__ delta) {
  double a = 0.0;
  for (int j = 0; j < s; ++j) a = a + dsq[j];
  *delta = a;
}

__global__ void __launch_bounds__(256)
mb_copy_centres_kernel(const double* __restrict__ C, int s, int64_t ldc, int d, double* __restrict__ Cold) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= s * d) return;
  const int q = t / s, j = t - q * s;
  Cold[t] = C[j + ldc * q];
}

__global__ void __launch_bounds__(256)
mb_start_kernel(const double* __restrict__ X, int64_t ldx, int d, const int32_t* __restrict__ init, int s,
                double* __restrict__ C, int64_t ldc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= s * d) return;
  const int q = t / s, j = t - q * s;
  C[j + ldc * q] = X[init[j] + ldx * q];
}

__global__ void __launch_bounds__(256)
mb_label_hist_kernel(const int32_t* __restrict__ label, int64_t n, int s, unsigned long long* __restrict__ hist) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = label[i];
  if (j >= 0 && j < s) atomicAdd(&hist[j], 1ull);
}

__global__ void mb_sizes_kernel(const unsigned long long* __restrict__ hist, int s, double* __restrict__ sizes) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < s) sizes[j] = (double)hist[j];
}

}  // namespace

// U: s x (d+1) column-major: centroids, then the number of rows whose nearest centroid (KNN_cpp, r = 1) it is.
void minibatch_kmeans_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, int s, const int32_t* init_idx_h,
                          int max_iters, uint64_t seed, double* U, int* iters_out) {
  if (s < 1 || s > n || d < 1) fail(2, "minibatchkmeans: need 1 <= s <= n");
  if (c->nranks != 1) fail(2, "subsample=\"minibatchkmeans\" is single-GPU only");
  const int dp = (d + MB_QCHUNK - 1) / MB_QCHUNK * MB_QCHUNK;
  const int ng = std::max(1, std::min(16, 4096 / (MB_GROUP * dp)));  // groups of centres staged per barrier pair
  const size_t smem = ((size_t)ng * MB_GROUP * dp + ng * MB_GROUP) * sizeof(double);
  if (smem > 200 * 1024) fail(2, "minibatchkmeans: d=%d exceeds the supported maximum", d);
  if (smem > 48 * 1024)
    FLGP_CUDA(cudaFuncSetAttribute(mb_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int j = 0; j < s; ++j)
    if (init_idx_h[j] < 0 || init_idx_h[j] >= n) fail(2, "initial index out of range");
  const int64_t b = std::min<int64_t>((int64_t)10 * s, n);  // batch_size = s * 10 (src/Utils.cpp:52)
  const int64_t ldc = s;                                    // the centroids live in U's first d columns
  const double maxabs = maxabs_run(c, X, n, ldx, d);
  const double m2 = (2.0 * d) * (maxabs * maxabs);
  DevBuf<int32_t> init(s), assign(b);
  DevBuf<double> Xb((size_t)b * d), Cold((size_t)s * d), dsq(s), delta(1);
  DevBuf<long long> cnt(s);
  init.upload(init_idx_h, s, c->stream);
  cnt.zero(c->stream);
  FLGP_LAUNCH(c, mb_start_kernel, ceil_div((int64_t)s * d, 256), 256, 0, X, ldx, d, init.p, s, U, ldc);
  const int early_stop_iter = 10;  // ClusterR defaults
  const double tol = 1e-4;
  int it = 0, calm = 0;
  while (it < max_iters) {
    const uint64_t key = mb_batch_key(seed, it);
    FLGP_LAUNCH(c, mb_gather_kernel, ceil_div(b, 256), 256, 0, X, n, ldx, d, b, key, Xb.p);
    FLGP_LAUNCH(c, mb_assign_kernel, ceil_div(b, MB_THREADS), MB_THREADS, smem, Xb.p, b, d, dp, ng, U, s, ldc, m2,
                assign.p);
    FLGP_LAUNCH(c, mb_copy_centres_kernel, ceil_div((int64_t)s * d, 256), 256, 0, U, s, ldc, d, Cold.p);
    FLGP_LAUNCH(c, mb_update_kernel, ceil_div((int64_t)s * 32, 256), 256, 0, Xb.p, b, d, assign.p, U, s, ldc, Cold.p,
                cnt.p, dsq.p);
    FLGP_LAUNCH(c, mb_delta_kernel, 1, 1, 0, dsq.p, s, delta.p);
    double dl = 0.0;
    delta.download(&dl, 1, c->stream);
    sync(c);
    ++it;
    calm = (dl < tol) ? calm + 1 : 0;
    if (calm >= early_stop_iter) break;
  }
  if (iters_out) *iters_out = it;
  // sizes (src/Utils.cpp:57-62): labels = KNN_cpp(X, U.leftCols(d), 1)["ind_knn"]; U(i, d) = (labels == i).count()
  DevBuf<int32_t> label(n);
  DevBuf<unsigned long long> hist(s);
  hist.zero(c->stream);
  knn_run(c, X, n, ldx, d, U, s, ldc, 1, label.p, nullptr);
  FLGP_LAUNCH(c, mb_label_hist_kernel, ceil_div(n, 256), 256, 0, label.p, n, s, hist.p);
  FLGP_LAUNCH(c, mb_sizes_kernel, ceil_div(s, 256), 256, 0, hist.p, s, U + (size_t)s * d);
  sync(c);
}

}  // namespace flgp
