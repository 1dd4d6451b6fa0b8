// eigh.cu — on-device top-K symmetric eigensolver for the s x s Gram A^T A
// (replaces the RSpectra::svds / Eigen::BDCSVD call of truncated_SVD_cpp,
//  /root/reference/src/TruncatedSVD.cpp:9-34; sigma^2 = eigenvalues of the Gram, right vectors lifted
//  later by sparse.cu).  No host LAPACK, no cuSOLVER: everything stays in HBM/L2.
//
//   1. Householder tridiagonalisation  Q^T G Q = T      one persistent cooperative kernel; the
//      trailing matrix (<= 32 MB at s = 2000) stays L2-resident; 2 grid barriers per column.
//   2. top-K eigenvalues of T by warp-wide multisection of the Sturm count (33-way per pass).
//   3. eigenvectors of T by inverse iteration (pivoted tridiagonal LU per eigenvalue, one thread
//      each, interleaved storage) with modified Gram-Schmidt inside clusters of close eigenvalues.
//   4. back-transformation Y = Q X, one CTA per eigenvector, vector resident in shared memory.
//
// Work: (4/3) s^3 flop for step 1 (BLAS-2, L2-bandwidth bound: 24 bytes per trailing element per
// column => 8 s^3 bytes), 2 s^2 K flop for step 4.
#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace flgp {

namespace {

constexpr int TD_THREADS = 512;

// deterministic block-wide sum; every thread receives the result.  red: >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // protect red from the previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : 0.0;
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// ---- 1. tridiagonalisation -----------------------------------------------------------------------
// A: s x s symmetric, full storage (row i contiguous).  Vh: s x s, row k receives the Householder
// vector of column k (v[0] = 1 at index 0, length s-k-1).  d (s), e (s-1), tau (s-1).
__global__ void __launch_bounds__(TD_THREADS)
tridiag_kernel(double* __restrict__ A, int s, double* __restrict__ Vh, double* __restrict__ dd, double* __restrict__ ee,
               double* __restrict__ tau_out, double* __restrict__ pbuf) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double sm[];
  double* vs = sm;           // s
  double* ws = sm + s;       // s
  double* red = sm + 2 * s;  // 32
  const int tid = threadIdx.x, lane = tid & 31;
  const int gwarp = (blockIdx.x * TD_THREADS + tid) >> 5;
  const int nwarp = (gridDim.x * TD_THREADS) >> 5;

  for (int k = 0; k < s - 1; ++k) {
    const int m = s - k - 1;                      // length of the column below the diagonal
    const double* xrow = A + (size_t)k * s + (k + 1);
    // --- phase A (redundant in every CTA): Householder vector of column k
    double part = 0.0;
    for (int j = tid; j < m; j += TD_THREADS) {
      double x = xrow[j];
      vs[j] = x;
      if (j > 0) part = fma(x, x, part);
    }
    const double xnorm2 = block_sum(part, red);  // also orders the vs[] writes
    const double alpha = vs[0];
    double beta, tau, scale;
    if (xnorm2 == 0.0) {
      beta = alpha;
      tau = 0.0;
      scale = 0.0;
    } else {
      double nrm = sqrt(fma(alpha, alpha, xnorm2));
      beta = (alpha >= 0.0) ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    __syncthreads();
    for (int j = tid; j < m; j += TD_THREADS) vs[j] = (j == 0) ? 1.0 : vs[j] * scale;
    __syncthreads();
    if (blockIdx.x == 0) {
      if (tid == 0) {
        dd[k] = A[(size_t)k * s + k];
        ee[k] = beta;
        tau_out[k] = tau;
        if (k == s - 2) dd[s - 1] = A[(size_t)(s - 1) * s + (s - 1)];
      }
      for (int j = tid; j < m; j += TD_THREADS) Vh[(size_t)k * s + j] = vs[j];
    }
    if (tau == 0.0) continue;  // uniform across the grid: H = I, nothing to update
    // --- phase B: p = tau * A22 v, one warp per row
    for (int row = gwarp; row < m; row += nwarp) {
      const double* ar = A + (size_t)(k + 1 + row) * s + (k + 1);
      double acc = 0.0;
      for (int j = lane; j < m; j += 32) acc = fma(ar[j], vs[j], acc);
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) pbuf[row] = tau * acc;
    }
    grid.sync();
    // --- phase C (redundant): w = p - (tau/2)(p.v) v
    part = 0.0;
    for (int j = tid; j < m; j += TD_THREADS) {
      double p = pbuf[j];
      ws[j] = p;
      part = fma(p, vs[j], part);
    }
    const double pv = block_sum(part, red);
    const double a2 = -0.5 * tau * pv;
    for (int j = tid; j < m; j += TD_THREADS) ws[j] = fma(a2, vs[j], ws[j]);
    __syncthreads();
    // --- phase D: A22 -= v w^T + w v^T (symmetric rounding: separate products, one add)
    for (int row = gwarp; row < m; row += nwarp) {
      double* ar = A + (size_t)(k + 1 + row) * s + (k + 1);
      const double vi = vs[row], wi = ws[row];
      for (int j = lane; j < m; j += 32) {
        double t = __dadd_rn(__dmul_rn(vi, ws[j]), __dmul_rn(wi, vs[j]));
        ar[j] = __dsub_rn(ar[j], t);
      }
    }
    grid.sync();
  }
}

// ---- 2. eigenvalues: multisection on the Sturm count ------------------------------------------------
// one warp per wanted eigenvalue; lam_out[kk], kk = 0..K-1 descending  <=>  ascending index s-1-kk.
__global__ void __launch_bounds__(256)
bisect_kernel(const double* __restrict__ dd, const double* __restrict__ ee, int s, int K, double* __restrict__ lam_out,
              double* __restrict__ tnorm_out) {
  extern __shared__ __align__(16) double sm[];
  double* d = sm;       // s
  double* e2 = sm + s;  // s
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31;
  double gl = DBL_MAX, gu = -DBL_MAX, emax = 0.0;
  for (int i = tid; i < s; i += 256) {
    double di = dd[i];
    double el = (i > 0) ? fabs(ee[i - 1]) : 0.0, er = (i < s - 1) ? fabs(ee[i]) : 0.0;
    d[i] = di;
    e2[i] = (i < s - 1) ? ee[i] * ee[i] : 0.0;
    gl = fmin(gl, di - el - er);
    gu = fmax(gu, di + el + er);
    emax = fmax(emax, er * er);
  }
  // block min / max
  for (int o = 16; o; o >>= 1) {
    gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
    gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
    emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  }
  __shared__ double rgl[8], rgu[8], rem[8];
  if (lane == 0) {
    rgl[tid >> 5] = gl;
    rgu[tid >> 5] = gu;
    rem[tid >> 5] = emax;
  }
  __syncthreads();
  gl = rgl[0];
  gu = rgu[0];
  emax = rem[0];
  for (int w = 1; w < 8; ++w) {
    gl = fmin(gl, rgl[w]);
    gu = fmax(gu, rgu[w]);
    emax = fmax(emax, rem[w]);
  }
  (void)red;
  const double ulp = DBL_EPSILON, safemin = DBL_MIN;
  const double pivmin = safemin * fmax(1.0, emax);
  const double tnorm = fmax(fabs(gl), fabs(gu));
  gl = gl - 2.1 * tnorm * ulp * s - 2.1 * pivmin;
  gu = gu + 2.1 * tnorm * ulp * s + 2.1 * pivmin;
  if (blockIdx.x == 0 && tid == 0 && tnorm_out) *tnorm_out = tnorm;

  const int kk = blockIdx.x * 8 + (tid >> 5);
  if (kk >= K) return;
  const int t = s - 1 - kk;  // ascending index of the wanted eigenvalue
  double lo = gl, hi = gu;   // count(lo) <= t < count(hi)
  for (int pass = 0; pass < 40; ++pass) {
    const double width = hi - lo;
    const double tol = fmax(2.0 * ulp * fmax(fabs(lo), fabs(hi)), pivmin);
    if (width <= tol) break;
    const double x = lo + width * ((double)(lane + 1) / 33.0);
    const int cnt = sturm_count(d, e2, s, x, pivmin);
    const unsigned ball = __ballot_sync(0xffffffffu, cnt >= t + 1);
    const int f = ball ? (__ffs(ball) - 1) : 32;  // first lane whose point is above lambda_t
    const double xhi = __shfl_sync(0xffffffffu, x, f & 31);
    const double xlo = __shfl_sync(0xffffffffu, x, (f - 1) & 31);
    if (f < 32) hi = xhi;
    if (f > 0) lo = xlo;
  }
  if (lane == 0) lam_out[kk] = 0.5 * (lo + hi);
}

// ---- 3. inverse iteration ------------------------------------------------------------------------
// Interleaved work arrays: element i of eigenvector kk at [i*K + kk] (coalesced across threads).
__device__ __forceinline__ double hash_uniform(unsigned i, unsigned kk) {
  unsigned long long z = ((unsigned long long)i << 32 | kk) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;  // (-1, 1)
}

// factor T - lam*I = P L U  (dgttrf recurrence);  dg: U diagonal, du: U first super-diagonal,
// du2: second super-diagonal, dl: multipliers, piv: 1 where rows i,i+1 were interchanged.
__global__ void __launch_bounds__(64)
invit_factor_kernel(const double* __restrict__ dd, const double* __restrict__ ee, int s, int K,
                    const double* __restrict__ lam, double tiny, double* __restrict__ dg, double* __restrict__ du,
                    double* __restrict__ du2, double* __restrict__ dl, unsigned char* __restrict__ piv, double* __restrict__ X) {
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= K) return;
  const double l = lam[kk];
#define AT(i) ((size_t)(i) * K + kk)
  double dcur = dd[0] - l;            // current diagonal entry of row i
  double ucur = (s > 1) ? ee[0] : 0;  // current super-diagonal entry of row i
  for (int i = 0; i < s - 1; ++i) {
    const double sub = ee[i];                       // sub-diagonal entry of row i+1
    double dnext = dd[i + 1] - l;                   // diagonal of row i+1
    double unext = (i + 1 < s - 1) ? ee[i + 1] : 0; // super-diagonal of row i+1
    if (fabs(dcur) >= fabs(sub)) {
      if (fabs(dcur) < tiny) dcur = (dcur < 0.0) ? -tiny : tiny;
      const double fact = sub / dcur;
      dg[AT(i)] = dcur;
      du[AT(i)] = ucur;
      du2[AT(i)] = 0.0;
      dl[AT(i)] = fact;
      piv[AT(i)] = 0;
      dcur = dnext - fact * ucur;
      ucur = unext;
    } else {
      const double fact = dcur / sub;
      dg[AT(i)] = sub;
      du[AT(i)] = dnext;
      du2[AT(i)] = unext;
      dl[AT(i)] = fact;
      piv[AT(i)] = 1;
      dcur = ucur - fact * dnext;
      ucur = -fact * unext;
    }
  }
  if (fabs(dcur) < tiny) dcur = (dcur < 0.0) ? -tiny : tiny;
  dg[AT(s - 1)] = dcur;
  for (int i = 0; i < s; ++i) X[AT(i)] = hash_uniform((unsigned)i, (unsigned)kk);
#undef AT
}

__global__ void __launch_bounds__(64)
invit_solve_kernel(int s, int K, const double* __restrict__ dg, const double* __restrict__ du,
                   const double* __restrict__ du2, const double* __restrict__ dl, const unsigned char* __restrict__ piv,
                   double* __restrict__ X) {
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= K) return;
#define AT(i) ((size_t)(i) * K + kk)
  // forward: L with interchanges
  double bi = X[AT(0)];
  double nrm = 0.0;
  for (int i = 0; i < s - 1; ++i) {
    double bn = X[AT(i + 1)];
    if (piv[AT(i)]) {
      X[AT(i)] = bn;
      bi = bi - dl[AT(i)] * bn;
    } else {
      X[AT(i)] = bi;
      bi = bn - dl[AT(i)] * bi;
    }
  }
  X[AT(s - 1)] = bi;
  // backward: U with two super-diagonals
  double x1 = X[AT(s - 1)] / dg[AT(s - 1)];
  X[AT(s - 1)] = x1;
  nrm = fma(x1, x1, nrm);
  double x2 = 0.0;
  for (int i = s - 2; i >= 0; --i) {
    double xi = (X[AT(i)] - du[AT(i)] * x1 - du2[AT(i)] * x2) / dg[AT(i)];
    X[AT(i)] = xi;
    nrm = fma(xi, xi, nrm);
    x2 = x1;
    x1 = xi;
  }
  // normalise (guards against overflow in the next solve); clusters are re-orthogonalised next
  const double inv = 1.0 / sqrt(nrm);
  for (int i = 0; i < s; ++i) X[AT(i)] *= inv;
#undef AT
}

// modified Gram-Schmidt inside each cluster of close eigenvalues; one CTA per cluster start.
// cstart[kk] = first index of the cluster that kk belongs to.
__global__ void __launch_bounds__(256)
invit_mgs_kernel(int s, int K, const int* __restrict__ cstart, double* __restrict__ X) {
  __shared__ double red[32];
  const int k0 = blockIdx.x;
  if (cstart[k0] != k0) return;
  int k1 = k0 + 1;
  while (k1 < K && cstart[k1] == k0) ++k1;
  if (k1 == k0 + 1) return;  // singleton: already normalised by the solve
  const int tid = threadIdx.x;
  for (int j = k0; j < k1; ++j) {
    for (int i = k0; i < j; ++i) {
      double part = 0.0;
      for (int q = tid; q < s; q += 256) part = fma(X[(size_t)q * K + i], X[(size_t)q * K + j], part);
      const double dot = block_sum(part, red);
      for (int q = tid; q < s; q += 256) X[(size_t)q * K + j] = fma(-dot, X[(size_t)q * K + i], X[(size_t)q * K + j]);
      __syncthreads();
    }
    double part = 0.0;
    for (int q = tid; q < s; q += 256) {
      double x = X[(size_t)q * K + j];
      part = fma(x, x, part);
    }
    const double inv = 1.0 / sqrt(block_sum(part, red));
    for (int q = tid; q < s; q += 256) X[(size_t)q * K + j] *= inv;
    __syncthreads();
  }
}

// ---- 4. back-transformation ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
backtransform_kernel(int s, int K, const double* __restrict__ Vh, const double* __restrict__ tau,
                     const double* __restrict__ X, double* __restrict__ Y) {
  extern __shared__ __align__(16) double sm[];
  double* x = sm;  // s
  __shared__ double red[32];
  const int kk = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < s; i += 256) x[i] = X[(size_t)i * K + kk];
  __syncthreads();
  for (int k = s - 2; k >= 0; --k) {
    const double tk = tau[k];
    if (tk == 0.0) continue;
    const int m = s - k - 1;
    const double* v = Vh + (size_t)k * s;
    double part = 0.0;
    for (int j = tid; j < m; j += 256) part = fma(v[j], x[k + 1 + j], part);
    const double dot = block_sum(part, red) * tk;
    for (int j = tid; j < m; j += 256) x[k + 1 + j] = fma(-dot, v[j], x[k + 1 + j]);
    __syncthreads();  // the next reflector reads elements other threads just wrote
  }
  for (int i = tid; i < s; i += 256) Y[(size_t)i + (size_t)s * kk] = x[i];
}

}  // namespace

void eigh_topk_run(Ctx* c, double* G, int s, int K, double* lam, double* Y) {
  if (K < 1 || K > s) fail(2, "eigh: need 1 <= K <= s (K=%d, s=%d)", K, s);
  DevBuf<double> dd(s), ee(s), tau(s), pbuf(s), tnorm(1);
  ee.zero(c->stream);
  tau.zero(c->stream);
  if (s == 1) {
    FLGP_CUDA(cudaMemcpyAsync(lam, G, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    double one = 1.0;
    FLGP_CUDA(cudaMemcpyAsync(Y, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    sync(c);
    return;
  }
  DevBuf<double> Vh((size_t)s * s);
  // 1. tridiagonalisation (cooperative, persistent)
  {
    size_t smem = (size_t)(2 * s + 32) * sizeof(double);
    FLGP_CUDA(cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FLGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tridiag_kernel, TD_THREADS, smem));
    if (per_sm < 1) fail(2, "eigh: s=%d needs more shared memory than one SM has", s);
    int grid = c->sm_count;  // one CTA per SM
    void* args[] = {&G, &s, &Vh.p, &dd.p, &ee.p, &tau.p, &pbuf.p};
    FLGP_CUDA(cudaLaunchCooperativeKernel((void*)tridiag_kernel, dim3(grid), dim3(TD_THREADS), args, smem, c->stream));
    c->launches++;
  }
  // 2. eigenvalues
  {
    size_t smem = (size_t)2 * s * sizeof(double);
    if (smem > 48 * 1024)
      FLGP_CUDA(cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLGP_LAUNCH(c, bisect_kernel, ceil_div(K, 8), 256, smem, dd.p, ee.p, s, K, lam, tnorm.p);
  }
  // clusters of close eigenvalues (host decides; K doubles)
  std::vector<double> lam_h(K);
  double tn = 0.0;
  FLGP_CUDA(cudaMemcpyAsync(lam_h.data(), lam, sizeof(double) * K, cudaMemcpyDeviceToHost, c->stream));
  FLGP_CUDA(cudaMemcpyAsync(&tn, tnorm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  sync(c);
  const double ortol = 1e-6 * tn;
  std::vector<int> cstart(K);
  cstart[0] = 0;
  for (int k = 1; k < K; ++k) cstart[k] = (lam_h[k - 1] - lam_h[k] <= ortol) ? cstart[k - 1] : k;
  DevBuf<int> cs(K);
  cs.upload(cstart.data(), K, c->stream);
  // 3. inverse iteration
  const size_t sk = (size_t)s * K;
  DevBuf<double> dg(sk), du(sk), du2(sk), dl(sk), X(sk);
  DevBuf<unsigned char> piv(sk);
  const double tiny = DBL_EPSILON * std::max(tn, DBL_MIN / DBL_EPSILON);
  FLGP_LAUNCH(c, invit_factor_kernel, ceil_div(K, 64), 64, 0, dd.p, ee.p, s, K, lam, tiny, dg.p, du.p, du2.p, dl.p,
              piv.p, X.p);
  for (int it = 0; it < 3; ++it) {
    FLGP_LAUNCH(c, invit_solve_kernel, ceil_div(K, 64), 64, 0, s, K, dg.p, du.p, du2.p, dl.p, piv.p, X.p);
    FLGP_LAUNCH(c, invit_mgs_kernel, K, 256, 0, s, K, cs.p, X.p);
  }
  // 4. back-transformation
  {
    size_t smem = (size_t)s * sizeof(double);
    if (smem > 48 * 1024)
      FLGP_CUDA(cudaFuncSetAttribute(backtransform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLGP_LAUNCH(c, backtransform_kernel, K, 256, smem, s, K, Vh.p, tau.p, X.p, Y);
  }
  sync(c);
}

}  // namespace flgp
