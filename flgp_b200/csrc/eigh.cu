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
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace flgp {

namespace {

constexpr int TD_THREADS = 1024;
constexpr int TD_MAXSEG = 16;  // a row of the trailing matrix is split over up to this many warps
constexpr int TD_U = 4;        // loads in flight per lane in the fused pass

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

// All-to-all flag barrier for a co-resident grid (cooperative launch): every CTA publishes its epoch with one
// release store and polls everybody else's slot.  No atomic is serialised on one L2 address (148 arrivals on one
// counter cost more than the rest of a column step), and no L1 state is relied on: every datum that crosses SMs
// inside the kernel is read with ld.global.cg.
constexpr int FLAG_PER = 8;      // slots watched per lane of the polling warp: grids of up to 256 CTAs
constexpr int FLAG_STRIDE = 32;  // one 128-byte line per CTA: polls of different flags go to different L2 slices
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void flag_barrier(unsigned* flags, unsigned epoch) {
  __syncthreads();  // this CTA's stores are all issued
  if (threadIdx.x == 0) {
    __threadfence();
    st_release(flags + (size_t)blockIdx.x * FLAG_STRIDE, epoch);
  }
  if (threadIdx.x < 32) {
    // every lane watches up to FLAG_PER slots and keeps all its polls in flight together: one round trip per try
    bool done;
    do {
      unsigned seen[FLAG_PER];
#pragma unroll
      for (int q = 0; q < FLAG_PER; ++q) {
        const unsigned b = threadIdx.x + 32 * q;
        seen[q] = (b < gridDim.x) ? ld_relaxed(flags + (size_t)b * FLAG_STRIDE) : epoch;
      }
      done = true;
#pragma unroll
      for (int q = 0; q < FLAG_PER; ++q) done &= (seen[q] >= epoch);
    } while (!__all_sync(0xffffffffu, done));
    __threadfence();  // acquire side: the polls above are ordered before everything that follows
  }
  __syncthreads();
}

// ---- 1. tridiagonalisation -----------------------------------------------------------------------
// A: s x s symmetric, full storage (row i contiguous).  Vh: s x s, row k receives the Householder
// vector of column k (v[0] = 1 at index 0, length s-k-1).  d (s), e (s-1), tau (s-1).
//
// One pass over the trailing matrix and ONE grid barrier per column: the rank-2 update of column k-1
// (v_{k-1}, w_{k-1}, kept in shared memory by every CTA) is applied while the product A v_k of column k is
// accumulated from the freshly updated values.  Column k itself is formed redundantly by every CTA from
// row k in memory minus the pending update, so no barrier is needed between the Householder vector and
// the pass.  pbuf is double-buffered by the parity of k.
__global__ void __launch_bounds__(TD_THREADS)
tridiag_kernel(double* __restrict__ A, int s, double* __restrict__ Vh, double* __restrict__ dd, double* __restrict__ ee,
               double* __restrict__ tau_out, double* __restrict__ pbuf, unsigned* __restrict__ flags,
               long long* __restrict__ prof) {
  long long t0 = 0, tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool timing = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define TD_TICK(i)            \
  if (timing) {               \
    long long t_ = clock64(); \
    tacc[i] += t_ - t0;       \
    t0 = t_;                  \
  }
  extern __shared__ __align__(16) double sm[];
  double* vp = sm;           // s: pending v_{k-1}; entry q <-> global index k + q
  double* wp = sm + s;       // s: pending w_{k-1}
  double* vs = sm + 2 * s;   // s: current  v_k;     entry q <-> global index k + 1 + q
  double* red = sm + 3 * s;  // 32
  const int tid = threadIdx.x, lane = tid & 31;
  const int gwarp = (blockIdx.x * TD_THREADS + tid) >> 5;
  const int nwarp = (gridDim.x * TD_THREADS) >> 5;
  bool pending = false;      // is there an update (vp, wp) not yet applied to memory?
  constexpr int PER = (4096 + TD_THREADS - 1) / TD_THREADS;  // row k is prefetched into registers (s <= 4096)
  double xpre[PER];
  const bool use_pre = s <= PER * TD_THREADS;
  if (use_pre) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int j = tid + q * TD_THREADS;
      xpre[q] = (j < s - 1) ? __ldcg(A + 1 + j) : 0.0;  // row 0, columns 1..
    }
  }

  for (int k = 0; k < s - 1; ++k) {
    const int m = s - k - 1;  // length of column k below the diagonal
    const double* arow = A + (size_t)k * s;
    if (timing) t0 = clock64();
    // --- column k of the up-to-date matrix: memory row k minus the pending rank-2 update (redundant per CTA)
    const double v0 = pending ? vp[0] : 0.0, w0 = pending ? wp[0] : 0.0;
    double part = 0.0;
    if (use_pre) {
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int j = tid + q * TD_THREADS;
        if (j < m) {
          double x = xpre[q];
          if (pending) x = __dsub_rn(x, __dadd_rn(__dmul_rn(v0, wp[j + 1]), __dmul_rn(w0, vp[j + 1])));
          vs[j] = x;
          if (j > 0) part = fma(x, x, part);
        }
      }
    } else {
      for (int j = tid; j < m; j += TD_THREADS) {
        double x = __ldcg(arow + k + 1 + j);
        if (pending) x = __dsub_rn(x, __dadd_rn(__dmul_rn(v0, wp[j + 1]), __dmul_rn(w0, vp[j + 1])));
        vs[j] = x;
        if (j > 0) part = fma(x, x, part);
      }
    }
    TD_TICK(0);
    const double xnorm2 = block_sum(part, red);  // also orders the vs[] writes
    TD_TICK(1);
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
    TD_TICK(2);
    for (int j = tid; j < m; j += TD_THREADS) {
      const double v = (j == 0) ? 1.0 : vs[j] * scale;
      vs[j] = v;
      if ((j >> 5) % gridDim.x == blockIdx.x) Vh[(size_t)k * s + j] = v;  // every CTA stores its share of v_k
    }
    __syncthreads();
    TD_TICK(3);
    if (blockIdx.x == (unsigned)(k % gridDim.x) && tid == 0) {
      double dk = __ldcg(arow + k);
      if (pending) dk = __dsub_rn(dk, __dadd_rn(__dmul_rn(v0, w0), __dmul_rn(w0, v0)));
      dd[k] = dk;
      ee[k] = beta;
      tau_out[k] = tau;
    }
    TD_TICK(4);
    // --- fused pass: apply the pending update to rows k+1.., accumulate the products (A22 v_k).  Work item =
    // (row, segment): a row is split over nseg warps when there are more warps than rows, so that the longest
    // dependent chain of the step stays short; loads are issued TD_U at a time before any store.
    int nseg = 1;
    while (nseg < TD_MAXSEG && (nseg * 2) * m <= nwarp) nseg *= 2;
    const int seglen = ((m + nseg - 1) / nseg + 31) & ~31;
    double* pk = pbuf + (size_t)(k & 1) * TD_MAXSEG * s;
    for (int item = gwarp; item < m * nseg; item += nwarp) {
      const int row = item / nseg, seg = item - row * nseg;
      const int c0 = seg * seglen, c1 = min(m, c0 + seglen);
      double* ar = A + (size_t)(k + 1 + row) * s + (k + 1);
      const double vi = pending ? vp[row + 1] : 0.0, wi = pending ? wp[row + 1] : 0.0;
      double acc = 0.0;
      for (int j0 = c0 + lane; j0 < c1; j0 += 32 * TD_U) {
        double a[TD_U];
#pragma unroll
        for (int u = 0; u < TD_U; ++u) {
          const int j = j0 + 32 * u;
          a[u] = (j < c1) ? __ldcg(ar + j) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < TD_U; ++u) {
          const int j = j0 + 32 * u;
          if (j < c1) {
            double av = a[u];
            if (pending) {
              av = __dsub_rn(av, __dadd_rn(__dmul_rn(vi, wp[j + 1]), __dmul_rn(wi, vp[j + 1])));
              __stcg(ar + j, av);
            }
            acc = fma(av, vs[j], acc);
          }
        }
      }
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) __stcg(pk + (size_t)seg * s + row, acc);
    }
    TD_TICK(5);
    flag_barrier(flags, (unsigned)(k + 1));
    TD_TICK(6);
    // --- next column's row from memory (row k+1 now carries every update but (v_k, w_k)): issued together with
    // the loads of the products so that the two L2 round trips overlap
    if (use_pre) {
      const double* anext = A + (size_t)(k + 1) * s + (k + 2);
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int j = tid + q * TD_THREADS;
        xpre[q] = (j < m - 1) ? __ldcg(anext + j) : 0.0;
      }
    }
    // --- w_k = p - (tau/2)(p.v) v with p = tau A22 v; (v_k, w_k) become the pending update
    part = 0.0;
    for (int j = tid; j < m; j += TD_THREADS) {
      double pg[TD_MAXSEG];
#pragma unroll
      for (int g = 0; g < TD_MAXSEG; ++g) pg[g] = (g < nseg) ? __ldcg(pk + (size_t)g * s + j) : 0.0;  // all in flight
      double p = 0.0;
#pragma unroll
      for (int g = 0; g < TD_MAXSEG; ++g) p += pg[g];  // fixed order: deterministic
      p *= tau;
      wp[j] = p;
      part = fma(p, vs[j], part);
    }
    TD_TICK(7);
    const double pv = block_sum(part, red);
    TD_TICK(8);
    const double a2 = -0.5 * tau * pv;
    for (int j = tid; j < m; j += TD_THREADS) {
      wp[j] = fma(a2, vs[j], wp[j]);
      vp[j] = vs[j];
    }
    pending = true;
    __syncthreads();
    TD_TICK(9);
  }
  if (timing)
    for (int i = 0; i < 12; ++i) prof[i] = tacc[i];
#undef TD_TICK
  // the last pending update has tau == 0 (a 1 x 1 column): memory already holds the final corner
  if (blockIdx.x == 0 && tid == 0) dd[s - 1] = __ldcg(A + (size_t)(s - 1) * s + (s - 1));
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
  // Both sweeps are sequential recurrences; their operands are fetched eight steps ahead so that the chain
  // waits on arithmetic, not on L2 round trips.
  constexpr int PF = 8;
  // forward: L with interchanges
  double bi = X[AT(0)];
  double nrm = 0.0;
  for (int i0 = 0; i0 < s - 1; i0 += PF) {
    const int cnt = min(PF, s - 1 - i0);
    double l[PF], xn[PF];
    unsigned char pv[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < cnt) {
        l[u] = dl[AT(i0 + u)];
        pv[u] = piv[AT(i0 + u)];
        xn[u] = X[AT(i0 + u + 1)];
      }
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < cnt) {
        if (pv[u]) {
          X[AT(i0 + u)] = xn[u];
          bi = bi - l[u] * xn[u];
        } else {
          X[AT(i0 + u)] = bi;
          bi = xn[u] - l[u] * bi;
        }
      }
  }
  X[AT(s - 1)] = bi;
  // backward: U with two super-diagonals
  double x1 = X[AT(s - 1)] / dg[AT(s - 1)];
  X[AT(s - 1)] = x1;
  nrm = fma(x1, x1, nrm);
  double x2 = 0.0;
  for (int i0 = s - 2; i0 >= 0; i0 -= PF) {
    const int cnt = min(PF, i0 + 1);
    double g[PF], u1[PF], u2[PF], b[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < cnt) {
        g[u] = dg[AT(i0 - u)];
        u1[u] = du[AT(i0 - u)];
        u2[u] = du2[AT(i0 - u)];
        b[u] = X[AT(i0 - u)];
      }
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < cnt) {
        double xi = (b[u] - u1[u] * x1 - u2[u] * x2) / g[u];
        X[AT(i0 - u)] = xi;
        nrm = fma(xi, xi, nrm);
        x2 = x1;
        x1 = xi;
      }
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
  DevBuf<double> dd(s), ee(s), tau(s), pbuf((size_t)2 * TD_MAXSEG * s), tnorm(1);
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
    size_t smem = (size_t)(3 * s + 32) * sizeof(double);
    FLGP_CUDA(cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FLGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tridiag_kernel, TD_THREADS, smem));
    if (per_sm < 1) fail(2, "eigh: s=%d needs more shared memory than one SM has", s);
    int grid = std::min(c->sm_count, 32 * FLAG_PER);  // one CTA per SM
    const bool prof_on = std::getenv("FLGP_EIGH_PROF") != nullptr;
    DevBuf<long long> prof(12);
    long long* profp = prof_on ? prof.p : nullptr;
    DevBuf<unsigned> flags((size_t)grid * FLAG_STRIDE);
    flags.zero(c->stream);
    void* args[] = {&G, &s, &Vh.p, &dd.p, &ee.p, &tau.p, &pbuf.p, &flags.p, &profp};
    FLGP_CUDA(cudaLaunchCooperativeKernel((void*)tridiag_kernel, dim3(grid), dim3(TD_THREADS), args, smem, c->stream));
    c->launches++;
    if (prof_on) {
      long long h[12];
      prof.download(h, 12, c->stream);
      sync(c);
      fprintf(stderr, "[flgp eigh prof] tridiag s=%d, cycles per column (CTA 0):", s);
      for (int i = 0; i < 10; ++i) fprintf(stderr, " %d:%.0f", i, (double)h[i] / (s - 1));
      fprintf(stderr, "\n");
    }
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
