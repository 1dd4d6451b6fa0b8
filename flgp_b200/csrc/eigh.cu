// eigh.cu — on-device top-K symmetric eigensolver for the s x s Gram A^T A: dispatcher + the direct route
// (replaces the RSpectra::svds / Eigen::BDCSVD call of truncated_SVD_cpp,
//  /root/reference/src/TruncatedSVD.cpp:9-34; sigma^2 = eigenvalues of the Gram, right vectors lifted
//  later by sparse.cu).  No host LAPACK, no cuSOLVER: everything stays in HBM/L2.
//
// eigh_topk_run: K <= s/5 and s >= 1024 -> Chebyshev-filtered subspace iteration on the FP64 tensor cores (chfsi.cu);
// anything else, and whatever that route declines or does not converge on -> the direct route below:
//   1. Householder tridiagonalisation  Q^T G Q = T, ONE barrier and one pass over the trailing matrix per column
//      (the rank-2 update of column k-1 is applied while A v_k is accumulated):
//        s <= 832   one thread-block cluster (8 or 16 CTAs), matrix resident in distributed shared memory,
//                   hardware cluster barrier (tridiag_cluster_kernel);
//        larger     cooperative grid, all-to-all flag barrier: streaming through L2 until the trailing matrix fits
//                   the grid's shared memory, resident after that (tridiag_kernel<false|true>).
//   2. top-K eigenvalues of T: one CTA per eigenvalue, 129-way multisection of the Sturm count.
//   3. eigenvectors of T by inverse iteration (pivoted tridiagonal LU per eigenvalue, one thread each, interleaved
//      storage) with modified Gram-Schmidt inside clusters of close eigenvalues, CGS2 for clusters > 12.
//   4. back-transformation Y = Q X with compact-WY blocks of 32 reflectors.
//
// Work: (4/3) s^3 flop for step 1 (BLAS-2; latency bound: s-1 dependent steps), 2 s^2 K flop for step 4.
#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace flgp {

namespace {

constexpr int TD_THREADS_STREAM = 1024;  // streaming pass: as many warps in flight as possible
constexpr int TD_THREADS_RES = 512;      // resident pass: fewer, fatter threads (cheaper reductions and barriers)
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
  if (threadIdx.x == 0) st_release(flags + (size_t)blockIdx.x * FLAG_STRIDE, epoch);  // cumulative over the bar.sync
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
// row k (published by the pass of column k-1, prefetched across the barrier) minus the pending update, so no
// barrier is needed between the Householder vector and the pass.  Buffers are double-buffered by parity of k.
//
// Two pass modes, columns [k_begin, k_end) per launch:
//   RESIDENT = false  the trailing matrix streams through L2 (read-modify-write, 16 bytes per element and column:
//                     L2-bandwidth bound); a row is split over several warps.  On exit the last pending update is
//                     applied to memory, so that the next launch starts from an up-to-date matrix.
//   RESIDENT = true   once the trailing matrix fits in the shared memory of the grid (rows dealt round-robin to the
//                     CTAs, <= TD_RPMAX rows each), every CTA keeps its rows on chip for the rest of the
//                     factorisation: per column it touches only its own shared memory, publishes |rows| products
//                     and one row; L2 only carries the vectors.
constexpr int TD_RPMAX = 12;  // rows per CTA in resident mode (shared-memory capacity)
constexpr int TD_RACC = 16;   // accumulators per lane, padded to a power of two for the butterfly reduction

template <bool RESIDENT, int TD_THREADS>
__global__ void __launch_bounds__(TD_THREADS)
tridiag_kernel(double* __restrict__ A, int s, int k_begin, int k_end, double* __restrict__ Vh, double* __restrict__ dd,
               double* __restrict__ ee, double* __restrict__ tau_out, double* __restrict__ pbuf,
               double* __restrict__ rowbuf, unsigned* __restrict__ flags, long long* __restrict__ prof) {
  long long t0 = 0, tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool timing = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define TD_TICK(i)            \
  if (timing) {               \
    long long t_ = clock64(); \
    tacc[i] += t_ - t0;       \
    t0 = t_;                  \
  }
  extern __shared__ __align__(16) double sm[];
  const int M0 = s - k_begin;  // order of the trailing matrix this launch starts from
  double* vp = sm;             // M0: pending v_{k-1}; entry q <-> global index k + q
  double* wp = sm + M0;        // M0: pending w_{k-1}
  double* vs = sm + 2 * M0;    // M0: current  v_k;     entry q <-> global index k + 1 + q
  double* red = sm + 3 * M0;   // 32
  double* scratch = red + 32;  // TD_RACC x 32 (RESIDENT)
  double* As = scratch + TD_RACC * 32;  // RP x LD (RESIDENT): own rows, column c <-> global column k_begin + c
  const int LD = (M0 + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = gridDim.x, b = blockIdx.x;
  const int gwarp = (b * TD_THREADS + tid) >> 5;
  const int nwarp = (G * TD_THREADS) >> 5;
  const int RP = (M0 + G - 1) / G;  // rows per CTA (RESIDENT)
  bool pending = false;             // is there an update (vp, wp) not yet applied to the matrix?
  constexpr int PER = 4;  // the next column's row is prefetched into registers when it fits PER per thread
  double xpre[PER];
  const bool use_pre = M0 <= PER * TD_THREADS;
  if (RESIDENT) {
    for (int l = 0; l < RP; ++l) {
      const int i = k_begin + b + l * G;
      if (i < s)
        for (int c = tid; c < M0; c += TD_THREADS) As[(size_t)l * LD + c] = __ldcg(A + (size_t)i * s + k_begin + c);
    }
  }
  if (use_pre) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int j = tid + q * TD_THREADS;
      xpre[q] = (j < M0 - 1) ? __ldcg(A + (size_t)k_begin * s + k_begin + 1 + j) : 0.0;  // row k_begin
    }
  }
  __syncthreads();

  for (int k = k_begin; k < k_end; ++k) {
    const int m = s - k - 1;  // length of column k below the diagonal
    const double* arow = A + (size_t)k * s;
    if (timing) t0 = clock64();
    // --- column k of the up-to-date matrix: its row minus the pending rank-2 update (redundant per CTA)
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
      if ((j >> 5) % G == b) Vh[(size_t)k * s + j] = v;  // every CTA stores its share of v_k
    }
    __syncthreads();
    TD_TICK(3);
    if (b == (RESIDENT ? (k - k_begin) % G : k % G) && tid == 0) {  // RESIDENT: the owner of row k
      double dk = RESIDENT ? As[(size_t)((k - k_begin) / G) * LD + (k - k_begin)] : __ldcg(arow + k);
      if (pending) dk = __dsub_rn(dk, __dadd_rn(__dmul_rn(v0, w0), __dmul_rn(w0, v0)));
      dd[k] = dk;
      ee[k] = beta;
      tau_out[k] = tau;
    }
    TD_TICK(4);
    double* pk = pbuf + (size_t)(k & 1) * TD_MAXSEG * s;
    double* rnext = rowbuf + (size_t)((k + 1) & 1) * s;  // row k+1 as the next column needs it
    int nseg = 1;
    if (RESIDENT) {
      // --- resident pass: a thread owns column strips, walks this CTA's live rows, keeps one partial product per row
      const int cbase = (k + 1) - k_begin;
      double acc[TD_RACC];
#pragma unroll
      for (int l = 0; l < TD_RACC; ++l) acc[l] = 0.0;
      for (int j = tid; j < m; j += TD_THREADS) {
        const double wpj = pending ? wp[j + 1] : 0.0, vpj = pending ? vp[j + 1] : 0.0, vsj = vs[j];
#pragma unroll
        for (int l = 0; l < TD_RPMAX; ++l) {
          const int i = k_begin + b + l * G;
          if (l < RP && i > k && i < s) {  // uniform over the CTA
            const int row = i - (k + 1);
            double* ap = As + (size_t)l * LD + cbase + j;
            double av = *ap;
            if (pending) {
              av = __dsub_rn(av, __dadd_rn(__dmul_rn(vp[row + 1], wpj), __dmul_rn(wp[row + 1], vpj)));
              *ap = av;
            }
            acc[l] = fma(av, vsj, acc[l]);
            if (row == 0 && j > 0) __stcg(rnext + (j - 1), av);  // publish row k+1 for the next column
          }
        }
      }
      // 16 partial products per lane -> one per lane pair: butterfly that halves the live values at every step
      // (16 exchanges instead of 60; the shuffle unit, not the fp64 pipe, was the limiter), then across warps
#pragma unroll
      for (int half = TD_RACC / 2, o = 16; half >= 1; half >>= 1, o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int q = 0; q < half; ++q) {
          const double send = upper ? acc[q] : acc[q + half];
          const double keep = upper ? acc[q + half] : acc[q];
          acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
      if ((lane & 1) == 0) scratch[(lane >> 1) * 32 + wid] = acc[0];  // row slot lane/2, this warp's column share
      __syncthreads();
      if (wid < RP) {
        double v = (lane < TD_THREADS / 32) ? scratch[wid * 32 + lane] : 0.0;
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int i = k_begin + b + wid * G;
        if (lane == 0 && i > k && i < s) __stcg(pk + (i - (k + 1)), v);
      }
    } else {
      // --- streaming pass: apply the pending update to rows k+1.., accumulate the products (A22 v_k).  Work item =
      // (row, segment): a row is split over nseg warps when there are more warps than rows, so that the longest
      // dependent chain of the step stays short; loads are issued TD_U at a time before any store.
      while (nseg < TD_MAXSEG && (nseg * 2) * m <= nwarp) nseg *= 2;
      const int seglen = ((m + nseg - 1) / nseg + 31) & ~31;
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
              if (row == 0 && j > 0) __stcg(rnext + (j - 1), av);  // publish row k+1 for the next column
            }
          }
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) __stcg(pk + (size_t)seg * s + row, acc);
      }
    }
    TD_TICK(5);
    flag_barrier(flags, (unsigned)(k + 1));
    TD_TICK(6);
    // --- the next column's row (it carries every update but (v_k, w_k)): issued together with the loads of the
    // products so that the two L2 round trips overlap
    if (use_pre) {
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int j = tid + q * TD_THREADS;
        xpre[q] = (j < m - 1) ? __ldcg(rnext + j) : 0.0;
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
  if (k_end >= s - 1) {
    // the last pending update has tau == 0 (a 1 x 1 column): the matrix already holds the final corner
    if (RESIDENT) {
      const int o = (s - 1) - k_begin;
      if (b == o % G && tid == 0) dd[s - 1] = As[(size_t)(o / G) * LD + o];
    } else if (b == 0 && tid == 0) {
      dd[s - 1] = __ldcg(A + (size_t)(s - 1) * s + (s - 1));
    }
  } else if (!RESIDENT && pending) {
    // hand-over: apply the pending update (entries q <-> global index k_end + q) to the rest of the matrix
    const int mr = s - k_end;
    for (int row = gwarp; row < mr; row += nwarp) {
      double* ar = A + (size_t)(k_end + row) * s + k_end;
      const double vi = vp[row], wi = wp[row];
      for (int j = lane; j < mr; j += 32)
        __stcg(ar + j, __dsub_rn(__ldcg(ar + j), __dadd_rn(__dmul_rn(vi, wp[j]), __dmul_rn(wi, vp[j]))));
    }
  }
}

// ---- 1b. tridiagonalisation of a SMALL matrix inside one thread-block cluster ----------------------------------------
// The Rayleigh-Ritz problems of chfsi.cu (order 256 .. 384) and the small Grams of configs 1 / 2 are too small for the
// grid-wide kernel above: its per-column cost is the ~2 us flag barrier across 148 CTAs.  Here the whole matrix lives
// in the shared memory of 8 or 16 CTAs (rows dealt round-robin), the products and the next column's row travel through
// distributed shared memory, and the one barrier per column is the hardware cluster barrier (~0.2 us).  Same
// algorithm and the same outputs (Vh, d, e, tau) as tridiag_kernel: pending rank-2 update applied inside the pass.
constexpr int TDC_THREADS = 512;  // measured at s = 256: 256 threads 0.80 ms, 512 threads 0.69 ms, 1024 threads 0.82 ms
constexpr int TDC_SMAX8 = 416;   // 8 CTAs (portable cluster size): 52 rows x 416 doubles + vectors per CTA
constexpr int TDC_SMAX16 = 832;  // 16 CTAs (non-portable size, B200 allows it): 52 rows x 832 doubles per CTA
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int TDC_NC>
__global__ void __launch_bounds__(TDC_THREADS)
tridiag_cluster_kernel(const double* __restrict__ A, int s, double* __restrict__ Vh, double* __restrict__ dd,
                       double* __restrict__ ee, double* __restrict__ tau_out) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) double sm[];
  const int LD = (s + 1) & ~1;
  const int RP = (s + TDC_NC - 1) / TDC_NC;
  double* vp = sm;            // pending v_{k-1}; entry q <-> global index k + q
  double* wp = vp + LD;       // pending w_{k-1}
  double* vs = wp + LD;       // current v_k; entry q <-> global index k + 1 + q
  double* xrow = vs + LD;     // 2 x LD: row k of the matrix (columns k+1 ..), written by the owner of the row
  double* pb = xrow + 2 * LD; // 2 x LD: products A v_k, written by the owners of the rows
  double* red = pb + 2 * LD;  // 32
  double* As = red + 32;      // RP x LD: rows rank, rank + NC, ...
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int l = 0; l < RP; ++l) {
    const int i = rank + l * TDC_NC;
    if (i < s)
      for (int c = tid; c < s; c += TDC_THREADS) As[(size_t)l * LD + c] = A[(size_t)i * s + c];
  }
  for (int j = tid; j < s - 1; j += TDC_THREADS) xrow[j] = A[1 + j];  // row 0
  double* xrow_r[TDC_NC];
#pragma unroll
  for (int r = 0; r < TDC_NC; ++r) xrow_r[r] = cluster.map_shared_rank(xrow, r);
  double* pb_mine = cluster.map_shared_rank(pb, lane & (TDC_NC - 1));  // the peer this lane delivers products to
  __syncthreads();
  cluster_barrier();  // every CTA of the cluster is resident before the first remote store
  bool pending = false;
  for (int k = 0; k < s - 1; ++k) {
    const int m = s - k - 1, par = k & 1;
    const double* xr = xrow + par * LD;
    // --- column k of the up-to-date matrix (redundant per CTA)
    const double v0 = pending ? vp[0] : 0.0, w0 = pending ? wp[0] : 0.0;
    double part = 0.0;
    for (int j = tid; j < m; j += TDC_THREADS) {
      double x = xr[j];
      if (pending) x = __dsub_rn(x, __dadd_rn(__dmul_rn(v0, wp[j + 1]), __dmul_rn(w0, vp[j + 1])));
      vs[j] = x;
      if (j > 0) part = fma(x, x, part);
    }
    const double xnorm2 = block_sum(part, red);
    const double alpha = vs[0];
    double beta, tau, scale;
    if (xnorm2 == 0.0) {
      beta = alpha;
      tau = 0.0;
      scale = 0.0;
    } else {
      const double nrm = sqrt(fma(alpha, alpha, xnorm2));
      beta = (alpha >= 0.0) ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    __syncthreads();
    for (int j = tid; j < m; j += TDC_THREADS) {
      const double v = (j == 0) ? 1.0 : vs[j] * scale;
      vs[j] = v;
      if ((j >> 5) % TDC_NC == rank) Vh[(size_t)k * s + j] = v;
    }
    __syncthreads();
    if (rank == k % TDC_NC && tid == 0) {
      double dk = As[(size_t)(k / TDC_NC) * LD + k];
      if (pending) dk = __dsub_rn(dk, __dadd_rn(__dmul_rn(v0, w0), __dmul_rn(w0, v0)));
      dd[k] = dk;
      ee[k] = beta;
      tau_out[k] = tau;
    }
    // --- pass over the own rows below k: apply the pending update, accumulate the product with v_k (warp per row)
    for (int l = wid; l < RP; l += TDC_THREADS / 32) {
      const int i = rank + l * TDC_NC;
      if (i <= k || i >= s) continue;  // warp-uniform
      const int row = i - (k + 1);
      const double vi = pending ? vp[row + 1] : 0.0, wi = pending ? wp[row + 1] : 0.0;
      double* ar = As + (size_t)l * LD + (k + 1);
      double acc = 0.0;
      for (int j = lane; j < m; j += 32) {
        double av = ar[j];
        if (pending) {
          av = __dsub_rn(av, __dadd_rn(__dmul_rn(vi, wp[j + 1]), __dmul_rn(wi, vp[j + 1])));
          ar[j] = av;
        }
        acc = fma(av, vs[j], acc);
        if (row == 0 && j > 0) {  // row k+1 as the next column needs it, to every CTA
#pragma unroll
          for (int r = 0; r < TDC_NC; ++r) xrow_r[r][(par ^ 1) * LD + (j - 1)] = av;
        }
      }
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane < TDC_NC) pb_mine[par * LD + row] = acc;  // lane r delivers the product to CTA r
    }
    cluster_barrier();
    // --- w_k = p - (tau/2)(p.v) v with p = tau A22 v; (v_k, w_k) become the pending update
    part = 0.0;
    for (int j = tid; j < m; j += TDC_THREADS) {
      const double p = pb[par * LD + j] * tau;
      wp[j] = p;
      part = fma(p, vs[j], part);
    }
    const double pv = block_sum(part, red);
    const double a2 = -0.5 * tau * pv;
    for (int j = tid; j < m; j += TDC_THREADS) {
      wp[j] = fma(a2, vs[j], wp[j]);
      vp[j] = vs[j];
    }
    pending = true;
    __syncthreads();
  }
  // the last pending update has tau == 0 (a 1 x 1 column): the matrix already holds the final corner
  if (rank == (s - 1) % TDC_NC && tid == 0) dd[s - 1] = As[(size_t)((s - 1) / TDC_NC) * LD + (s - 1)];
  cluster_barrier();  // nobody leaves while a peer could still address its shared memory
}

constexpr int BI_WARPS = 4;  // warps per eigenvalue: 128 probes per pass
// 1/q to ~1 ulp: hardware seed (20 bits) + two Newton steps; |q| >= pivmin (normal), so no special cases arise.
// The division is the whole dependent chain of a Sturm step; the IEEE sequence is ~3x longer.
__device__ __forceinline__ double fast_rcp(double q) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(q));
  r = fma(fma(-q, r, 1.0), r, r);
  r = fma(fma(-q, r, 1.0), r, r);
  return r;
}
// ---- 2. eigenvalues: multisection on the Sturm count ------------------------------------------------
// one CTA per wanted eigenvalue; lam_out[kk], kk = 0..K-1 descending  <=>  ascending index s-1-kk.
__global__ void __launch_bounds__(BI_WARPS * 32)
bisect_kernel(const double* __restrict__ dd, const double* __restrict__ ee, int s, int K, double* __restrict__ lam_out,
              double* __restrict__ tnorm_out) {
  extern __shared__ __align__(16) double sm[];
  double* d = sm;       // s
  double* e2 = sm + s;  // s
  const int tid = threadIdx.x, lane = tid & 31;
  double gl = DBL_MAX, gu = -DBL_MAX, emax = 0.0;
  for (int i = tid; i < s; i += BI_WARPS * 32) {
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
  __shared__ double rgl[BI_WARPS], rgu[BI_WARPS], rem[BI_WARPS];
  if (lane == 0) {
    rgl[tid >> 5] = gl;
    rgu[tid >> 5] = gu;
    rem[tid >> 5] = emax;
  }
  __syncthreads();
  gl = rgl[0];
  gu = rgu[0];
  emax = rem[0];
  for (int w = 1; w < BI_WARPS; ++w) {
    gl = fmin(gl, rgl[w]);
    gu = fmax(gu, rgu[w]);
    emax = fmax(emax, rem[w]);
  }
  const double ulp = DBL_EPSILON, safemin = DBL_MIN;
  const double pivmin = safemin * fmax(1.0, emax);
  const double tnorm = fmax(fabs(gl), fabs(gu));
  gl = gl - 2.1 * tnorm * ulp * s - 2.1 * pivmin;
  gu = gu + 2.1 * tnorm * ulp * s + 2.1 * pivmin;
  if (blockIdx.x == 0 && tid == 0 && tnorm_out) *tnorm_out = tnorm;

  const int kk = blockIdx.x;  // one CTA per wanted eigenvalue
  const int t = s - 1 - kk;   // ascending index of the wanted eigenvalue
  double lo = gl, hi = gu;    // count(lo) <= t < count(hi)
  // (NP + 1)-way multisection, one probe per thread: the Sturm recurrence is a pure latency chain (reciprocal,
  // multiply, subtract, clamp), so the probes of a pass run on different warps rather than interleaved in one
  constexpr int NP = BI_WARPS * 32;
  for (int pass = 0; pass < 40; ++pass) {
    const double width = hi - lo;
    const double tol = fmax(2.0 * ulp * fmax(fabs(lo), fabs(hi)), pivmin);
    if (width <= tol) break;  // uniform over the CTA
    const double x = lo + width * ((double)(tid + 1) / (double)(NP + 1));
    double q = 1.0;
    int cnt = 0;
    for (int i = 0; i < s; ++i) {
      const double ei = (i > 0) ? e2[i - 1] : 0.0;
      double qq = (d[i] - x) - ei * fast_rcp(q);  // the recurrence of sturm_count (core_math.cuh)
      qq = (fabs(qq) < pivmin) ? -pivmin : qq;
      cnt += (qq < 0.0);
      q = qq;
    }
    const int f = __syncthreads_count(cnt < t + 1);  // probes 0..f-1 lie below lambda_t, f.. above
    const double xlo = lo + width * ((double)f / (double)(NP + 1));
    const double xhi = lo + width * ((double)(f + 1) / (double)(NP + 1));
    if (f < NP) hi = xhi;
    if (f > 0) lo = xlo;
  }
  if (tid == 0) lam_out[kk] = 0.5 * (lo + hi);
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

// factor T - lam*I = P L U  (dgttrf recurrence);  dg: RECIPROCAL of the U diagonal, du: U first super-diagonal,
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
      dg[AT(i)] = 1.0 / dcur;  // the solves multiply by the reciprocal: no division on their dependent chain
      du[AT(i)] = ucur;
      du2[AT(i)] = 0.0;
      dl[AT(i)] = fact;
      piv[AT(i)] = 0;
      dcur = dnext - fact * ucur;
      ucur = unext;
    } else {
      const double fact = dcur / sub;
      dg[AT(i)] = 1.0 / sub;
      du[AT(i)] = dnext;
      du2[AT(i)] = unext;
      dl[AT(i)] = fact;
      piv[AT(i)] = 1;
      dcur = ucur - fact * dnext;
      ucur = -fact * unext;
    }
  }
  if (fabs(dcur) < tiny) dcur = (dcur < 0.0) ? -tiny : tiny;
  dg[AT(s - 1)] = 1.0 / dcur;
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
  double x1 = X[AT(s - 1)] * dg[AT(s - 1)];
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
        double xi = (b[u] - u1[u] * x1 - u2[u] * x2) * g[u];
        X[AT(i0 - u)] = xi;
        nrm = fma(xi, xi, nrm);
        x2 = x1;
        x1 = xi;
      }
  }
  // normalise (guards against overflow in the next solve); clusters are re-orthogonalised next
  const double inv = 1.0 / sqrt(nrm);
  for (int i0 = 0; i0 < s; i0 += PF) {  // loads batched ahead of the stores: one L2 round trip per PF rows
    double t[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (i0 + u < s) t[u] = X[AT(i0 + u)];
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (i0 + u < s) X[AT(i0 + u)] = t[u] * inv;
  }
#undef AT
}

// modified Gram-Schmidt inside each cluster of close eigenvalues; one CTA per cluster start.
// cstart[kk] = first index of the cluster that kk belongs to.
__global__ void __launch_bounds__(256)
invit_mgs_kernel(int s, int K, const int* __restrict__ cstart, double* __restrict__ X, int big_min) {
  __shared__ double red[32];
  const int k0 = blockIdx.x;
  if (cstart[k0] != k0) return;
  int k1 = k0 + 1;
  while (k1 < K && cstart[k1] == k0) ++k1;
  if (k1 == k0 + 1) return;  // singleton: already normalised by the solve
  if (k1 - k0 > big_min) return;  // large cluster: invit_cgs2_kernel
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

// Large clusters (more than CGS_MIN members): classical Gram-Schmidt applied twice ("twice is enough"), one CTA of
// 1024 threads per cluster.  Per vector j the inner products against ALL earlier members are formed in one pass
// (lane = member, warp = row slice; fixed-order cross-warp sum) and subtracted in one pass (thread = row), so the
// sequential depth is 5 block steps per vector instead of j: the 87-member cluster of the C3-shaped Gram costs
// 1.3 ms instead of 9.4 ms per sweep.  Same arithmetic on every run (fixed reduction orders).
constexpr int CGS_MIN = 12, CGS_THREADS = 1024;

__global__ void __launch_bounds__(CGS_THREADS)
invit_cgs2_kernel(int s, int K, const int* __restrict__ cstart, double* __restrict__ X, int min_size) {
  extern __shared__ double cg_sm[];  // part[32][mc] then coef[mc]
  __shared__ double red[32];
  const int k0 = blockIdx.x;
  if (cstart[k0] != k0) return;
  int k1 = k0 + 1;
  while (k1 < K && cstart[k1] == k0) ++k1;
  const int mc = k1 - k0;
  if (mc <= CGS_MIN || mc < min_size) return;  // small clusters: invit_mgs_kernel; mid-sized: Cholesky-QR (host loop)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double* part = cg_sm;                    // 32 x mc
  double* coef = cg_sm + (size_t)32 * mc;  // mc
  for (int jj = 0; jj < mc; ++jj) {
    const int j = k0 + jj;
    for (int pass = 0; pass < 2 && jj > 0; ++pass) {
      // inner products c_i = X_i . X_j, i < jj
      for (int g = 0; g < jj; g += 32) {
        const int i = g + lane;
        double a = 0.0;
        if (i < jj)
          for (int q = wid; q < s; q += 32) a = fma(X[(size_t)q * K + k0 + i], X[(size_t)q * K + j], a);
        if (i < jj) part[(size_t)wid * mc + i] = a;
      }
      __syncthreads();
      for (int i = tid; i < jj; i += CGS_THREADS) {
        double c = 0.0;
        for (int w = 0; w < 32; ++w) c += part[(size_t)w * mc + i];
        coef[i] = c;
      }
      __syncthreads();
      for (int q = tid; q < s; q += CGS_THREADS) {
        const double* xr = X + (size_t)q * K + k0;
        double v = xr[jj];
        for (int i = 0; i < jj; ++i) v = fma(-coef[i], xr[i], v);
        X[(size_t)q * K + j] = v;
      }
      __syncthreads();
    }
    double p2 = 0.0;
    for (int q = tid; q < s; q += CGS_THREADS) {
      const double x = X[(size_t)q * K + j];
      p2 = fma(x, x, p2);
    }
    const double inv = 1.0 / sqrt(block_sum(p2, red));
    for (int q = tid; q < s; q += CGS_THREADS) X[(size_t)q * K + j] *= inv;
    __syncthreads();
  }
}

// ---- 4. back-transformation ----------------------------------------------------------------------
// Y = H_0 H_1 ... H_{s-2} X with the reflectors grouped WY_NB at a time (compact WY, LAPACK dlarft/dlarfb):
//   H_{k0} ... H_{k0+nb-1} = I - V T V^T,  T upper triangular, T(i,i) = tau_i,
//   T(0:i, i) = -tau_i T(0:i, 0:i) (V(:, 0:i)^T v_i).
// The sequential depth drops from s reflectors (one reduction each) to s / WY_NB blocks (two streamed passes over
// the block's reflectors each); columns of X are independent, so CTAs never synchronise with each other.
// Block b, reflector i <-> column k0 + i; row offset t <-> global row k0 + 1 + t; V(t, i) = Vh[(k0+i) s + t - i], t >= i.
constexpr int WY_NB = 32;

__global__ void __launch_bounds__(256)
wy_T_kernel(int s, const double* __restrict__ Vh, const double* __restrict__ tau, double* __restrict__ Tg) {
  constexpr int RT = 64;
  __shared__ double Vs[RT][WY_NB + 1];
  __shared__ double Gs[WY_NB][WY_NB + 1];
  __shared__ double Ts[WY_NB][WY_NB + 1];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * WY_NB;
  const int nbk = min(WY_NB, (s - 1) - k0);
  const int mb = s - 1 - k0;
  const int a2 = (tid >> 4) * 2, b2 = (tid & 15) * 2;  // this thread's 2 x 2 tile of G = V^T V
  double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0;
  for (int t0 = 0; t0 < mb; t0 += RT) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RT * WY_NB / 256; ++q) {
      const int e = tid + q * 256, i = e / RT, tt = e % RT, t = t0 + tt;
      Vs[tt][i] = (i < nbk && t < mb && t >= i) ? Vh[(size_t)(k0 + i) * s + (t - i)] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int tt = 0; tt < RT; ++tt) {
      const double va0 = Vs[tt][a2], va1 = Vs[tt][a2 + 1], vb0 = Vs[tt][b2], vb1 = Vs[tt][b2 + 1];
      g00 = fma(va0, vb0, g00);
      g01 = fma(va0, vb1, g01);
      g10 = fma(va1, vb0, g10);
      g11 = fma(va1, vb1, g11);
    }
  }
  Gs[a2][b2] = g00;
  Gs[a2][b2 + 1] = g01;
  Gs[a2 + 1][b2] = g10;
  Gs[a2 + 1][b2 + 1] = g11;
  for (int e = tid; e < WY_NB * WY_NB; e += 256) Ts[e / WY_NB][e % WY_NB] = 0.0;
  __syncthreads();
  // column i of T from columns 0..i-1 (one warp; row r per lane)
  if (tid < 32) {
    for (int i = 0; i < nbk; ++i) {
      const double ti = tau[k0 + i];
      double acc = 0.0;
      if (tid < i)
        for (int cc = tid; cc < i; ++cc) acc = fma(Ts[tid][cc], Gs[cc][i], acc);
      __syncwarp();
      if (tid < i) Ts[tid][i] = -ti * acc;
      if (tid == i) Ts[i][i] = ti;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < WY_NB * WY_NB; e += 256) Tg[(size_t)blockIdx.x * WY_NB * WY_NB + e] = Ts[e / WY_NB][e % WY_NB];
}

constexpr int WY_NC = 2;         // eigenvector columns per CTA
constexpr int WY_THREADS = 512;
constexpr int WY_U = 8;          // loads in flight per thread

__global__ void __launch_bounds__(WY_THREADS)
wy_apply_kernel(int s, int K, const double* __restrict__ Vh, const double* __restrict__ Tg,
                const double* __restrict__ X, double* __restrict__ Y) {
  extern __shared__ __align__(16) double sm[];
  double* Xs = sm;                      // WY_NC x s
  double* Ts = Xs + (size_t)WY_NC * s;  // WY_NB x WY_NB
  double* Ws = Ts + WY_NB * WY_NB;      // WY_NB x WY_NC : V^T x
  double* W2 = Ws + WY_NB * WY_NC;      // WY_NB x WY_NC : T V^T x
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int c0 = blockIdx.x * WY_NC;
  for (int e = tid; e < WY_NC * s; e += WY_THREADS) {
    const int c = e / s, i = e - c * s;
    Xs[e] = (c0 + c < K) ? X[(size_t)i * K + c0 + c] : 0.0;
  }
  const int nblocks = (s - 1 + WY_NB - 1) / WY_NB;
  for (int blk = nblocks - 1; blk >= 0; --blk) {
    const int k0 = blk * WY_NB;
    const int nbk = min(WY_NB, (s - 1) - k0);
    const int mb = s - 1 - k0;
    __syncthreads();  // Xs of the previous block is final; Ts / Ws may be overwritten
    for (int e = tid; e < WY_NB * WY_NB; e += WY_THREADS) Ts[e] = Tg[(size_t)blk * WY_NB * WY_NB + e];
    // phase 1: Ws(i, c) = v_i . x_c ; a warp per reflector, lanes along the rows
    for (int i = wid; i < WY_NB; i += WY_THREADS / 32) {
      double acc[WY_NC];
#pragma unroll
      for (int c = 0; c < WY_NC; ++c) acc[c] = 0.0;
      if (i < nbk) {
        const int k = k0 + i, mi = s - k - 1;
        const double* v = Vh + (size_t)k * s;
        for (int j0 = lane; j0 < mi; j0 += 32 * WY_U) {
          double vv[WY_U];
#pragma unroll
          for (int u = 0; u < WY_U; ++u) {
            const int j = j0 + 32 * u;
            vv[u] = (j < mi) ? v[j] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < WY_U; ++u) {
            const int j = j0 + 32 * u;
            if (j < mi) {
#pragma unroll
              for (int c = 0; c < WY_NC; ++c) acc[c] = fma(vv[u], Xs[(size_t)c * s + k + 1 + j], acc[c]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < WY_NC; ++c) {
        double a = acc[c];
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) Ws[i * WY_NC + c] = a;
      }
    }
    __syncthreads();
    // W2 = T Ws (T upper triangular)
    if (tid < WY_NB * WY_NC) {
      const int i = tid / WY_NC, c = tid % WY_NC;
      double a = 0.0;
      for (int ip = i; ip < nbk; ++ip) a = fma(Ts[i * WY_NB + ip], Ws[ip * WY_NC + c], a);
      W2[tid] = a;
    }
    __syncthreads();
    // phase 2: x_c(t) -= sum_i V(t, i) W2(i, c) ; a thread per row, reflectors in batches
    for (int t = tid; t < mb; t += WY_THREADS) {
      double xa[WY_NC];
#pragma unroll
      for (int c = 0; c < WY_NC; ++c) xa[c] = Xs[(size_t)c * s + k0 + 1 + t];
      for (int i0 = 0; i0 < nbk; i0 += WY_U) {
        double vv[WY_U];
#pragma unroll
        for (int u = 0; u < WY_U; ++u) {
          const int i = i0 + u;
          vv[u] = (i < nbk && t >= i) ? Vh[(size_t)(k0 + i) * s + (t - i)] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < WY_U; ++u) {
          const int i = i0 + u;
          if (i < nbk) {
#pragma unroll
            for (int c = 0; c < WY_NC; ++c) xa[c] = fma(-vv[u], W2[i * WY_NC + c], xa[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < WY_NC; ++c) Xs[(size_t)c * s + k0 + 1 + t] = xa[c];
    }
  }
  __syncthreads();
  for (int e = tid; e < WY_NC * s; e += WY_THREADS) {
    const int c = e / s, i = e - c * s;
    if (c0 + c < K) Y[(size_t)i + (size_t)s * (c0 + c)] = Xs[e];
  }
}

}  // namespace

void eigh_topk_run(Ctx* c, double* G, int s, int K, double* lam, double* Y, bool psd) {
  if (K < 1 || K > s) fail(2, "eigh: need 1 <= K <= s (K=%d, s=%d)", K, s);
  // K << s: Chebyshev-filtered subspace iteration on the tensor cores (chfsi.cu); it certifies its own result
  // (residuals) and leaves G untouched, so anything it declines or fails on takes the direct route below
  static const bool no_chfsi = std::getenv("FLGP_EIGH_DIRECT") != nullptr;
  if (!no_chfsi && s >= 1024 && 5 * K <= s) {
    bool ok;
    {
      StageScope st(c, "eigh_chfsi");
      ok = chfsi_topk_run(c, G, s, K, lam, Y, psd);
    }
    if (ok) return;
  }
  eigh_direct_run(c, G, s, K, lam, Y);
}

void eigh_direct_run(Ctx* c, double* G, int s, int K, double* lam, double* Y) {
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
  // 1. tridiagonalisation (cooperative, persistent): streaming launch while the trailing matrix is larger than the
  //    grid's shared memory, resident launch for the rest
  static const bool no_cluster = std::getenv("FLGP_EIGH_NO_CLUSTER") != nullptr;
  if (s <= TDC_SMAX16 && !no_cluster) {
    const int nc = s <= TDC_SMAX8 ? 8 : 16;
    const int LD = (s + 1) & ~1, RP = (s + nc - 1) / nc;
    const size_t smem = ((size_t)7 * LD + 32 + (size_t)RP * LD) * sizeof(double);
    double f = 0.0;
    for (int k = 0; k < s - 1; ++k) f += 4.0 * (double)(s - k - 1) * (double)(s - k - 1);
    StageScope st(c, "eigh_tridiag_cluster", f, 0.0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc);
    cfg.blockDim = dim3(TDC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = nc;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const double* Gc = G;
    if (nc == 8) {
      FLGP_CUDA(cudaFuncSetAttribute(tridiag_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FLGP_CUDA(cudaLaunchKernelEx(&cfg, tridiag_cluster_kernel<8>, Gc, s, Vh.p, dd.p, ee.p, tau.p));
    } else {
      FLGP_CUDA(cudaFuncSetAttribute(tridiag_cluster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FLGP_CUDA(cudaFuncSetAttribute(tridiag_cluster_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      FLGP_CUDA(cudaLaunchKernelEx(&cfg, tridiag_cluster_kernel<16>, Gc, s, Vh.p, dd.p, ee.p, tau.p));
    }
    c->launches++;
  } else {
    const int grid = std::min(c->sm_count, 32 * FLAG_PER);  // one CTA per SM
    int smem_max = 0;
    FLGP_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    auto smem_need = [&](int M0, bool resident) {
      size_t words = (size_t)3 * M0 + 32 + TD_RACC * 32;
      if (resident) words += (size_t)((M0 + grid - 1) / grid) * ((M0 + 1) & ~1);
      return words * sizeof(double);
    };
    int M_res = 0;  // largest trailing order that can be kept resident
    for (int M0 = std::min(s, 4 * TD_THREADS_RES); M0 >= 2; --M0)
      if ((M0 + grid - 1) / grid <= TD_RPMAX && smem_need(M0, true) + 1024 <= (size_t)smem_max) {
        M_res = M0;
        break;
      }
    if (std::getenv("FLGP_EIGH_NO_RESIDENT")) M_res = 0;
    const int k_switch = (M_res >= 2) ? std::max(0, s - M_res) : s - 1;  // columns [0, k_switch) stream through L2
    const bool prof_on = std::getenv("FLGP_EIGH_PROF") != nullptr;
    DevBuf<long long> prof(24);
    DevBuf<unsigned> flags((size_t)grid * FLAG_STRIDE);
    DevBuf<double> rowbuf((size_t)2 * s);
    flags.zero(c->stream);
    if (prof_on) prof.zero(c->stream);
    auto launch = [&](bool resident, int kb, int ke, long long* profp) {
      const size_t smem = smem_need(s - kb, resident);
      if (smem + 1024 > (size_t)smem_max) fail(2, "eigh: s=%d needs more shared memory than one SM has", s);
      const void* fn = resident ? (const void*)tridiag_kernel<true, TD_THREADS_RES>
                                : (const void*)tridiag_kernel<false, TD_THREADS_STREAM>;
      const int threads = resident ? TD_THREADS_RES : TD_THREADS_STREAM;
      FLGP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      FLGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
      if (per_sm < 1) fail(2, "eigh: the tridiagonalisation kernel does not fit on an SM (s=%d)", s);
      int s_ = s, kb_ = kb, ke_ = ke;
      void* args[] = {&G, &s_, &kb_, &ke_, &Vh.p, &dd.p, &ee.p, &tau.p, &pbuf.p, &rowbuf.p, &flags.p, &profp};
      FLGP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, c->stream));
      c->launches++;
    };
    // flops of columns [a, b): sum over k of 4 (s-k-1)^2  (symmetric product + rank-2 update)
    auto td_flops = [&](int a, int b) {
      double f = 0.0;
      for (int k = a; k < b; ++k) f += 4.0 * (double)(s - k - 1) * (double)(s - k - 1);
      return f;
    };
    if (k_switch > 0) {
      const int ke = std::min(k_switch, s - 1);
      StageScope st(c, "eigh_tridiag_streaming", td_flops(0, ke), 4.0 * td_flops(0, ke));  // 16 B per element and column
      launch(false, 0, ke, prof_on ? prof.p : nullptr);
    }
    if (k_switch < s - 1) {
      StageScope st(c, "eigh_tridiag_resident", td_flops(k_switch, s - 1), 0.0);
      launch(true, k_switch, s - 1, prof_on ? prof.p + 12 : nullptr);
    }
    if (prof_on) {
      long long h[24];
      prof.download(h, 24, c->stream);
      sync(c);
      fprintf(stderr, "[flgp eigh prof] tridiag s=%d, streaming columns [0,%d), resident [%d,%d); cycles per column (CTA 0)\n",
              s, k_switch, k_switch, s - 1);
      for (int ph = 0; ph < 2; ++ph) {
        const int cols = ph == 0 ? std::min(k_switch, s - 1) : (s - 1 - k_switch);
        if (cols <= 0) continue;
        fprintf(stderr, "  %s:", ph == 0 ? "streaming" : "resident ");
        for (int i = 0; i < 10; ++i) fprintf(stderr, " %d:%.0f", i, (double)h[12 * ph + i] / cols);
        fprintf(stderr, "\n");
      }
    }
  }
  // 2. eigenvalues
  {
    size_t smem = (size_t)2 * s * sizeof(double);
    if (smem > 48 * 1024)
      FLGP_CUDA(cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageScope st(c, "eigh_bisect");
    FLGP_LAUNCH(c, bisect_kernel, K, BI_WARPS * 32, smem, dd.p, ee.p, s, K, lam, tnorm.p);
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
  int max_cluster = 1;
  for (int k = 0, run = 0; k < K; ++k) {
    run = (cstart[k] == k) ? 1 : run + 1;
    max_cluster = std::max(max_cluster, run);
  }
  const size_t cgs_smem = (size_t)33 * max_cluster * sizeof(double);
  const bool cgs_ok = max_cluster > CGS_MIN && cgs_smem <= 200 * 1024;
  if (cgs_ok && cgs_smem > 40 * 1024)
    FLGP_CUDA(cudaFuncSetAttribute(invit_cgs2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cgs_smem));
  // 3. inverse iteration
  const size_t sk = (size_t)s * K;
  DevBuf<double> dg(sk), du(sk), du2(sk), dl(sk), X(sk);
  DevBuf<unsigned char> piv(sk);
  const double tiny = DBL_EPSILON * std::max(tn, DBL_MIN / DBL_EPSILON);
  {
    StageScope st(c, "eigh_inverse_iteration");
    FLGP_LAUNCH(c, invit_factor_kernel, ceil_div(K, 64), 64, 0, dd.p, ee.p, s, K, lam, tiny, dg.p, du.p, du2.p, dl.p,
                piv.p, X.p);
    // clusters of 13 .. 512 close eigenvalues: Cholesky-QR (Gram + cluster Cholesky + rotation: GEMM-shaped, all SMs)
    // instead of Gram-Schmidt by one CTA, which was 67 of the 87 ms of the Nystrom anchor eigensolve (150-member
    // cluster of an exponentially decaying spectrum).  Any orthonormal basis of the cluster's span serves inverse
    // iteration; a Gram that is numerically singular falls back to the Gram-Schmidt kernel.
    std::vector<std::pair<int, int>> mid;  // (first column, size)
    for (int k = 0; k < K;) {
      int e = k + 1;
      while (e < K && cstart[e] == k) ++e;
      if (e - k > CGS_MIN && e - k <= 512) mid.emplace_back(k, e - k);
      k = e;
    }
    int mid_max = 0;
    for (auto& m_ : mid) mid_max = std::max(mid_max, m_.second);
    DevBuf<double> Sc((size_t)std::max(mid_max, 1) * std::max(mid_max, 1)), Lc((size_t)std::max(mid_max, 1) * std::max(mid_max, 1)),
        Tc((size_t)s * std::max(mid_max, 1));
    auto cluster_cholqr = [&](int k0, int mc) -> bool {
      for (int pass = 0; pass < 2; ++pass) {
        // S = Xc^T Xc (Xc = columns k0 .. k0+mc-1 of the row-major s x K array X)
        gemm_general_splitk_run(c, X.p + k0, 1, K, X.p + k0, K, 1, mc, mc, s, Sc.p);
        double ratio = 0.0;
        if (!chol_inv_run(c, Sc.p, mc, Lc.p, &ratio)) return false;
        // Xc <- Xc L^-T : T(q, j) = sum_i Xc(q, i) Linv(j, i)
        gemm_general_run(c, X.p + k0, K, 1, Lc.p, 1, mc, s, mc, mc, Tc.p, mc, 1);
        FLGP_CUDA(cudaMemcpy2DAsync(X.p + k0, sizeof(double) * K, Tc.p, sizeof(double) * mc, sizeof(double) * mc, s,
                                    cudaMemcpyDeviceToDevice, c->stream));
        if (ratio > 0.3) break;
      }
      return true;
    };
    for (int it = 0; it < 3; ++it) {
      FLGP_LAUNCH(c, invit_solve_kernel, ceil_div(K, 64), 64, 0, s, K, dg.p, du.p, du2.p, dl.p, piv.p, X.p);
      FLGP_LAUNCH(c, invit_mgs_kernel, K, 256, 0, s, K, cs.p, X.p, cgs_ok ? CGS_MIN : K);
      bool chol_ok = true;
      for (auto& m_ : mid) chol_ok = chol_ok && cluster_cholqr(m_.first, m_.second);
      // clusters beyond 512 members, and everything mid-sized when a Gram was singular: Gram-Schmidt kernel
      if (cgs_ok && (max_cluster > 512 || !chol_ok))
        FLGP_LAUNCH(c, invit_cgs2_kernel, K, CGS_THREADS, cgs_smem, s, K, cs.p, X.p, chol_ok ? 513 : 0);
    }
  }
  // 4. back-transformation (blocked compact WY)
  {
    StageScope st(c, "eigh_backtransform", 4.0 * s * (double)s * K / 2.0, 0.0);
    const int nblocks = (s - 1 + WY_NB - 1) / WY_NB;
    DevBuf<double> Tg((size_t)nblocks * WY_NB * WY_NB);
    FLGP_LAUNCH(c, wy_T_kernel, nblocks, 256, 0, s, Vh.p, tau.p, Tg.p);
    size_t smem = ((size_t)WY_NC * s + WY_NB * WY_NB + 2 * WY_NB * WY_NC) * sizeof(double);
    if (smem > 48 * 1024)
      FLGP_CUDA(cudaFuncSetAttribute(wy_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLGP_LAUNCH(c, wy_apply_kernel, ceil_div(K, WY_NC), WY_THREADS, smem, s, K, Vh.p, Tg.p, X.p, Y);
    sync(c);  // Tg is released on return
  }
  sync(c);
}

}  // namespace flgp
