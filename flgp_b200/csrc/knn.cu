// knn.cu — point-to-anchor squared distances + top-r selection
// (replaces KNN_cpp / KNN_Index, /root/reference/src/Utils.cpp:72-97, 102-192).
//
//   D(i,j) = ((-2 * sum_k x_ik u_jk) + |x_i|^2) + |u_j|^2          (src/Utils.cpp:121)
// evaluated with separately rounded multiplies and adds in ascending k — the exact fp64 value
// the oracle computes — and fed IN ANCHOR ORDER to a per-point emulation of libstdc++'s
// std::partial_sort (__heap_select + __sort_heap), so indices are bit-exact including ties.
// The r-entry heap lives with the thread that owns the point; only its top key is compared in
// the hot loop.  No n x s distance matrix is ever materialised (the reference batches 100 rows).
//
// Roofline: FP64 pipe (2*s*d flop per point); bytes 8d read + 4r (+8r) written.
#include <algorithm>
#include <cmath>

#include <cstdlib>

#include "kernels.cuh"

namespace flgp {

namespace {

constexpr int KNN_RMAX = 32;

// |row|^2, sequential, separately rounded (X.rowwise().squaredNorm())
__global__ void rownorm_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, double* out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0;
  for (int k = 0; k < d; ++k) {
    double x = X[i + ldx * k];
    a = __dadd_rn(a, __dmul_rn(x, x));
  }
  out[i] = a;
}

// anchors -> records [u_0 .. u_{d-1}, |u|^2, pad]
__global__ void knn_prep_kernel(const double* __restrict__ U, int s, int64_t ldu, int d, int str, double* rec) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s) return;
  double a = 0.0;
  for (int k = 0; k < d; ++k) {
    double u = U[j + ldu * k];
    a = __dadd_rn(a, __dmul_rn(u, u));
    rec[(size_t)j * str + k] = u;
  }
  rec[(size_t)j * str + d] = a;
  for (int k = d + 1; k < str; ++k) rec[(size_t)j * str + k] = 0.0;
}

struct TopR {
  double hk[KNN_RMAX];
  int hi[KNN_RMAX];
  double top;
  int r;
  __device__ __forceinline__ void push(double dist, int j) {
    // elements 0..r-1 fill the heap range [first, middle); make_heap once it is full
    if (j < r) {
      hk[j] = dist;
      hi[j] = j;
      if (j == r - 1) {
        heap_make(hk, hi, r);
        top = hk[0];
      }
    } else if (dist < top) {
      heap_adjust(hk, hi, 0, r, dist, j);
      top = hk[0];
    }
  }
  __device__ __forceinline__ void finish(int64_t i, int64_t n, int32_t* ind, double* dist) {
    heap_sort(hk, hi, r);
    for (int a = 0; a < r; ++a) {
      ind[i + n * a] = hi[a];
      if (dist) dist[i + n * a] = hk[a];
    }
  }
};

// ---- small d: thread-per-point(s), anchor records broadcast from shared memory ----------------------
// R > 0: the r-entry heap is a RegHeap<R> (registers only, compile-time r); R == 0: run-time r, local arrays.
// Hot loop per (point, anchor): 8 fp64-pipe ops for the reference's distance expression, then an INTEGER
// compare of the high word against the heap top's (for non-negative doubles a < b implies hi(a) <= hi(b);
// a negative distance has a negative high word and always passes; a negative top disables the filter).
// Only the rare candidate pays the exact fp64 compare and the heap update.
template <int R>
struct HeapState {
  RegHeap<R> h;
  __device__ __forceinline__ void fill(int j, double d) { h.set(j, d, j); }
  __device__ __forceinline__ void make(int) { heap_make_acc(h, R); }
  __device__ __forceinline__ double top() const { return h.k[0]; }
  __device__ __forceinline__ void replace(int, double d, int j) { heap_adjust_acc(h, 0, R, d, j); }
  __device__ __forceinline__ void finish(int, int64_t i, int64_t n, int32_t* ind, double* dist) {
    heap_sort_acc(h, R);
#pragma unroll
    for (int a = 0; a < R; ++a) {
      ind[i + n * a] = h.id[a];
      if (dist) dist[i + n * a] = h.k[a];
    }
  }
};
template <>
struct HeapState<0> {
  double hk[KNN_RMAX];
  int hi[KNN_RMAX];
  __device__ __forceinline__ void fill(int j, double d) {
    hk[j] = d;
    hi[j] = j;
  }
  __device__ __forceinline__ void make(int r) { heap_make(hk, hi, r); }
  __device__ __forceinline__ double top() const { return hk[0]; }
  __device__ __forceinline__ void replace(int r, double d, int j) { heap_adjust(hk, hi, 0, r, d, j); }
  __device__ __forceinline__ void finish(int r, int64_t i, int64_t n, int32_t* ind, double* dist) {
    heap_sort(hk, hi, r);
    for (int a = 0; a < r; ++a) {
      ind[i + n * a] = hi[a];
      if (dist) dist[i + n * a] = hk[a];
    }
  }
};

__device__ __forceinline__ int top_filter(double top) { return top < 0.0 ? 0x7fffffff : __double2hiint(top); }

// sel / nsel / perm (all or none): the kernel processes the *nsel source rows sel[0..), row q of X, and writes
// the result to output row perm[q] (the pruned path's fallback for the few points it cannot decide, below).
template <int D, int R, int P>
__global__ void __launch_bounds__(256)
knn_small_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, const double* __restrict__ rec, int s, int r_in,
                 int32_t* __restrict__ ind, double* __restrict__ dist, int chunk, int one,
                 const int32_t* __restrict__ sel, const int* __restrict__ nsel, const int32_t* __restrict__ perm,
                 int64_t n_out) {
  constexpr int STR = (D + 2) / 2 * 2;
  const int r = R ? R : r_in;
  extern __shared__ __align__(16) double srec[];
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * (256 * P);
  const int64_t count = sel ? (int64_t)*nsel : n;
  if (base >= count) return;
  double x[P][D], xn[P];
  HeapState<R> hs[P];
  int th[P];
  int64_t orow[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int64_t q = base + (int64_t)p * 256 + tid;
    const int64_t i = (q < count) ? (sel ? (int64_t)sel[q] : q) : -1;
    orow[p] = (i < 0) ? -1 : ((sel && perm) ? (int64_t)perm[i] : i);
    xn[p] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      x[p][k] = (i >= 0) ? X[i + ldx * k] : 0.0;
      xn[p] = __dadd_rn(xn[p], __dmul_rn(x[p][k], x[p][k]));
    }
    th[p] = 0x7fffffff;
  }
  for (int c0 = 0; c0 < s; c0 += chunk) {
    const int cnt = min(chunk, s - c0);
    __syncthreads();
    for (int t = tid; t < cnt * STR; t += 256) srec[t] = rec[(size_t)c0 * STR + t];
    __syncthreads();
    int j = 0;
    if (c0 == 0) {
      // the first r anchors are the heap range [first, middle) of std::partial_sort
      for (; j < r; ++j) {
        const double* ur = srec + (size_t)j * STR;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) acc = __dadd_rn(acc, __dmul_rn(x[p][k], ur[k]));
          hs[p].fill(j, __dadd_rn(__dadd_rn(__dmul_rn(-2.0, acc), xn[p]), ur[D]));
        }
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        hs[p].make(r);
        th[p] = top_filter(hs[p].top());
      }
    }
#pragma unroll 2
    for (; j < cnt; ++j) {
      double ur[STR];
      const double2* rj = reinterpret_cast<const double2*>(srec + (size_t)j * STR);
#pragma unroll
      for (int q = 0; q < STR / 2; ++q) {
        double2 t = rj[q];
        ur[2 * q] = t.x;
        ur[2 * q + 1] = t.y;
      }
      double dd[P];
      bool cand = false;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) acc = __dadd_rn(acc, __dmul_rn(x[p][k], ur[k]));
        dd[p] = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, acc), xn[p]), ur[D]);
        cand |= (__double2hiint(dd[p]) <= th[p]);
      }
      if (cand) {
#pragma unroll 1
        for (int q = 0; q < one; ++q) {  // `one` == 1: keeps this a real, rarely taken branch
#pragma unroll
          for (int p = 0; p < P; ++p)
            if (dd[p] < hs[p].top()) {
              hs[p].replace(r, dd[p], c0 + j);
              th[p] = top_filter(hs[p].top());
            }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p)
    if (orow[p] >= 0) hs[p].finish(r, orow[p], n_out, ind, dist);
}

// ---- exact pruning on the cluster-sorted layout that k-means leaves behind (small d) -----------------------
// Rows are grouped in segments; segment a holds rows that are close to anchor a (its k-means cluster at the last
// sort - but correctness below does not depend on that).  For a row x of segment a let
//   ub   >= |x - U_a|                       (computed from the row's own distance to U_a, plus rounding slack)
//   rho_a = |U_a - its r-th nearest anchor|, U_a itself counted first (so r anchors lie within rho_a of U_a).
// Those r anchors are within ub + rho_a of x, hence x's r-th smallest distance is <= ub + rho_a, and an anchor j with
//   |U_a - U_j| >= 2 ub + rho_a + eta   has   |x - U_j| >= ub + rho_a + eta,
// i.e. its true distance exceeds that of r other anchors by eta, and its COMPUTED squared distance (rounding error
// <= Delta, eta^2 > 2 Delta) is strictly larger than theirs: j is neither among the r nearest nor tied with them.
// Scanning {j : |U_a - U_j| < 2 ub + rho_a + eta} therefore yields the same r smallest values as the full scan.
// std::partial_sort's answer is unique when those r values and the next one are pairwise distinct; if the row sees
// an exact tie among its r+1 smallest, or its bound reaches past the staged list, it is handed to the brute-force
// kernel above (libstdc++ heap emulation over all anchors).  Lists are sorted by anchor-anchor distance so that a
// row stops at its own bound.
constexpr int KNN_LMAX = 256;

template <int R>
struct TopSorted {  // the R + 1 smallest (key, id) pairs seen, ascending, ties by id: registers only
  double k[R + 1];
  int id[R + 1];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int a = 0; a <= R; ++a) {
      k[a] = INFINITY;
      id[a] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ void insert(double d, int j) {
    if (d < k[R] || (d == k[R] && j < id[R])) {
      k[R] = d;
      id[R] = j;
#pragma unroll
      for (int u = R; u > 0; --u) {
        const bool lt = k[u] < k[u - 1] || (k[u] == k[u - 1] && id[u] < id[u - 1]);
        const double tk = k[u];
        const int ti = id[u];
        if (lt) {
          k[u] = k[u - 1];
          id[u] = id[u - 1];
          k[u - 1] = tk;
          id[u - 1] = ti;
        }
      }
    }
  }
  __device__ __forceinline__ bool tie() const {
    bool t = false;
#pragma unroll
    for (int a = 0; a < R; ++a) t |= (k[a] == k[a + 1]);  // +inf padding never equals a finite key
    return t;
  }
};

// per anchor: rho_a, then the anchors within thr = 2 R_a + rho_a + eta sorted by distance (a itself first);
// if more than KNN_LMAX qualify the radius is halved until they fit.  lrad[a] = the radius the list is complete to.
__global__ void __launch_bounds__(256)
knn_lists_kernel(const double* __restrict__ U, int s, int64_t ldu, int d, int r,
                 const unsigned long long* __restrict__ Rbits, const double* __restrict__ move, double eta,
                 int32_t* __restrict__ list_j, double* __restrict__ list_cc, int32_t* __restrict__ len,
                 double* __restrict__ lrad, double* __restrict__ rho_out) {
  __shared__ double kcc[KNN_LMAX];
  __shared__ int kj[KNN_LMAX];
  __shared__ double rk[8];
  __shared__ int rj[8];
  __shared__ int count;
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double ua[4];
  for (int k = 0; k < d; ++k) ua[k] = U[a + ldu * k];
  auto ccof = [&](int j) {
    double cc = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = ua[k] - U[j + ldu * k];
      cc = fma(df, df, cc);
    }
    return sqrt(cc);
  };
  // r rounds of "smallest (cc, j) above the previous one": the last is rho_a
  double pk = -1.0;
  int pj = -1;
  for (int round = 0; round < r; ++round) {
    double bk = INFINITY;
    int bj = 0x7fffffff;
    for (int j = tid; j < s; j += 256) {
      const double cc = ccof(j);
      const bool above = cc > pk || (cc == pk && j > pj);
      if (above && (cc < bk || (cc == bk && j < bj))) {
        bk = cc;
        bj = j;
      }
    }
    for (int o = 16; o; o >>= 1) {
      const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ok < bk || (ok == bk && oj < bj)) {
        bk = ok;
        bj = oj;
      }
    }
    __syncthreads();
    if (lane == 0) {
      rk[wid] = bk;
      rj[wid] = bj;
    }
    __syncthreads();
    bk = rk[0];
    bj = rj[0];
    for (int w = 1; w < 8; ++w)
      if (rk[w] < bk || (rk[w] == bk && rj[w] < bj)) {
        bk = rk[w];
        bj = rj[w];
      }
    pk = bk;
    pj = bj;
  }
  const double rho = pk * (1.0 + 1e-9);
  const double Ra = Rbits ? __longlong_as_double((long long)Rbits[a]) + move[a] : 0.0;
  double thr = (2.0 * Ra + rho + eta) * (1.0 + 1e-9);
  int cnt;
  while (true) {
    __syncthreads();
    if (tid == 0) count = 0;
    __syncthreads();
    for (int j = tid; j < s; j += 256) {
      const double cc = ccof(j);
      if (cc < thr) {
        const int pos = atomicAdd(&count, 1);
        if (pos < KNN_LMAX) {
          kcc[pos] = cc;
          kj[pos] = j;
        }
      }
    }
    __syncthreads();
    cnt = count;
    if (cnt <= KNN_LMAX) break;
    thr *= 0.5;  // too many: a shorter list is still exact for the rows whose bound fits (the others fall back)
  }
  int npow = 1;
  while (npow < cnt) npow <<= 1;
  for (int t = cnt + tid; t < npow; t += 256) {
    kcc[t] = INFINITY;
    kj[t] = 0x7fffffff;
  }
  __syncthreads();
  for (int k2 = 2; k2 <= npow; k2 <<= 1)
    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      for (int t = tid; t < npow; t += 256) {
        const int u = t ^ j2;
        if (u > t) {
          const bool up = (t & k2) == 0;
          const double c0 = kcc[t], c1 = kcc[u];
          const int i0 = kj[t], i1 = kj[u];
          const bool gt = (c0 > c1) || (c0 == c1 && i0 > i1);
          if (gt == up) {
            kcc[t] = c1;
            kcc[u] = c0;
            kj[t] = i1;
            kj[u] = i0;
          }
        }
      }
      __syncthreads();
    }
  for (int t = tid; t < cnt; t += 256) {
    list_j[(size_t)a * KNN_LMAX + t] = kj[t];
    list_cc[(size_t)a * KNN_LMAX + t] = kcc[t];
  }
  if (tid == 0) {
    len[a] = cnt;
    lrad[a] = thr;
    rho_out[a] = rho;
  }
}

template <int D, int R>
__global__ void __launch_bounds__(256)
knn_segment_kernel(const double* __restrict__ Xs, int64_t n, const double* __restrict__ rec,
                   const int32_t* __restrict__ perm, const int* __restrict__ seg_start,
                   const int32_t* __restrict__ list_j, const double* __restrict__ list_cc,
                   const int32_t* __restrict__ len, const double* __restrict__ lrad, const double* __restrict__ rho,
                   double delta2, double eta, int32_t* __restrict__ ind, double* __restrict__ dist,
                   int32_t* __restrict__ strag, int* __restrict__ nstrag) {
  constexpr int STR = (D + 2) / 2 * 2;
  __shared__ __align__(16) double srec[KNN_LMAX * STR];
  __shared__ double scc[KNN_LMAX];
  __shared__ int sj[KNN_LMAX];
  const int a = blockIdx.x, tid = threadIdx.x;
  const int beg = seg_start[a], end = seg_start[a + 1];
  if (beg == end) return;
  const int L = len[a];
  for (int q = tid; q < L; q += 256) {
    sj[q] = list_j[(size_t)a * KNN_LMAX + q];
    scc[q] = list_cc[(size_t)a * KNN_LMAX + q];
  }
  __syncthreads();
  for (int t = tid; t < L * STR; t += 256) srec[t] = rec[(size_t)sj[t / STR] * STR + (t % STR)];
  __syncthreads();
  const double radius = lrad[a], rho_a = rho[a];
  for (int p = beg + tid; p < end; p += 256) {
    double x[D], xn = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      x[k] = Xs[p + n * k];
      xn = __dadd_rn(xn, __dmul_rn(x[k], x[k]));
    }
    auto sqdist = [&](int q) {  // the reference's expression, src/Utils.cpp:121
      const double* ur = srec + (size_t)q * STR;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) acc = __dadd_rn(acc, __dmul_rn(x[k], ur[k]));
      return __dadd_rn(__dadd_rn(__dmul_rn(-2.0, acc), xn), ur[D]);
    };
    TopSorted<R> top;
    top.init();
    const double d0 = sqdist(0);  // entry 0 is anchor a itself
    top.insert(d0, sj[0]);
    const double ub = sqrt(fmax(d0, 0.0) + delta2) * (1.0 + 1e-14);
    const double thr = (2.0 * ub + rho_a + eta) * (1.0 + 1e-9);
    bool fallback = !(thr <= radius) || L < 1;
    if (!fallback) {
      for (int q = 1; q < L; ++q) {
        if (scc[q] >= thr) break;  // sorted by anchor-anchor distance: nothing further can reach the top r
        top.insert(sqdist(q), sj[q]);
      }
      fallback = top.tie();
    }
    if (fallback) {  // warp-aggregated append
      const unsigned act = __activemask();
      const int lane = tid & 31, lead = __ffs(act) - 1;
      int base = 0;
      if (lane == lead) base = atomicAdd(nstrag, __popc(act));
      base = __shfl_sync(act, base, lead);
      strag[base + __popc(act & ((1u << lane) - 1))] = p;
      continue;
    }
    const int64_t i = perm ? (int64_t)perm[p] : (int64_t)p;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      ind[i + n * q] = top.id[q];
      if (dist) dist[i + n * q] = top.k[q];
    }
  }
}

// ---- any d: 64 x 64 distance tiles in shared memory, scanned in anchor order ----------------------
constexpr int NT_TP = 64, NT_TC = 64, NT_K = 8;  // 4+4+32.5 KB static shared memory

__global__ void __launch_bounds__(256)
knn_tiled_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, const double* __restrict__ U, int s,
                 int64_t ldu, const double* __restrict__ xn, const double* __restrict__ un, int r,
                 int32_t* __restrict__ ind, double* __restrict__ dist) {
  __shared__ double Xs[NT_K][NT_TP];
  __shared__ double Us[NT_K][NT_TC];
  __shared__ double Ds[NT_TP][NT_TC + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * NT_TP;
  TopR h;
  h.r = r;
  h.top = INFINITY;
  const int64_t iscan = i0 + tid;  // threads 0..63 own one point each for the scan
  const double my_xn = (tid < NT_TP && iscan < n) ? xn[iscan] : 0.0;
  for (int c0 = 0; c0 < s; c0 += NT_TC) {
    double av[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) av[a][b] = 0.0;
    for (int k0 = 0; k0 < d; k0 += NT_K) {
      __syncthreads();
#pragma unroll
      for (int q = 0; q < NT_K * 64 / 256; ++q) {
        int e = tid + q * 256, kk = e >> 6, p = e & 63;
        int k = k0 + kk;
        int64_t i = i0 + p;
        Xs[kk][p] = (k < d && i < n) ? X[i + ldx * k] : 0.0;
        int j = c0 + p;
        Us[kk][p] = (k < d && j < s) ? U[j + ldu * k] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < NT_K; ++kk) {  // zero padding adds +0 exactly
        double xa[4], ub[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) xa[a] = Xs[kk][ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) ub[b] = Us[kk][tx * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) av[a][b] = __dadd_rn(av[a][b], __dmul_rn(xa[a], ub[b]));
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) Ds[ty * 4 + a][tx * 4 + b] = av[a][b];
    __syncthreads();
    if (tid < NT_TP && iscan < n) {
      const int cnt = min(NT_TC, s - c0);
      for (int j = 0; j < cnt; ++j) {
        double dd = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, Ds[tid][j]), my_xn), un[c0 + j]);
        h.push(dd, c0 + j);
      }
    }
    // the next chunk's first __syncthreads orders these reads before Ds is rewritten
  }
  if (tid < NT_TP && iscan < n) h.finish(iscan, n, ind, dist);
}

template <int D, int R>
void launch_small_r(Ctx* c, const double* X, int64_t n, int64_t ldx, const double* rec, int s, int r,
                    int32_t* ind, double* dist) {
  constexpr int STR = (D + 2) / 2 * 2;
  constexpr int P = R ? 2 : 1;
  int chunk = std::min(s, 1024);
  size_t smem = (size_t)chunk * STR * sizeof(double);
  int grid = ceil_div(n, 256 * P);
  FLGP_LAUNCH(c, (knn_small_kernel<D, R, P>), grid, 256, smem, X, n, ldx, rec, s, r, ind, dist, chunk, 1,
              (const int32_t*)nullptr, (const int*)nullptr, (const int32_t*)nullptr, n);
}

// pruned scan over the cluster-sorted layout + brute-force fallback for the undecided rows
template <int D, int R>
void launch_pruned_r(Ctx* c, const KMeansSorted& so, int64_t n, const double* rec, const double* U, int s, int64_t ldu,
                     int32_t* ind, double* dist, bool keep_sorted) {
  const int32_t* perm = keep_sorted ? nullptr : so.perm.p;  // null: output row = sorted position
  constexpr int STR = (D + 2) / 2 * 2;
  constexpr int P = 2;
  const double m = so.maxabs;
  const double Delta = 8.0 * (D + 4) * 1.1102230246251565e-16 * (4.0 * D * m * m);  // rounding of one computed distance
  const double delta2 = 4.0 * Delta;
  const double eta = 2.0 * std::sqrt(2.0 * Delta) + 1e-12 * m;  // eta^2 > 2 Delta
  DevBuf<int32_t> lj((size_t)s * KNN_LMAX), llen(s), strag((size_t)n);
  DevBuf<double> lcc((size_t)s * KNN_LMAX), lrad(s), rho(s);
  DevBuf<int> nstrag(1);
  nstrag.zero(c->stream);
  FLGP_LAUNCH(c, knn_lists_kernel, s, 256, 0, U, s, ldu, D, R, so.Rbits.p, so.move.p, eta, lj.p, lcc.p, llen.p, lrad.p,
              rho.p);
  FLGP_LAUNCH(c, (knn_segment_kernel<D, R>), s, 256, 0, so.Xs.p, n, rec, perm, so.seg_start.p, lj.p, lcc.p, llen.p,
              lrad.p, rho.p, delta2, eta, ind, dist, strag.p, nstrag.p);
  int chunk = std::min(s, 1024);
  size_t smem = (size_t)chunk * STR * sizeof(double);
  int grid = ceil_div(n, 256 * P);  // blocks beyond *nstrag exit at once
  FLGP_LAUNCH(c, (knn_small_kernel<D, R, P>), grid, 256, smem, so.Xs.p, n, n, rec, s, R, ind, dist, chunk, 1, strag.p,
              nstrag.p, perm, n);
  sync(c);  // the work buffers are freed on return
}
template <int D>
bool launch_pruned(Ctx* c, const KMeansSorted& so, int64_t n, const double* rec, const double* U, int s, int64_t ldu,
                   int r, int32_t* ind, double* dist, bool ks) {
  switch (r) {
    case 1: launch_pruned_r<D, 1>(c, so, n, rec, U, s, ldu, ind, dist, ks); return true;
    case 2: launch_pruned_r<D, 2>(c, so, n, rec, U, s, ldu, ind, dist, ks); return true;
    case 3: launch_pruned_r<D, 3>(c, so, n, rec, U, s, ldu, ind, dist, ks); return true;
    case 4: launch_pruned_r<D, 4>(c, so, n, rec, U, s, ldu, ind, dist, ks); return true;
    case 5: launch_pruned_r<D, 5>(c, so, n, rec, U, s, ldu, ind, dist, ks); return true;
    default: return false;
  }
}
template <int D>
void launch_small(Ctx* c, const double* X, int64_t n, int64_t ldx, const double* rec, int s, int r,
                  int32_t* ind, double* dist) {
  switch (r) {
    case 1: launch_small_r<D, 1>(c, X, n, ldx, rec, s, r, ind, dist); break;
    case 2: launch_small_r<D, 2>(c, X, n, ldx, rec, s, r, ind, dist); break;
    case 3: launch_small_r<D, 3>(c, X, n, ldx, rec, s, r, ind, dist); break;
    case 4: launch_small_r<D, 4>(c, X, n, ldx, rec, s, r, ind, dist); break;
    case 5: launch_small_r<D, 5>(c, X, n, ldx, rec, s, r, ind, dist); break;
    default: launch_small_r<D, 0>(c, X, n, ldx, rec, s, r, ind, dist); break;
  }
}
// ---- large d on the tensor cores (distsel.cu): thresholds, exact re-scoring, uncertified rows ------------------
__global__ void knn_maxbits_kernel(const double* __restrict__ v, int n, unsigned long long* out) {
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmax(m, v[i]);
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));  // m >= 0
}
// thr_i = coef * (|x_i|^2 + max_j |u_j|^2): twice the distance between an oracle-order and a tensor-core value
__global__ void knn_thr_kernel(const double* __restrict__ xn, int64_t n, const unsigned long long* __restrict__ unmax,
                               double coef, double* __restrict__ thr) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) thr[i] = coef * (xn[i] + __longlong_as_double((long long)*unmax));
}
// distances of the selected anchors in the reference's own order: ((-2 * sum_k x_k u_k) + |x|^2) + |u|^2
__global__ void knn_rescore_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d,
                                   const double* __restrict__ U, int s, int64_t ldu, const double* __restrict__ xn,
                                   const double* __restrict__ un, int r, const int32_t* __restrict__ ind,
                                   double* __restrict__ dist) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n * r) return;
  const int64_t i = e % n;
  const int j = ind[e];
  if (j < 0 || j >= s) return;
  double dot = 0.0;
  for (int k = 0; k < d; ++k) dot = __dadd_rn(dot, __dmul_rn(X[i + ldx * k], U[j + ldu * k]));
  dist[e] = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), xn[i]), un[j]);
}
// One CTA per uncertified row: all s distances in the reference's order, then the literal heap selection.
// dynamic shared memory: d + s doubles.
__global__ void __launch_bounds__(256)
knn_exact_rows_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, const double* __restrict__ U, int s,
                      int64_t ldu, const double* __restrict__ xn, const double* __restrict__ un, int r,
                      const int* __restrict__ und_count, const int32_t* __restrict__ und_list,
                      int32_t* __restrict__ ind, double* __restrict__ dist) {
  extern __shared__ double sm[];
  double* xs = sm;
  double* Ds = sm + d;
  const int tid = threadIdx.x, cnt = *und_count;
  for (int u = blockIdx.x; u < cnt; u += gridDim.x) {
    const int64_t i = und_list[u];
    __syncthreads();
    for (int k = tid; k < d; k += 256) xs[k] = X[i + ldx * k];
    __syncthreads();
    const double my_xn = xn[i];
    for (int j = tid; j < s; j += 256) {
      double dot = 0.0;
      for (int k = 0; k < d; ++k) dot = __dadd_rn(dot, __dmul_rn(xs[k], U[j + ldu * k]));
      Ds[j] = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), my_xn), un[j]);
    }
    __syncthreads();
    if (tid == 0) {
      TopR h;
      h.r = r;
      h.top = INFINITY;
      for (int j = 0; j < s; ++j) h.push(Ds[j], j);
      h.finish(i, n, ind, dist);
    }
  }
}

}  // namespace

void knn_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, const double* U, int s, int64_t ldu,
             int r, int32_t* ind, double* dist, const KMeansSorted* sorted, bool* out_sorted) {
  if (out_sorted) *out_sorted = false;
  if (r < 1 || r > s) fail(2, "KNN: need 1 <= r <= s (r=%d, s=%d)", r, s);
  if (r > KNN_RMAX) fail(2, "KNN: r=%d exceeds the supported maximum %d", r, KNN_RMAX);
  if (n <= 0) return;
  if (d <= 4) {
    const int str = (d + 2) / 2 * 2;
    DevBuf<double> rec((size_t)s * str);
    FLGP_LAUNCH(c, knn_prep_kernel, ceil_div(s, 128), 128, 0, U, s, ldu, d, str, rec.p);
    if (sorted && sorted->valid && r <= 5 && n < ((int64_t)1 << 31)) {
      bool done = false;
      switch (d) {
        case 1: done = launch_pruned<1>(c, *sorted, n, rec.p, U, s, ldu, r, ind, dist, out_sorted != nullptr); break;
        case 2: done = launch_pruned<2>(c, *sorted, n, rec.p, U, s, ldu, r, ind, dist, out_sorted != nullptr); break;
        case 3: done = launch_pruned<3>(c, *sorted, n, rec.p, U, s, ldu, r, ind, dist, out_sorted != nullptr); break;
        default: done = launch_pruned<4>(c, *sorted, n, rec.p, U, s, ldu, r, ind, dist, out_sorted != nullptr); break;
      }
      if (done) {
        if (out_sorted) *out_sorted = true;
        return;
      }
    }
    switch (d) {
      case 1: launch_small<1>(c, X, n, ldx, rec.p, s, r, ind, dist); break;
      case 2: launch_small<2>(c, X, n, ldx, rec.p, s, r, ind, dist); break;
      case 3: launch_small<3>(c, X, n, ldx, rec.p, s, r, ind, dist); break;
      default: launch_small<4>(c, X, n, ldx, rec.p, s, r, ind, dist); break;
    }
    sync(c);  // rec is freed on return
  } else {
    DevBuf<double> xn(n), un(s);
    FLGP_LAUNCH(c, rownorm_kernel, ceil_div(n, 256), 256, 0, X, n, ldx, d, xn.p);
    FLGP_LAUNCH(c, rownorm_kernel, ceil_div(s, 256), 256, 0, U, (int64_t)s, ldu, d, un.p);
    const size_t xsm = sizeof(double) * ((size_t)d + s);
    if (dist_select_supported(n, s, r) && n < ((int64_t)1 << 31) && xsm <= 200 * 1024 &&
        std::getenv("FLGP_NO_DMMA_DIST") == nullptr) {
      // tensor-core inner products + certified selection (distsel.cu); uncertified rows in the reference's order
      const int dp = (d + 1) / 2 * 2;
      DevBuf<double> Xr((size_t)n * dp), Ur((size_t)s * dp), thr(n);
      DevBuf<unsigned long long> unmax(1);
      DevBuf<int> und_count(1);
      DevBuf<int32_t> und_list(n);
      to_rowmajor_run(c, X, n, ldx, d, dp, Xr.p);
      to_rowmajor_run(c, U, s, ldu, d, dp, Ur.p);
      unmax.zero(c->stream);
      und_count.zero(c->stream);
      FLGP_LAUNCH(c, knn_maxbits_kernel, 8, 256, 0, un.p, s, unmax.p);
      // |oracle value - true| and |tensor value - true| are each below 2^-53 (4d + 10)(|x|^2 + |u|^2): see distsel.cu
      const double coef = 4.0 * (4.0 * d + 10.0) * 1.1102230246251565e-16;
      FLGP_LAUNCH(c, knn_thr_kernel, ceil_div(n, 256), 256, 0, xn.p, n, unmax.p, coef, thr.p);
      dist_select_run(c, Xr.p, n, Ur.p, s, dp, un.p, r, 0.0, thr.p, ind, n, und_count.p, und_list.p);
      if (dist)
        FLGP_LAUNCH(c, knn_rescore_kernel, ceil_div(n * r, 256), 256, 0, X, n, ldx, d, U, s, ldu, xn.p, un.p, r, ind,
                    dist);
      if (xsm > 40 * 1024)
        FLGP_CUDA(cudaFuncSetAttribute(knn_exact_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
      FLGP_LAUNCH(c, knn_exact_rows_kernel, c->sm_count * 4, 256, xsm, X, n, ldx, d, U, s, ldu, xn.p, un.p, r,
                  und_count.p, und_list.p, ind, dist);
      sync(c);
      return;
    }
    FLGP_LAUNCH(c, knn_tiled_kernel, ceil_div(n, NT_TP), 256, 0, X, n, ldx, d, U, s, ldu, xn.p, un.p, r, ind,
                dist);
    sync(c);
  }
}

}  // namespace flgp
