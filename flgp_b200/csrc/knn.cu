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

template <int D, int R, int P>
__global__ void __launch_bounds__(256)
knn_small_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, const double* __restrict__ rec, int s, int r_in,
                 int32_t* __restrict__ ind, double* __restrict__ dist, int chunk, int one) {
  constexpr int STR = (D + 2) / 2 * 2;
  const int r = R ? R : r_in;
  extern __shared__ __align__(16) double srec[];
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * (256 * P);
  double x[P][D], xn[P];
  HeapState<R> hs[P];
  int th[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int64_t i = base + (int64_t)p * 256 + tid;
    xn[p] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      x[p][k] = (i < n) ? X[i + ldx * k] : 0.0;
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
  for (int p = 0; p < P; ++p) {
    const int64_t i = base + (int64_t)p * 256 + tid;
    if (i < n) hs[p].finish(r, i, n, ind, dist);
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
  FLGP_LAUNCH(c, (knn_small_kernel<D, R, P>), grid, 256, smem, X, n, ldx, rec, s, r, ind, dist, chunk, 1);
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

}  // namespace

void knn_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, const double* U, int s, int64_t ldu,
             int r, int32_t* ind, double* dist) {
  if (r < 1 || r > s) fail(2, "KNN: need 1 <= r <= s (r=%d, s=%d)", r, s);
  if (r > KNN_RMAX) fail(2, "KNN: r=%d exceeds the supported maximum %d", r, KNN_RMAX);
  if (n <= 0) return;
  if (d <= 4) {
    const int str = (d + 2) / 2 * 2;
    DevBuf<double> rec((size_t)s * str);
    FLGP_LAUNCH(c, knn_prep_kernel, ceil_div(s, 128), 128, 0, U, s, ldu, d, str, rec.p);
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
    FLGP_LAUNCH(c, knn_tiled_kernel, ceil_div(n, NT_TP), 256, 0, X, n, ldx, d, U, s, ldu, xn.p, un.p, r, ind,
                dist);
    sync(c);
  }
}

}  // namespace flgp
