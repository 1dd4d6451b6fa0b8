// kmeans.cu — Lloyd k-means anchors (replaces the stats::kmeans callback at
// /root/reference/src/Utils.cpp:36-45; contract in oracle/flgp_oracle.cpp orc_kmeans_*).
//
// The contract per iteration: score every (point, centre) pair in fp64 FMA form, arg-min with lowest-index ties,
// compare with the previous assignment, accumulate the centroid sums as two-limb int64 fixed point with integer
// atomics (associative => the sums are bit-identical for any thread schedule and any number of GPUs).  The only
// collective is one int64 all-reduce of 2*s*d + s + 1 words per iteration.
//
// How the passes are run (all of it returns the brute-force result bit for bit):
//   d <= 4   pass 1: brute force (kmeans_assign_small), through 128 pivot centres when s >= 512 (kmeans_assign_listed);
//            passes 2..: rows sorted by cluster, Hamerly bounds, ONE kernel per pass (kmeans_lists_kernel ->
//            kmeans_pass_fused -> kmeans_update_kernel), persistent integer sums (a reassignment is -x / +x);
//   d > 4    distances on the FP64 tensor cores with a certified selection (distsel.cu), the uncertified rows in the
//            oracle's order, Hamerly bounds around it (kmeans_hbounds / kmeans_hcommit).
//
// Roofline: brute force = FP64 FMA pipe, 2*s*d flop per point per iteration, bytes 8d + 4 per point;
// the pruned pass = HBM (20 B of bounds and centre numbers per point streamed, one 32-byte record per survivor).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

namespace flgp {

namespace {

constexpr int KM_THREADS = 256;

__global__ void maxabs_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d,
                              unsigned long long* out) {
  double m = 0.0;
  const int64_t total = n * d;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = e / n, i = e - k * n;
    double a = fabs(X[i + ldx * k]);
    if (a > m) m = a;
  }
  for (int o = 16; o; o >>= 1) {
    double t = __shfl_xor_sync(0xffffffffu, m, o);
    if (t > m) m = t;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));  // m >= 0
}

// centres -> records [ -2c_0 .. -2c_{d-1}, |c|^2, pad ]  (stride STR = even(d+1))
__global__ void kmeans_prep_kernel(const double* __restrict__ C, int s, int d, int str, double M, double* rec) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s) return;
  double a = 0.0;
  for (int k = 0; k < d; ++k) {
    double c = C[j + (size_t)s * k];
    a = fma(c, c, a);
    rec[(size_t)j * str + k] = -2.0 * c;
  }
  rec[(size_t)j * str + d] = __dadd_rn(a, M);  // M keeps every score positive (see the contract in the oracle)
  for (int k = d + 1; k < str; ++k) rec[(size_t)j * str + k] = 0.0;
}

// split records into the tiled kernel's operands: C2 = -2 C (s x d col-major, element-wise) and
// cn[j] = (sum_k fma(c,c,.)) + M (one thread per centre, ascending k as the oracle; loads batched 8 deep so that the
// d-long chain waits on arithmetic, not on memory)
__global__ void kmeans_prep_c2_kernel(const double* __restrict__ C, size_t len, double* __restrict__ C2) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e < len) C2[e] = -2.0 * C[e];
}
__global__ void kmeans_prep_cn_kernel(const double* __restrict__ C, int s, int d, double M, double* __restrict__ cn) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s) return;
  double a = 0.0;
  int k = 0;
  for (; k + 8 <= d; k += 8) {
    double c[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q] = C[j + (size_t)s * (k + q)];
#pragma unroll
    for (int q = 0; q < 8; ++q) a = fma(c[q], c[q], a);
  }
  for (; k < d; ++k) {
    const double c = C[j + (size_t)s * k];
    a = fma(c, c, a);
  }
  cn[j] = __dadd_rn(a, M);
}

// upper bound on the true distance |x - c_j| from a computed score (score = |x-c|^2 - |x|^2 + M up to rounding)
__device__ __forceinline__ double km_ub(double score, double M, double xn, double delta2) {
  return sqrt(fmax((score - M) + xn, 0.0) + delta2) * (1.0 + 1e-14);
}
// The same two bounds through a single-precision square root (a handful of instructions instead of ~35): they only
// steer the pruning, so being 2^-21 looser costs nothing; outside the float range the fp64 root is used.
__device__ __forceinline__ double km_ub_fast(double score, double M, double xn, double delta2) {
  const double v = fmax((score - M) + xn, 0.0) + delta2;
  if (!(v > 1e-30 && v < 1e30)) return sqrt(v) * (1.0 + 1e-14);
  return (double)sqrtf((float)v) * (1.0 + 4.8e-7);  // float conversion and root: 2 half-ulp errors of 6e-8 each
}
__device__ __forceinline__ double km_lb_fast(double score, double M, double xn, double delta2) {
  const double v = fmax(((score - M) + xn) - delta2, 0.0);
  if (!(v > 1e-30 && v < 1e30)) return (v < 1e30 || v != v) ? 0.0 : sqrt(v) * (1.0 - 1e-14);
  return (double)sqrtf((float)v) * (1.0 - 4.8e-7);
}
__device__ __forceinline__ void km_radius(unsigned long long* Rbits, int a, double ub) {
  // non-negative doubles order like their bits.  The radius usually covers the point already (a possibly stale,
  // i.e. smaller, cached value only costs a redundant atomic): one atomic per point on 2000 addresses is the
  // difference between a compute-bound and an atomics-bound pass.
  const unsigned long long b = (unsigned long long)__double_as_longlong(ub);
  if (b > Rbits[a]) atomicMax(&Rbits[a], b);
}

__device__ __forceinline__ void km_commit(const Fx& fx, int s, int d, int bj, const double* x, int32_t* assign,
                                          int64_t i, unsigned long long* acc, int& changed) {
  if (assign[i] != bj) ++changed;
  assign[i] = bj;
  if (!acc) return;  // pre-pass against the pivot centres: only the label and the radius are wanted
  for (int k = 0; k < d; ++k) {
    long long h, l;
    fx_encode(fx, x[k], &h, &l);
    atomicAdd(&acc[bj + (size_t)s * k], (unsigned long long)h);
    atomicAdd(&acc[(size_t)s * d + bj + (size_t)s * k], (unsigned long long)l);
  }
  atomicAdd(&acc[(size_t)2 * s * d + bj], 1ull);
}

// ---- small d: thread-per-point(s), centres broadcast from shared memory -------------------------
template <int D, int P>
__global__ void __launch_bounds__(KM_THREADS)
kmeans_assign_small(const double* __restrict__ X, int64_t n, int64_t ldx, const double* __restrict__ rec, int s,
                    Fx fx, int32_t* __restrict__ assign, unsigned long long* __restrict__ acc, int chunk, int one,
                    unsigned long long* __restrict__ Rbits, double M, double delta2) {
  constexpr int STR = (D + 2) / 2 * 2;
  extern __shared__ __align__(16) double srec[];
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * (KM_THREADS * P);
  double x[P][D], best[P];
  int bj[P], bh[P];  // bh = high word of best: scores are positive, so they order like their high words
#pragma unroll
  for (int p = 0; p < P; ++p) {
    int64_t i = base + (int64_t)p * KM_THREADS + tid;
#pragma unroll
    for (int k = 0; k < D; ++k) x[p][k] = (i < n) ? X[i + ldx * k] : 0.0;
    // seed with the centre this point had last iteration (centre 0 on the first): the running best is
    // then almost always final already, so the exact-compare path below is hardly ever entered
    int seed = (i < n) ? assign[i] : 0;
    if (seed < 0) seed = 0;
    const double* rs = rec + (size_t)seed * STR;
    double e = rs[D];
#pragma unroll
    for (int k = 0; k < D; ++k) e = fma(x[p][k], rs[k], e);
    best[p] = e;
    bh[p] = __double2hiint(e);
    bj[p] = seed;
  }
  for (int c0 = 0; c0 < s; c0 += chunk) {
    const int cnt = min(chunk, s - c0);
    __syncthreads();
    for (int t = tid; t < cnt * STR; t += KM_THREADS) srec[t] = rec[(size_t)c0 * STR + t];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      double cr[STR];
      const double2* rj = reinterpret_cast<const double2*>(srec + (size_t)j * STR);
#pragma unroll
      for (int q = 0; q < STR / 2; ++q) {
        double2 t = rj[q];
        cr[2 * q] = t.x;
        cr[2 * q + 1] = t.y;
      }
      double e[P];
      bool cand = false;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        e[p] = cr[D];
#pragma unroll
        for (int k = 0; k < D; ++k) e[p] = fma(x[p][k], cr[k], e[p]);
        // integer pre-filter on the ALU pipe: for positive doubles a < b implies hi(a) <= hi(b) (a
        // non-positive score has a negative high word and always passes); the fp64 pipe keeps the FMAs
        cand |= (__double2hiint(e[p]) <= bh[p]);
      }
      if (cand) {
#pragma unroll 1
        for (int q = 0; q < one; ++q) {  // `one` == 1: a loop cannot be if-converted => one real, rare branch
          const int jj = c0 + j;
#pragma unroll
          for (int p = 0; p < P; ++p)
            if (e[p] < best[p] || (e[p] == best[p] && jj < bj[p])) {  // lowest index among equal minima
              best[p] = e[p];
              bh[p] = __double2hiint(e[p]);
              bj[p] = jj;
            }
        }
      }
    }
  }
  int changed = 0;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    int64_t i = base + (int64_t)p * KM_THREADS + tid;
    if (i < n) {
      km_commit(fx, s, D, bj[p], x[p], assign, i, acc, changed);
      if (Rbits) {  // cluster radius for the pruned passes that follow
        double xn = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) xn = fma(x[p][k], x[p][k], xn);
        km_radius(Rbits, bj[p], km_ub(best[p], M, xn, delta2));
      }
    }
  }
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if ((tid & 31) == 0 && changed && acc) atomicAdd(&acc[(size_t)2 * s * D + s], (unsigned long long)changed);
}

// ---- first pass through pivots (small d) ---------------------------------------------------------------------------
// A brute-force pass costs 2 s d flop per point.  Instead: (1) label every point with the nearest of np << s PIVOT
// centres (a subset of the centres; the kernel above with acc == nullptr) and record each pivot group's radius R_p;
// (2) sort the points by pivot; (3) per pivot list the centres within 2 R_p + eta of the pivot centre: a member at
// distance ub <= R_p from the pivot centre has its nearest centre within ub, hence within 2 ub of the pivot centre,
// and every unlisted centre is at least ub + eta away, so its COMPUTED score is strictly larger than the pivot
// centre's (eta^2 > 2 Delta) - it can neither win nor tie; (4) run the brute-force inner loop of the kernel above
// over that list only (kmeans_assign_listed).  The arg-min, its lowest-index tie rule, the sums and the radii are
// exactly those of the full scan.
template <int D, int P>
__global__ void __launch_bounds__(KM_THREADS)
kmeans_assign_listed(const double4* __restrict__ Xs4, const int32_t* __restrict__ perm, const int* __restrict__ items,
                     const int* __restrict__ n_items, const double* __restrict__ rec, const int32_t* __restrict__ list_j,
                     const int32_t* __restrict__ len, int lmax, int s, Fx fx, int32_t* __restrict__ assign,
                     unsigned long long* __restrict__ acc, unsigned long long* __restrict__ Rbits, double M,
                     double delta2) {
  constexpr int STR = (D + 2) / 2 * 2;
  extern __shared__ __align__(16) double srec[];  // lmax records, then lmax indices
  if ((int)blockIdx.x >= *n_items) return;
  const int tid = threadIdx.x;
  const int pv = items[3 * blockIdx.x], beg = items[3 * blockIdx.x + 1], end = items[3 * blockIdx.x + 2];
  int* sj = reinterpret_cast<int*>(srec + (size_t)lmax * STR);
  const int L0 = len[pv];
  const int L = (L0 < 0) ? 0 : L0;
  for (int q = tid; q < L; q += KM_THREADS) sj[q] = list_j[(size_t)pv * lmax + q];
  __syncthreads();
  for (int t = tid; t < L * STR; t += KM_THREADS) srec[t] = rec[(size_t)sj[t / STR] * STR + (t % STR)];
  __syncthreads();
  double x[P][D], best[P];
  int bj[P], bh[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int q = beg + p * KM_THREADS + tid;
    double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
    if (q < end) v = Xs4[q];
    const double xa[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < D; ++k) x[p][k] = xa[k];
    best[p] = INFINITY;
    bh[p] = 0x7fffffff;
    bj[p] = 0x7fffffff;
  }
  auto scan = [&](const double* cr, int jj) {
    double e[P];
    bool cand = false;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      e[p] = cr[D];
#pragma unroll
      for (int k = 0; k < D; ++k) e[p] = fma(x[p][k], cr[k], e[p]);
      cand |= (__double2hiint(e[p]) <= bh[p]);  // integer pre-filter: scores are positive (see kmeans_assign_small)
    }
    if (cand) {
#pragma unroll
      for (int p = 0; p < P; ++p)
        if (e[p] < best[p] || (e[p] == best[p] && jj < bj[p])) {  // lowest index among equal minima, any scan order
          best[p] = e[p];
          bh[p] = __double2hiint(e[p]);
          bj[p] = jj;
        }
    }
  };
  if (L0 >= 0) {
#pragma unroll 2
    for (int j = 0; j < L; ++j) {
      double cr[STR];
      const double2* rj = reinterpret_cast<const double2*>(srec + (size_t)j * STR);
#pragma unroll
      for (int q = 0; q < STR / 2; ++q) {
        double2 t = rj[q];
        cr[2 * q] = t.x;
        cr[2 * q + 1] = t.y;
      }
      scan(cr, sj[j]);
    }
  } else {  // list overflow: this pivot's members scan every centre (records from global memory; rare)
    for (int j = 0; j < s; ++j) {
      double cr[STR];
      const double2* rj = reinterpret_cast<const double2*>(rec + (size_t)j * STR);
#pragma unroll
      for (int q = 0; q < STR / 2; ++q) {
        double2 t = rj[q];
        cr[2 * q] = t.x;
        cr[2 * q + 1] = t.y;
      }
      scan(cr, j);
    }
  }
  int changed = 0;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int q = beg + p * KM_THREADS + tid;
    if (q < end) {
      assign[perm[q]] = -1;  // the label written by the pre-pass was a pivot number
      km_commit(fx, s, D, bj[p], x[p], assign, perm[q], acc, changed);
      double xn = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) xn = fma(x[p][k], x[p][k], xn);
      km_radius(Rbits, bj[p], km_ub(best[p], M, xn, delta2));
    }
  }
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if ((tid & 31) == 0 && changed) atomicAdd(&acc[(size_t)2 * s * D + s], (unsigned long long)changed);
}

// per pivot: every centre within 2 R_p + eta of the pivot centre (unsorted; the pivot centre itself included)
__global__ void __launch_bounds__(256)
kmeans_pivot_lists_kernel(const double* __restrict__ C, int s, int d, const int32_t* __restrict__ pivots,
                          const unsigned long long* __restrict__ Rp, double eta, int lmax, int32_t* __restrict__ list_j,
                          int32_t* __restrict__ len) {
  __shared__ int count;
  const int pv = blockIdx.x, tid = threadIdx.x, a = pivots[pv];
  if (tid == 0) count = 0;
  __syncthreads();
  const double thr = (2.0 * (__longlong_as_double((long long)Rp[pv]) + eta) + eta) * (1.0 + 1e-9);
  for (int j = tid; j < s; j += 256) {
    double cc = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = C[a + (size_t)s * k] - C[j + (size_t)s * k];
      cc = fma(df, df, cc);
    }
    if (sqrt(cc) < thr) {
      const int pos = atomicAdd(&count, 1);
      if (pos < lmax) list_j[(size_t)pv * lmax + pos] = j;
    }
  }
  __syncthreads();
  if (tid == 0) len[pv] = (count <= lmax) ? count : -1;
}

// pivot records = the centre records of the chosen centres
__global__ void kmeans_pivot_rec_kernel(const double* __restrict__ rec, int str, const int32_t* __restrict__ pivots, int np,
                                        double* __restrict__ prec) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < np * str) prec[e] = rec[(size_t)pivots[e / str] * str + (e % str)];
}

__global__ void kmeans_hist_kernel(const int32_t* __restrict__ label, int64_t n, int nb, long long* __restrict__ cnt) {
  extern __shared__ int hs[];
  for (int t = threadIdx.x; t < nb; t += blockDim.x) hs[t] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&hs[label[i]], 1);
  __syncthreads();
  for (int t = threadIdx.x; t < nb; t += blockDim.x)
    if (hs[t]) atomicAdd(reinterpret_cast<unsigned long long*>(&cnt[t]), (unsigned long long)hs[t]);
}

// segment offsets and work items of at most chunk points (item = segment, first, last + 1); one thread per segment
__global__ void kmeans_items_kernel(const long long* __restrict__ cnt, int nb, int chunk, int* __restrict__ cursor,
                                    int* __restrict__ items, int* __restrict__ n_items) {
  extern __shared__ int sh[];  // nb segment starts, nb item starts
  int* seg0 = sh;
  int* it0 = sh + nb;
  if (threadIdx.x == 0) {
    int run = 0, ni = 0;
    for (int j = 0; j < nb; ++j) {
      seg0[j] = run;
      it0[j] = ni;
      run += (int)cnt[j];
      ni += ((int)cnt[j] + chunk - 1) / chunk;
    }
    *n_items = ni;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    cursor[j] = seg0[j];
    const int end = seg0[j] + (int)cnt[j];
    int ni = it0[j];
    for (int b = seg0[j]; b < end; b += chunk, ++ni) {
      items[3 * ni] = j;
      items[3 * ni + 1] = b;
      items[3 * ni + 2] = min(end, b + chunk);
    }
  }
}

// counting-sort scatter for FEW keys (the pivot groups): a CTA counts its 2048 points per key in shared memory,
// reserves one range per key with one global atomic each, and places its points inside those ranges
constexpr int KM_FK_PTS = 8;
__global__ void __launch_bounds__(256)
kmeans_scatter_fewkeys_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, const int32_t* __restrict__ key,
                              int nkeys, int* __restrict__ cursor, double4* __restrict__ Xdst,
                              int32_t* __restrict__ dst_perm) {
  extern __shared__ int sh[];  // nkeys counts, nkeys bases
  int* cntk = sh;
  int* base = sh + nkeys;
  const int tid = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * (256 * KM_FK_PTS);
  for (int t = tid; t < nkeys; t += 256) cntk[t] = 0;
  __syncthreads();
  int k[KM_FK_PTS], rank[KM_FK_PTS];
#pragma unroll
  for (int q = 0; q < KM_FK_PTS; ++q) {
    const int64_t i = c0 + q * 256 + tid;
    k[q] = (i < n) ? key[i] : -1;
    if (k[q] >= 0) rank[q] = atomicAdd(&cntk[k[q]], 1);
  }
  __syncthreads();
  for (int t = tid; t < nkeys; t += 256) base[t] = cntk[t] ? atomicAdd(&cursor[t], cntk[t]) : 0;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < KM_FK_PTS; ++q) {
    const int64_t i = c0 + q * 256 + tid;
    if (k[q] < 0) continue;
    const int pos = base[k[q]] + rank[q];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int kk = 0; kk < d; ++kk) v[kk] = X[i + ldx * kk];
    Xdst[pos] = make_double4(v[0], v[1], v[2], v[3]);
    dst_perm[pos] = (int32_t)i;
  }
}

// ---- any d: 64 points x 64 centres register-tiled FMA kernel ------------------------------------
constexpr int KT_TP = 64, KT_TC = 64, KT_K = 16;

__global__ void __launch_bounds__(256)
kmeans_assign_tiled(const double* __restrict__ X, int64_t n, int64_t ldx, int d, const double* __restrict__ C2,
                    const double* __restrict__ cn, int s, Fx fx, int32_t* __restrict__ assign,
                    unsigned long long* __restrict__ acc) {
  __shared__ double Xs[KT_K][KT_TP];
  __shared__ double Cs[KT_K][KT_TC];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * KT_TP;
  double best[4];
  int bj[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    best[a] = INFINITY;
    bj[a] = 0;
  }
  for (int c0 = 0; c0 < s; c0 += KT_TC) {
    double av[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int j = c0 + tx * 4 + b;
      double init = (j < s) ? cn[j] : INFINITY;
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a][b] = init;
    }
    for (int k0 = 0; k0 < d; k0 += KT_K) {
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int e = tid + q * 256, kk = e >> 6, p = e & 63;
        int k = k0 + kk;
        int64_t i = i0 + p;
        Xs[kk][p] = (k < d && i < n) ? X[i + ldx * k] : 0.0;
        int j = c0 + p;
        Cs[kk][p] = (k < d && j < s) ? C2[j + (size_t)s * k] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KT_K; ++kk) {  // zero padding: fma(0,0,e) == e
        double xa[4], cb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) xa[a] = Xs[kk][ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) cb[b] = Cs[kk][tx * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) av[a][b] = fma(xa[a], cb[b], av[a][b]);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double lv = INFINITY;
      int lj = 0x7fffffff;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        int j = c0 + tx * 4 + b;
        if (j < s && av[a][b] < lv) {
          lv = av[a][b];
          lj = j;
        }
      }
      // lowest value, then lowest index, across the 16 lanes that share this point
#pragma unroll
      for (int o = 8; o; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, lv, o);
        int oj = __shfl_xor_sync(0xffffffffu, lj, o);
        if (ov < lv || (ov == lv && oj < lj)) {
          lv = ov;
          lj = oj;
        }
      }
      if (lv < best[a]) {
        best[a] = lv;
        bj[a] = lj;
      }
    }
  }
  int changed = 0;
  if (tx == 0) {
    for (int a = 0; a < 4; ++a) {
      int64_t i = i0 + ty * 4 + a;
      if (i >= n) continue;
      if (assign[i] != bj[a]) ++changed;
      assign[i] = bj[a];
      for (int k = 0; k < d; ++k) {
        long long h, l;
        fx_encode(fx, X[i + ldx * k], &h, &l);
        atomicAdd(&acc[bj[a] + (size_t)s * k], (unsigned long long)h);
        atomicAdd(&acc[(size_t)s * d + bj[a] + (size_t)s * k], (unsigned long long)l);
      }
      atomicAdd(&acc[(size_t)2 * s * d + bj[a]], 1ull);
    }
  }
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if ((tid & 31) == 0 && changed) atomicAdd(&acc[(size_t)2 * s * d + s], (unsigned long long)changed);
}

// ---- large d, passes 2..: Hamerly bounds around the tensor-core evaluation -----------------------------------------
// Every point carries u >= |x - c_a| and l <= |x - c_j| for all j != a.  After the centres move: u += move[a],
// l -= (largest move of any OTHER centre).  u + eta <= l  =>  the oracle's computed scores keep c_a strictly first
// (eta^2 > 2 Delta, as in the small-d path), so the point is skipped without touching its coordinates; the rest first
// tighten u with one exact score (d flops), and only what still fails is gathered and re-evaluated against all centres
// by dist_select_kernel, which also returns the two smallest values for the new bounds.
__global__ void kmeans_xnorm_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, double* __restrict__ xn) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0;
  for (int k = 0; k < d; ++k) {
    const double x = X[i + ldx * k];
    a = fma(x, x, a);
  }
  xn[i] = a;
}
// move[j] >= |c_new - c_old|, one warp per centre
__global__ void kmeans_move_kernel(const double* __restrict__ C, const double* __restrict__ Cold, int s, int d,
                                   double* __restrict__ move) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= s) return;
  double a = 0.0;
  for (int k = lane; k < d; k += 32) {
    const double df = C[j + (size_t)s * k] - Cold[j + (size_t)s * k];
    a = fma(df, df, a);
  }
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) move[j] = sqrt(a) * (1.0 + 1e-12) * (1.0 + 1e-12);  // summation-order slack
}
// mm = {largest move, second largest move, index of the largest}
__global__ void kmeans_move_max_kernel(const double* __restrict__ move, int s, double* __restrict__ mm,
                                       unsigned long long* __restrict__ maxbits) {
  __shared__ double v1[256], v2[256];
  __shared__ int i1[256];
  const int tid = threadIdx.x;
  double a = -1.0, b = -1.0;
  int ia = -1;
  for (int j = tid; j < s; j += 256) {
    const double m = move[j];
    if (m > a) {
      b = a;
      a = m;
      ia = j;
    } else if (m > b) {
      b = m;
    }
  }
  v1[tid] = a;
  v2[tid] = b;
  i1[tid] = ia;
  __syncthreads();
  if (tid == 0) {
    for (int t = 1; t < 256; ++t) {
      if (v1[t] > a) {
        b = fmax(a, v2[t]);
        a = v1[t];
        ia = i1[t];
      } else {
        b = fmax(b, v1[t]);
      }
    }
    mm[0] = fmax(a, 0.0);
    mm[1] = fmax(b, 0.0);
    mm[2] = (double)ia;
    *maxbits = (unsigned long long)__double_as_longlong(fmax(a, 0.0));
  }
}
__global__ void __launch_bounds__(256)
kmeans_hbounds_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, const double* __restrict__ C2,
                      const double* __restrict__ cn, int s, const int32_t* __restrict__ assign,
                      const double4* __restrict__ cl, double eta, double M,
                      double delta2, const double* __restrict__ xn, double* __restrict__ UB, double* __restrict__ LB,
                      int32_t* __restrict__ surv, int* __restrict__ nsurv) {
  const int lane = threadIdx.x & 31;
  for (int64_t i0 = (int64_t)blockIdx.x * 256; i0 < n; i0 += (int64_t)gridDim.x * 256) {  // uniform per warp
    const int64_t i = i0 + threadIdx.x;
    bool need = false;
    if (i < n) {
      const int a = assign[i];
      // cl[a] = {move of c_a, largest move among the centres of a's neighbour list, radius the list is complete to}
      const double4 ca = cl[a];
      const double u = (UB[i] + ca.x) * (1.0 + 1e-15);
      const double l = fmin((LB[i] - ca.y) * (1.0 - 1e-15) - 1e-300, (ca.z - u * (1.0 + 1e-15)) * (1.0 - 1e-15));
      if (u + eta <= l) {
        UB[i] = u;
        LB[i] = l;
      } else {
        double e = cn[a];
        for (int k = 0; k < d; ++k) e = fma(X[i + ldx * k], C2[a + (size_t)s * k], e);
        const double ub0 = km_ub_fast(e, M, xn[i], delta2);
        if (ub0 + eta <= l) {
          UB[i] = ub0;
          LB[i] = l;
        } else {
          need = true;  // the evaluation rewrites both bounds
        }
      }
    }
    const unsigned mw = __ballot_sync(0xffffffffu, need);
    if (mw) {
      const int lead = __ffs(mw) - 1;
      int base = 0;
      if (lane == lead) base = atomicAdd(nsurv, __popc(mw));
      base = __shfl_sync(0xffffffffu, base, lead);
      if (need) surv[base + __popc(mw & ((1u << lane) - 1))] = (int32_t)i;
    }
  }
}
// survivors' rows, row-major dp pitch, one warp per row
__global__ void kmeans_gather_rows_kernel(const double* __restrict__ Xr, int dp, const int32_t* __restrict__ surv,
                                          const int* __restrict__ nsurv, double* __restrict__ Xs) {
  const int lane = threadIdx.x & 31, cnt = *nsurv;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < cnt; q += nw) {
    const double* src = Xr + (int64_t)surv[q] * dp;
    for (int k = lane; k < dp; k += 32) Xs[q * dp + k] = src[k];
  }
}
// rows the selection could not certify (positions in the survivor list): every score in the oracle's order; best
// (lowest index on ties) and the runner-up value.  dynamic shared memory: d doubles.
__global__ void __launch_bounds__(256)
kmeans_exact_rows2_kernel(const double* __restrict__ X, int64_t ldx, int d, const double* __restrict__ C2,
                          const double* __restrict__ cn, int s, const int* __restrict__ und_count,
                          const int32_t* __restrict__ und_list, const int32_t* __restrict__ surv,
                          int32_t* __restrict__ sidx, double* __restrict__ svals) {
  extern __shared__ double xs[];
  __shared__ double rv[8], rs2[8];
  __shared__ int rj[8];
  const int tid = threadIdx.x, cnt = *und_count;
  for (int u = blockIdx.x; u < cnt; u += gridDim.x) {
    const int q = und_list[u];
    const int64_t i = surv ? surv[q] : q;
    __syncthreads();
    for (int k = tid; k < d; k += 256) xs[k] = X[i + ldx * k];
    __syncthreads();
    double lv = INFINITY, ls = INFINITY;
    int lj = 0x7fffffff;
    for (int j = tid; j < s; j += 256) {
      double e = cn[j];
      for (int k = 0; k < d; ++k) e = fma(xs[k], C2[j + (size_t)s * k], e);
      if (e < lv) {
        ls = lv;
        lv = e;
        lj = j;
      } else if (e < ls) {
        ls = e;
      }
    }
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, lv, o), os = __shfl_xor_sync(0xffffffffu, ls, o);
      const int oj = __shfl_xor_sync(0xffffffffu, lj, o);
      if (ov < lv || (ov == lv && oj < lj)) {
        ls = fmin(lv, os);
        lv = ov;
        lj = oj;
      } else {
        ls = fmin(ls, ov);
      }
    }
    if ((tid & 31) == 0) {
      rv[tid >> 5] = lv;
      rs2[tid >> 5] = ls;
      rj[tid >> 5] = lj;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w) {
        if (rv[w] < lv || (rv[w] == lv && rj[w] < lj)) {
          ls = fmin(lv, rs2[w]);
          lv = rv[w];
          lj = rj[w];
        } else {
          ls = fmin(ls, rv[w]);
        }
      }
      sidx[q] = (lj < s) ? lj : 0;
      svals[2 * (size_t)q] = lv;
      svals[2 * (size_t)q + 1] = ls;
    }
  }
}
// evaluated rows: new bounds from the two smallest values; a changed assignment is an exact -enc(x) / +enc(x).
// One warp per evaluated row q (row surv[q] of the matrix; Xrows holds row q at pitch dp).
__global__ void __launch_bounds__(256)
kmeans_hcommit_kernel(const double* __restrict__ Xrows, int d, int dp, const int32_t* __restrict__ surv,
                      const int* __restrict__ nsurv, const int32_t* __restrict__ sidx, const double* __restrict__ svals,
                      const double* __restrict__ xn, int32_t* __restrict__ assign, int s, Fx fx,
                      unsigned long long* __restrict__ acc, double M, double delta2, double* __restrict__ UB,
                      double* __restrict__ LB, unsigned long long* __restrict__ Rcur) {
  const int lane = threadIdx.x & 31, cnt = *nsurv;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int changed = 0;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < cnt; q += nw) {
    const int64_t i = surv ? surv[q] : q;
    const int a_old = assign[i], a_new = sidx[q];
    if (lane == 0) {
      const double x2 = xn[i];
      const double ub = km_ub_fast(svals[2 * q], M, x2, delta2);
      UB[i] = ub;
      LB[i] = km_lb_fast(svals[2 * q + 1], M, x2, delta2);
      if (Rcur) km_radius(Rcur, a_new, ub);  // radius of the (new) cluster for the next pass's neighbour lists
    }
    if (a_old == a_new) continue;
    for (int k = lane; k < d; k += 32) {
      long long h, l;
      fx_encode(fx, Xrows[q * dp + k], &h, &l);
      atomicAdd(&acc[a_new + (size_t)s * k], (unsigned long long)h);
      atomicAdd(&acc[(size_t)s * d + a_new + (size_t)s * k], (unsigned long long)l);
      if (a_old >= 0) {
        atomicAdd(&acc[a_old + (size_t)s * k], (unsigned long long)(-h));
        atomicAdd(&acc[(size_t)s * d + a_old + (size_t)s * k], (unsigned long long)(-l));
      }
    }
    if (lane == 0) {
      atomicAdd(&acc[(size_t)2 * s * d + a_new], 1ull);
      if (a_old >= 0) atomicAdd(&acc[(size_t)2 * s * d + a_old], ~0ull);
      assign[i] = a_new;
      ++changed;
    }
  }
  if (changed) atomicAdd(&acc[(size_t)2 * s * d + s], (unsigned long long)changed);
}

// centroid update, one thread per centre; move[j] >= |c_new - c_old| (for the pruned passes' radius bound)
// kstate[0]: first pass in which no assignment changed (0 = none yet); kstate[1]: assignments changed so far.
// The host reads them every few passes only: a pass after convergence changes nothing (same sums, same centres), so
// running ahead of the check is harmless and the reported iteration count is still the exact one.
__global__ void kmeans_update_kernel(const long long* __restrict__ acc, int s, int d, Fx fx, double* C, double* sizes,
                                     double* move, unsigned long long* maxmove_bits, int it, long long* kstate) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) {
    const long long ch = acc[(size_t)2 * s * d + s];
    if (ch == 0 && kstate[0] == 0) kstate[0] = it;
    kstate[1] += ch;
  }
  if (j >= s) return;
  long long cnt = acc[(size_t)2 * s * d + j];
  sizes[j] = (double)cnt;
  double mv = 0.0;
  if (cnt != 0) {  // an empty cluster keeps its centre
    for (int k = 0; k < d; ++k) {
      double sum = fx_decode(fx, acc[j + (size_t)s * k], acc[(size_t)s * d + j + (size_t)s * k]);
      double cnew = sum / (double)cnt;
      double df = cnew - C[j + (size_t)s * k];
      mv = fma(df, df, mv);
      C[j + (size_t)s * k] = cnew;
    }
  }
  if (move) {
    const double mj = sqrt(mv) * (1.0 + 1e-12);
    move[j] = mj;
    atomicMax(maxmove_bits, (unsigned long long)__double_as_longlong(mj));  // non-negative doubles order like their bits
  }
}

// the same for large d, one thread per (centre, coordinate); no displacement bookkeeping (brute-force passes only)
__global__ void kmeans_update_wide_kernel(const long long* __restrict__ acc, int s, int d, Fx fx, double* C,
                                          double* sizes, int it, long long* kstate) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e == 0) {
    const long long ch = acc[(size_t)2 * s * d + s];
    if (ch == 0 && kstate[0] == 0) kstate[0] = it;
    kstate[1] += ch;
  }
  if (e >= (size_t)s * d) return;
  const int j = (int)(e % s);
  const long long cnt = acc[(size_t)2 * s * d + j];
  if (e < (size_t)s) sizes[j] = (double)cnt;
  if (cnt != 0) C[e] = fx_decode(fx, acc[e], acc[(size_t)s * d + e]) / (double)cnt;  // an empty cluster keeps its centre
}

// rows of X that this rank owns -> bit patterns in the (zeroed) centre buffer
__global__ void kmeans_init_kernel(const double* __restrict__ X, int64_t n_local, int64_t ldx, int d, int s,
                                   int64_t row_offset, const int32_t* __restrict__ init_idx, long long* Cbits) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s * d) return;
  int j = e % s, k = e / s;
  int64_t i = (int64_t)init_idx[j] - row_offset;
  if (i >= 0 && i < n_local) Cbits[j + (size_t)s * k] = __double_as_longlong(X[i + ldx * k]);
}

// ---- exact pruned passes on cluster-sorted points (small d) -------------------------------------------
// After the first brute-force pass every point knows a centre a (its previous assignment).  With
// ub >= |x - c_a| (true distance), a centre j with |c_a - c_j| >= 2 ub + eta satisfies
//   |x - c_j| >= |c_a - c_j| - |x - c_a| >= |x - c_a| + eta   =>   score_j - score_a >= eta^2 > 2 Delta,
// where Delta bounds the rounding error of a computed score, so the COMPUTED score of j is strictly larger
// than that of a: j can neither win nor tie.  Scanning {a} + {j : |c_a - c_j| < 2 ub + eta} therefore returns
// exactly the arg-min of the brute-force scan, including its lowest-index tie rule.
//   * per centre: candidates within 2 R_a + eta, SORTED by centre-centre distance (R_a >= every member's ub:
//     max ub of the previous pass + how far the centre has moved since);
//   * per point: walk that list until the centre-centre distance exceeds 2 ub_i + eta (its own bound);
//   * points are kept sorted by cluster (a permutation + a gathered copy of X), so a warp walks ONE list and
//     every list / record load is a broadcast; the order is refreshed when enough points have moved;
//   * sums are integer limbs, so a reassignment is an exact -x / +x on persistent accumulators.
constexpr int KM_NPIVOT = 128;        // pivot centres of the first pass
constexpr int KM_PIVOT_LMAX = 1024;   // longest candidate list of a pivot group
constexpr int KM_PIVOT_MIN_S = 512;   // below this many centres the plain full scan is as fast
constexpr int KM_CHECK = 8;   // passes between two host checks of the convergence flag
constexpr int KM_LMAX = 512;  // longest neighbour list; longer => that cluster's members scan everything

__global__ void __launch_bounds__(256)
kmeans_lists_kernel(const double* __restrict__ C, int s, int d, const unsigned long long* __restrict__ Rprev,
                    const double* __restrict__ move, double eta, int32_t* __restrict__ list_j,
                    double* __restrict__ list_cc, int32_t* __restrict__ len, double* __restrict__ lthr,
                    double4* __restrict__ cl, const unsigned long long* __restrict__ maxmove_bits,
                    unsigned long long* __restrict__ Rcur, float4* __restrict__ clf = nullptr,
                    double* __restrict__ rec = nullptr, int str = 0, double M = 0.0,
                    long long* __restrict__ zero_changed = nullptr, int* __restrict__ zero_count = nullptr) {
  __shared__ double kcc[KM_LMAX];
  __shared__ int kj[KM_LMAX];
  __shared__ int count;
  const int a = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) count = 0;
  // folded into this launch (a pruned pass is ~10 short launches otherwise; at 8 GPUs their host cost shows): the
  // centre's score record (kmeans_prep_kernel) and the zeroing of the pass's counters
  if (rec && tid == 32) {
    double acc2 = 0.0;
    for (int k = 0; k < d; ++k) {
      const double cc = C[a + (size_t)s * k];
      acc2 = fma(cc, cc, acc2);
      rec[(size_t)a * str + k] = -2.0 * cc;
    }
    rec[(size_t)a * str + d] = __dadd_rn(acc2, M);
    for (int k = d + 1; k < str; ++k) rec[(size_t)a * str + k] = 0.0;
  }
  if (a == 0 && tid == 64) {
    if (zero_changed) *zero_changed = 0;
    if (zero_count) *zero_count = 0;
  }
  __syncthreads();
  const double Ra = __longlong_as_double((long long)Rprev[a]) + move[a];
  // A member's freshly computed bound ub0 can exceed the carried radius by the rounding slack of a score
  // (sqrt(5 Delta) < eta): the list is made complete to 2 (Ra + eta) + eta, so that no member ever reaches past it
  // (a full scan over all centres would stall its whole warp).  Keeping more centres is always safe.
  const double thr = (2.0 * (Ra + eta) + eta) * (1.0 + 1e-9);
  if (tid == 0) {
    lthr[a] = thr;  // every centre outside the list is at least this far from centre a
    // members that keep centre a are within Ra of it (their bound grew by move[a]); evaluated points add theirs
    Rcur[a] = (unsigned long long)__double_as_longlong(Ra * (1.0 + 1e-15));
  }
  const double thr2 = thr * thr * (1.0 + 1e-15);  // sqrt(cc2) < thr implies cc2 < thr2: the root only where it can matter
  double ca4[4] = {0.0, 0.0, 0.0, 0.0};  // d <= 4: the centre's own coordinates stay in registers
  if (d <= 4)
    for (int k = 0; k < d; ++k) ca4[k] = C[a + (size_t)s * k];
  for (int j = tid; j < s; j += 256) {
    if (j == a) continue;
    double cc = 0.0;
    if (d <= 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < d) {
          const double df = ca4[k] - C[j + s * k];
          cc = fma(df, df, cc);
        }
    } else {
      for (int k = 0; k < d; ++k) {
        const double df = C[a + (size_t)s * k] - C[j + (size_t)s * k];
        cc = fma(df, df, cc);
      }
    }
    if (!(cc < thr2)) continue;
    cc = sqrt(cc);
    if (cc < thr) {
      const int pos = atomicAdd(&count, 1);
      if (pos < KM_LMAX) {
        kcc[pos] = cc;
        kj[pos] = j;
      }
    }
  }
  __syncthreads();
  const int cnt = count;
  if (cnt > KM_LMAX) {  // no list: every centre counts as a neighbour
    if (tid == 0) {
      len[a] = -1;
      const double mm_ = __longlong_as_double((long long)*maxmove_bits);
      cl[a] = make_double4(move[a], mm_, INFINITY, 0.0);
      if (clf) clf[a] = make_float4(__double2float_ru(move[a]), __double2float_ru(mm_), INFINITY, 0.f);
    }
    return;
  }
  // sort of (cc, j) ascending, ties broken by j => deterministic lists.  Short lists (the usual case: a few dozen
  // neighbours) by ranking: entry t goes to position #{u : key_u < key_t}, one pass over the list per thread and two
  // barriers instead of the ~log^2 barriers of the bitonic network below (the sort was most of this kernel's time).
  if (cnt <= 256) {
    double ct = 0.0;
    int jt = 0, rank = 0;
    if (tid < cnt) {
      ct = kcc[tid];
      jt = kj[tid];
      for (int u = 0; u < cnt; ++u) {
        const double cu = kcc[u];
        const int ju = kj[u];
        rank += (cu < ct) || (cu == ct && ju < jt);
      }
    }
    __syncthreads();
    if (tid < cnt) {
      kcc[rank] = ct;
      kj[rank] = jt;
    }
    __syncthreads();
  } else {
  // bitonic sort of (cc, j) ascending; padding is (+inf, INT_MAX)
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
  }
  double mv = 0.0;  // largest displacement among the listed centres (lower-bound decay of this cluster's points)
  for (int t = tid; t < cnt; t += 256) {
    list_j[(size_t)a * KM_LMAX + t] = kj[t];
    list_cc[(size_t)a * KM_LMAX + t] = kcc[t];
    mv = fmax(mv, move[kj[t]]);
  }
  for (int o = 16; o; o >>= 1) mv = fmax(mv, __shfl_xor_sync(0xffffffffu, mv, o));
  __shared__ double smv[8];
  if ((tid & 31) == 0) smv[tid >> 5] = mv;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) mv = fmax(mv, smv[w]);
    cl[a] = make_double4(move[a], mv, thr, 0.0);
    // the streaming bound test works in single precision with directed rounding (moves up, radius down)
    if (clf) clf[a] = make_float4(__double2float_ru(move[a]), __double2float_ru(mv), __double2float_rd(thr), 0.f);
    len[a] = cnt;
  }
}

__device__ __forceinline__ float2 km_pack(double u, double l) {  // u rounded up, l rounded down
  return make_float2(__double2float_ru(u), __double2float_rd(l));
}
// ---- one kernel per pruned pass -----------------------------------------------------------------------------------
// Hamerly-style bounds are carried from pass to pass: u >= |x - c_a| and l <= |x - c_j| for every j != a, stored as
// floats rounded outwards (u up, l down; the update below is single-precision arithmetic with directed rounding, so
// every step errs on the safe side).  The centres have moved since the bounds were taken: c_a by move[a]; the centres
// of a's current neighbour list by at most nbmove[a]; a centre outside the list is at least lthr[a] from c_a, hence
// lthr[a] - u from x.  If u + eta <= l still holds, every other centre is farther than c_a by eta, its COMPUTED score
// is strictly larger (eta^2 > 2 Delta) and the point keeps its centre - without its coordinates being read.
// cl[a] = {move, nbmove, lthr, -}.
//
// One CTA takes a tile of KF_TILE consecutive (cluster-sorted) points through all of it, with the tile's bounds and
// centre numbers staged in shared memory:
//   1. every point: the moved bounds and the test; the points that fail are compacted into `sv`;
//   2. the survivors, densely over the threads: u from one exact score against the point's own centre (one 32-byte
//      sector per point); what still fails is compacted into `wk`;
//   3. the walkers, densely again: the neighbour list of the point's centre (sorted by centre-centre distance) up to
//      the point's own 2 ub + eta; new bounds; a reassignment is an exact -x / +x on the persistent integer sums;
//   4. the tile's bounds are written back once, coalesced.
// The earlier form (a bound-test kernel that wrote 16-byte work records, and a grid-stride evaluation kernel over
// them) moved 557 MB per pass at C4 and walked the lists with 6 of 32 lanes; this one moves ~350 MB, never updates a
// single 8-byte pair inside a 32-byte sector, and starts every walk with full warps.  Same arithmetic per point.
// measured at C4 (99 pruned passes): 128 threads x 8 points 13.9 ms, 256 x 8 14.2, 128 x 4 14.7, 256 x 4 15.1, 128 x 16 19.2
constexpr int KF_T = 128, KF_Q = 8, KF_TILE = KF_T * KF_Q, KF_WT = 32 * KF_Q, KF_R = 2;
typedef unsigned short kf_idx;  // index of a point inside its warp's tile

template <int D>
__global__ void __launch_bounds__(KF_T, 8)
kmeans_pass_fused(int64_t n, const double4* __restrict__ Xs4, const double* __restrict__ rec, int s, Fx fx,
                  int32_t* __restrict__ as, unsigned long long* __restrict__ acc, const int32_t* __restrict__ list_j,
                  const double* __restrict__ list_cc, const int32_t* __restrict__ len,
                  const double* __restrict__ lthr, const float4* __restrict__ cl, double M, double delta2, double eta,
                  float eta_up, unsigned long long* __restrict__ Rcur, float2* __restrict__ UL,
                  unsigned long long* __restrict__ zero_maxmove, unsigned long long* __restrict__ prof,
                  int* __restrict__ prof_surv) {
  constexpr int STR = (D + 2) / 2 * 2;
  __shared__ __align__(16) float2 ul_s[KF_TILE];
  __shared__ __align__(16) int as_s[KF_TILE];
  __shared__ kf_idx sv_s[KF_TILE], wk_s[KF_TILE];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (blockIdx.x == 0 && tid == 1 && zero_maxmove) *zero_maxmove = 0;  // read by the lists kernel before, written by the update after
  // every warp owns KF_WT consecutive points and its slice of the staging arrays: no CTA-wide barrier anywhere
  const int64_t c0 = (int64_t)blockIdx.x * KF_TILE + wid * KF_WT;
  if (c0 >= n) return;
  float2* ul_t = ul_s + wid * KF_WT;
  int* as_t = as_s + wid * KF_WT;
  kf_idx* sv = sv_s + wid * KF_WT;
  kf_idx* wk = wk_s + wid * KF_WT;
  const unsigned lt_mask = (1u << lane) - 1;
  // ---- 1. bound test, four consecutive points per lane and group (16-byte loads; `as` and `UL` are padded to a
  // whole tile).  The survivors are placed with one prefix sum over the warp.
  int nsv;
  {
    constexpr int NG = KF_Q / 4;
    int4 a4[NG];
    float4 u4[NG][2];
#pragma unroll
    for (int g = 0; g < NG; ++g) {  // everything is requested before the first value is looked at
      const int64_t p = c0 + g * 128 + lane * 4;
      a4[g] = __ldcs(reinterpret_cast<const int4*>(as + p));
      u4[g][0] = __ldcs(reinterpret_cast<const float4*>(UL + p));
      u4[g][1] = __ldcs(reinterpret_cast<const float4*>(UL + p) + 1);
    }
    unsigned fail = 0;  // bit 4 g + k: point k of group g needs its coordinates
    int nvalid = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int i0 = g * 128 + lane * 4;
      const int64_t left = n - (c0 + i0);  // points of this group inside the array
      int av[4] = {a4[g].x, a4[g].y, a4[g].z, a4[g].w};
      const float uu[4] = {u4[g][0].x, u4[g][0].z, u4[g][1].x, u4[g][1].z};
      const float ll[4] = {u4[g][0].y, u4[g][0].w, u4[g][1].y, u4[g][1].w};
      float nu[4], nl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool valid = k < left;
        if (!valid) av[k] = 0;
        const float4 ca = cl[av[k]];  // {move of c_a (rounded up), largest move in a's list (up), list radius (down)}
        nu[k] = __fadd_ru(uu[k], ca.x);
        nl[k] = fminf(__fsub_rd(ll[k], ca.y), __fsub_rd(ca.z, nu[k]));
        if (valid) {
          ++nvalid;
          if (!(__fadd_ru(nu[k], eta_up) <= nl[k])) fail |= 1u << (4 * g + k);
        }
      }
      *reinterpret_cast<float4*>(&ul_t[i0]) = make_float4(nu[0], nl[0], nu[1], nl[1]);
      *reinterpret_cast<float4*>(&ul_t[i0 + 2]) = make_float4(nu[2], nl[2], nu[3], nl[3]);
      *reinterpret_cast<int4*>(&as_t[i0]) = make_int4(av[0], av[1], av[2], av[3]);
    }
    const int mine = __popc(fail);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    nsv = __shfl_sync(0xffffffffu, incl, 31);
    int base = incl - mine;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (fail & (1u << (4 * g + k))) sv[base++] = (kf_idx)(g * 128 + lane * 4 + k);
    if (prof) {
      int skipped = nvalid - mine;
      for (int o = 16; o; o >>= 1) skipped += __shfl_xor_sync(0xffffffffu, skipped, o);
      if (lane == 0 && skipped) atomicAdd(prof, (unsigned long long)skipped);
      if (lane == 0 && nsv) atomicAdd(prof_surv, nsv);
    }
  }
  __syncwarp();
  auto score = [&](const double (&x)[D], int j) {
    const double2* rj = reinterpret_cast<const double2*>(rec + (size_t)j * STR);
    double cr[STR];
#pragma unroll
    for (int q = 0; q < STR / 2; ++q) {
      double2 t = rj[q];
      cr[2 * q] = t.x;
      cr[2 * q + 1] = t.y;
    }
    double e = cr[D];
#pragma unroll
    for (int k = 0; k < D; ++k) e = fma(x[k], cr[k], e);
    return e;
  };
  // ---- 2. tightening: u from one exact score against the point's own centre; KF_R survivors per lane and trip,
  // their coordinates requested together
  int nwk = 0;
  for (int t0 = 0; t0 < nsv; t0 += 32 * KF_R) {
    double4 xv[KF_R];
    int ii[KF_R];
#pragma unroll
    for (int r = 0; r < KF_R; ++r) {
      const int t = t0 + r * 32 + lane;
      ii[r] = (t < nsv) ? (int)sv[t] : -1;
      if (ii[r] >= 0) xv[r] = Xs4[c0 + ii[r]];  // one 32-byte sector per point (a walker reads it again: kept cacheable)
    }
#pragma unroll
    for (int r = 0; r < KF_R; ++r) {
      bool walk = false;
      const int i = ii[r];
      if (i >= 0) {
        const double xa[4] = {xv[r].x, xv[r].y, xv[r].z, xv[r].w};
        double x[D], xn = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          x[k] = xa[k];
          xn = fma(x[k], x[k], xn);
        }
        const double ub0 = km_ub_fast(score(x, as_t[i]), M, xn, delta2);
        const double l = (double)ul_t[i].y;
        if (ub0 + eta <= l) ul_t[i] = km_pack(ub0, l);  // nothing can beat or tie centre a
        else walk = true;
      }
      const unsigned mw = __ballot_sync(0xffffffffu, walk);
      if (walk) wk[nwk + __popc(mw & lt_mask)] = (kf_idx)i;
      nwk += __popc(mw);
    }
  }
  if (prof && lane == 0 && nwk) atomicAdd(prof + 5, (unsigned long long)nwk);
  __syncwarp();
  // ---- 3. the neighbour-list walk of what is left.  The list entries are requested ahead of their use (indices two
  // entries ahead, score records one ahead), so that an entry costs one load latency instead of three in a row.
  int changed = 0;
  for (int t = lane; t < nwk; t += 32) {
    const int i = wk[t];
    const int64_t p = c0 + i;
    const int a = as_t[i];
    const double4 xv = Xs4[p];
    const double xa[4] = {xv.x, xv.y, xv.z, xv.w};
    double x[D], xn = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      x[k] = xa[k];
      xn = fma(x[k], x[k], xn);
    }
    double best = score(x, a), sec = INFINITY;  // sec: second smallest score seen
    int bj = a;
    const double ub0 = km_ub_fast(best, M, xn, delta2);
    const double thr = (2.0 * ub0 + eta) * (1.0 + 1e-9);
    const int L = len[a];
    double ccb = lthr[a];             // every centre that was NOT scanned is at least this far from centre a
    const bool full = (L < 0) || !(thr <= ccb);  // no list, or the bound reaches past the radius the list is complete to
    if (!full) {
      const int32_t* lj = list_j + (size_t)a * KM_LMAX;
      const double* lc = list_cc + (size_t)a * KM_LMAX;
      // entries past L are never used, but always inside the KM_LMAX-long row
      double c1 = lc[0], c2 = lc[1];
      int j1 = lj[0], j2 = lj[1];
      double r1[STR];
      {
        const double2* rj = reinterpret_cast<const double2*>(rec + (size_t)(L > 0 ? j1 : a) * STR);
#pragma unroll
        for (int k = 0; k < STR / 2; ++k) {
          const double2 tt = rj[k];
          r1[2 * k] = tt.x;
          r1[2 * k + 1] = tt.y;
        }
      }
      for (int q = 0; q < L; ++q) {
        const double cq = c1;
        const int j = j1;
        double rq[STR];
#pragma unroll
        for (int k = 0; k < STR; ++k) rq[k] = r1[k];
        c1 = c2;
        j1 = j2;
        if (q + 2 < L) {
          c2 = lc[q + 2];
          j2 = lj[q + 2];
        }
        if (cq >= thr) {  // sorted by centre-centre distance: nothing further can win or tie
          ccb = cq;
          break;
        }
        if (q + 1 < L) {
          const double2* rj = reinterpret_cast<const double2*>(rec + (size_t)j1 * STR);
#pragma unroll
          for (int k = 0; k < STR / 2; ++k) {
            const double2 tt = rj[k];
            r1[2 * k] = tt.x;
            r1[2 * k + 1] = tt.y;
          }
        }
        double e = rq[D];
#pragma unroll
        for (int k = 0; k < D; ++k) e = fma(x[k], rq[k], e);
        if (e < best || (e == best && j < bj)) {
          sec = best;
          best = e;
          bj = j;
        } else if (e < sec) {
          sec = e;
        }
      }
    } else {  // scan every centre (list overflow; otherwise excluded by the slack in the list radius)
      if (prof) atomicAdd(prof + 1, 1ull);
      ccb = INFINITY;
      for (int j = 0; j < s; ++j) {
        if (j == a) continue;
        const double e = score(x, j);
        if (e < best || (e == best && j < bj)) {
          sec = best;
          best = e;
          bj = j;
        } else if (e < sec) {
          sec = e;
        }
      }
    }
    // bounds for the passes that follow: u >= |x - c_bj|, l <= distance to every other centre
    const double ub = (bj == a) ? ub0 : km_ub_fast(best, M, xn, delta2);
    ul_t[i] = km_pack(ub, fmin(km_lb_fast(sec, M, xn, delta2), (ccb - ub0) * (1.0 - 1e-14)));
    if (bj != a) {
      ++changed;
      as[p] = bj;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        long long h, l;
        fx_encode(fx, x[k], &h, &l);
        atomicAdd(&acc[a + (size_t)s * k], (unsigned long long)(-h));
        atomicAdd(&acc[(size_t)s * D + a + (size_t)s * k], (unsigned long long)(-l));
        atomicAdd(&acc[bj + (size_t)s * k], (unsigned long long)h);
        atomicAdd(&acc[(size_t)s * D + bj + (size_t)s * k], (unsigned long long)l);
      }
      atomicAdd(&acc[(size_t)2 * s * D + a], (unsigned long long)(-1ll));
      atomicAdd(&acc[(size_t)2 * s * D + bj], 1ull);
    }
    km_radius(Rcur, bj, ub);  // radius of the (new) cluster for the next pass
  }
  __syncwarp();
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if (lane == 0 && changed) atomicAdd(&acc[(size_t)2 * s * D + s], (unsigned long long)changed);
  // ---- 4. the warp's bounds, written once
#pragma unroll
  for (int q = 0; q < KF_Q / 2; ++q) {
    const int i = (q * 32 + lane) * 2;
    if (c0 + i + 1 < n) *reinterpret_cast<float4*>(UL + c0 + i) = *reinterpret_cast<const float4*>(&ul_t[i]);
    else if (c0 + i < n) UL[c0 + i] = ul_t[i];
  }
}

// ---- cluster-sorted layout: counting sort by assignment -------------------------------------------------
__global__ void __launch_bounds__(1024)
kmeans_offsets_kernel(const long long* __restrict__ cnt, int s, int* __restrict__ cursor, int* __restrict__ seg_start) {
  // exclusive prefix sum of the cluster sizes by ONE CTA: every thread sums a contiguous run, the run totals are
  // scanned through shared memory (a serial scan by one thread took 70 us at s = 2000, per re-sort)
  __shared__ int part[1024];
  const int tid = threadIdx.x, per = (s + 1023) / 1024;
  const int j0 = tid * per, j1 = min(s, j0 + per);
  int sum = 0;
  for (int j = j0; j < j1; ++j) sum += (int)cnt[j];
  part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = (tid >= o) ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - sum;
  for (int j = j0; j < j1; ++j) {
    cursor[j] = run;
    seg_start[j] = run;
    run += (int)cnt[j];
  }
  if (tid == 1023) seg_start[s] = part[1023];
}

// src_perm == nullptr: source is the original order (row i <-> index i)
__global__ void __launch_bounds__(256)
kmeans_scatter_kernel(const double* __restrict__ Xsrc, int64_t n, int64_t ldsrc, int d, const int32_t* __restrict__ asrc,
                      const int32_t* __restrict__ src_perm, int* __restrict__ cursor, double4* __restrict__ Xdst,
                      int32_t* __restrict__ adst, int32_t* __restrict__ dst_perm, const float2* __restrict__ ULsrc,
                      float2* __restrict__ ULdst, const double4* __restrict__ X4src,
                      unsigned long long* __restrict__ Rnew) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  // everything the point carries is requested before its position is known (the atomic below is an L2 round trip)
  const int a = asrc[i];
  const int32_t pv = src_perm ? src_perm[i] : (int32_t)i;
  double4 xv;
  if (X4src) {
    xv = X4src[i];
  } else {  // first sort: from the caller's column-major matrix to one 32-byte record per point
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < d; ++k) v[k] = Xsrc[i + ldsrc * k];
    xv = make_double4(v[0], v[1], v[2], v[3]);
  }
  // bounds travel with the point; before the first sort there are none: u = inf, l = 0 force a full evaluation
  const float2 ul = ULsrc ? ULsrc[i] : make_float2(INFINITY, 0.f);
  // one atomic per group of equal keys in the warp
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, a);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&cursor[a], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  const int pos = base + __popc(peers & ((1u << lane) - 1));
  dst_perm[pos] = pv;
  adst[pos] = a;
  Xdst[pos] = xv;
  ULdst[pos] = ul;
  // the cluster radii are re-tightened at every sort (between sorts they only grow: R += move): the largest upper
  // bound of the group in one warp reduction (non-negative floats order like their bits), one atomic per group
  if (Rnew) {
    const unsigned mb = __reduce_max_sync(peers, __float_as_uint(ul.x));
    if (lane == leader) atomicMax(&Rnew[a], (unsigned long long)__double_as_longlong((double)__uint_as_float(mb)));
  }
}

// the KNN stage reads the sorted rows column-major
__global__ void kmeans_x4_to_colmajor_kernel(const double4* __restrict__ X4, int64_t n, int d, double* __restrict__ Xs) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double4 v = X4[p];
  const double a[4] = {v.x, v.y, v.z, v.w};
  for (int k = 0; k < d; ++k) Xs[p + n * k] = a[k];
}

__global__ void kmeans_unpermute_kernel(const int32_t* __restrict__ as, const int32_t* __restrict__ perm, int64_t n,
                                        int32_t* __restrict__ assign) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) assign[perm[p]] = as[p];
}

template <int D>
void launch_small(Ctx* c, const double* X, int64_t n, int64_t ldx, const double* rec, int s, const Fx& fx,
                  int32_t* assign, unsigned long long* acc, unsigned long long* Rbits, double M, double delta2) {
  constexpr int P = 4;
  constexpr int STR = (D + 2) / 2 * 2;
  int chunk = std::min(s, 1024);
  size_t smem = (size_t)chunk * STR * sizeof(double);
  int grid = ceil_div(n, (int64_t)KM_THREADS * P);
  if (grid < 1) return;
  FLGP_LAUNCH(c, (kmeans_assign_small<D, P>), grid, KM_THREADS, smem, X, n, ldx, rec, s, fx, assign, acc, chunk, 1,
              Rbits, M, delta2);
}

}  // namespace

// total within-cluster sum of squares of an assignment, as exact fixed-point limbs (order independent: the same
// bits for every thread schedule and every rank count) -- what stats::kmeans compares its nstart runs by
__global__ void kmeans_withinss_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d,
                                       const double* __restrict__ U, int s, const int32_t* __restrict__ assign, Fx fx,
                                       unsigned long long* __restrict__ out) {
  long long h = 0, l = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = assign[i];
    double w = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = X[i + ldx * k] - U[a + (size_t)s * k];
      w = fma(df, df, w);
    }
    long long hh, ll;
    fx_encode(fx, w, &hh, &ll);
    h += hh;
    l += ll;
  }
  for (int o = 16; o; o >>= 1) {
    h += __shfl_xor_sync(0xffffffffu, h, o);
    l += __shfl_xor_sync(0xffffffffu, l, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], (unsigned long long)h);
    atomicAdd(&out[1], (unsigned long long)l);
  }
}

double kmeans_withinss_run(Ctx* c, const double* X, int64_t n_local, int64_t ldx, int d, const double* U, int s,
                           const int32_t* assign, int64_t n_total) {
  const double maxabs = maxabs_run(c, X, n_local, ldx, d);
  Fx fx;
  if (fx_make(4.0 * d * maxabs * maxabs + 1e-300, n_total, &fx)) fail(2, "kmeans: non-finite input");
  DevBuf<unsigned long long> acc(2);
  acc.zero(c->stream);
  if (n_local > 0) {
    const int grid = (int)std::min<int64_t>((n_local + 255) / 256, (int64_t)c->sm_count * 8);
    FLGP_LAUNCH(c, kmeans_withinss_kernel, grid, 256, 0, X, n_local, ldx, d, U, s, assign, fx, acc.p);
  }
  comm_allreduce_i64(c, reinterpret_cast<int64_t*>(acc.p), 2);
  long long h[2];
  acc.download(reinterpret_cast<unsigned long long*>(h), 2, c->stream);
  sync(c);
  return fx_todouble(h[0]) * fx.q1 + fx_todouble(h[1]) * fx.q2;
}

double maxabs_run(Ctx* c, const double* X, int64_t n_local, int64_t ldx, int d) {
  DevBuf<unsigned long long> m(1);
  m.zero(c->stream);
  if (n_local > 0) {
    int grid = std::min<int64_t>((n_local * d + 255) / 256, (int64_t)c->sm_count * 8);
    FLGP_LAUNCH(c, maxabs_kernel, grid, 256, 0, X, n_local, ldx, d, m.p);
  }
  comm_allreduce_max_f64(c, reinterpret_cast<double*>(m.p), 1);  // non-negative doubles order like their bits
  double h = 0.0;
  m.download(reinterpret_cast<unsigned long long*>(&h), 1, c->stream);
  sync(c);
  return h;
}

void kmeans_run(Ctx* c, const double* X, int64_t n_local, int64_t ldx, int d, int s, int64_t n_total,
                int64_t row_offset, const int32_t* init_idx_h, int iter_max, double* U, int32_t* assign,
                int* iters_out, KMeansSorted* sorted_out) {
  if (s < 1 || s > n_total || d < 1) fail(2, "kmeans: need 1 <= s <= n (s=%d, n=%lld)", s, (long long)n_total);
  for (int j = 0; j < s; ++j)
    if (init_idx_h[j] < 0 || init_idx_h[j] >= n_total) fail(2, "kmeans: initial index out of range");
  double maxabs = maxabs_run(c, X, n_local, ldx, d);
  Fx fx;
  if (fx_make(maxabs, n_total, &fx)) fail(2, "kmeans: non-finite input");
  const double Moff = (2.0 * d) * (maxabs * maxabs);  // score offset of the k-means contract

  const size_t words = (size_t)2 * s * d + s + 1;
  DevBuf<long long> acc(words);
  DevBuf<int32_t> init(s);
  init.upload(init_idx_h, s, c->stream);
  double* C = U;                     // first s*d entries of U are the centres (column-major)
  double* sizes = U + (size_t)s * d; // last column
  FLGP_CUDA(cudaMemsetAsync(U, 0, sizeof(double) * (size_t)s * (d + 1), c->stream));
  FLGP_LAUNCH(c, kmeans_init_kernel, ceil_div(s * d, 256), 256, 0, X, n_local, ldx, d, s, row_offset, init.p,
              reinterpret_cast<long long*>(C));
  comm_allreduce_i64(c, reinterpret_cast<int64_t*>(C), (size_t)s * d);
  if (n_local > 0) FLGP_CUDA(cudaMemsetAsync(assign, 0xFF, sizeof(int32_t) * n_local, c->stream));

  const bool small = d <= 4;
  const int str = (d + 2) / 2 * 2;
  DevBuf<double> rec, C2, cn;
  if (small) rec.alloc((size_t)s * str);
  else {
    C2.alloc((size_t)s * d);
    cn.alloc(s);
  }
  // large d: assignment on the FP64 tensor cores (distsel.cu) over row-major copies, certified rows only; the rest
  // in the oracle's order.  Accumulators are persistent (exact integer -x / +x per changed row).
  const bool dmma = !small && dist_select_supported(n_local, s, 1) && n_local < ((int64_t)1 << 31) &&
                    std::getenv("FLGP_NO_DMMA_DIST") == nullptr;
  const int dp = (d + 1) / 2 * 2;
  DevBuf<double> Xr, Cr, Xsv, hxn, hUB, hLB, hvals, Cold, hmove, hmm;
  DevBuf<int32_t> newa, und_list, hsurv;
  DevBuf<int> und_count, hnsurv;
  const bool hamerly = dmma && std::getenv("FLGP_NO_HAMERLY") == nullptr;
  if (dmma) {
    Xr.alloc((size_t)n_local * dp);
    Cr.alloc((size_t)s * dp);
    newa.alloc(n_local);
    und_list.alloc(n_local);
    und_count.alloc(1);
    hnsurv.alloc(1);
    hvals.alloc((size_t)2 * n_local);
    hxn.alloc(n_local);
    hUB.alloc(n_local);
    hLB.alloc(n_local);
    if (hamerly) {
      Xsv.alloc((size_t)n_local * dp);
      hsurv.alloc(n_local);
      Cold.alloc((size_t)s * d);
      hmove.alloc(s);
      hmm.alloc(4);
      hmove.zero(c->stream);
    }
    FLGP_LAUNCH(c, kmeans_xnorm_kernel, ceil_div(n_local, 256), 256, 0, X, n_local, ldx, d, hxn.p);
    to_rowmajor_run(c, X, n_local, ldx, d, dp, Xr.p);
    const size_t xsm = sizeof(double) * d;
    if (xsm > 200 * 1024) fail(2, "kmeans: d=%d exceeds the supported maximum", d);
    if (xsm > 40 * 1024) {
      FLGP_CUDA(cudaFuncSetAttribute(kmeans_exact_rows2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
    }
  }
  unsigned long long* uacc = reinterpret_cast<unsigned long long*>(acc.p);
  // pruned passes (small d): persistent local accumulators, neighbour lists, cluster-sorted points
  const bool pruned = small && n_total >= 2;
  const bool persistent = pruned || dmma;  // local accumulators survive the pass; a copy is all-reduced
  const double bound = Moff + 4.0 * d * (maxabs * maxabs);                    // >= the magnitude of every score term
  const double Delta = 8.0 * (d + 4) * 1.1102230246251565e-16 * bound;        // generous bound on a score's rounding error
  const double delta2 = 4.0 * Delta;                                          // slack added to squared distances
  const double eta = 2.0 * std::sqrt(2.0 * Delta) + 1e-12 * maxabs;           // eta^2 > 2 Delta with room to spare
  DevBuf<unsigned long long> Rbits[2];
  DevBuf<int32_t> nlist, nlen, perm[2], as[2];
  DevBuf<double> ncc, move, lthr;
  DevBuf<double4> Xs4[2];
  DevBuf<float2> UL[2];
  DevBuf<double4> cl;
  DevBuf<float4> clf;
  const bool prof_skip = std::getenv("FLGP_KMEANS_PROF") != nullptr;
  DevBuf<int> prof_hist(128);
  prof_hist.zero(c->stream);
  DevBuf<unsigned long long> maxmove, nskip;
  DevBuf<int> cursor, seg_start, nstrag;
  DevBuf<long long> acc_red;  // all-reduced copy of the local accumulators (multi-GPU)
  if (pruned) {
    Rbits[0].alloc(s);
    Rbits[1].alloc(s);
    nlist.alloc((size_t)s * KM_LMAX);
    ncc.alloc((size_t)s * KM_LMAX);
    nlen.alloc(s);
    move.alloc(s);
    cursor.alloc(s);
    seg_start.alloc(s + 1);
    nstrag.alloc(1);
    for (int b = 0; b < 2; ++b) {
      perm[b].alloc(std::max<int64_t>(n_local, 1));
      as[b].alloc((int64_t)ceil_div(std::max<int64_t>(n_local, 1), KF_TILE) * KF_TILE);  // whole tiles: kmeans_pass_fused
      Xs4[b].alloc(std::max<int64_t>(n_local, 1));
      UL[b].alloc((int64_t)ceil_div(std::max<int64_t>(n_local, 1), KF_TILE) * KF_TILE);
    }
    lthr.alloc(s);
    cl.alloc(s);
    clf.alloc(s);
    maxmove.alloc(1);
    nskip.alloc(8);
    maxmove.zero(c->stream);
    nskip.zero(c->stream);
    Rbits[0].zero(c->stream);
  }
  if (persistent && c->nranks > 1) acc_red.alloc(words);
  if (hamerly) {  // neighbour lists of the centres, as in the small-d pruned passes
    Rbits[0].alloc(s);
    Rbits[1].alloc(s);
    nlist.alloc((size_t)s * KM_LMAX);
    ncc.alloc((size_t)s * KM_LMAX);
    nlen.alloc(s);
    lthr.alloc(s);
    cl.alloc(s);
    maxmove.alloc(1);
    maxmove.zero(c->stream);
    Rbits[0].zero(c->stream);
  }
  int cur = 0;              // which sorted buffer is live
  bool have_sorted = false;
  int64_t moved_since_sort = 0, moved_base = 0;  // assignments changed (all ranks) since the last sort
  DevBuf<long long> kstate(2);
  kstate.zero(c->stream);
  int rsel = 0;  // Rbits[rsel]: radii gathered during the previous pass
  int filt_pause = 0, filt_backoff = KM_CHECK;  // large-d bound filter: passes left without it / next pause length
  bool last_filtered = false;
  auto resort = [&]() {
    // counting sort by cluster of the current assignment (local counts live in acc)
    StageScope st(c, "kmeans_sort");
    FLGP_LAUNCH(c, kmeans_offsets_kernel, 1, 1024, 0, acc.p + (size_t)2 * s * d, s, cursor.p, seg_start.p);
    const int nxt = have_sorted ? 1 - cur : 0;
    if (n_local > 0) {
      if (have_sorted) FLGP_CUDA(cudaMemsetAsync(Rbits[rsel].p, 0, sizeof(unsigned long long) * s, c->stream));
      if (have_sorted)
        FLGP_LAUNCH(c, kmeans_scatter_kernel, ceil_div(n_local, 256), 256, 0, (const double*)nullptr, n_local, n_local, d,
                    as[cur].p, perm[cur].p, cursor.p, Xs4[nxt].p, as[nxt].p, perm[nxt].p, UL[cur].p, UL[nxt].p,
                    Xs4[cur].p, Rbits[rsel].p);
      else
        FLGP_LAUNCH(c, kmeans_scatter_kernel, ceil_div(n_local, 256), 256, 0, X, n_local, ldx, d, assign,
                    (const int32_t*)nullptr, cursor.p, Xs4[nxt].p, as[nxt].p, perm[nxt].p, (const float2*)nullptr,
                    UL[nxt].p, (const double4*)nullptr, (unsigned long long*)nullptr);
    }
    cur = nxt;
    have_sorted = true;
    moved_base += moved_since_sort;
    moved_since_sort = 0;
  };
  int it = 0;
  while (it < iter_max) {
    ++it;
    const bool brute = !pruned || it == 1;
    const bool fused_pass = small && !brute;  // pruned pass: record, counters are handled inside its own kernels
    if (brute && !(dmma && it > 1)) acc.zero(c->stream);
    else if (!fused_pass)
      FLGP_CUDA(cudaMemsetAsync(acc.p + (words - 1), 0, sizeof(long long), c->stream));  // the `changed` slot
    if (fused_pass) {
    } else if (small) FLGP_LAUNCH(c, kmeans_prep_kernel, ceil_div(s, 128), 128, 0, C, s, d, str, Moff, rec.p);
    else {
      FLGP_LAUNCH(c, kmeans_prep_c2_kernel, ceil_div((int64_t)s * d, 256), 256, 0, C, (size_t)s * d, C2.p);
      FLGP_LAUNCH(c, kmeans_prep_cn_kernel, ceil_div(s, 64), 64, 0, C, s, d, Moff, cn.p);
    }
    static const int resort_div = std::getenv("FLGP_KM_RESORT") ? std::atoi(std::getenv("FLGP_KM_RESORT")) : 4;
    if (!brute && (!have_sorted || moved_since_sort * resort_div > n_total)) resort();
    // when timing is on, the assign+accumulate work gets its own CUDA-event pair per pass
    StageScope kst(c, brute ? "kmeans_assign_kernel" : "kmeans_pruned_pass", 2.0 * s * d * (double)n_local,
                   (8.0 * d + 4.0) * (double)n_local);
    if (!brute) {
      unsigned long long* Rprev = Rbits[rsel].p;
      unsigned long long* Rcur = Rbits[1 - rsel].p;
      FLGP_LAUNCH(c, kmeans_lists_kernel, s, 256, 0, C, s, d, Rprev, move.p, eta, nlist.p, ncc.p, nlen.p, lthr.p, cl.p,
                  maxmove.p, Rcur, clf.p, rec.p, str, Moff, acc.p + (words - 1), nstrag.p);
      if (n_local == 0) maxmove.zero(c->stream);  // otherwise kmeans_pass_fused does it
      if (n_local > 0) {
#define FLGP_FUSED(D_)                                                                                            \
  FLGP_LAUNCH(c, (kmeans_pass_fused<D_>), ceil_div(n_local, KF_TILE), KF_T, 0, n_local, Xs4[cur].p, rec.p, s, fx,  \
              as[cur].p, uacc, nlist.p, ncc.p, nlen.p, lthr.p, clf.p, Moff, delta2, eta,                           \
              std::nextafterf((float)eta, INFINITY), Rcur, UL[cur].p, maxmove.p, prof_skip ? nskip.p : nullptr,    \
              prof_skip ? nstrag.p : nullptr)
        switch (d) {
          case 1: FLGP_FUSED(1); break;
          case 2: FLGP_FUSED(2); break;
          case 3: FLGP_FUSED(3); break;
          default: FLGP_FUSED(4); break;
        }
#undef FLGP_FUSED
      }
      if (prof_skip && it < 128)
        FLGP_CUDA(cudaMemcpyAsync(prof_hist.p + it, nstrag.p, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
      rsel = 1 - rsel;
    } else if (small && pruned && s >= KM_PIVOT_MIN_S && n_local > 0) {
      // first pass through pivots (see kmeans_assign_listed): np pivot centres, then each pivot group against the
      // centres that can reach it
      const int np = KM_NPIVOT, lmax = KM_PIVOT_LMAX, chunk = KM_THREADS * 4;
      std::vector<int32_t> pv_h(np);
      for (int q = 0; q < np; ++q) pv_h[q] = (int32_t)((int64_t)q * s / np);
      DevBuf<int32_t> pv(np), plist((size_t)np * lmax), plen(np);
      DevBuf<double> prec((size_t)np * str);
      DevBuf<unsigned long long> Rp(np);
      DevBuf<long long> pcnt(np);
      const int max_items = (int)(n_local / chunk) + np + 1;
      DevBuf<int> items((size_t)3 * max_items), n_items(1);
      pv.upload(pv_h.data(), np, c->stream);
      Rp.zero(c->stream);
      pcnt.zero(c->stream);
      FLGP_LAUNCH(c, kmeans_pivot_rec_kernel, ceil_div(np * str, 128), 128, 0, rec.p, str, pv.p, np, prec.p);
      switch (d) {  // labels (pivot numbers) into `assign`, group radii into Rp; nothing is accumulated
        case 1: launch_small<1>(c, X, n_local, ldx, prec.p, np, fx, assign, nullptr, Rp.p, Moff, delta2); break;
        case 2: launch_small<2>(c, X, n_local, ldx, prec.p, np, fx, assign, nullptr, Rp.p, Moff, delta2); break;
        case 3: launch_small<3>(c, X, n_local, ldx, prec.p, np, fx, assign, nullptr, Rp.p, Moff, delta2); break;
        default: launch_small<4>(c, X, n_local, ldx, prec.p, np, fx, assign, nullptr, Rp.p, Moff, delta2); break;
      }
      FLGP_LAUNCH(c, kmeans_hist_kernel, c->sm_count * 4, 256, np * sizeof(int), assign, n_local, np, pcnt.p);
      FLGP_LAUNCH(c, kmeans_items_kernel, 1, 128, 2 * np * sizeof(int), pcnt.p, np, chunk, cursor.p, items.p,
                  n_items.p);
      FLGP_LAUNCH(c, kmeans_scatter_fewkeys_kernel, ceil_div(n_local, 256 * KM_FK_PTS), 256, 2 * np * sizeof(int), X,
                  n_local, ldx, d, assign, np, cursor.p, Xs4[1].p, perm[1].p);
      FLGP_LAUNCH(c, kmeans_pivot_lists_kernel, np, 256, 0, C, s, d, pv.p, Rp.p, eta, lmax, plist.p, plen.p);
      const size_t lsm = (size_t)lmax * (str * sizeof(double) + sizeof(int));
#define FLGP_LISTED(D_)                                                                                             \
  do {                                                                                                              \
    if (lsm > 48 * 1024)                                                                                            \
      FLGP_CUDA(cudaFuncSetAttribute((kmeans_assign_listed<D_, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                     (int)lsm));                                                                    \
    FLGP_LAUNCH(c, (kmeans_assign_listed<D_, 4>), max_items, KM_THREADS, lsm, Xs4[1].p, perm[1].p, items.p,         \
                n_items.p, rec.p, plist.p, plen.p, lmax, s, fx, assign, uacc, Rbits[rsel].p, Moff, delta2);         \
  } while (0)
      switch (d) {
        case 1: FLGP_LISTED(1); break;
        case 2: FLGP_LISTED(2); break;
        case 3: FLGP_LISTED(3); break;
        default: FLGP_LISTED(4); break;
      }
#undef FLGP_LISTED
      sync(c);  // the pivot work buffers are released here
    } else if (small) {
      unsigned long long* R0 = pruned ? Rbits[rsel].p : nullptr;
      switch (d) {
        case 1: launch_small<1>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc, R0, Moff, delta2); break;
        case 2: launch_small<2>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc, R0, Moff, delta2); break;
        case 3: launch_small<3>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc, R0, Moff, delta2); break;
        default: launch_small<4>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc, R0, Moff, delta2); break;
      }
    } else if (dmma) {
      to_rowmajor_run(c, C, s, s, d, dp, Cr.p);
      und_count.zero(c->stream);
      // pass 1 evaluates every row; so does a pass while the filter is paused (it is paused, with doubling back-off,
      // whenever more than half of the rows survive it: on data without low-dimensional structure the bound test,
      // the tightening and the gather then cost more than they save).  A full evaluation refreshes every bound, so
      // the filter can resume at any pass.
      const bool filtered = hamerly && it > 1 && filt_pause == 0;
      if (filt_pause > 0) --filt_pause;
      last_filtered = filtered;
      const double* rows = Xr.p;
      const int32_t* surv = nullptr;
      if (filtered) {
        hnsurv.zero(c->stream);
        FLGP_LAUNCH(c, kmeans_lists_kernel, s, 256, 0, C, s, d, Rbits[rsel].p, hmove.p, eta, nlist.p, ncc.p, nlen.p,
                    lthr.p, cl.p, maxmove.p, Rbits[1 - rsel].p);
        FLGP_LAUNCH(c, kmeans_hbounds_kernel, c->sm_count * 8, 256, 0, X, n_local, ldx, d, C2.p, cn.p, s, assign, cl.p,
                    eta, Moff, delta2, hxn.p, hUB.p, hLB.p, hsurv.p, hnsurv.p);
        FLGP_LAUNCH(c, kmeans_gather_rows_kernel, c->sm_count * 8, 256, 0, Xr.p, dp, hsurv.p, hnsurv.p, Xsv.p);
        rows = Xsv.p;
        surv = hsurv.p;
      } else {
        const int all = (int)n_local;
        FLGP_CUDA(cudaMemcpyAsync(hnsurv.p, &all, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        if (hamerly) FLGP_CUDA(cudaMemsetAsync(Rbits[rsel].p, 0, sizeof(unsigned long long) * s, c->stream));
      }
      dist_select_run(c, rows, n_local, Cr.p, s, dp, cn.p, 1, 4.0 * Delta, nullptr, newa.p, n_local, und_count.p,
                      und_list.p, filtered ? hnsurv.p : nullptr, hvals.p);
      FLGP_LAUNCH(c, kmeans_exact_rows2_kernel, c->sm_count * 4, 256, sizeof(double) * d, X, ldx, d, C2.p, cn.p, s,
                  und_count.p, und_list.p, surv, newa.p, hvals.p);
      // radii: pass 1 collects them into Rbits[rsel]; a filtered pass into the buffer its lists kernel initialised
      unsigned long long* Rw = hamerly ? (filtered ? Rbits[1 - rsel].p : Rbits[rsel].p) : nullptr;
      FLGP_LAUNCH(c, kmeans_hcommit_kernel, c->sm_count * 8, 256, 0, rows, d, dp, surv, hnsurv.p, newa.p, hvals.p, hxn.p,
                  assign, s, fx, uacc, Moff, delta2, hUB.p, hLB.p, Rw);
      if (filtered) rsel = 1 - rsel;
    } else {
      int grid = ceil_div(n_local, KT_TP);
      if (grid > 0)
        FLGP_LAUNCH(c, kmeans_assign_tiled, grid, 256, 0, X, n_local, ldx, d, C2.p, cn.p, s, fx, assign, uacc);
    }
    kst.stop();
    long long* red = acc.p;
    if (persistent && c->nranks > 1) {  // keep the local sums intact: reduce out of place
      red = acc_red.p;
      comm_allreduce_i64_to(c, reinterpret_cast<const int64_t*>(acc.p), reinterpret_cast<int64_t*>(red), words);
    } else {
      comm_allreduce_i64(c, reinterpret_cast<int64_t*>(red), words);
    }
    if (pruned && brute) maxmove.zero(c->stream);  // a pruned pass zeroes it in kmeans_pass_fused
    if (small) {
      FLGP_LAUNCH(c, kmeans_update_kernel, ceil_div(s, 128), 128, 0, red, s, d, fx, C, sizes, pruned ? move.p : nullptr,
                  pruned ? maxmove.p : nullptr, it, kstate.p);
    } else {
      if (hamerly)
        FLGP_CUDA(cudaMemcpyAsync(Cold.p, C, sizeof(double) * (size_t)s * d, cudaMemcpyDeviceToDevice, c->stream));
      FLGP_LAUNCH(c, kmeans_update_wide_kernel, ceil_div((int64_t)s * d, 256), 256, 0, red, s, d, fx, C, sizes, it,
                  kstate.p);
      if (hamerly) {
        FLGP_LAUNCH(c, kmeans_move_kernel, ceil_div((int64_t)s * 32, 256), 256, 0, C, Cold.p, s, d, hmove.p);
        FLGP_LAUNCH(c, kmeans_move_max_kernel, 1, 256, 0, hmove.p, s, hmm.p, maxmove.p);
      }
    }
    // one host round trip per KM_CHECK passes (and after the first, which decides about the sorted layout)
    if (it == 1 || it % KM_CHECK == 0 || it == iter_max || (hamerly && it == 3)) {
      FLGP_CUDA(cudaMemcpyAsync(c->pinned, kstate.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
      if (hamerly) FLGP_CUDA(cudaMemcpyAsync(c->pinned + 2, hnsurv.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      sync(c);
      if (hamerly && last_filtered) {
        const int nsv = *reinterpret_cast<const int*>(c->pinned + 2);
        if ((int64_t)nsv * 2 > n_local) {
          filt_pause = filt_backoff;
          filt_backoff = std::min(filt_backoff * 2, 64);
        }
      }
      moved_since_sort = c->pinned[1] - moved_base;
      if (c->pinned[0] != 0) {  // no assignment changed anywhere in pass pinned[0]
        it = (int)c->pinned[0];
        break;
      }
    }
  }
  if (have_sorted && n_local > 0)
    FLGP_LAUNCH(c, kmeans_unpermute_kernel, ceil_div(n_local, 256), 256, 0, as[cur].p, perm[cur].p, n_local, assign);
  if (pruned && std::getenv("FLGP_KMEANS_PROF")) {
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    nskip.download(h, 8, c->stream);
    sync(c);
    fprintf(stderr, "[flgp kmeans prof] %d passes, %.1f%% of the point-passes after the first skipped by their bounds, "
            "%llu full scans\n", it, it > 1 ? 100.0 * (double)h[0] / ((double)n_local * (it - 1)) : 0.0, h[1]);
    fprintf(stderr, "[flgp kmeans prof] neighbour lists walked for %.1f%% of the point-passes\n",
            100.0 * (double)h[5] / ((double)std::max<int64_t>(n_local, 1) * std::max(it - 1, 1)));
    int hist[128];
    prof_hist.download(hist, 128, c->stream);
    sync(c);
    fprintf(stderr, "[flgp kmeans prof] survivors of the bound test per pass (%% of points):");
    for (int q = 2; q <= it && q < 128; ++q) fprintf(stderr, " %.0f", 100.0 * hist[q] / (double)std::max<int64_t>(n_local, 1));
    fprintf(stderr, "\n");
  }
  if (iters_out) *iters_out = it;
  if (sorted_out) {
    sorted_out->valid = have_sorted;
    if (have_sorted) {
      // hand the cluster-sorted layout to the next stage (KNN): radii are those of the last pass; the centres
      // have moved by at most move[] since (zero when the loop stopped because nothing changed)
      sorted_out->Xs.alloc(std::max<int64_t>(n_local * d, 1));
      if (n_local > 0)
        FLGP_LAUNCH(c, kmeans_x4_to_colmajor_kernel, ceil_div(n_local, 256), 256, 0, Xs4[cur].p, n_local, d,
                    sorted_out->Xs.p);
      sorted_out->perm = std::move(perm[cur]);
      sorted_out->as = std::move(as[cur]);
      sorted_out->Rbits = std::move(Rbits[rsel]);
      sorted_out->move = std::move(move);
      sorted_out->seg_start = std::move(seg_start);
      sorted_out->eta = eta;
      sorted_out->maxabs = maxabs;
    }
  }
  sync(c);
}

}  // namespace flgp
