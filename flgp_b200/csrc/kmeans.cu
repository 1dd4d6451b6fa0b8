// kmeans.cu — Lloyd k-means anchors (replaces the stats::kmeans callback at
// /root/reference/src/Utils.cpp:36-45; contract in oracle/flgp_oracle.cpp orc_kmeans_*).
//
// Per iteration ONE fused kernel does: score every (point, centre) pair in fp64 FMA form,
// arg-min with lowest-index ties, compare with the previous assignment, and accumulate the
// centroid sums as two-limb int64 fixed point with integer atomics (associative => the sums are
// bit-identical for any thread schedule and any number of GPUs).  The only collective is one
// int64 all-reduce of 2*s*d + s + 1 words per iteration.
//
// Roofline: FP64 FMA pipe.  Algorithmic work = 2*s*d flop per point per iteration
// (SURVEY.md §8d); bytes 8d + 4 per point (negligible: intensity = s/4 flop/byte).
#include <algorithm>
#include <cfloat>
#include <cmath>

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

// split records into the tiled kernel's operands: C2 (s x d col-major) and cn (s)
__global__ void kmeans_prep_tiled_kernel(const double* __restrict__ C, int s, int d, double M, double* C2, double* cn) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s) return;
  double a = 0.0;
  for (int k = 0; k < d; ++k) {
    double c = C[j + (size_t)s * k];
    a = fma(c, c, a);
    C2[j + (size_t)s * k] = -2.0 * c;
  }
  cn[j] = __dadd_rn(a, M);
}

__device__ __forceinline__ void km_commit(const Fx& fx, int s, int d, int bj, const double* x, int32_t* assign,
                                          int64_t i, unsigned long long* acc, int& changed) {
  if (assign[i] != bj) ++changed;
  assign[i] = bj;
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
                    Fx fx, int32_t* __restrict__ assign, unsigned long long* __restrict__ acc, int chunk, int one) {
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
    if (i < n) km_commit(fx, s, D, bj[p], x[p], assign, i, acc, changed);
  }
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if ((tid & 31) == 0 && changed) atomicAdd(&acc[(size_t)2 * s * D + s], (unsigned long long)changed);
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

__global__ void kmeans_update_kernel(const long long* __restrict__ acc, int s, int d, Fx fx, double* C,
                                     double* sizes) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s * d) return;
  int j = e % s, k = e / s;
  long long cnt = acc[(size_t)2 * s * d + j];
  if (k == 0) sizes[j] = (double)cnt;
  if (cnt == 0) return;  // empty cluster keeps its centre
  double sum = fx_decode(fx, acc[j + (size_t)s * k], acc[(size_t)s * d + j + (size_t)s * k]);
  C[j + (size_t)s * k] = sum / (double)cnt;
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

// ---- exact pruned iterations (small d) ------------------------------------------------------------------
// After the first brute-force pass every point knows a centre a (its previous assignment).  With
// ub >= |x - c_a| (true distance), a centre j with |c_a - c_j| >= 2 ub + eta satisfies
//   |x - c_j| >= |c_a - c_j| - |x - c_a| >= |x - c_a| + eta   =>   score_j - score_a >= eta^2 > 2 Delta,
// where Delta bounds the rounding error of a computed score, so the COMPUTED score of j is strictly larger
// than that of a: j can neither win nor tie.  Scanning only {a} + {j : |c_a - c_j| < 2 R_a + eta}
// (R_a = max ub over the members of a) therefore returns exactly the arg-min of the brute-force scan,
// including its lowest-index tie rule.  Sums are integer limbs, so a reassignment is an exact -x / +x.
constexpr int KM_LMAX = 512;  // longest neighbour list; longer => that cluster falls back to the full scan

template <int D>
__global__ void __launch_bounds__(256)
kmeans_radius_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, const double* __restrict__ rec,
                     const int32_t* __restrict__ assign, double M, double delta2, unsigned long long* __restrict__ Rbits) {
  constexpr int STR = (D + 2) / 2 * 2;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int a = assign[i];
  const double* ra = rec + (size_t)a * STR;
  double e = ra[D], xn = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double x = X[i + ldx * k];
    e = fma(x, ra[k], e);
    xn = fma(x, x, xn);
  }
  const double d2 = fmax((e - M) + xn, 0.0) + delta2;    // >= true squared distance
  const double ub = sqrt(d2) * (1.0 + 1e-14);
  atomicMax(&Rbits[a], (unsigned long long)__double_as_longlong(ub));  // non-negative doubles order like their bits
}

__global__ void __launch_bounds__(128)
kmeans_lists_kernel(const double* __restrict__ C, int s, int d, const unsigned long long* __restrict__ Rbits, double eta,
                    int32_t* __restrict__ list, int32_t* __restrict__ len) {
  __shared__ int count;
  const int a = blockIdx.x;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  const double Ra = __longlong_as_double((long long)Rbits[a]);
  const double thr = (2.0 * Ra + eta) * (1.0 + 1e-9);   // inflated: keeping more centres is always safe
  for (int j = threadIdx.x; j < s; j += 128) {
    if (j == a) continue;
    double cc = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = C[a + (size_t)s * k] - C[j + (size_t)s * k];
      cc = fma(df, df, cc);
    }
    if (sqrt(cc) < thr) {
      const int pos = atomicAdd(&count, 1);
      if (pos < KM_LMAX) list[(size_t)a * KM_LMAX + pos] = j;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) len[a] = (count <= KM_LMAX) ? count : -1;
}

template <int D>
__global__ void __launch_bounds__(256)
kmeans_assign_pruned(const double* __restrict__ X, int64_t n, int64_t ldx, const double* __restrict__ rec, int s, Fx fx,
                     int32_t* __restrict__ assign, unsigned long long* __restrict__ acc,
                     const int32_t* __restrict__ list, const int32_t* __restrict__ len) {
  constexpr int STR = (D + 2) / 2 * 2;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int changed = 0;
  if (i < n) {
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = X[i + ldx * k];
    const int a = assign[i];
    auto score = [&](int j) {
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
    double best = score(a);
    int bj = a;
    const int L = len[a];
    if (L >= 0) {
      const int32_t* la = list + (size_t)a * KM_LMAX;
      for (int q = 0; q < L; ++q) {
        const int j = la[q];
        const double e = score(j);
        if (e < best || (e == best && j < bj)) {
          best = e;
          bj = j;
        }
      }
    } else {  // list overflow: full scan for this cluster's members (rare)
      for (int j = 0; j < s; ++j) {
        const double e = score(j);
        if (e < best || (e == best && j < bj)) {
          best = e;
          bj = j;
        }
      }
    }
    if (bj != a) {
      changed = 1;
      assign[i] = bj;
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
  }
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if ((threadIdx.x & 31) == 0 && changed) atomicAdd(&acc[(size_t)2 * s * D + s], (unsigned long long)changed);
}

template <int D>
void launch_pruned(Ctx* c, const double* X, int64_t n, int64_t ldx, const double* rec, const double* C, int s,
                   const Fx& fx, int32_t* assign, unsigned long long* acc, double M, double delta2, double eta,
                   unsigned long long* Rbits, int32_t* list, int32_t* len) {
  FLGP_CUDA(cudaMemsetAsync(Rbits, 0, sizeof(unsigned long long) * s, c->stream));
  const int grid = ceil_div(n, 256);
  if (grid > 0) FLGP_LAUNCH(c, (kmeans_radius_kernel<D>), grid, 256, 0, X, n, ldx, rec, assign, M, delta2, Rbits);
  FLGP_LAUNCH(c, kmeans_lists_kernel, s, 128, 0, C, s, D, Rbits, eta, list, len);
  if (grid > 0) FLGP_LAUNCH(c, (kmeans_assign_pruned<D>), grid, 256, 0, X, n, ldx, rec, s, fx, assign, acc, list, len);
}

template <int D>
void launch_small(Ctx* c, const double* X, int64_t n, int64_t ldx, const double* rec, int s, const Fx& fx,
                  int32_t* assign, unsigned long long* acc) {
  constexpr int P = 4;
  constexpr int STR = (D + 2) / 2 * 2;
  int chunk = std::min(s, 1024);
  size_t smem = (size_t)chunk * STR * sizeof(double);
  int grid = ceil_div(n, (int64_t)KM_THREADS * P);
  if (grid < 1) return;
  FLGP_LAUNCH(c, (kmeans_assign_small<D, P>), grid, KM_THREADS, smem, X, n, ldx, rec, s, fx, assign, acc, chunk, 1);
}

}  // namespace

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
                int* iters_out) {
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
  unsigned long long* uacc = reinterpret_cast<unsigned long long*>(acc.p);
  // pruned iterations (small d): persistent local accumulators + neighbour lists; see the proof above
  const bool pruned = small && n_total >= 2;
  const double bound = Moff + 4.0 * d * (maxabs * maxabs);                    // >= |score| terms
  const double Delta = 8.0 * (d + 4) * 1.1102230246251565e-16 * bound;        // generous bound on a score's rounding error
  const double delta2 = 4.0 * Delta;                                          // slack added to squared distances
  const double eta = 2.0 * std::sqrt(2.0 * Delta) + 1e-12 * maxabs;           // eta^2 > 2 Delta with room to spare
  DevBuf<unsigned long long> Rbits;
  DevBuf<int32_t> nlist, nlen;
  DevBuf<long long> acc_red;  // all-reduced copy of the local accumulators (multi-GPU)
  if (pruned) {
    Rbits.alloc(s);
    nlist.alloc((size_t)s * KM_LMAX);
    nlen.alloc(s);
    if (c->nranks > 1) acc_red.alloc(words);
  }
  int it = 0;
  while (it < iter_max) {
    ++it;
    const bool brute = !pruned || it == 1;
    if (brute) acc.zero(c->stream);
    else FLGP_CUDA(cudaMemsetAsync(acc.p + (words - 1), 0, sizeof(long long), c->stream));  // the `changed` slot
    if (small) FLGP_LAUNCH(c, kmeans_prep_kernel, ceil_div(s, 128), 128, 0, C, s, d, str, Moff, rec.p);
    else FLGP_LAUNCH(c, kmeans_prep_tiled_kernel, ceil_div(s, 128), 128, 0, C, s, d, Moff, C2.p, cn.p);
    // when timing is on, the assign+accumulate work gets its own CUDA-event pair per iteration
    StageScope kst(c, brute ? "kmeans_assign_kernel" : "kmeans_pruned_iteration", 2.0 * s * d * (double)n_local,
                   (8.0 * d + 4.0) * (double)n_local);
    if (!brute) {
      switch (d) {
        case 1: launch_pruned<1>(c, X, n_local, ldx, rec.p, C, s, fx, assign, uacc, Moff, delta2, eta, Rbits.p, nlist.p, nlen.p); break;
        case 2: launch_pruned<2>(c, X, n_local, ldx, rec.p, C, s, fx, assign, uacc, Moff, delta2, eta, Rbits.p, nlist.p, nlen.p); break;
        case 3: launch_pruned<3>(c, X, n_local, ldx, rec.p, C, s, fx, assign, uacc, Moff, delta2, eta, Rbits.p, nlist.p, nlen.p); break;
        default: launch_pruned<4>(c, X, n_local, ldx, rec.p, C, s, fx, assign, uacc, Moff, delta2, eta, Rbits.p, nlist.p, nlen.p); break;
      }
    } else if (small) {
      switch (d) {
        case 1: launch_small<1>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc); break;
        case 2: launch_small<2>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc); break;
        case 3: launch_small<3>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc); break;
        default: launch_small<4>(c, X, n_local, ldx, rec.p, s, fx, assign, uacc); break;
      }
    } else {
      int grid = ceil_div(n_local, KT_TP);
      if (grid > 0)
        FLGP_LAUNCH(c, kmeans_assign_tiled, grid, 256, 0, X, n_local, ldx, d, C2.p, cn.p, s, fx, assign, uacc);
    }
    kst.stop();
    long long* red = acc.p;
    if (pruned && c->nranks > 1) {  // keep the local sums intact; reduce a copy
      FLGP_CUDA(cudaMemcpyAsync(acc_red.p, acc.p, sizeof(long long) * words, cudaMemcpyDeviceToDevice, c->stream));
      red = acc_red.p;
    }
    comm_allreduce_i64(c, reinterpret_cast<int64_t*>(red), words);
    FLGP_LAUNCH(c, kmeans_update_kernel, ceil_div(s * d, 256), 256, 0, red, s, d, fx, C, sizes);
    FLGP_CUDA(cudaMemcpyAsync(c->pinned, red + (words - 1), sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    sync(c);
    if (c->pinned[0] == 0) break;  // no assignment changed anywhere
  }
  if (iters_out) *iters_out = it;
}

}  // namespace flgp
