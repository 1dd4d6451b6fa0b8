// hostcheck.cpp — TEST-ONLY host compilation of flgp_b200/csrc/core_math.cuh.
// The per-thread arithmetic of the CUDA kernels (heap emulation of std::partial_sort, simplex
// projection, LAE solver, fixed-point codec, Sturm count) is written `__host__ __device__`; this shim
// compiles the very same source with g++ so that the not-gpu test suite can compare it with the
// oracle bit for bit before any GPU time is spent.  It is never part of the product library.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../flgp_b200/csrc/core_math.cuh"

using namespace flgp;

namespace {
struct PtrX {
  const double* x;
  double operator()(int k) const { return x[k]; }
};
struct PtrU {
  const double* U;
  int ld;
  double operator()(int a, int k) const { return U[a + (size_t)ld * k]; }
};
template <int R, int D>
void lae_fixed(const double* x, const double* U, double* z, int* it, int* bt) {
  PtrX xa{x};
  PtrU ua{U, R};
  static double tab[100];
  static bool init = (lae_alpha_table(tab), true);
  (void)init;
  LaeStats st = lae_solve<R, D>(R, D, xa, ua, z, tab);
  *it = st.iters;
  *bt = st.backtracks;
}
}  // namespace

// the register-heap variant used by the small-d KNN kernel (compile-time r)
template <int R>
static void topr_reg(const double* row, int s, int32_t* ind, double* key) {
  RegHeap<R> h;
  for (int j = 0; j < R; ++j) h.set(j, row[j], j);
  heap_make_acc(h, R);
  double top = h.getk(0);
  for (int j = R; j < s; ++j)
    if (row[j] < top) {
      heap_adjust_acc(h, 0, R, row[j], j);
      top = h.getk(0);
    }
  heap_sort_acc(h, R);
  for (int a = 0; a < R; ++a) {
    ind[a] = h.geti(a);
    key[a] = h.getk(a);
  }
}

extern "C" {

// the TopR state machine of knn.cu, fed a precomputed distance row in anchor order
void hc_topr(const double* row, int s, int r, int32_t* ind, double* key) {
  std::vector<double> hk(r);
  std::vector<int> hi(r);
  double top = INFINITY;
  for (int j = 0; j < s; ++j) {
    double dist = row[j];
    if (j < r) {
      hk[j] = dist;
      hi[j] = j;
      if (j == r - 1) {
        heap_make(hk.data(), hi.data(), r);
        top = hk[0];
      }
    } else if (dist < top) {
      heap_adjust(hk.data(), hi.data(), 0, r, dist, j);
      top = hk[0];
    }
  }
  heap_sort(hk.data(), hi.data(), r);
  for (int a = 0; a < r; ++a) {
    ind[a] = hi[a];
    key[a] = hk[a];
  }
}

int hc_topr_reg(const double* row, int s, int r, int32_t* ind, double* key) {
  switch (r) {
    case 1: topr_reg<1>(row, s, ind, key); return 0;
    case 2: topr_reg<2>(row, s, ind, key); return 0;
    case 3: topr_reg<3>(row, s, ind, key); return 0;
    case 4: topr_reg<4>(row, s, ind, key); return 0;
    case 5: topr_reg<5>(row, s, ind, key); return 0;
    case 6: topr_reg<6>(row, s, ind, key); return 0;
    case 7: topr_reg<7>(row, s, ind, key); return 0;
    case 8: topr_reg<8>(row, s, ind, key); return 0;
  }
  return 1;
}

void hc_simplex(const double* v, int r, double* z) {
  std::vector<double> scratch(r);
  simplex_project<0>(v, r, z, scratch.data());
}

// generic (run-time r, d) and the compile-time instantiations the kernels use
int hc_lae(const double* x, int d, const double* Ur, int r, int fixed, double* z, int* it, int* bt) {
  if (r > LAE_RMAX) return 1;
  if (fixed) {
    if (r == 3 && d == 3) { lae_fixed<3, 3>(x, Ur, z, it, bt); return 0; }
    if (r == 3 && d == 2) { lae_fixed<3, 2>(x, Ur, z, it, bt); return 0; }
    if (r == 5 && d == 3) { lae_fixed<5, 3>(x, Ur, z, it, bt); return 0; }
    if (r == 2 && d == 2) { lae_fixed<2, 2>(x, Ur, z, it, bt); return 0; }
    return 2;
  }
  PtrX xa{x};
  PtrU ua{Ur, r};
  double zz[LAE_RMAX];
  LaeStats st = lae_solve<0, 0>(r, d, xa, ua, zz);
  *it = st.iters;
  *bt = st.backtracks;
  for (int a = 0; a < r; ++a) z[a] = zz[a];
  return 0;
}

int hc_fx_roundtrip(double maxabs, int64_t count, const double* x, int64_t len, long long* hi, long long* lo,
                    double* back) {
  Fx fx;
  if (fx_make(maxabs, count, &fx)) return 1;
  for (int64_t i = 0; i < len; ++i) {
    fx_encode(fx, x[i], hi + i, lo + i);
    back[i] = fx_decode(fx, hi[i], lo[i]);
  }
  return 0;
}

int hc_sturm(const double* d, const double* e2, int n, double x, double pivmin) {
  return sturm_count(d, e2, n, x, pivmin);
}

// ---- mini-batch k-means (csrc/minibatch.cu): the kernels' per-thread pieces run by "virtual threads" ----------------
int64_t hc_mb_perm(int64_t k, int64_t n, uint64_t key) { return mb_perm(k, n, key); }
uint64_t hc_mb_batch_key(uint64_t seed, int it) { return mb_batch_key(seed, it); }

namespace {
struct HostBatchRow {
  const double* Xb;
  int64_t k, b;
  double operator()(int q) const { return Xb[k + b * q]; }
};
}  // namespace

// mb_assign_kernel: every batch row walks the centre groups exactly as a CUDA thread does (same staging of the
// zero-padded records, same group / chunk sizes as the kernel: 8 centres x 16 coordinates)
void hc_mb_assign(const double* Xb, int64_t b, int d, const double* C, int s, int64_t ldc, double m2,
                  int32_t* assign) {
  constexpr int G = 8, QC = 16;
  const int dp = (d + QC - 1) / QC * QC;
  std::vector<double> rec((size_t)G * dp), cn(G);
  std::vector<double> best(b, 0.0);
  std::vector<int> bj(b, 0);
  for (int j0 = 0; j0 < s; j0 += G) {
    for (int t = 0; t < G * dp; ++t) {
      const int jj = t / dp, q = t - jj * dp;
      const int j = j0 + jj;
      rec[t] = (j < s && q < d) ? -2.0 * C[j + ldc * q] : 0.0;
    }
    for (int g = 0; g < G; ++g) cn[g] = (j0 + g < s) ? mb_centre_norm(C, ldc, j0 + g, d, m2) : 0.0;
    for (int64_t k = 0; k < b; ++k) {
      double e[G];
      HostBatchRow xk{Xb, k, b};
      mb_score_group<G, QC>(xk, d, dp, rec.data(), cn.data(), e);
      mb_argmin_group<G>(e, j0, s, &best[k], &bj[k]);
    }
  }
  for (int64_t k = 0; k < b; ++k) assign[k] = bj[k];
}

// minibatch_kmeans_run's loop with the kernels replaced by their host images: gather by mb_perm, hc_mb_assign,
// mb_update_kernel's warp (32 virtual lanes per centre, members in batch order), mb_delta_kernel's sums.
int hc_minibatch_kmeans(const double* X, int64_t n, int64_t ldx, int d, int s, const int32_t* init, int max_iters,
                        uint64_t seed, double maxabs, double* C /* s x d, ld s */) {
  const int64_t b = std::min<int64_t>((int64_t)10 * s, n), ldc = s;
  const double m2 = (2.0 * d) * (maxabs * maxabs);
  for (int t = 0; t < s * d; ++t) {
    const int q = t / s, j = t - q * s;
    C[j + ldc * q] = X[init[j] + ldx * q];
  }
  std::vector<double> Xb((size_t)b * d), Cold((size_t)s * d), dsq(s);
  std::vector<int32_t> assign(b);
  std::vector<long long> cnt(s, 0);
  int it = 0, calm = 0;
  while (it < max_iters) {
    const uint64_t key = mb_batch_key(seed, it);
    for (int64_t k = 0; k < b; ++k) {
      const int64_t row = mb_perm(k, n, key);
      for (int q = 0; q < d; ++q) Xb[k + b * q] = X[row + ldx * q];
    }
    hc_mb_assign(Xb.data(), b, d, C, s, ldc, m2, assign.data());
    for (int t = 0; t < s * d; ++t) {
      const int q = t / s, j = t - q * s;
      Cold[t] = C[j + ldc * q];
    }
    for (int j = 0; j < s; ++j) {  // one warp per centre
      long long n_j = cnt[j];
      bool touched = false;
      for (int64_t k0 = 0; k0 < b; k0 += 32) {
        unsigned bal = 0;
        for (int lane = 0; lane < 32; ++lane) {
          const int64_t k = k0 + lane;
          const int a = (k < b) ? assign[k] : -1;
          if (a == j) bal |= 1u << lane;
        }
        while (bal) {
          const int src = __builtin_ffs((int)bal) - 1;
          bal &= bal - 1;
          const int64_t km = k0 + src;
          n_j += 1;
          const double eta = 1.0 / (double)n_j;
          for (int lane = 0; lane < 32; ++lane)
            for (int q = lane; q < d; q += 32) C[j + ldc * q] = mb_update_coord(C[j + ldc * q], Xb[km + b * q], eta);
          touched = true;
        }
      }
      cnt[j] = n_j;
      double dj = 0.0;
      if (touched)
        for (int q = 0; q < d; ++q) {
          const double df = C[j + ldc * q] - Cold[j + (int64_t)s * q];
          dj = dj + df * df;
        }
      dsq[j] = dj;
    }
    double dl = 0.0;
    for (int j = 0; j < s; ++j) dl = dl + dsq[j];
    ++it;
    calm = (dl < 1e-4) ? calm + 1 : 0;
    if (calm >= 10) break;
  }
  return it;
}

}  // extern "C"
