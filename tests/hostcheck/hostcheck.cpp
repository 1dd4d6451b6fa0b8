// hostcheck.cpp — TEST-ONLY host compilation of flgp_b200/csrc/core_math.cuh.
// The per-thread arithmetic of the CUDA kernels (heap emulation of std::partial_sort, simplex
// projection, LAE solver, fixed-point codec, Sturm count) is written `__host__ __device__`; this shim
// compiles the very same source with g++ so that the not-gpu test suite can compare it with the
// oracle bit for bit before any GPU time is spent.  It is never part of the product library.
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

}  // extern "C"
