// debug harness 2: the real lae.cu kernels vs host execution of lae_solve on the same rows
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "../../flgp_b200/csrc/lae.cu"
namespace flgp {
void comm_allreduce_i64(Ctx*, int64_t*, size_t) {}
void comm_allreduce_f64(Ctx*, double*, size_t) {}
void comm_allreduce_max_f64(Ctx*, double*, size_t) {}
void comm_destroy(Ctx*) {}
}
using namespace flgp;
struct PX { const double* x; double operator()(int k) const { return x[k]; } };
struct PU { const double* U; int ld; double operator()(int a, int k) const { return U[a + (size_t)ld * k]; } };

int main() {
  Ctx c; cudaStreamCreate(&c.stream);
  const int n = 1500, d = 6, s = 60, r = 3;
  std::vector<double> X((size_t)n * d), U((size_t)s * d);
  srand(3);
  auto rnd = [] { return rand() / (double)RAND_MAX * 2 - 1; };
  for (auto& x : X) x = rnd() * 2;
  for (int j = 0; j < s; ++j) for (int k = 0; k < d; ++k) U[j + (size_t)s * k] = X[(j * 25) + (size_t)n * k] + 0.05 * rnd();
  std::vector<int32_t> ind((size_t)n * r);
  for (int i = 0; i < n; ++i) {
    std::vector<double> D(s); std::vector<int> id(s); std::iota(id.begin(), id.end(), 0);
    for (int j = 0; j < s; ++j) { double a = 0; for (int k = 0; k < d; ++k) { double t = X[i + (size_t)n * k] - U[j + (size_t)s * k]; a += t * t; } D[j] = a; }
    std::partial_sort(id.begin(), id.begin() + r, id.end(), [&](int a, int b) { return D[a] < D[b]; });
    for (int a = 0; a < r; ++a) ind[i + (size_t)n * a] = id[a];
  }
  DevBuf<double> dX(X.size()), dU(U.size()), dZx((size_t)n * r), dW((size_t)n * r);
  DevBuf<int32_t> dind(ind.size()), dZj((size_t)n * r);
  DevBuf<long long> st(2); st.zero(c.stream);
  dX.upload(X.data(), X.size(), c.stream); dU.upload(U.data(), U.size(), c.stream); dind.upload(ind.data(), ind.size(), c.stream);
  lae_run(&c, dX.p, n, n, d, dU.p, s, s, r, dind.p, dZj.p, dZx.p, dW.p, st.p);
  std::vector<double> Zx((size_t)n * r), W((size_t)n * r); std::vector<int32_t> Zj((size_t)n * r);
  dZx.download(Zx.data(), Zx.size(), c.stream); dZj.download(Zj.data(), Zj.size(), c.stream); dW.download(W.data(), W.size(), c.stream);
  cudaStreamSynchronize(c.stream);
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  int badW = 0, badZ = 0;
  for (int i = 0; i < n; ++i) {
    std::vector<double> Ur((size_t)r * d), x(d);
    for (int k = 0; k < d; ++k) { x[k] = X[i + (size_t)n * k]; for (int a = 0; a < r; ++a) Ur[a + (size_t)r * k] = U[ind[i + (size_t)n * a] + (size_t)s * k]; }
    double zz[LAE_RMAX]; int it, bt;
    PX xa{x.data()}; PU ua{Ur.data(), r};
    LaeStats ls = lae_solve<0, 0>(r, d, xa, ua, zz); it = ls.iters; bt = ls.backtracks;
    for (int a = 0; a < r; ++a) {
      if (W[i + (size_t)n * a] != zz[a]) { if (badW++ < 6) printf(" W mismatch i=%d a=%d host %.17g dev %.17g diff %.3g it=%d\n", i, a, zz[a], W[i + (size_t)n * a], zz[a] - W[i + (size_t)n * a], it); }
      // find in CSR
      bool found = false;
      for (int b = 0; b < r; ++b) if (Zj[(size_t)i * r + b] == ind[i + (size_t)n * a]) { found = true; if (Zx[(size_t)i * r + b] != zz[a]) { if (badZ++ < 6) printf(" Z mismatch i=%d a=%d host %.17g dev %.17g\n", i, a, zz[a], Zx[(size_t)i * r + b]); } }
      if (!found) printf(" column missing i=%d a=%d\n", i, a);
    }
  }
  printf("bad W %d bad Z %d of %d rows\n", badW, badZ, n);
  return 0;
}
