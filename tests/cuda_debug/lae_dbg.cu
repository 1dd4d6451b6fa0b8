// debug harness: device vs host execution of the SAME core_math.cuh functions
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../flgp_b200/csrc/core_math.cuh"
using namespace flgp;

struct PX { const double* x; __host__ __device__ double operator()(int k) const { return x[k]; } };
struct PU { const double* U; int ld; __host__ __device__ double operator()(int a, int k) const { return U[a + (size_t)ld * k]; } };

__global__ void k_simplex(const double* v, int r, int n, double* z) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double vv[LAE_RMAX], zz[LAE_RMAX], sc[LAE_RMAX];
  for (int a = 0; a < r; ++a) vv[a] = v[(size_t)i * r + a];
  simplex_project<0>(vv, r, zz, sc);
  for (int a = 0; a < r; ++a) z[(size_t)i * r + a] = zz[a];
}
__global__ void k_lae(const double* X, const double* U, int r, int d, int n, double* Z, int* its) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PX x{X + (size_t)i * d};
  PU u{U + (size_t)i * r * d, r};
  double zz[LAE_RMAX];
  int it = 0, bt = 0;
  lae_solve<0, 0>(r, d, x, u, zz, &it, &bt);
  for (int a = 0; a < r; ++a) Z[(size_t)i * r + a] = zz[a];
  its[2 * i] = it; its[2 * i + 1] = bt;
}

int main() {
  const int n = 4096;
  for (int cfg = 0; cfg < 3; ++cfg) {
    int r = cfg == 0 ? 3 : cfg == 1 ? 7 : 5, d = cfg == 0 ? 6 : cfg == 1 ? 3 : 40;
    std::vector<double> X((size_t)n * d), U((size_t)n * r * d), V((size_t)n * r);
    srand(1 + cfg);
    auto rnd = [] { return rand() / (double)RAND_MAX * 2 - 1; };
    for (auto& u : U) u = rnd() * 2;
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < d; ++k) {
        double m = 0; for (int a = 0; a < r; ++a) m += U[(size_t)i * r * d + a + (size_t)r * k];
        X[(size_t)i * d + k] = m / r + 0.3 * rnd();
      }
    for (auto& v : V) v = rnd() * 3;
    double *dX, *dU, *dV, *dZ, *dZs; int* dI;
    cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dU, U.size() * 8); cudaMalloc(&dV, V.size() * 8);
    cudaMalloc(&dZ, (size_t)n * r * 8); cudaMalloc(&dZs, (size_t)n * r * 8); cudaMalloc(&dI, n * 8);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dU, U.data(), U.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, V.data(), V.size() * 8, cudaMemcpyHostToDevice);
    k_simplex<<<n / 128, 128>>>(dV, r, n, dZs);
    k_lae<<<n / 128, 128>>>(dX, dU, r, d, n, dZ, dI);
    std::vector<double> Z((size_t)n * r), Zs((size_t)n * r); std::vector<int> I(2 * n);
    cudaMemcpy(Z.data(), dZ, Z.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Zs.data(), dZs, Zs.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(I.data(), dI, n * 8, cudaMemcpyDeviceToHost);
    printf("cfg r=%d d=%d: %s\n", r, d, cudaGetErrorString(cudaGetLastError()));
    int bad_s = 0, bad_l = 0, bad_it = 0;
    for (int i = 0; i < n; ++i) {
      double zz[LAE_RMAX], sc[LAE_RMAX];
      simplex_project<0>(&V[(size_t)i * r], r, zz, sc);
      for (int a = 0; a < r; ++a) if (zz[a] != Zs[(size_t)i * r + a]) { if (bad_s++ < 3) printf(" simplex mismatch i=%d a=%d host %.17g dev %.17g\n", i, a, zz[a], Zs[(size_t)i * r + a]); }
      PX x{&X[(size_t)i * d]}; PU u{&U[(size_t)i * r * d], r};
      int it, bt;
      lae_solve<0, 0>(r, d, x, u, zz, &it, &bt);
      if (it != I[2 * i] || bt != I[2 * i + 1]) { if (bad_it++ < 3) printf(" iters mismatch i=%d host (%d,%d) dev (%d,%d)\n", i, it, bt, I[2 * i], I[2 * i + 1]); }
      for (int a = 0; a < r; ++a) if (zz[a] != Z[(size_t)i * r + a]) { if (bad_l++ < 5) printf(" lae mismatch i=%d a=%d host %.17g dev %.17g (it %d/%d)\n", i, a, zz[a], Z[(size_t)i * r + a], it, I[2 * i]); }
    }
    printf(" simplex bad %d, lae bad %d, iters bad %d of %d\n", bad_s, bad_l, bad_it, n);
  }
  return 0;
}
