// gemm.cu — small/medium dense fp64 products of the path: the heat-kernel covariance
// H = V0 diag(exp(-t(1-lambda))) V1^T  (HK_from_spectrum_cpp, /root/reference/src/Spectrum.cpp:83-94),
// the K x K blocks of the GPR tail and the folds W M W^T.  One strided, shared-memory tiled FMA
// kernel (64 x 64 tile, 4 x 4 per thread, ascending-k accumulation => deterministic).
// The diagonal scaling is fused into the A-operand load, as the reference scales V0's columns first.
#include "kernels.cuh"

namespace flgp {

namespace {

constexpr int GT = 64, GK = 16, GPAD = 66;

// C(i,j) = sum_k (A(i,k) * sc[k]) * B(k,j);  element strides for every operand.
__global__ void __launch_bounds__(256)
gemm_strided_kernel(const double* __restrict__ A, int64_t ars, int64_t acs, const double* __restrict__ B,
                    int64_t brs, int64_t bcs, const double* __restrict__ sc, int64_t M, int64_t N, int K,
                    double* __restrict__ C, int64_t crs, int64_t ccs) {
  __shared__ __align__(16) double As[GK][GPAD];
  __shared__ __align__(16) double Bs[GK][GPAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * GT, j0 = (int64_t)blockIdx.x * GT;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  const bool a_kfast = (acs == 1), b_kfast = (brs == 1);
  for (int k0 = 0; k0 < K; k0 += GK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + q * 256;
      int kk = a_kfast ? (e & 15) : (e >> 6), m = a_kfast ? (e >> 4) : (e & 63);
      int k = k0 + kk;
      int64_t i = i0 + m;
      double v = 0.0;
      if (k < K && i < M) {
        v = A[i * ars + k * acs];
        if (sc) v = __dmul_rn(v, sc[k]);
      }
      As[kk][m] = v;
      kk = b_kfast ? (e & 15) : (e >> 6);
      m = b_kfast ? (e >> 4) : (e & 63);
      k = k0 + kk;
      int64_t j = j0 + m;
      Bs[kk][m] = (k < K && j < N) ? B[k * brs + j * bcs] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double xa[4], xb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = As[kk][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) xb[b] = Bs[kk][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(xa[a], xb[b], acc[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int64_t i = i0 + ty * 4 + a, j = j0 + tx * 4 + b;
      if (i < M && j < N) C[i * crs + j * ccs] = acc[a][b];
    }
}

void gemm_strided(Ctx* c, const double* A, int64_t ars, int64_t acs, const double* B, int64_t brs, int64_t bcs,
                  const double* sc, int64_t M, int64_t N, int K, double* C, int64_t crs, int64_t ccs) {
  if (M <= 0 || N <= 0) return;
  dim3 grid(ceil_div(N, GT), ceil_div(M, GT));
  FLGP_LAUNCH(c, gemm_strided_kernel, grid, 256, 0, A, ars, acs, B, brs, bcs, sc, M, N, K, C, crs, ccs);
}

}  // namespace

void gemm_nt_run(Ctx* c, const double* A, const double* B, const double* sc, int64_t M, int64_t N, int K,
                 double* C, int64_t ldc) {
  gemm_strided(c, A, K, 1, B, 1, K, sc, M, N, K, C, 1, ldc);
}

void gemm_nn_run(Ctx* c, const double* A, const double* B, int64_t M, int64_t N, int K, double* C) {
  gemm_strided(c, A, K, 1, B, N, 1, nullptr, M, N, K, C, N, 1);
}

void gemv_run(Ctx* c, const double* A, const double* x, int64_t M, int K, double* y) {
  gemm_strided(c, A, K, 1, x, 1, 0, nullptr, M, 1, K, y, 1, 0);
}

void gram_small_run(Ctx* c, const double* V, const double* y, int64_t n_rows, int K, double* G, double* g) {
  // G = V^T V : A(i,k) = V[k*K + i], B(k,j) = V[k*K + j]
  if (n_rows > INT32_MAX) fail(2, "too many training rows");
  gemm_strided(c, V, 1, K, V, K, 1, nullptr, K, K, (int)n_rows, G, 1, K);
  if (y && g) gemm_strided(c, V, 1, K, y, 1, 0, nullptr, K, 1, (int)n_rows, g, 1, 0);
}

}  // namespace flgp
