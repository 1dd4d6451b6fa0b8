// tail.cu — the K x K algebra of the GPR tail at fixed hyper-parameters, on the device
// (m > K branch of predict_regression_cpp, /root/reference/src/Predict.cpp:61-74, and of
//  posterior_covariance_regression, /root/reference/src/Utils.cpp:237-247; the reference uses Eigen LLT).
//
// With G1 = V1^T V1 (K x K), g1 = V1^T Y, Ls = diag(exp(-t(1-lambda)/2)), Lam = Ls^2, ns = noise + sigma:
//   Q     = Ls G1 Ls + ns I                                  (symmetric positive definite)
//   X     = Q^-1 [ Ls G1 | Ls g1 ]                           (K x (K+1))
//   coef  = Lam (g1 - G1 Ls X[:,K]) / ns                     (= Lam V1^T alpha)
//   M     = Lam - Lam (G1 - G1 Ls X[:, :K]) Lam / ns         (posterior covariance in eigen-coordinates)
// One CTA factors Q (packed lower triangle, shared memory when it fits), the K+1 triangular solves run one warp per
// right-hand side, and a last kernel forms coef and M.  Nothing leaves the device, so the tail needs no
// host round trip between the Gram all-reduce and the n-sized folded products.
#include "kernels.cuh"

namespace flgp {

namespace {

__device__ __forceinline__ size_t pk(int i, int k, int K) {  // packed lower triangle, column-major: i >= k
  return (size_t)k * K - (size_t)k * (k - 1) / 2 + (i - k);
}

// Gg: K_ld x K_ld column-major G1 followed by g1 (K_ld).  Xt: K x NR row-major right-hand sides, NR = K + 1.
__global__ void __launch_bounds__(256)
tail_build_kernel(const double* __restrict__ Gg, int K_ld, int K, const double* __restrict__ ls, double ns,
                  double* __restrict__ Qp, double* __restrict__ Xt) {
  const int NR = K + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * NR) return;
  const int i = e / NR, j = e % NR;
  const double* g1 = Gg + (size_t)K_ld * K_ld;
  if (j < K) {
    const double g = Gg[i + (size_t)K_ld * j];
    Xt[(size_t)i * NR + j] = ls[i] * g;
    if (i >= j) Qp[pk(i, j, K)] = ls[i] * g * ls[j] + (i == j ? ns : 0.0);
  } else {
    Xt[(size_t)i * NR + j] = ls[i] * g1[i];
  }
}

// in-place Cholesky of the packed matrix (right-looking: scale column j, then rank-1 update of the trailing
// triangle, one warp per trailing column so that every shared-memory access runs along a column).
// One CTA; flag = 1 if a pivot is not positive.
__global__ void __launch_bounds__(1024)
tail_chol_kernel(double* __restrict__ Qp, int K, int use_smem, int* __restrict__ flag) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const size_t np = (size_t)K * (K + 1) / 2;
  double* L = Qp;
  if (use_smem) {
    for (size_t t = tid; t < np; t += nt) sm[t] = Qp[t];
    L = sm;
    __syncthreads();
  }
  for (int j = 0; j < K; ++j) {
    const double dj = L[pk(j, j, K)];
    if (!(dj > 0.0)) {  // uniform: every thread reads the same value
      if (tid == 0) *flag = 1;
      return;
    }
    const double rt = sqrt(dj), inv = 1.0 / rt;
    __syncthreads();  // everyone has read the pivot
    double* colj = L + pk(j, j, K);  // colj[q] = L(j + q, j)
    for (int q = tid; q < K - j; q += nt) colj[q] = (q == 0) ? rt : colj[q] * inv;
    __syncthreads();
    for (int k = j + 1 + wid; k < K; k += nw) {
      const double ljk = colj[k - j];
      double* colk = L + pk(k, k, K);  // colk[q] = L(k + q, k)
      for (int q = lane; q < K - k; q += 32) colk[q] = fma(-colj[k - j + q], ljk, colk[q]);
    }
    __syncthreads();
  }
  if (use_smem)
    for (size_t t = tid; t < np; t += nt) Qp[t] = sm[t];
}

// L L^T x = b for the columns of Xt: one warp per right-hand side (its vector in shared memory), 8 per CTA, L staged
// in shared memory when it fits.  Forward sweep in axpy form, backward sweep in dot form: both walk columns of L.
constexpr int TS_WARPS = 8;
__global__ void __launch_bounds__(TS_WARPS * 32)
tail_solve_kernel(const double* __restrict__ Lp, int K, double* __restrict__ Xt, int use_smem) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NR = K + 1;
  const size_t np = (size_t)K * (K + 1) / 2;
  double* bvec = sm + (size_t)wid * K;  // this warp's right-hand side
  const double* L = Lp;
  if (use_smem) {
    double* Ls = sm + (size_t)TS_WARPS * K;
    for (size_t t = tid; t < np; t += TS_WARPS * 32) Ls[t] = Lp[t];
    L = Ls;
  }
  const int c = blockIdx.x * TS_WARPS + wid;
  const bool live = c < NR;
  if (live)
    for (int i = lane; i < K; i += 32) bvec[i] = Xt[(size_t)i * NR + c];
  __syncthreads();
  if (!live) return;
  for (int k = 0; k < K; ++k) {  // L y = b
    const double* colk = L + pk(k, k, K);
    const double xk = bvec[k] / colk[0];
    __syncwarp();
    if (lane == 0) bvec[k] = xk;
    for (int q = 1 + lane; q < K - k; q += 32) bvec[k + q] = fma(-colk[q], xk, bvec[k + q]);
    __syncwarp();
  }
  for (int i = K - 1; i >= 0; --i) {  // L^T x = y
    const double* coli = L + pk(i, i, K);
    double acc = 0.0;
    for (int q = 1 + lane; q < K - i; q += 32) acc = fma(coli[q], bvec[i + q], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const double xi = (bvec[i] - acc) / coli[0];
    __syncwarp();
    if (lane == 0) bvec[i] = xi;
    __syncwarp();
  }
  for (int i = lane; i < K; i += 32) Xt[(size_t)i * NR + c] = bvec[i];
}

// coef (K_ld, zero padded) and M (K_ld x K_ld, zero padded; symmetric)
__global__ void __launch_bounds__(256)
tail_finish_kernel(const double* __restrict__ Gg, int K_ld, int K, const double* __restrict__ ls,
                   const double* __restrict__ lam, double ns, const double* __restrict__ Xt,
                   double* __restrict__ coef, double* __restrict__ M) {
  const int NR = K + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * NR) return;
  const int i = e / NR, j = e % NR;
  const double* g1 = Gg + (size_t)K_ld * K_ld;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc = fma(Gg[i + (size_t)K_ld * k] * ls[k], Xt[(size_t)k * NR + j], acc);
  if (j == K) {
    coef[i] = lam[i] * ((g1[i] - acc) / ns);
  } else {
    const double am = lam[i] * (Gg[i + (size_t)K_ld * j] - acc) * lam[j] / ns;
    M[i + (size_t)K_ld * j] = (i == j ? lam[i] : 0.0) - am;
  }
}

}  // namespace

void tail_woodbury_run(Ctx* c, const double* Gg, int K_ld, int K, const double* ls, const double* lam, double ns,
                       double* coef, double* M, int* flag) {
  const int NR = K + 1;
  const size_t np = (size_t)K * (K + 1) / 2;
  DevBuf<double> Qp(np), Xt((size_t)K * NR);
  FLGP_CUDA(cudaMemsetAsync(coef, 0, sizeof(double) * K_ld, c->stream));
  FLGP_CUDA(cudaMemsetAsync(M, 0, sizeof(double) * (size_t)K_ld * K_ld, c->stream));
  FLGP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  FLGP_LAUNCH(c, tail_build_kernel, ceil_div((int64_t)K * NR, 256), 256, 0, Gg, K_ld, K, ls, ns, Qp.p, Xt.p);
  {
    const size_t smem = np * sizeof(double);
    const int use_smem = smem <= 200 * 1024;
    if (use_smem)
      FLGP_CUDA(cudaFuncSetAttribute(tail_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLGP_LAUNCH(c, tail_chol_kernel, 1, 1024, use_smem ? smem : 0, Qp.p, K, use_smem, flag);
  }
  {
    const size_t vec = (size_t)TS_WARPS * K * sizeof(double);
    const int use_smem = vec + np * sizeof(double) <= 200 * 1024;
    const size_t smem = vec + (use_smem ? np * sizeof(double) : 0);
    if (smem > 48 * 1024)
      FLGP_CUDA(cudaFuncSetAttribute(tail_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLGP_LAUNCH(c, tail_solve_kernel, ceil_div(NR, TS_WARPS), TS_WARPS * 32, smem, Qp.p, K, Xt.p, use_smem);
  }
  FLGP_LAUNCH(c, tail_finish_kernel, ceil_div((int64_t)K * NR, 256), 256, 0, Gg, K_ld, K, ls, lam, ns, Xt.p, coef, M);
  sync(c);  // Qp / Xt are released on return
}

}  // namespace flgp
