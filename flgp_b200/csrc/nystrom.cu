// nystrom.cu — the Nystrom-extension baseline of FLGP (fit_nystrom_regression_gp_cpp,
// /root/reference/src/Fit.cpp:222-357), SURVEY.md §8f row 4.
//
//   anchors U (s x d)  ->  Z_UU = exp(-D_UU / (a2 mean D_UU)), doubly normalised to W_UU (s x s, dense, symmetric)
//   top-K eigenpairs of W_UU (the RSpectra::eigs_sym callback: eigh.cu) -> rescaled anchor eigenvectors
//   every row x: z = exp(-D(x,U) / (a2 mean)), normalised twice -> w (s);  v(x) = w . vecs / (|lambda| + 1e-9)
//
// The n x s kernel block is never kept: rows are processed in blocks; one CTA turns a row of X into its normalised
// weight row (distances in the reference's operation order, exp, both row sums — all in shared memory) and the
// block's n_b x s weights meet the s x K anchor eigenvectors in the FP64 tensor-core GEMM of gemm.cu (2 n s K flop:
// the one genuinely GEMM-bound stage of the reference).
#include "kernels.cuh"

namespace flgp {

namespace {

__device__ __forceinline__ double nys_block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : 0.0;
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// squared distances in the reference's order: ((-2 sum_k x_k u_k) + |x|^2) + |u|^2, sequential separately rounded sums
__device__ __forceinline__ double nys_dist(const double* xs, double xn, const double* __restrict__ U, int64_t ldu, int j,
                                           int d, double unj) {
  double dot = 0.0;
  for (int k = 0; k < d; ++k) dot = __dadd_rn(dot, __dmul_rn(xs[k], U[j + ldu * k]));
  return __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), xn), unj);
}

__global__ void nys_rownorm_kernel(const double* __restrict__ U, int s, int64_t ldu, int d, double* un) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s) return;
  double a = 0.0;
  for (int k = 0; k < d; ++k) {
    const double u = U[j + ldu * k];
    a = __dadd_rn(a, __dmul_rn(u, u));
  }
  un[j] = a;
}

// D_UU (s x s, row-major = row i contiguous), one CTA per row.  dynamic shared memory: d doubles
__global__ void __launch_bounds__(256)
nys_dist_uu_kernel(const double* __restrict__ U, int s, int64_t ldu, int d, const double* __restrict__ un,
                   double* __restrict__ D) {
  extern __shared__ double nsm[];
  const int i = blockIdx.x, tid = threadIdx.x;
  for (int k = tid; k < d; k += 256) nsm[k] = U[i + ldu * k];
  __syncthreads();
  const double xn = un[i];
  for (int j = tid; j < s; j += 256) D[(size_t)i * s + j] = nys_dist(nsm, xn, U, ldu, j, d, un[j]);
}

// Z = exp(-D / denom) in place; rs[i] = rowsum + 1e-9
__global__ void __launch_bounds__(256)
nys_uu_z_kernel(const double* __restrict__ D, int s, double denom, double* __restrict__ Z, double* __restrict__ rs) {
  __shared__ double red[32];
  const int i = blockIdx.x, tid = threadIdx.x;
  double part = 0.0;
  for (int j = tid; j < s; j += 256) {
    const double z = exp(-D[(size_t)i * s + j] / denom);
    Z[(size_t)i * s + j] = z;
    part += z;
  }
  const double t = nys_block_sum(part, red);
  if (tid == 0) rs[i] = t + 1e-9;
}
// A = diag(1/rs) Z diag(1/rs) in place (lower-triangle arithmetic mirrored, so A is exactly symmetric);
// dsum[i] = rowsum(A) + 1e-9
__global__ void __launch_bounds__(256)
nys_uu_a_kernel(const double* __restrict__ Z, int s, const double* __restrict__ rs, double* __restrict__ A,
                double* __restrict__ dsum) {
  __shared__ double red[32];
  const int i = blockIdx.x, tid = threadIdx.x;
  double part = 0.0;
  for (int j = tid; j < s; j += 256) {
    const int hi = max(i, j), lo = min(i, j);
    const double a = ((1.0 / rs[hi]) * Z[(size_t)hi * s + lo]) * (1.0 / rs[lo]);
    A[(size_t)i * s + j] = a;
    part += a;
  }
  const double t = nys_block_sum(part, red);
  if (tid == 0) dsum[i] = t + 1e-9;
}
// W = diag(1/sqrt(dsum)) A diag(1/sqrt(dsum)), symmetric by the same mirroring
__global__ void nys_uu_w_kernel(const double* __restrict__ A, int s, const double* __restrict__ dsum,
                                double* __restrict__ W) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e >= (size_t)s * s) return;
  const int i = (int)(e / s), j = (int)(e % s);
  const int hi = max(i, j), lo = min(i, j);
  W[e] = ((1.0 / sqrt(dsum[hi])) * A[(size_t)hi * s + lo]) * (1.0 / sqrt(dsum[lo]));
}
// anchor eigenvectors as the extension uses them: vecs = sqrt(s) (sdi Y) / (colnorm + 1e-9), then / (|lambda| + 1e-9);
// written transposed, Bt (K x s row-major), the operand layout of the GEMM.  One CTA per eigenvector.
__global__ void __launch_bounds__(256)
nys_vecs_kernel(const double* __restrict__ Y, int s, int K, const double* __restrict__ dsum,
                const double* __restrict__ lam, double* __restrict__ Bt) {
  __shared__ double red[32];
  const int k = blockIdx.x, tid = threadIdx.x;
  double part = 0.0;
  for (int j = tid; j < s; j += 256) {
    const double v = (1.0 / sqrt(dsum[j])) * Y[j + (size_t)s * k];
    part = fma(v, v, part);
  }
  const double nrm = sqrt(nys_block_sum(part, red));
  const double sc = 1.0 / (nrm + 1e-9), il = 1.0 / (fabs(lam[k]) + 1e-9), rt = sqrt((double)s);
  for (int j = tid; j < s; j += 256) {
    const double v = (1.0 / sqrt(dsum[j])) * Y[j + (size_t)s * k];
    Bt[(size_t)k * s + j] = ((rt * v) * sc) * il;
  }
}

// One CTA per row of X (rows row0 .. row0 + nb - 1): the row's normalised extension weights
//   z_j = exp(-D(x, u_j) / denom);  a_j = (z_j / (sum z + 1e-9)) * (1 / rs_j);  w_j = a_j / (sum a + 1e-9)
// Wx: nb x s row-major.  dynamic shared memory: d + s doubles.
__global__ void __launch_bounds__(256)
nys_rows_kernel(const double* __restrict__ X, int64_t ldx, int64_t row0, int d, const double* __restrict__ U, int s,
                int64_t ldu, const double* __restrict__ un, const double* __restrict__ rs, double denom,
                double* __restrict__ Wx) {
  extern __shared__ double nsm[];
  __shared__ double red[32];
  double* xs = nsm;
  double* zs = nsm + d;
  const int tid = threadIdx.x;
  const int64_t i = row0 + blockIdx.x;
  for (int k = tid; k < d; k += 256) xs[k] = X[i + ldx * k];
  __syncthreads();
  double xn = 0.0;
  for (int k = 0; k < d; ++k) xn = __dadd_rn(xn, __dmul_rn(xs[k], xs[k]));
  double part = 0.0;
  for (int j = tid; j < s; j += 256) {
    const double z = exp(-nys_dist(xs, xn, U, ldu, j, d, un[j]) / denom);
    zs[j] = z;
    part += z;
  }
  const double izr = 1.0 / (nys_block_sum(part, red) + 1e-9);
  part = 0.0;
  for (int j = tid; j < s; j += 256) {
    const double a = (izr * zs[j]) * (1.0 / rs[j]);
    zs[j] = a;
    part += a;
  }
  const double iar = 1.0 / (nys_block_sum(part, red) + 1e-9);
  for (int j = tid; j < s; j += 256) Wx[(size_t)blockIdx.x * s + j] = iar * zs[j];
}

// out[i] = add + sum_k T(i,k) V(i,k), row-major n x K operands
__global__ void nys_rowdot_kernel(const double* __restrict__ T, const double* __restrict__ V, int64_t n, int K, double add,
                                  double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0;
  for (int k = 0; k < K; ++k) a = fma(T[i * K + k], V[i * K + k], a);
  out[i] = add + a;
}

__global__ void __launch_bounds__(256) nys_sum_partial_kernel(const double* __restrict__ v, int64_t len, double* part) {
  __shared__ double red[32];
  double a = 0.0;
  for (int64_t e = blockIdx.x * 256 + threadIdx.x; e < len; e += (int64_t)gridDim.x * 256) a += v[e];
  const double t = nys_block_sum(a, red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

}  // namespace

void nys_anchor_distances_run(Ctx* c, const double* U, int s, int64_t ldu, int d, double* un, double* D, double* mean_h) {
  FLGP_LAUNCH(c, nys_rownorm_kernel, ceil_div(s, 128), 128, 0, U, s, ldu, d, un);
  const size_t sm = sizeof(double) * d;
  if (sm > 40 * 1024) FLGP_CUDA(cudaFuncSetAttribute(nys_dist_uu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  FLGP_LAUNCH(c, nys_dist_uu_kernel, s, 256, sm, U, s, ldu, d, un, D);
  const int np = 128;
  DevBuf<double> part(np);
  FLGP_LAUNCH(c, nys_sum_partial_kernel, np, 256, 0, D, (int64_t)s * s, part.p);
  std::vector<double> h(np);
  part.download(h.data(), np, c->stream);
  sync(c);
  double t = 0.0;
  for (int q = 0; q < np; ++q) t += h[q];
  *mean_h = t / ((double)s * s);  // distances_UU.array().sum() / (s * s), src/Fit.cpp:247
}

void nys_anchor_operator_run(Ctx* c, const double* D, int s, int K, double denom, double* rs, double* lam, double* Bt) {
  DevBuf<double> Z((size_t)s * s), A((size_t)s * s), dsum(s), Y((size_t)s * K);
  FLGP_LAUNCH(c, nys_uu_z_kernel, s, 256, 0, D, s, denom, Z.p, rs);
  FLGP_LAUNCH(c, nys_uu_a_kernel, s, 256, 0, Z.p, s, rs, A.p, dsum.p);
  FLGP_LAUNCH(c, nys_uu_w_kernel, ceil_div((int64_t)s * s, 256), 256, 0, A.p, s, dsum.p, Z.p);  // W into Z
  {
    StageScope st(c, "eigh", (4.0 / 3.0) * s * (double)s * s + 2.0 * s * (double)s * K, 8.0 * s * (double)s * s);
    eigh_topk_run(c, Z.p, s, K, lam, Y.p);
  }
  FLGP_LAUNCH(c, nys_vecs_kernel, K, 256, 0, Y.p, s, K, dsum.p, lam, Bt);
  sync(c);
}

void nys_extend_rows_run(Ctx* c, const double* X, int64_t ldx, int64_t row0, int64_t nb, int d, const double* U, int s,
                         int64_t ldu, const double* un, const double* rs, double denom, const double* Bt, int K,
                         double* Wx, double* V) {
  if (nb <= 0) return;
  const size_t sm = sizeof(double) * ((size_t)d + s);
  if (sm > 200 * 1024) fail(2, "nystrom: d + s too large for one row in shared memory");
  if (sm > 40 * 1024) FLGP_CUDA(cudaFuncSetAttribute(nys_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  {
    StageScope st(c, "nystrom_rows", (2.0 * d + 40.0) * s * (double)nb, 8.0 * s * (double)nb);
    FLGP_LAUNCH(c, nys_rows_kernel, (int)nb, 256, sm, X, ldx, row0, d, U, s, ldu, un, rs, denom, Wx);
  }
  // V (nb x K row-major) = Wx (nb x s) Bt^T: C(k + K i) = sum_j Bt(k, j) Wx(i, j)
  StageScope st(c, "nystrom_gemm", 2.0 * s * (double)K * (double)nb, 8.0 * s * (double)nb);
  gemm_nt_ld_run(c, Bt, s, Wx, s, nullptr, K, nb, s, V, K);
}

void nys_rowdot_run(Ctx* c, const double* T, const double* V, int64_t n, int K, double add, double* out) {
  if (n <= 0) return;
  FLGP_LAUNCH(c, nys_rowdot_kernel, ceil_div(n, 256), 256, 0, T, V, n, K, add, out);
}

}  // namespace flgp
