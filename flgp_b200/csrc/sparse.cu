// sparse.cu — everything that streams the fixed-r CSR cross-similarity matrix Z once:
// column sums, graph-Laplacian scaling (graphLaplacian_cpp, /root/reference/src/Utils.cpp:195-212),
// the Lambda^{-1/2} column scale and the s x s Gram of spectrum_from_Z_cpp (src/Spectrum.cpp:146-161),
// the eigenvector lift (u = A v / sigma, src/TruncatedSVD.cpp:23-30) and the folded prediction tail.
//
// All of these are HBM-bound (12r bytes per point per pass).  Reductions over the n
// rows (column sums, Gram) are accumulated as two-limb int64 fixed point with integer atomics, so
// they are deterministic, identical to the oracle's exact=1 flavour bit for bit, and independent of
// how the rows are sharded over GPUs (the limbs are all-reduced as int64).
#include "kernels.cuh"

namespace flgp {

namespace {

constexpr int SP_THREADS = 256;

// ---- column sums: shared-memory privatised limbs, flushed once per CTA ---------------------------
__global__ void __launch_bounds__(SP_THREADS)
colsum_kernel(int64_t nnz, int s, const int32_t* __restrict__ Zj, const double* __restrict__ Zx, Fx fx,
              unsigned long long* __restrict__ acc, int use_smem) {
  extern __shared__ unsigned long long sacc[];  // 2*s limbs when use_smem
  if (use_smem) {
    for (int t = threadIdx.x; t < 2 * s; t += SP_THREADS) sacc[t] = 0ull;
    __syncthreads();
  }
  for (int64_t e = blockIdx.x * (int64_t)SP_THREADS + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * SP_THREADS) {
    double z = Zx[e];
    if (z == 0.0) continue;
    long long h, l;
    fx_encode(fx, z, &h, &l);
    int j = Zj[e];
    if (use_smem) {
      atomicAdd(&sacc[j], (unsigned long long)h);
      if (l) atomicAdd(&sacc[s + j], (unsigned long long)l);
    } else {
      atomicAdd(&acc[j], (unsigned long long)h);
      if (l) atomicAdd(&acc[s + j], (unsigned long long)l);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * s; t += SP_THREADS)
      if (sacc[t]) atomicAdd(&acc[t], sacc[t]);
  }
}

__global__ void decode_vec_kernel(const long long* __restrict__ acc, int s, Fx fx, double* out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < s) out[j] = fx_decode(fx, acc[j], acc[s + j]);
}

// ---- graphLaplacian_cpp: per-row scaling, in place -----------------------------------------------
template <int RT>
__global__ void __launch_bounds__(SP_THREADS)
gl_apply_kernel(int64_t n, int r_in, const int32_t* __restrict__ Zj, double* __restrict__ Zx, int mode,
                const double* __restrict__ colsum, const double* __restrict__ num_class) {
  const int r = RT ? RT : r_in;
  const int64_t i = blockIdx.x * (int64_t)SP_THREADS + threadIdx.x;
  if (i >= n) return;
  constexpr int RA = RT ? RT : 32;
  double z[RA];
  double rs = 0.0;
#pragma unroll
  for (int a = 0; a < RA; ++a)
    if (a < r) {
      int64_t e = i * r + a;
      double v = Zx[e];
      int j = Zj[e];
      if (mode >= 1) v = __dmul_rn(v, __ddiv_rn(1.0, __dadd_rn(colsum[j], 1e-9)));  // src/Utils.cpp:201,204
      if (mode == 2) v = __dmul_rn(v, num_class[j]);                                 // :205
      z[a] = v;
      rs = __dadd_rn(rs, v);                                                         // :210
    }
  const double ir = __ddiv_rn(1.0, __dadd_rn(rs, 1e-9));                             // :211
#pragma unroll
  for (int a = 0; a < RA; ++a)
    if (a < r) Zx[i * r + a] = __dmul_rn(ir, z[a]);
}

__global__ void spectrum_scale_kernel(int s, const double* __restrict__ colsum, double* w) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < s) w[j] = __ddiv_rn(1.0, sqrt(__dadd_rn(fabs(colsum[j]), 1e-9)));          // src/Spectrum.cpp:150
}

// ---- Gram: upper-triangle limbs by integer atomics -----------------------------------------------
template <int RT>
__global__ void __launch_bounds__(SP_THREADS)
gram_kernel(int64_t n, int s, int r_in, const int32_t* __restrict__ Zj, const double* __restrict__ Zx,
            const double* __restrict__ w, Fx fx, unsigned long long* __restrict__ hi,
            unsigned long long* __restrict__ lo) {
  const int r = RT ? RT : r_in;
  constexpr int RA = RT ? RT : 32;
  const int64_t i = blockIdx.x * (int64_t)SP_THREADS + threadIdx.x;
  if (i >= n) return;
  double a[RA];
  int cj[RA];
#pragma unroll
  for (int p = 0; p < RA; ++p)
    if (p < r) {
      cj[p] = Zj[i * r + p];
      a[p] = __dmul_rn(Zx[i * r + p], w[cj[p]]);
    }
#pragma unroll
  for (int p = 0; p < RA; ++p)
#pragma unroll
    for (int q = p; q < RA; ++q)
      if (q < r) {
        double pr = __dmul_rn(a[p], a[q]);
        if (pr == 0.0) continue;
        long long h, l;
        fx_encode(fx, pr, &h, &l);
        size_t at = (size_t)cj[p] + (size_t)s * cj[q];  // rows are column-sorted: cj[p] <= cj[q]
        atomicAdd(&hi[at], (unsigned long long)h);
        if (l) atomicAdd(&lo[at], (unsigned long long)l);
      }
}

__global__ void gram_decode_kernel(const long long* __restrict__ hi, const long long* __restrict__ lo, int s, Fx fx,
                                   double* __restrict__ G) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;  // row (fast index of column-major G)
  int b = blockIdx.y;
  if (a >= s) return;
  size_t at = (a <= b) ? (size_t)a + (size_t)s * b : (size_t)b + (size_t)s * a;
  G[(size_t)a + (size_t)s * b] = fx_decode(fx, hi[at], lo[at]);
}

// ---- lift / folded tail ------------------------------------------------------------------------
// one warp per output row; lanes stride over k (Wm rows are contiguous in k)
__global__ void __launch_bounds__(SP_THREADS)
lift_rows_kernel(int r, const int32_t* __restrict__ Zj, const double* __restrict__ Zx, const double* __restrict__ w,
                 const double* __restrict__ Wm, int K, const int32_t* __restrict__ idx, int64_t n_rows,
                 double* __restrict__ out, int64_t ldo, int colmajor) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)SP_THREADS + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int64_t i = idx ? (int64_t)idx[row] : row;
  for (int k = lane; k < K; k += 32) {
    double acc = 0.0;
    for (int p = 0; p < r; ++p) {
      int cp = Zj[i * r + p];
      double ap = __dmul_rn(Zx[i * r + p], w[cp]);
      acc = fma(ap, Wm[(size_t)cp * K + k], acc);
    }
    if (colmajor) out[row + ldo * k] = acc;
    else out[row * (int64_t)K + k] = acc;
  }
}

__global__ void __launch_bounds__(SP_THREADS)
sparse_rowdot_kernel(int64_t n, int r, const int32_t* __restrict__ Zj, const double* __restrict__ Zx,
                     const double* __restrict__ w, const double* __restrict__ v, double* __restrict__ y) {
  const int64_t i = blockIdx.x * (int64_t)SP_THREADS + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int p = 0; p < r; ++p) {
    int cp = Zj[i * r + p];
    acc = fma(__dmul_rn(Zx[i * r + p], w[cp]), v[cp], acc);
  }
  y[i] = acc;
}

__global__ void __launch_bounds__(SP_THREADS)
sparse_quadform_kernel(int64_t n, int s, int r, const int32_t* __restrict__ Zj, const double* __restrict__ Zx,
                       const double* __restrict__ w, const double* __restrict__ B, double add, double* __restrict__ q) {
  const int64_t i = blockIdx.x * (int64_t)SP_THREADS + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int p = 0; p < r; ++p) {
    int cp = Zj[i * r + p];
    double ap = __dmul_rn(Zx[i * r + p], w[cp]);
    if (ap == 0.0) continue;
    double inner = 0.0;
    for (int t = 0; t < r; ++t) {
      int ct = Zj[i * r + t];
      inner = fma(__dmul_rn(Zx[i * r + t], w[ct]), B[(size_t)cp + (size_t)s * ct], inner);
    }
    acc = fma(ap, inner, acc);
  }
  q[i] = add + acc;
}

// caller-supplied CSR (flgp_spectrum_from_z, flgp_graph_laplacian): the kernels above rely on 0 <= column < s, strictly
// ascending columns inside a row (the Gram writes (c_p, c_q), p <= q, into the upper triangle only) and finite values.
// flags[0] |= 1 bad column, 2 unsorted / duplicate column, 4 non-finite value;  maxbits = bits of max |value|.
__global__ void csr_validate_kernel(int64_t n, int s, int r, const int32_t* __restrict__ Zj, const double* __restrict__ Zx,
                                    int* __restrict__ flags, unsigned long long* __restrict__ maxbits) {
  int bad = 0;
  double mx = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int prev = -1;
    for (int p = 0; p < r; ++p) {
      const int cj = Zj[i * r + p];
      const double v = Zx[i * r + p];
      if (cj < 0 || cj >= s) bad |= 1;
      if (cj <= prev) bad |= 2;
      if (!isfinite(v)) bad |= 4;
      prev = cj;
      mx = fmax(mx, fabs(v));
    }
  }
  for (int o = 16; o; o >>= 1) {
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicOr(flags, bad);
    atomicMax(maxbits, (unsigned long long)__double_as_longlong(mx));  // non-negative doubles order like their bits
  }
}

}  // namespace

double csr_validate_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx) {
  DevBuf<int> flags(1);
  DevBuf<unsigned long long> mb(1);
  flags.zero(c->stream);
  mb.zero(c->stream);
  if (n > 0) {
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
    FLGP_LAUNCH(c, csr_validate_kernel, grid, 256, 0, n, s, r, Zj, Zx, flags.p, mb.p);
  }
  int f = 0;
  double mx = 0.0;
  flags.download(&f, 1, c->stream);
  mb.download(reinterpret_cast<unsigned long long*>(&mx), 1, c->stream);
  sync(c);
  if (f & 1) fail(2, "sparse matrix: a column index is outside [0, s)");
  if (f & 2) fail(2, "sparse matrix: the column indices of a row must be strictly ascending (sorted, no duplicates)");
  if (f & 4) fail(2, "sparse matrix: non-finite value");
  return mx;
}

namespace {
}  // namespace

void colsum_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx, int64_t n_total,
                double* colsum, double vmax) {
  Fx fx;
  if (fx_make(vmax, n_total, &fx)) fail(2, "colsum: bad scale");
  DevBuf<long long> acc((size_t)2 * s);
  acc.zero(c->stream);
  const int64_t nnz = n * r;
  if (nnz > 0) {
    size_t smem = (size_t)2 * s * sizeof(unsigned long long);
    int use_smem = smem <= 96 * 1024;
    if (use_smem && smem > 48 * 1024)
      FLGP_CUDA(cudaFuncSetAttribute(colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)std::min<int64_t>((nnz + SP_THREADS - 1) / SP_THREADS, (int64_t)c->sm_count * 4);
    FLGP_LAUNCH(c, colsum_kernel, grid, SP_THREADS, use_smem ? smem : 0, nnz, s, Zj, Zx, fx,
                reinterpret_cast<unsigned long long*>(acc.p), use_smem);
  }
  comm_allreduce_i64(c, reinterpret_cast<int64_t*>(acc.p), (size_t)2 * s);
  FLGP_LAUNCH(c, decode_vec_kernel, ceil_div(s, 256), 256, 0, acc.p, s, fx, colsum);
  sync(c);
}

void gl_apply_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, double* Zx, int mode,
                  const double* colsum, const double* num_class) {
  (void)s;
  if (mode < 0 || mode > 2) fail(2, "Error: the type of graph Laplacian is not supported!");
  if (r > 32) fail(2, "r=%d exceeds 32", r);
  if (n <= 0) return;
  int grid = ceil_div(n, SP_THREADS);
  if (r == 3) FLGP_LAUNCH(c, gl_apply_kernel<3>, grid, SP_THREADS, 0, n, r, Zj, Zx, mode, colsum, num_class);
  else if (r == 5) FLGP_LAUNCH(c, gl_apply_kernel<5>, grid, SP_THREADS, 0, n, r, Zj, Zx, mode, colsum, num_class);
  else FLGP_LAUNCH(c, gl_apply_kernel<0>, grid, SP_THREADS, 0, n, r, Zj, Zx, mode, colsum, num_class);
}

void spectrum_scale_run(Ctx* c, int s, const double* colsum, double* w) {
  FLGP_LAUNCH(c, spectrum_scale_kernel, ceil_div(s, 256), 256, 0, s, colsum, w);
}

void gram_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx, const double* w,
              int64_t n_total, double* G, double pmax) {
  Fx fx;
  if (fx_make(pmax, n_total, &fx)) fail(2, "gram: bad scale");
  if (r > 32) fail(2, "r=%d exceeds 32", r);
  const size_t ss = (size_t)s * s;
  DevBuf<long long> limbs(2 * ss);
  limbs.zero(c->stream);
  unsigned long long* hi = reinterpret_cast<unsigned long long*>(limbs.p);
  unsigned long long* lo = hi + ss;
  if (n > 0) {
    int grid = ceil_div(n, SP_THREADS);
    if (r == 3) FLGP_LAUNCH(c, gram_kernel<3>, grid, SP_THREADS, 0, n, s, r, Zj, Zx, w, fx, hi, lo);
    else if (r == 5) FLGP_LAUNCH(c, gram_kernel<5>, grid, SP_THREADS, 0, n, s, r, Zj, Zx, w, fx, hi, lo);
    else FLGP_LAUNCH(c, gram_kernel<0>, grid, SP_THREADS, 0, n, s, r, Zj, Zx, w, fx, hi, lo);
  }
  comm_allreduce_i64(c, reinterpret_cast<int64_t*>(limbs.p), 2 * ss);
  dim3 grid(ceil_div(s, 256), s);
  FLGP_LAUNCH(c, gram_decode_kernel, grid, 256, 0, limbs.p, limbs.p + ss, s, fx, G);
  sync(c);
}

void lift_rows_run(Ctx* c, int r, const int32_t* Zj, const double* Zx, const double* w, const double* Wm,
                   int K, const int32_t* idx, int64_t n_rows, double* out, int64_t ldo, bool colmajor) {
  if (n_rows <= 0) return;
  int grid = ceil_div(n_rows * 32, SP_THREADS);
  FLGP_LAUNCH(c, lift_rows_kernel, grid, SP_THREADS, 0, r, Zj, Zx, w, Wm, K, idx, n_rows, out, ldo, colmajor ? 1 : 0);
}

void sparse_rowdot_run(Ctx* c, int64_t n, int r, const int32_t* Zj, const double* Zx, const double* w,
                       const double* v, double* y) {
  if (n <= 0) return;
  FLGP_LAUNCH(c, sparse_rowdot_kernel, ceil_div(n, SP_THREADS), SP_THREADS, 0, n, r, Zj, Zx, w, v, y);
}

void sparse_quadform_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx, const double* w,
                         const double* B, double add, double* q) {
  if (n <= 0) return;
  FLGP_LAUNCH(c, sparse_quadform_kernel, ceil_div(n, SP_THREADS), SP_THREADS, 0, n, s, r, Zj, Zx, w, B, add, q);
}

}  // namespace flgp
