// distsel.cu — large-d (d > 4) point-to-centre distances on the FP64 tensor cores, with EXACT selection.
//
// Serves the two distance stages of the path when d is large:
//   * the Lloyd assignment of subsample_cpp's k-means (/root/reference/src/Utils.cpp:32-45; contract in
//     oracle/flgp_oracle.cpp orc_kmeans_*): arg-min_j of  (|c_j|^2 + M) - 2 x.c_j ;
//   * KNN_cpp (/root/reference/src/Utils.cpp:102-192): the r smallest of ((-2 x.u_j) + |x|^2) + |u_j|^2 in ascending
//     order.
// Both are "the R1 smallest of  add_j - 2 <x_i, c_j>  over j".  The inner products are a GEMM, so they run as DMMA
// (mma.sync.m8n8k4.f64; tcgen05 has no fp64 kind) on TMA-staged, 128B-swizzled operand boxes — the same main loop as
// gemm.cu — and the selection is fused into the epilogue: every thread keeps the R1 = r+1 smallest (value, index)
// pairs of its rows in registers, merged across the row's lanes / warps at the end.  Nothing of size n x s is written.
//
// Exactness.  The tensor-core sums are in a different order than the oracle's (sequential) sums, so the values differ
// in the last bits.  Both differ from the true value by at most E (bound below), hence from each other by 2E: when the
// r+1 smallest tensor-core values are pairwise more than thr = 2 * 2E apart, the oracle's values are ordered the same
// way and its answer (set AND order; ties impossible) is certified.  Rows that cannot be certified (exact ties,
// duplicated centres, lattice data) are appended to a list and re-done by the caller with the oracle-order kernels.
// Distances that are OUTPUT (KNN_cpp's distances_sp) are always re-scored in the oracle's order (knn.cu).
#include "kernels.cuh"
#include "tma_dmma.cuh"

namespace flgp {

namespace {

// column-major n x d (ld ldx) -> row-major n x dp (zero padded columns d..dp-1)
__global__ void __launch_bounds__(256)
to_rowmajor_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, int dp, double* __restrict__ Xr) {
  __shared__ double t[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int q = ty; q < 32; q += 8) {
    const int k = k0 + q;
    const int64_t i = i0 + tx;
    t[q][tx] = (k < d && i < n) ? X[i + ldx * k] : 0.0;
  }
  __syncthreads();
  for (int q = ty; q < 32; q += 8) {
    const int64_t i = i0 + q;
    const int k = k0 + tx;
    if (i < n && k < dp) Xr[i * dp + k] = t[tx][q];
  }
}

template <int R1>
__device__ __forceinline__ void sel_insert(double (&v)[R1], int (&ix)[R1], double x, int j) {
  // precondition: x < v[R1-1]
  v[R1 - 1] = x;
  ix[R1 - 1] = j;
#pragma unroll
  for (int q = R1 - 1; q > 0; --q) {
    if (v[q] < v[q - 1]) {
      const double tv = v[q];
      v[q] = v[q - 1];
      v[q - 1] = tv;
      const int ti = ix[q];
      ix[q] = ix[q - 1];
      ix[q - 1] = ti;
    }
  }
}

// One CTA = 64 rows of X against all s centres, in chunks of 64 centres; 4 warps, each a 32 x 32 sub-tile.
template <int R1>
__global__ void __launch_bounds__(128)
dist_select_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapC,
                   const double* __restrict__ add, int64_t n, int s, int nk, double thr0,
                   const double* __restrict__ thr_row, int r, int32_t* __restrict__ out_idx, int64_t ldo,
                   int* __restrict__ und_count, int32_t* __restrict__ und_list, const int* __restrict__ n_rows_dev,
                   double* __restrict__ out_val) {
  // n_rows_dev (optional): only the first *n_rows_dev rows are live (a device-side count, e.g. the survivors of a
  // bound test); CTAs beyond it leave at once, so the launch needs no host round trip
  if (n_rows_dev) {
    const int64_t live = *n_rows_dev;
    if ((int64_t)blockIdx.x * DG_T >= live) return;
    n = n < live ? n : live;
  }
  extern __shared__ __align__(1024) unsigned char dsm_raw[];
  double* boxes = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[DG_STAGES];
  __shared__ double mv[DG_T][R1];
  __shared__ int mi[DG_T][R1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wm = (wid >> 1) * 32, wn = (wid & 1) * 32;
  const int64_t i0 = (int64_t)blockIdx.x * DG_T;
  const int nchunk = (s + DG_T - 1) / DG_T;
  const int total = nchunk * nk;
  if (tid == 0) {
    for (int st = 0; st < DG_STAGES; ++st) mbar_init(&full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int q) {
    const int st = q % DG_STAGES;
    const int ch = q / nk, kt = q - ch * nk;
    double* a = boxes + (size_t)st * 2 * DG_T * DG_K;
    mbar_expect_tx(&full[st], DG_STAGE_BYTES);
    tma_load_2d(a, &mapX, kt * DG_K, (int)i0, &full[st]);  // rows / columns outside the matrix are zero filled
    tma_load_2d(a + DG_T * DG_K, &mapC, kt * DG_K, ch * DG_T, &full[st]);
  };
  if (tid == 0)
    for (int q = 0; q < DG_STAGES - 1 && q < total; ++q) issue(q);
  double v[4][R1];
  int ix[4][R1];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      v[a][q] = INFINITY;
      ix[a][q] = 0x7fffffff;
    }
  const int fr = lane >> 2, fk = lane & 3;
  int q = 0;
  for (int ch = 0; ch < nchunk; ++ch) {
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int kt = 0; kt < nk; ++kt, ++q) {
      const int st = q % DG_STAGES;
      mbar_wait(&full[st], (q / DG_STAGES) & 1);
      const double* As = boxes + (size_t)st * 2 * DG_T * DG_K;
      const double* Bs = As + DG_T * DG_K;
#pragma unroll
      for (int ks = 0; ks < DG_K; ks += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = box_at(As, wm + a * 8 + fr, ks + fk);
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = box_at(Bs, wn + b * 8 + fr, ks + fk);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
      }
      __syncthreads();  // the slot consumed one step ago is free: refill it
      if (tid == 0 && q + DG_STAGES - 1 < total) issue(q + DG_STAGES - 1);
    }
    // epilogue of the chunk: the thread holds rows wm + a*8 + fr, centres c0 + wn + b*8 + 2 fk + {0,1}
    const int c0 = ch * DG_T + wn + 2 * fk;
    double ad[4][2];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = c0 + b * 8 + h;
        ad[b][h] = (j < s) ? __ldg(add + j) : INFINITY;
      }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double val = fma(-2.0, acc[a][b][h], ad[b][h]);
          if (val < v[a][R1 - 1]) sel_insert<R1>(v[a], ix[a], val, c0 + b * 8 + h);
        }
  }
  // merge the lists of the 4 lanes that share a row
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double pv[R1];
      int pi[R1];
#pragma unroll
      for (int t = 0; t < R1; ++t) {
        pv[t] = __shfl_xor_sync(0xffffffffu, v[a][t], o);
        pi[t] = __shfl_xor_sync(0xffffffffu, ix[a][t], o);
      }
#pragma unroll
      for (int t = 0; t < R1; ++t)
        if (pv[t] < v[a][R1 - 1]) sel_insert<R1>(v[a], ix[a], pv[t], pi[t]);
    }
  // ... and of the two warps that share the rows (centre halves wn = 0 / 32)
  if ((wid & 1) && fk == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int t = 0; t < R1; ++t) {
        mv[wm + a * 8 + fr][t] = v[a][t];
        mi[wm + a * 8 + fr][t] = ix[a][t];
      }
  }
  __syncthreads();
  if (!(wid & 1) && fk == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = wm + a * 8 + fr;
      const int64_t i = i0 + row;
      if (i >= n) continue;
#pragma unroll
      for (int t = 0; t < R1; ++t) {
        const double pv = mv[row][t];
        if (pv < v[a][R1 - 1]) sel_insert<R1>(v[a], ix[a], pv, mi[row][t]);
      }
      const double thr = thr_row ? thr_row[i] : thr0;
      bool ok = true;
#pragma unroll
      for (int t = 1; t < R1; ++t)
        if (t <= r) ok = ok && (v[a][t] - v[a][t - 1] > thr);  // NaN / inf - inf => not certified
#pragma unroll
      for (int t = 0; t < R1 - 1; ++t)
        if (t < r) out_idx[i + ldo * t] = (ix[a][t] < s) ? ix[a][t] : 0;
      if (out_val) {  // the two smallest values (k-means bounds: best and runner-up)
        out_val[2 * i] = v[a][0];
        out_val[2 * i + 1] = v[a][1];
      }
      if (!ok) und_list[atomicAdd(und_count, 1)] = (int32_t)i;
    }
  }
}

template <int R1>
void launch_select(Ctx* c, const CUtensorMap& mapX, const CUtensorMap& mapC, const double* add, int64_t n, int s,
                   int nk, double thr0, const double* thr_row, int r, int32_t* out_idx, int64_t ldo, int* und_count,
                   int32_t* und_list, const int* n_rows_dev, double* out_val) {
  const size_t smem = (size_t)DG_STAGES * DG_STAGE_BYTES + 1024;
  FLGP_CUDA(cudaFuncSetAttribute(dist_select_kernel<R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FLGP_LAUNCH(c, dist_select_kernel<R1>, ceil_div(n, DG_T), 128, smem, mapX, mapC, add, n, s, nk, thr0, thr_row, r,
              out_idx, ldo, und_count, und_list, n_rows_dev, out_val);
}

}  // namespace

void to_rowmajor_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, int dp, double* Xr) {
  if (n <= 0) return;
  dim3 grid(ceil_div(n, 32), ceil_div(dp, 32));
  FLGP_LAUNCH(c, to_rowmajor_kernel, grid, 256, 0, X, n, ldx, d, dp, Xr);
}

bool dist_select_supported(int64_t n, int s, int r) { return n >= DG_T && s >= DG_T && r >= 1 && r <= 5; }

void dist_select_run(Ctx* c, const double* Xr, int64_t n, const double* Cr, int s, int dp, const double* add, int r,
                     double thr0, const double* thr_row, int32_t* out_idx, int64_t ldo, int* und_count,
                     int32_t* und_list, const int* n_rows_dev, double* out_val) {
  if (!dist_select_supported(n, s, r)) fail(2, "dist_select: unsupported shape (n=%lld, s=%d, r=%d)", (long long)n, s, r);
  if (dp % 2) fail(2, "dist_select: the row pitch must be even");
  const CUtensorMap mapX = make_map(Xr, n, dp, dp), mapC = make_map(Cr, s, dp, dp);
  const int nk = (dp + DG_K - 1) / DG_K;
  if (r == 1) launch_select<2>(c, mapX, mapC, add, n, s, nk, thr0, thr_row, r, out_idx, ldo, und_count, und_list, n_rows_dev,
                          out_val);
  else if (r == 2) launch_select<3>(c, mapX, mapC, add, n, s, nk, thr0, thr_row, r, out_idx, ldo, und_count, und_list, n_rows_dev,
                          out_val);
  else if (r == 3) launch_select<4>(c, mapX, mapC, add, n, s, nk, thr0, thr_row, r, out_idx, ldo, und_count, und_list, n_rows_dev,
                          out_val);
  else launch_select<6>(c, mapX, mapC, add, n, s, nk, thr0, thr_row, r, out_idx, ldo, und_count, und_list, n_rows_dev,
                          out_val);
}

}  // namespace flgp
