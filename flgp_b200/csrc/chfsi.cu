// chfsi.cu — top-K eigenpairs of the s x s Gram by Chebyshev-filtered subspace iteration on the FP64 tensor cores.
//
// Replaces, for K << s, the Householder route of eigh.cu (one grid-wide barrier per column: latency bound at 0.02 of
// the fp64 peak) behind the same seam: RSpectra::svds / eigs_sym of /root/reference/src/TruncatedSVD.cpp:23-28 and
// src/Fit.cpp:262-276 — the reference itself uses a Krylov method for K < s.  Everything s-sized is a GEMM:
//
//   0. spectral bounds + a density-of-states estimate of the (nb+1)-th eigenvalue from nv short Lanczos runs
//      (Gauss quadrature of the Lanczos tridiagonals), nb = K + guard columns, a multiple of 64;
//   1. X (s x nb) random;  repeat:  X <- p(G) X with p a scaled Chebyshev polynomial that is <= 1 on [a, c] (the
//      unwanted part of the spectrum) and grows above the cut c — one DMMA GEMM (G times the active columns, TMA
//      staged, three-term recurrence fused into the epilogue) per degree.  Columns are kept sorted by Ritz value and
//      every group of 64 columns gets only the degree its residual still needs, so the active block is a suffix
//      of the columns and the GEMM shrinks (CTA tile height 64/48/32/16 rows keeps ~all SMs busy);
//   2. Cholesky-QR (Gram GEMM, blocked Cholesky + triangular inverse, rotation GEMM), twice when ill conditioned;
//   3. Rayleigh-Ritz: W = G X, H = X^T W, small dense eigenproblem (eigh.cu, order nb), X <- X V, W <- W V,
//      residual norms |W_j - theta_j X_j| decide the next degrees.  Converged when all K residuals <= 1e-13 |G|.
//
// The filter never forms anything larger than s x nb; G is only read.  When the iteration does not converge within
// its budget (clustered spectra whose K-th gap the guards do not cover, K close to s) the caller falls back to the
// direct solver, so the result never depends on this path being applicable.  Deterministic: no atomics, every
// reduction has a fixed order, so replicated ranks compute bit-identical eigenvectors.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.cuh"
#include "tma_dmma.cuh"

namespace flgp {

namespace {

constexpr int CF_BN = 64;      // vector columns per CTA tile = column group of the degree schedule
constexpr int CF_K = DG_K;     // 16: one 128-byte TMA row per k step
// ring depth by tile height (stage = (8 H + 64) x 128 bytes): as deep as ~190 KB of shared memory allow; the short tiles
// of a sharded / mostly converged block move little data per stage and need the depth to cover the L2 latency
__host__ __device__ constexpr int cf_stages(int H) { return H <= 4 ? 16 : 12; }
constexpr int CF_CONSUMERS = 16;                  // consumer warps; one more warp feeds the TMA ring
constexpr int CF_THREADS = 32 * (CF_CONSUMERS + 1);

// rows x K row-major matrix, box = box_rows x 16 doubles, 128-byte swizzle
CUtensorMap make_map_box(const double* base, int64_t rows, int K, int64_t ld, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)CF_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(3, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return m;
}

// ---- the filter GEMM -----------------------------------------------------------------------------------------------
// Out(i, col0 + j) = c1 * sum_k G(i, k) X(k, col0 + j) + c2 * X(i, col0 + j) + c3 * P(i, col0 + j),  j < ncols.
// G: s x s symmetric, row-major (= column-major);  X, P, Out: s x nb column-major with leading dimension ld, i.e.
// every vector contiguous = the k-contiguous "col" operand of mma.sync.m8n8k4.f64.  CTA tile BM x 64, BM = 8 H
// (H = 2 .. 8): 16 consumer warps as 4 x 4 plus one producer warp; operands arrive as one TMA box each per 16-wide
// k slab (SWIZZLE_128B) through a full/empty mbarrier ring, so no CTA-wide barrier sits in the main loop.
// BM = 56 fills 144 of 148 SMs at s = 2000, nb = 256.
// The accumulation order over k is fixed (ascending), independent of the tile shape and of which CTA owns the tile.
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <int H>  // tile height in 8-row blocks (2 .. 8): BM = 8 H
__global__ void __launch_bounds__(CF_THREADS)
cheb_gemm_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX, int s, int col0,
                 int ncols, const double* __restrict__ X, const double* __restrict__ P, double* __restrict__ Out,
                 int64_t ld, double c1, double c2, double c3) {
  constexpr int BM = 8 * H;
  constexpr int MTX = (H + 3) / 4;  // most 8-row blocks any warp row owns
  constexpr int CF_STAGES = cf_stages(H);
  constexpr uint32_t STAGE_BYTES = (uint32_t)(BM + CF_BN) * CF_K * sizeof(double);
  extern __shared__ __align__(1024) unsigned char dsm_raw[];
  const uint32_t boxes = (smem_u32(dsm_raw) + 1023u) & ~1023u;  // SWIZZLE_128B: boxes on 1024-byte boundaries
  __shared__ uint64_t full[CF_STAGES], empty[CF_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = blockIdx.y * BM, j0 = col0 + blockIdx.x * CF_BN;
  const int nk = (s + CF_K - 1) / CF_K;
  if (tid == 0) {
    for (int st = 0; st < CF_STAGES; ++st) {
      mbar_init(&full[st], 1);
      mbar_init(&empty[st], CF_CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (wid == CF_CONSUMERS) {
    // producer warp: one lane keeps the ring full; a slot is refilled as soon as all consumer warps have released it
    if (lane == 0) {
      for (int kt = 0; kt < nk; ++kt) {
        const int st = kt % CF_STAGES;
        if (kt >= CF_STAGES) mbar_wait(&empty[st], ((kt / CF_STAGES) - 1) & 1);
        const uint32_t a = boxes + (uint32_t)st * STAGE_BYTES;
        mbar_expect_tx(&full[st], STAGE_BYTES);
        tma_load_2d_u32(a, &mapG, kt * CF_K, i0, &full[st]);
        tma_load_2d_u32(a + BM * CF_K * 8, &mapX, kt * CF_K, j0, &full[st]);
      }
    }
    return;
  }
  // 16 consumer warps as 4 x 4: warp row r owns H/4 (+1 for r < H%4) blocks of 8 rows, warp column 16 vector columns.
  // The 4 warps that share an SM sub-partition (same wid % 4) sit in the 4 different rows: equal tensor work each.
  const int wr = wid >> 2;
  const int mtw = H / 4 + (wr < H % 4 ? 1 : 0);
  const int wm = 8 * (wr * (H / 4) + min(wr, H % 4)), wn = (wid & 3) * 16;
  double acc[MTX][2][2];
#pragma unroll
  for (int a = 0; a < MTX; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  // Fragment row of a lane: the 8 rows (columns) of an m8n8k4 block may be dealt to the lane groups in any order as
  // long as the epilogue uses the same order.  A 64-bit shared load is served per half-warp (lanes 0-15 = groups
  // 0-3); with the 128-byte swizzle rows r and r^1 exchange the same pair of 16-byte chunks, so groups 0-3 must not
  // hold adjacent rows: group g takes row pr(g) = 0,2,4,6,1,3,5,7 -> every half-warp touches 8 distinct chunks.
  const int fg = lane >> 2, fk = lane & 3;
  const int fr = ((fg & 3) << 1) | (fg >> 2);
  // element (row, k) of a swizzled box: row * 128 bytes + 16-byte chunk ((k >> 1) ^ (row & 7)) + (k & 1) * 8;
  // row & 7 == fr for every fragment row of this thread, so the in-row offsets depend on q = k / 4 only
  uint32_t offA[4], offB[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t inrow = (uint32_t)((((2 * q + (fk >> 1)) ^ fr) << 4) | ((fk & 1) << 3));
    offA[q] = (uint32_t)(wm + fr) * 128u + inrow;
    offB[q] = (uint32_t)BM * 128u + (uint32_t)(wn + fr) * 128u + inrow;
  }
  for (int kt = 0; kt < nk; ++kt) {
    const int st = kt % CF_STAGES;
    mbar_wait(&full[st], (kt / CF_STAGES) & 1);
    const uint32_t sb = boxes + (uint32_t)st * STAGE_BYTES;
    double af[4][MTX], bf[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int a = 0; a < MTX; ++a)
        if (a < mtw) af[q][a] = lds_f64(sb + offA[q] + a * 1024);
#pragma unroll
      for (int b = 0; b < 2; ++b) bf[q][b] = lds_f64(sb + offB[q] + b * 1024);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);  // the fragments are in registers: the slot may be refilled
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int a = 0; a < MTX; ++a)
        if (a < mtw) {
#pragma unroll
          for (int b = 0; b < 2; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[q][a], bf[q][b]);
        }
  }
  const int jend = col0 + ncols;
#pragma unroll
  for (int a = 0; a < MTX; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + wm + a * 8 + fr;
      if (a >= mtw || i >= s) continue;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n = 2 * fk + q;  // accumulator column n was fed by lane group n, i.e. physical column pr(n)
        const int j = j0 + wn + b * 8 + (((n & 3) << 1) | (n >> 2));
        if (j >= jend) continue;
        const int64_t at = i + ld * (int64_t)j;
        double v = c1 * acc[a][b][q];
        if (c2 != 0.0) v = fma(c2, X[at], v);
        if (c3 != 0.0) v = fma(c3, P[at], v);
        Out[at] = v;
      }
    }
}

// ---- small helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double cf_hash(unsigned i, unsigned j) {
  unsigned long long z = ((unsigned long long)i << 32 | j) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
__global__ void cf_init_kernel(double* X, int s, int64_t ld, int ncols, unsigned salt) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)s * ncols) return;
  const int j = (int)(e / s), i = (int)(e - (int64_t)j * s);
  X[i + ld * j] = cf_hash((unsigned)i, (unsigned)j + salt);
}

__device__ __forceinline__ double cf_block_sum(double v, double* red) {  // fixed order; every thread gets the sum
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : 0.0;
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
// every column to unit length (one CTA per column)
__global__ void __launch_bounds__(256) cf_normalize_kernel(double* X, int s, int64_t ld) {
  __shared__ double red[32];
  double* x = X + ld * (int64_t)blockIdx.x;
  double p = 0.0;
  for (int i = threadIdx.x; i < s; i += 256) p = fma(x[i], x[i], p);
  const double n2 = cf_block_sum(p, red);
  const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
  for (int i = threadIdx.x; i < s; i += 256) x[i] *= inv;
}
// res[j] = |W_j - theta_j X_j|
__global__ void __launch_bounds__(256)
cf_resid_kernel(const double* X, const double* W, const double* theta, int s, int64_t ld, double* res) {
  __shared__ double red[32];
  const double* x = X + ld * (int64_t)blockIdx.x;
  const double* w = W + ld * (int64_t)blockIdx.x;
  const double th = theta[blockIdx.x];
  double p = 0.0;
  for (int i = threadIdx.x; i < s; i += 256) {
    const double d = fma(-th, x[i], w[i]);
    p = fma(d, d, p);
  }
  const double n2 = cf_block_sum(p, red);
  if (threadIdx.x == 0) res[blockIdx.x] = sqrt(n2);
}
// column-major s x ncols (ld) -> row-major s x ncols
__global__ void cf_transpose_kernel(const double* __restrict__ X, int s, int64_t ld, int ncols, double* __restrict__ Xr) {
  __shared__ double t[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int jj = threadIdx.y; jj < 32; jj += 8) {
    const int i = i0 + threadIdx.x, j = j0 + jj;
    t[jj][threadIdx.x] = (i < s && j < ncols) ? X[i + ld * j] : 0.0;
  }
  __syncthreads();
  for (int ii = threadIdx.y; ii < 32; ii += 8) {
    const int i = i0 + ii, j = j0 + threadIdx.x;
    if (i < s && j < ncols) Xr[(int64_t)i * ncols + j] = t[threadIdx.x][ii];
  }
}
// Multi-GPU: the filter is column-parallel, so every rank filters its share of every 64-column group and the blocks are
// all-gathered.  Rank q owns the cpg = 64 / R columns [g * 64 + q * cpg, ...) of group g; its local block keeps the
// groups in order (local column g * cpg + j), so the active columns are still a suffix.
__global__ void cf_pack_kernel(const double* __restrict__ X, int s, int64_t ld, int ng, int cpg, int rank,
                               double* __restrict__ Xl) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)s * ng * cpg) return;
  const int lc = (int)(e / s), i = (int)(e - (int64_t)lc * s);
  const int g = lc / cpg, j = lc - g * cpg;
  Xl[i + ld * lc] = X[i + ld * (g * CF_BN + rank * cpg + j)];
}
// recv: R blocks of (s x ng*cpg), block q from rank q  ->  X (s x nb)
__global__ void cf_unpack_kernel(const double* __restrict__ recv, int s, int64_t ld, int ng, int cpg, int R,
                                 double* __restrict__ X) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nbl = ng * cpg;
  if (e >= (int64_t)s * nbl * R) return;
  const int q = (int)(e / ((int64_t)s * nbl));
  const int64_t rem = e - (int64_t)q * s * nbl;
  const int lc = (int)(rem / s), i = (int)(rem - (int64_t)lc * s);
  const int g = lc / cpg, j = lc - g * cpg;
  X[i + ld * (g * CF_BN + q * cpg + j)] = recv[(size_t)q * ld * nbl + i + ld * lc];
}
// H <- (H + H^T) / 2  (nb x nb column-major)
__global__ void cf_symmetrize_kernel(double* H, int nb) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nb * nb) return;
  const int i = e % nb, j = e / nb;
  if (i > j) {
    const double v = 0.5 * (H[i + (int64_t)nb * j] + H[j + (int64_t)nb * i]);
    H[i + (int64_t)nb * j] = v;
    H[j + (int64_t)nb * i] = v;
  }
}

// ---- Cholesky + triangular inverse of the nb x nb Gram: one thread-block cluster ------------------------------------
// S: nb x nb column-major, symmetric positive definite (only read).  M: nb x nb ROW-major work matrix; on exit its
// lower triangle is L^{-1} with S = L L^T (the upper triangle is zero).  info[0] = 1 when a pivot was not positive,
// info[1..2] = min / max of diag(L).
// Right-looking over 32-wide panels with the identity carried along as right-hand side, in place: in row i the
// columns left of the current panel hold the running R = I - sum L21 Y (becoming L^{-1}), the columns from the panel
// on hold the Schur complement — the two live ranges never overlap.  Per panel p (D = M[p,p], Di = chol(D)^{-1}):
//   Y[p, J<p] = Di R[p, J],  Y[p,p] = Di            -> rows p of the result
//   L21(i)    = M[i, p] Di^T                        (rows i below the panel)
//   M[i, J]   = (J == p ? 0 : M[i, J]) - L21(i) Z[J],   Z[J] = Y[p, J] (J <= p),  L21(J)^T (J > p),   J <= i
// The rows below the panel are dealt round-robin to the CH_NC CTAs of the cluster; Z (32 x nb) is exchanged through
// global memory (L2) between two cluster barriers per panel.  Every CTA factors the 32 x 32 diagonal block itself.
constexpr int CH_NB = 32, CH_NC = 8, CH_THREADS = 512;
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__global__ void __cluster_dims__(CH_NC, 1, 1) __launch_bounds__(CH_THREADS)
cf_chol_inv_kernel(const double* __restrict__ S, int nb, double* M, double* Zg, double* __restrict__ info,
                   long long* __restrict__ prof) {
  long long t0 = clock64(), tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define CH_TICK(i)                                \
  if (prof && threadIdx.x == 0) {                 \
    const long long t_ = clock64();               \
    tacc[i] += t_ - t0;                           \
    t0 = t_;                                      \
  }
  extern __shared__ __align__(16) double csm[];
  double* D = csm;                        // 32 x 33: diagonal block, then its Cholesky factor
  double* Di = D + CH_NB * 33;            // 32 x 33: inverse of the factor
  double* Zs = Di + CH_NB * 33;           // 32 x nb: Z panel, Zs[k * nb + col]
  double* Ls = Zs + (size_t)CH_NB * nb;   // rows_per_cta x 33: this CTA's rows of L21
  __shared__ int bad;
  __shared__ double dmin_s, dmax_s, invd[CH_NB];
  const int tid = threadIdx.x, rank = (int)cluster_rank();
  const int gtid = rank * CH_THREADS + tid, gthreads = CH_NC * CH_THREADS;
  if (tid == 0) {
    bad = 0;
    dmin_s = DBL_MAX;
    dmax_s = 0.0;
  }
  // M = lower triangle of S, row-major (S symmetric: read the upper triangle column-wise = coalesced)
  for (int e = gtid; e < nb * nb; e += gthreads) {
    const int i = e / nb, col = e - i * nb;
    __stcg(M + e, col <= i ? S[col + (size_t)nb * i] : 0.0);
  }
  cluster_sync_all();
  CH_TICK(0);
  for (int j0 = 0; j0 < nb; j0 += CH_NB) {
    const int jb = min(CH_NB, nb - j0);
    const int r0 = j0 + jb;         // first row below the panel
    const int below = nb - r0;
    // (a) diagonal block -> smem, factor + invert (one warp; every CTA does it: no broadcast needed)
    for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
      const int i = e / CH_NB, j = e % CH_NB;
      D[i * 33 + j] = (i < jb && j <= i) ? __ldcg(M + (size_t)(j0 + i) * nb + j0 + j) : 0.0;
      Di[i * 33 + j] = 0.0;
    }
    __syncthreads();
    CH_TICK(1);
    // Cholesky and inverse of the 32 x 32 block by ONE warp, lane = row, the row in registers, loops fully unrolled:
    // a step is a shuffle (pivot), an rsqrt, and independent shared-memory broadcasts + FMAs - no block barrier inside
    // the 2 x 32 dependent steps (the all-thread version spent ~520 cycles per step on its barrier and its rsqrt).
    if (tid < 32) {
      const int lane = tid;
      double row[CH_NB];
#pragma unroll
      for (int j = 0; j < CH_NB; ++j) row[j] = D[lane * 33 + j];
      double lmin = DBL_MAX, lmax = 0.0;
      bool isbad = false;
#pragma unroll
      for (int k = 0; k < CH_NB; ++k) {
        if (k < jb) {  // warp-uniform
          double dkk = __shfl_sync(0xffffffffu, row[k], k);
          if (!(dkk > 0.0)) {
            isbad = true;
            dkk = DBL_MIN;
          }
          const double inv = rsqrt(dkk);
          const double lik = (lane >= k) ? row[k] * inv : 0.0;  // column k of the factor
          row[k] = lik;
          D[lane * 33 + k] = lik;
          if (lane == 0) invd[k] = inv;
          lmin = fmin(lmin, dkk * inv);
          lmax = fmax(lmax, dkk * inv);
          __syncwarp();
#pragma unroll
          for (int j = k + 1; j < CH_NB; ++j)
            if (lane >= j) row[j] = fma(-lik, D[j * 33 + k], row[j]);
        }
      }
      __syncwarp();
      // Di(r, c) = (delta_rc - sum_{q < r} L(r, q) Di(q, c)) / L(r, r): lane = column c, the column in registers
      double colv[CH_NB];
#pragma unroll
      for (int r = 0; r < CH_NB; ++r) {
        double v = (r == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < r; ++q) v = fma(-D[r * 33 + q], colv[q], v);
        colv[r] = (r < jb && lane <= r) ? v * invd[r] : 0.0;
        Di[r * 33 + lane] = colv[r];
      }
      if (lane == 0) {
        if (isbad) bad = 1;
        dmin_s = fmin(dmin_s, lmin);
        dmax_s = fmax(dmax_s, lmax);
      }
    }
    CH_TICK(2);
    __syncthreads();
    CH_TICK(3);
    // (b) Y[p, col] = sum_q Di(k, q) R(j0 + q, col) for col < j0: every CTA takes a column range, staged through
    //     smem (Zs is free until the barrier) so that no dependent L2 round trip sits in the inner loop; Y[p, p] = Di
    {
      const int wd = (j0 + CH_NC - 1) / CH_NC;
      const int c0 = min(j0, rank * wd), c1 = min(j0, c0 + wd), cw = c1 - c0;
      for (int e = tid; e < jb * cw; e += CH_THREADS) {
        const int q = e / cw, cx = e - q * cw;
        Zs[(size_t)q * nb + cx] = __ldcg(M + (size_t)(j0 + q) * nb + c0 + cx);
      }
      __syncthreads();
      for (int e = tid; e < jb * cw; e += CH_THREADS) {
        const int k = e / cw, cx = e - k * cw;
        double v = 0.0;
        for (int q = 0; q <= k; ++q) v = fma(Di[k * 33 + q], Zs[(size_t)q * nb + cx], v);
        __stcg(Zg + (size_t)k * nb + c0 + cx, v);
      }
      if (rank == 0)
        for (int e = tid; e < jb * jb; e += CH_THREADS) {
          const int k = e / jb, cx = e - k * jb;
          __stcg(Zg + (size_t)k * nb + j0 + cx, Di[k * 33 + cx]);
        }
    }
    // (c) L21 of this CTA's rows (row r0 + t with t % CH_NC == rank): kept in smem, published as Z[:, r0 + t]
    const int myrows = (below - rank + CH_NC - 1) / CH_NC;  // t = rank, rank + CH_NC, ...
    for (int e = tid; e < myrows * jb; e += CH_THREADS) {
      const int lr = e / jb, q = e - lr * jb;
      Ls[lr * 33 + q] = __ldcg(M + (size_t)(r0 + rank + lr * CH_NC) * nb + j0 + q);
    }
    __syncthreads();
    {
      double vreg[4];  // myrows * jb <= 64 * 32 = 4 * CH_THREADS
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = tid + u * CH_THREADS;
        vreg[u] = 0.0;
        if (e < myrows * jb) {
          const int lr = e / jb, k = e - lr * jb;
          for (int q = 0; q <= k; ++q) vreg[u] = fma(Ls[lr * 33 + q], Di[k * 33 + q], vreg[u]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = tid + u * CH_THREADS;
        if (e < myrows * jb) {
          const int lr = e / jb, k = e - lr * jb;
          Ls[lr * 33 + k] = vreg[u];
          __stcg(Zg + (size_t)k * nb + r0 + rank + lr * CH_NC, vreg[u]);
        }
      }
    }
    CH_TICK(4);
    cluster_sync_all();
    CH_TICK(5);
    // (d) Z panel -> smem (columns 0 .. nb-1 that exist: Y part [0, r0), L21^T part [r0, nb))
    for (int e0 = 0; e0 < jb * nb; e0 += 8 * CH_THREADS) {  // 8 loads in flight per thread
      double t8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * CH_THREADS + tid;
        t8[u] = (e < jb * nb) ? __ldcg(Zg + e) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * CH_THREADS + tid;
        if (e < jb * nb) Zs[e] = t8[u];
      }
    }
    __syncthreads();
    // rows of the result: M[p rows, 0 .. r0) = Y   (CTA 0; nobody reads these rows again)
    if (rank == 0)
      for (int e = tid; e < jb * r0; e += CH_THREADS) {
        const int k = e / r0, col = e - k * r0;
        __stcg(M + (size_t)(j0 + k) * nb + col, Zs[(size_t)k * nb + col]);
      }
    CH_TICK(6);
    // (e) update of this CTA's rows, columns 0 .. i: all (row, column) pairs at once, so that the L2 round trips of
    //     different rows overlap
    for (int e = tid; e < myrows * nb; e += CH_THREADS) {
      const int lr = e / nb, col = e - lr * nb;
      const int i = r0 + rank + lr * CH_NC;
      if (col > i) continue;
      // the old value is fetched first and consumed last: its L2 round trip hides behind the 32 FMAs
      const double old = (col >= j0 && col < r0) ? 0.0 : __ldcg(M + (size_t)i * nb + col);
      const double* lrow = Ls + lr * 33;
      double acc = 0.0;
#pragma unroll 8
      for (int k = 0; k < jb; ++k) acc = fma(lrow[k], Zs[(size_t)k * nb + col], acc);
      __stcg(M + (size_t)i * nb + col, old - acc);
    }
    cluster_sync_all();
    CH_TICK(7);
  }
  if (prof && rank == 0 && tid == 0)
    for (int i = 0; i < 8; ++i) prof[i] = tacc[i];
#undef CH_TICK
  if (rank == 0 && tid == 0) {
    info[0] = (double)bad;
    info[1] = dmin_s;
    info[2] = dmax_s;
  }
}

// ---- Lanczos (density of states) -----------------------------------------------------------------------------------
constexpr int LZ_NB = 2;       // n-blocks of the DMMA
constexpr int LZ_NV = 8 * LZ_NB;  // simultaneous, independent Lanczos runs (16: the quadrature's variance, not its
                               // resolution, limits the cut estimate - 16 runs x 20 steps beat 8 x 32 at 0.6 of the cost)
constexpr int LZ_BM = 16;      // rows of G per CTA (two consumer warps of 8 rows)
constexpr int LZ_SUB = 4;      // 16-wide k slabs per ring stage (a slab is consumed in ~100 cycles: fewer, fatter stages)
constexpr int LZ_STAGES = 8;
constexpr uint32_t LZ_SLAB_BYTES = (LZ_BM + LZ_NV) * CF_K * sizeof(double);  // G box 16 x 16, V box LZ_NV x 16
// U = G V for all runs at once on the tensor cores (V, Vp, U: s x LZ_NV column-major, ld s), plus this CTA's share of
// the inner products u.v, u.u, u.vp per run: partial[(cta * 3 + which) * LZ_NV + run].
__global__ void __launch_bounds__(96)
lanczos_symv_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapV, int s,
                    const double* __restrict__ V, const double* __restrict__ Vp, double* __restrict__ U,
                    double* __restrict__ partial) {
  extern __shared__ __align__(1024) unsigned char dsm_raw[];
  const uint32_t boxes = (smem_u32(dsm_raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[LZ_STAGES], empty[LZ_STAGES];
  __shared__ double wsum[2][3][LZ_NV];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = blockIdx.x * LZ_BM;
  const int nslab = (s + CF_K - 1) / CF_K, nst = (nslab + LZ_SUB - 1) / LZ_SUB;
  if (tid == 0) {
    for (int st = 0; st < LZ_STAGES; ++st) {
      mbar_init(&full[st], 1);
      mbar_init(&empty[st], 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (wid == 2) {
    if (lane == 0) {
      for (int t = 0; t < nst; ++t) {
        const int st = t % LZ_STAGES;
        if (t >= LZ_STAGES) mbar_wait(&empty[st], ((t / LZ_STAGES) - 1) & 1);
        const int nsub = min(LZ_SUB, nslab - t * LZ_SUB);
        mbar_expect_tx(&full[st], (uint32_t)nsub * LZ_SLAB_BYTES);
        for (int u = 0; u < nsub; ++u) {
          const uint32_t a = boxes + (uint32_t)(st * LZ_SUB + u) * LZ_SLAB_BYTES;
          tma_load_2d_u32(a, &mapG, (t * LZ_SUB + u) * CF_K, i0, &full[st]);
          tma_load_2d_u32(a + LZ_BM * CF_K * 8, &mapV, (t * LZ_SUB + u) * CF_K, 0, &full[st]);
        }
      }
    }
    return;
  }
  const int fg = lane >> 2, fk = lane & 3;
  const int fr = ((fg & 3) << 1) | (fg >> 2);  // conflict-free row order of the fragment loads (see cheb_gemm_kernel)
  uint32_t offA[4], offB[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t inrow = (uint32_t)((((2 * q + (fk >> 1)) ^ fr) << 4) | ((fk & 1) << 3));
    offA[q] = (uint32_t)(wid * 8 + fr) * 128u + inrow;
    offB[q] = (uint32_t)LZ_BM * 128u + (uint32_t)fr * 128u + inrow;
  }
  // one accumulator pair per k-slab position: a single pair would chain all 4 ceil(s/16) DMMAs of the row block
  // through the tensor pipe's latency (the kernel was 2x slower for it); the partial sums are added in a fixed order
  double acc[LZ_SUB][LZ_NB][2];
#pragma unroll
  for (int u = 0; u < LZ_SUB; ++u)
#pragma unroll
    for (int b = 0; b < LZ_NB; ++b) acc[u][b][0] = acc[u][b][1] = 0.0;
  for (int t = 0; t < nst; ++t) {
    const int st = t % LZ_STAGES;
    mbar_wait(&full[st], (t / LZ_STAGES) & 1);
    const int nsub = min(LZ_SUB, nslab - t * LZ_SUB);
    double af[LZ_SUB][4], bf[LZ_SUB][LZ_NB][4];
#pragma unroll
    for (int u = 0; u < LZ_SUB; ++u) {
      const uint32_t sb = boxes + (uint32_t)(st * LZ_SUB + u) * LZ_SLAB_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        af[u][q] = (u < nsub) ? lds_f64(sb + offA[q]) : 0.0;
#pragma unroll
        for (int b = 0; b < LZ_NB; ++b) bf[u][b][q] = (u < nsub) ? lds_f64(sb + offB[q] + b * 1024) : 0.0;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int u = 0; u < LZ_SUB; ++u)
#pragma unroll
        for (int b = 0; b < LZ_NB; ++b) dmma_m8n8k4(acc[u][b][0], acc[u][b][1], af[u][q], bf[u][b][q]);
  }
  static_assert(LZ_SUB == 4, "the final sum is written out for four partial accumulators");
  // epilogue: u(i, run) for run = 8 b + pr(2 fk + q); inner products over this warp's 8 rows, then over the two warps
  const int i = i0 + wid * 8 + fr;
#pragma unroll
  for (int b = 0; b < LZ_NB; ++b) {
    double p[3][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double u = (acc[0][b][q] + acc[1][b][q]) + (acc[2][b][q] + acc[3][b][q]);
      const int n = 2 * fk + q;
      const int run = 8 * b + (((n & 3) << 1) | (n >> 2));
      const int64_t at = i + (int64_t)s * run;
      const double v = (i < s) ? V[at] : 0.0, vp = (i < s) ? Vp[at] : 0.0;
      if (i < s) U[at] = u;
      p[0][q] = (i < s) ? u * v : 0.0;
      p[1][q] = (i < s) ? u * u : 0.0;
      p[2][q] = (i < s) ? u * vp : 0.0;
    }
#pragma unroll
    for (int w = 0; w < 3; ++w)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double x = p[w][q];
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        const int n = 2 * fk + q;
        if (fg == 0) wsum[wid][w][8 * b + (((n & 3) << 1) | (n >> 2))] = x;
      }
  }
  // the two consumer warps meet on a named barrier (the producer warp has left)
  asm volatile("bar.sync 1, 64;" ::: "memory");
  if (tid < 3 * LZ_NV) {
    const int w = tid / LZ_NV, run = tid % LZ_NV;
    partial[((size_t)blockIdx.x * 3 + w) * LZ_NV + run] = wsum[0][w][run] + wsum[1][w][run];
  }
}
// finish one Lanczos step from U = G V and the per-CTA inner products u.v, one CTA per run:
//   alpha = u.v;  w = u - alpha v - beta_prev vp;  beta = |w|;  v_new = w / beta  (the caller rotates vp <- v <- v_new).
// Fixed reduction orders: the result does not depend on the schedule.
__global__ void __launch_bounds__(1024)
lanczos_update_kernel(int s, int step, int nparts, const double* __restrict__ partial, const double* __restrict__ V,
                      const double* __restrict__ Vp, const double* __restrict__ U, double* __restrict__ Vnew,
                      double* __restrict__ alpha, double* __restrict__ beta) {
  __shared__ double red[32];
  __shared__ double al_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, run = blockIdx.x;
  if (wid == 0) {  // lanes stride over the CTAs' partial sums, then a fixed shuffle tree
    double a = 0.0;
    for (int cta = lane; cta < nparts; cta += 32) a += partial[((size_t)cta * 3) * LZ_NV + run];
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) al_s = a;
  }
  __syncthreads();
  const double al = al_s, bp = step > 0 ? beta[(step - 1) * LZ_NV + run] : 0.0;
  const double* v = V + (size_t)s * run;
  const double* vp = Vp + (size_t)s * run;
  const double* u = U + (size_t)s * run;
  constexpr int PER = 4;  // s <= 4096: a thread keeps its elements of w in registers
  double w[PER];
  double n2 = 0.0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int i = tid + q * 1024;
    w[q] = (i < s) ? fma(-bp, vp[i], fma(-al, v[i], u[i])) : 0.0;
    n2 = fma(w[q], w[q], n2);
  }
  const double be = sqrt(cf_block_sum(n2, red));
  const double inv = be > 0.0 ? 1.0 / be : 0.0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int i = tid + q * 1024;
    if (i < s) Vnew[(size_t)s * run + i] = w[q] * inv;
  }
  if (tid == 0) {
    alpha[step * LZ_NV + run] = al;
    beta[step * LZ_NV + run] = be;
  }
}

// eigenvalues of a k x k symmetric tridiagonal (d, e) with the FIRST components of its eigenvectors (implicit QL,
// EISPACK imtql2 restricted to one row of the eigenvector matrix).  Host, k <= 64.
bool tql_first_row(int k, std::vector<double> d, std::vector<double> e, std::vector<double>& theta, std::vector<double>& z) {
  z.assign(k, 0.0);
  z[0] = 1.0;
  e.resize(k);
  e[k - 1] = 0.0;
  for (int l = 0; l < k; ++l) {
    int iter = 0;
    while (true) {
      int m = l;
      for (; m < k - 1; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= DBL_EPSILON * dd) break;
      }
      if (m == l) break;
      if (++iter > 60) return false;
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = std::hypot(g, 1.0);
      g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
      double s = 1.0, c = 1.0, p = 0.0;
      int i = m - 1;
      for (; i >= l; --i) {
        double f = s * e[i], b = c * e[i];
        r = std::hypot(f, g);
        e[i + 1] = r;
        if (r == 0.0) {
          d[i + 1] -= p;
          e[m] = 0.0;
          break;
        }
        s = f / r;
        c = g / r;
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * c * b;
        p = s * r;
        d[i + 1] = g + p;
        g = c * r - b;
        f = z[i + 1];
        z[i + 1] = s * z[i] + c * f;
        z[i] = c * z[i] - s * f;
      }
      if (r == 0.0 && i >= l) continue;
      d[l] -= p;
      e[l] = g;
      e[m] = 0.0;
    }
  }
  theta = d;
  return true;
}

struct Filter {  // p(x) = scaled Chebyshev polynomial: small on [a, c], 1 at top
  double a, c, top;
};

const bool g_cf_debug = std::getenv("FLGP_CHFSI_DEBUG") != nullptr;

}  // namespace

bool chol_inv_run(Ctx* c, const double* S, int nb, double* Linv, double* ratio) {
  if (nb < 1 || nb > 512) fail(2, "chol_inv: order %d outside 1..512", nb);
  DevBuf<double> Zg((size_t)CH_NB * nb), info(4);
  const int rows_cta = ceil_div(nb, CH_NC);
  const size_t smem = ((size_t)2 * CH_NB * 33 + (size_t)CH_NB * nb + (size_t)rows_cta * 33) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    FLGP_CUDA(cudaFuncSetAttribute(cf_chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  FLGP_LAUNCH(c, cf_chol_inv_kernel, CH_NC, CH_THREADS, smem, S, nb, Linv, Zg.p, info.p, (long long*)nullptr);
  double ih[3];
  info.download(ih, 3, c->stream);
  sync(c);
  if (ratio) *ratio = ih[2] > 0.0 ? ih[1] / ih[2] : 0.0;
  return ih[0] == 0.0 && ih[1] > 0.0;
}

// Top-K eigenpairs of G (s x s, symmetric, full storage, only read).  lam (K, descending) and Y (s x K column-major,
// orthonormal) on the device.  psd: the caller knows G is positive semi-definite (a Gram).  Returns false when the
// iteration is not applicable or did not converge (outputs then undefined; G untouched).
bool chfsi_topk_run(Ctx* c, const double* G, int s, int K, double* lam, double* Y, bool psd) {
  if (s % 2 != 0 || reinterpret_cast<uintptr_t>(G) % 16 != 0) return false;  // TMA: 16-byte row pitch
  static const int env_guard = std::getenv("FLGP_CHFSI_NB") ? std::atoi(std::getenv("FLGP_CHFSI_NB")) : 0;
  int nb = ((K + std::max(32, K / 4) + CF_BN - 1) / CF_BN) * CF_BN;
  if (env_guard > 0) nb = ((std::max(env_guard, K + 8) + CF_BN - 1) / CF_BN) * CF_BN;
  if (nb > 512 || nb * 3 > s || s > 4096) return false;  // s <= 4096: lanczos_update_kernel keeps a run in registers
  const int ng = nb / CF_BN;
  const int64_t ld = s;
  const size_t blk = (size_t)s * nb;
  DevBuf<double> Xb[3] = {DevBuf<double>(blk), DevBuf<double>(blk), DevBuf<double>(blk)};
  DevBuf<double> Wb(blk), Xr(blk);
  DevBuf<double> S((size_t)nb * nb), Linv((size_t)nb * nb), Hs((size_t)nb * nb), Vs((size_t)nb * nb), Zg((size_t)CH_NB * nb), theta(nb), res(nb), info(4);
  // ---- tensor maps: G with 4 box heights, each vector block as the "col" operand
  CUtensorMap mapG[9];  // index = tile height / 8
  for (int h = 2; h <= 8; ++h) mapG[h] = make_map_box(G, s, s, s, 8 * h);
  // a set of three rotating column blocks (the recurrence needs Y_{j-2}, Y_{j-1}, Y_j) with their tensor maps
  struct BlockSet {
    CUtensorMap map[3];
    double* p[3];
    int ncols;
  };
  BlockSet full;
  full.ncols = nb;
  for (int q = 0; q < 3; ++q) {
    full.p[q] = Xb[q].p;
    full.map[q] = make_map_box(Xb[q].p, nb, s, ld, CF_BN);
  }
  // multi-GPU: this rank's columns of every group (see cf_pack_kernel)
  static const bool no_shard = std::getenv("FLGP_CHFSI_NO_SHARD") != nullptr;
  const int R = c->nranks;
  const bool shard = R > 1 && CF_BN % R == 0 && !no_shard;
  const int cpg = shard ? CF_BN / R : CF_BN, nbl = ng * cpg;
  DevBuf<double> Xl[3], recv;
  BlockSet loc;
  loc.ncols = nbl;
  if (shard) {
    recv.alloc(blk);
    const int nbl_map = std::max(nbl, CF_BN);  // the TMA box is 64 columns wide: keep the tensor at least that wide
    for (int q = 0; q < 3; ++q) {
      Xl[q].alloc((size_t)s * nbl_map);
      if (nbl_map > nbl) Xl[q].zero(c->stream);
      loc.p[q] = Xl[q].p;
      loc.map[q] = make_map_box(Xl[q].p, nbl_map, s, ld, CF_BN);
    }
  }
  auto gemm = [&](const BlockSet& bs, int src, const double* P, double* Out, int col0, int ncols, double c1, double c2,
                  double c3) {
    // tile height: the smallest that still gives at most one CTA per SM (the active block shrinks with the degrees)
    const int ct = ceil_div(ncols, CF_BN);
    int h = 8;
    while (h > 2 && ceil_div(s, 8 * (h - 1)) * ct <= c->sm_count) --h;
    const size_t smem = (size_t)cf_stages(h) * (8 * h + CF_BN) * CF_K * sizeof(double) + 1024;
    dim3 grid(ct, ceil_div(s, 8 * h));
#define CF_LAUNCH(HH)                                                                                                 \
  {                                                                                                                   \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      FLGP_CUDA(cudaFuncSetAttribute(cheb_gemm_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    FLGP_LAUNCH(c, cheb_gemm_kernel<HH>, grid, CF_THREADS, smem, mapG[h], bs.map[src], s, col0, ncols, bs.p[src], P,  \
                Out, ld, c1, c2, c3);                                                                                 \
  }
    switch (h) {
      case 8: CF_LAUNCH(8) break;
      case 7: CF_LAUNCH(7) break;
      case 6: CF_LAUNCH(6) break;
      case 5: CF_LAUNCH(5) break;
      case 4: CF_LAUNCH(4) break;
      case 3: CF_LAUNCH(3) break;
      default: CF_LAUNCH(2) break;
    }
#undef CF_LAUNCH
  };
  double cost = 0.0;  // in units of one full-width application of G
  // ---- 0. spectral bounds and the first cut
  Filter f;
  double gnorm;
  {
    StageScope st(c, "eigh_chfsi_dos");
    const int kl = std::min(20, s / 4);
    const int nparts = ceil_div(s, LZ_BM);
    DevBuf<double> Vb[3] = {DevBuf<double>((size_t)s * LZ_NV), DevBuf<double>((size_t)s * LZ_NV),
                            DevBuf<double>((size_t)s * LZ_NV)};
    DevBuf<double> W((size_t)s * LZ_NV), al((size_t)kl * LZ_NV), be((size_t)kl * LZ_NV), part((size_t)nparts * 3 * LZ_NV);
    CUtensorMap mapV[3];
    for (int q = 0; q < 3; ++q) mapV[q] = make_map_box(Vb[q].p, LZ_NV, s, s, LZ_NV);
    const size_t lz_smem = (size_t)LZ_STAGES * LZ_SUB * LZ_SLAB_BYTES + 1024;
    static bool lz_attr = false;
    if (!lz_attr) {
      FLGP_CUDA(cudaFuncSetAttribute(lanczos_symv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lz_smem));
      lz_attr = true;
    }
    // buffers rotate: vp = Vb[j % 3], v = Vb[(j + 1) % 3], v_new = Vb[(j + 2) % 3]
    Vb[0].zero(c->stream);
    FLGP_LAUNCH(c, cf_init_kernel, ceil_div((int64_t)s * LZ_NV, 256), 256, 0, Vb[1].p, s, (int64_t)s, LZ_NV, 0x51000000u);
    FLGP_LAUNCH(c, cf_normalize_kernel, LZ_NV, 256, 0, Vb[1].p, s, (int64_t)s);
    for (int j = 0; j < kl; ++j) {
      const int ip = j % 3, iv = (j + 1) % 3, in = (j + 2) % 3;
      FLGP_LAUNCH(c, lanczos_symv_kernel, nparts, 96, lz_smem, mapG[2], mapV[iv], s, Vb[iv].p, Vb[ip].p, W.p, part.p);
      FLGP_LAUNCH(c, lanczos_update_kernel, LZ_NV, 1024, 0, s, j, nparts, part.p, Vb[iv].p, Vb[ip].p, W.p,
                  Vb[in].p, al.p, be.p);
    }
    std::vector<double> alh((size_t)kl * LZ_NV), beh((size_t)kl * LZ_NV);
    al.download(alh.data(), alh.size(), c->stream);
    be.download(beh.data(), beh.size(), c->stream);
    sync(c);
    struct Node {
      double th, wt;
    };
    std::vector<Node> nodes;
    double tmin = DBL_MAX, tmax = -DBL_MAX, rmin = 0.0, rmax = 0.0;
    for (int v = 0; v < LZ_NV; ++v) {
      std::vector<double> d(kl), e(kl), th, z;
      int kk = kl;
      for (int j = 0; j < kl; ++j) {
        d[j] = alh[(size_t)j * LZ_NV + v];
        e[j] = beh[(size_t)j * LZ_NV + v];
        if (!(e[j] > 0.0) || !std::isfinite(e[j]) || !std::isfinite(d[j])) {  // invariant subspace found / breakdown
          kk = j + 1;
          break;
        }
      }
      d.resize(kk);
      e.resize(kk);
      const double blast = e[kk - 1];
      if (!tql_first_row(kk, d, e, th, z)) return false;
      for (int j = 0; j < kk; ++j) {
        nodes.push_back({th[j], z[j] * z[j] / LZ_NV});
        (void)blast;
        if (th[j] < tmin) tmin = th[j];
        if (th[j] > tmax) tmax = th[j];
      }
    }
    (void)rmin;
    (void)rmax;
    if (!(tmax > tmin)) return false;
    std::sort(nodes.begin(), nodes.end(), [](const Node& x, const Node& y) { return x.th > y.th; });
    static const double target_mul = std::getenv("FLGP_CHFSI_TARGET") ? std::atof(std::getenv("FLGP_CHFSI_TARGET")) : 1.0;
    const double target = target_mul * nb;
    double cum = 0.0, cut = nodes.back().th;
    for (const Node& nd : nodes) {
      cum += nd.wt * s;
      if (cum >= target) {
        cut = nd.th;
        break;
      }
    }
    const double width = tmax - tmin;
    f.a = tmin - 0.02 * width;
    if (psd && f.a < 0.0) f.a = 0.0;
    f.top = tmax + 0.02 * width;
    f.c = std::min(std::max(cut, f.a + 0.05 * width), f.top - 0.05 * width);
    gnorm = std::max(std::fabs(f.top), std::fabs(f.a));
    cost += 0.1 * kl;
    if (g_cf_debug)
      fprintf(stderr, "[chfsi] s=%d K=%d nb=%d dos: a=%.5f top=%.5f c0=%.5f (lanczos min %.5f max %.5f)\n", s, K, nb, f.a,
              f.top, f.c, tmin, tmax);
  }
  const double tol = 1e-13 * gnorm;
  int cur = 0;  // buffer that holds X
  FLGP_LAUNCH(c, cf_init_kernel, ceil_div((int64_t)s * nb, 256), 256, 0, Xb[0].p, s, ld, nb, 0u);

  // X <- p(G) X with per-group degrees deg[g] (non-decreasing in g); result back in Xb[cur].
  // The recurrence runs on block set bs with gw columns per group: the full block, or (multi-GPU) this rank's share,
  // which is packed from / all-gathered back into the full block around it.  Every column is computed by exactly one
  // rank with the same instruction sequence as in a single-GPU run, so the result does not depend on the rank count.
  auto recurrence = [&](const BlockSet& bs, int gw, int start, const std::vector<int>& deg) -> int {
    const int mm = deg[ng - 1];
    const double e = 0.5 * (f.c - f.a), cen = 0.5 * (f.c + f.a);
    const double sig1 = e / (f.top - cen);
    double sig = sig1;
    int bj = start;                     // buffer of Y_{j-1}
    int bjm = -1;                       // buffer of Y_{j-2}
    std::vector<int> where(ng, start);  // buffer holding the final columns of each group
    for (int j = 1; j <= mm; ++j) {
      int g0 = 0;
      while (deg[g0] < j) ++g0;
      const int col0 = g0 * gw, ncols = bs.ncols - col0;
      int bo = 0;
      while (bo == bj || bo == bjm) ++bo;
      if (j == 1) {
        gemm(bs, bj, nullptr, bs.p[bo], col0, ncols, sig1 / e, -cen * sig1 / e, 0.0);
      } else {
        const double sig2 = 1.0 / (2.0 / sig1 - sig);
        gemm(bs, bj, bs.p[bjm], bs.p[bo], col0, ncols, 2.0 * sig2 / e, -cen * 2.0 * sig2 / e, -sig * sig2);
        sig = sig2;
      }
      cost += (double)ncols / bs.ncols / (shard ? R : 1);
      for (int g = g0; g < ng; ++g)
        if (deg[g] == j) where[g] = bo;
      bjm = bj;
      bj = bo;
    }
    // gather the finished groups into the buffer that holds most of them
    int cnt[3] = {0, 0, 0};
    for (int g = 0; g < ng; ++g) cnt[where[g]]++;
    int dst = 0;
    for (int q = 1; q < 3; ++q)
      if (cnt[q] > cnt[dst]) dst = q;
    for (int g = 0; g < ng; ++g)
      if (where[g] != dst)
        FLGP_CUDA(cudaMemcpyAsync(bs.p[dst] + (size_t)g * gw * ld, bs.p[where[g]] + (size_t)g * gw * ld,
                                  sizeof(double) * gw * ld, cudaMemcpyDeviceToDevice, c->stream));
    return dst;
  };
  auto filter = [&](const std::vector<int>& deg) {
    if (deg[ng - 1] <= 0) return;
    if (!shard) {
      cur = recurrence(full, CF_BN, cur, deg);
      return;
    }
    FLGP_LAUNCH(c, cf_pack_kernel, ceil_div((int64_t)s * nbl, 256), 256, 0, Xb[cur].p, s, ld, ng, cpg, c->rank, Xl[0].p);
    const int dst = recurrence(loc, cpg, 0, deg);
    comm_allgather_f64(c, Xl[dst].p, recv.p, (size_t)s * nbl);
    FLGP_LAUNCH(c, cf_unpack_kernel, ceil_div((int64_t)s * nbl * R, 256), 256, 0, recv.p, s, ld, ng, cpg, R, Xb[cur].p);
  };
  // W = G X (Rayleigh-Ritz), the same way
  auto apply_G = [&](double* W) {
    if (!shard) {
      gemm(full, cur, nullptr, W, 0, nb, 1.0, 0.0, 0.0);
      return;
    }
    FLGP_LAUNCH(c, cf_pack_kernel, ceil_div((int64_t)s * nbl, 256), 256, 0, Xb[cur].p, s, ld, ng, cpg, c->rank, Xl[0].p);
    gemm(loc, 0, nullptr, Xl[1].p, 0, nbl, 1.0, 0.0, 0.0);
    comm_allgather_f64(c, Xl[1].p, recv.p, (size_t)s * nbl);
    FLGP_LAUNCH(c, cf_unpack_kernel, ceil_div((int64_t)s * nbl * R, 256), 256, 0, recv.p, s, ld, ng, cpg, R, W);
  };
  // algorithmic work of one filter call: 2 s^2 flop per active column and degree; G (s^2) + 4 column blocks per launch
  auto filter_work = [&](const std::vector<int>& deg, double* bytes) {
    double fl = 0.0, by = 0.0;
    for (int j = 1; j <= deg[ng - 1]; ++j) {
      int g0 = 0;
      while (deg[g0] < j) ++g0;
      const double ncols = nb - g0 * CF_BN;
      fl += 2.0 * s * (double)s * ncols;
      by += 8.0 * (s * (double)s + 4.0 * s * ncols);
    }
    *bytes = by;
    return fl;
  };
  // Out(s x nb col-major) = In * M with B(j, k) = M(k, j) given row-major (N = nb rows of length nb)
  auto rotate = [&](const double* In, const double* Brm, double* Out) {
    dim3 tg(ceil_div(s, 32), ceil_div(nb, 32));
    FLGP_LAUNCH(c, cf_transpose_kernel, tg, dim3(32, 8), 0, In, s, ld, nb, Xr.p);
    gemm_nt_ld_run(c, Xr.p, nb, Brm, nb, nullptr, s, nb, nb, Out, ld);
  };
  // Cholesky-QR of Xb[cur]; returns false when the Gram is numerically singular
  double last_ratio = 1.0;
  auto cholqr = [&]() -> bool {
    FLGP_LAUNCH(c, cf_normalize_kernel, nb, 256, 0, Xb[cur].p, s, ld);
    for (int pass = 0; pass < 3; ++pass) {
      gemm_tn_splitk_run(c, Xb[cur].p, ld, Xb[cur].p, ld, nb, nb, s, S.p);
      const int rows_cta = ceil_div(nb, CH_NC);
      const size_t smem = ((size_t)2 * CH_NB * 33 + (size_t)CH_NB * nb + (size_t)rows_cta * 33) * sizeof(double);
      static bool attr = false;
      if (!attr) {
        FLGP_CUDA(cudaFuncSetAttribute(cf_chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
      }
      static const bool chprof = std::getenv("FLGP_CHOL_PROF") != nullptr;
      DevBuf<long long> prof(8);
      FLGP_LAUNCH(c, cf_chol_inv_kernel, CH_NC, CH_THREADS, smem, S.p, nb, Linv.p, Zg.p, info.p, chprof ? prof.p : nullptr);
      if (chprof) {
        long long ph[8];
        prof.download(ph, 8, c->stream);
        sync(c);
        fprintf(stderr, "[chol prof] init %lld | load D %lld, chol %lld, inv %lld, Y+L21 %lld, sync %lld, Z load %lld, update+sync %lld\n",
                ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]);
      }
      double ih[3];
      info.download(ih, 3, c->stream);
      sync(c);
      if (ih[0] != 0.0 || !(ih[1] > 0.0)) return false;
      rotate(Xb[cur].p, Linv.p, Xb[cur].p);
      const double ratio = ih[1] / ih[2];
      if (g_cf_debug) fprintf(stderr, "[chfsi]   cholqr pass %d: min/max diag(L) = %.2e\n", pass, ratio);
      last_ratio = ratio;
      // one pass leaves |X^T X - I| ~ eps / ratio^2: enough for the next filter segment or a rough Rayleigh-Ritz
      // step as long as that stays below ~1e-6 (the next Cholesky-QR repairs it); the pass before an accepted result
      // must have ratio >= 0.05
      if (ratio > 2e-5) break;
    }
    return true;
  };
  std::vector<double> th(nb), rs(nb);
  auto rayleigh_ritz = [&]() {
    double* X = Xb[cur].p;
    apply_G(Wb.p);
    cost += shard ? 1.0 / R : 1.0;
    gemm_tn_splitk_run(c, X, ld, Wb.p, ld, nb, nb, s, Hs.p);
    FLGP_LAUNCH(c, cf_symmetrize_kernel, ceil_div(nb * nb, 256), 256, 0, Hs.p, nb);
    eigh_direct_run(c, Hs.p, nb, nb, theta.p, Vs.p);
    rotate(X, Vs.p, X);       // B(j, k) = V(k, j): V column-major is exactly that
    rotate(Wb.p, Vs.p, Wb.p);
    FLGP_LAUNCH(c, cf_resid_kernel, nb, 256, 0, X, Wb.p, theta.p, s, ld, res.p);
    theta.download(th.data(), nb, c->stream);
    res.download(rs.data(), nb, c->stream);
    sync(c);
  };

  // ---- 1. first pass from the random block: as many degrees as the conditioning of the filtered block allows
  std::vector<int> deg(ng);
  {
    const double e = 0.5 * (f.c - f.a), cen = 0.5 * (f.c + f.a);
    const double rate1 = std::acosh(std::max(1.0, (f.top - cen) / e));
    const int m1 = (int)std::max(6.0, std::min(40.0, std::log(1e6) / std::max(rate1, 1e-3)));
    for (int g = 0; g < ng; ++g) deg[g] = m1;
  }
  const int max_outer = 8;
  const double max_cost = 400.0;
  bool converged = false;
  std::vector<std::vector<int>> pending;
  for (int it = 1; it <= max_outer && cost < max_cost; ++it) {
    {
      double by = 0.0;
      const double fl = filter_work(deg, &by);
      StageScope st(c, "eigh_chfsi_filter", fl, by);
      filter(deg);
    }
    {
      StageScope st(c, "eigh_chfsi_cholqr");
      if (!cholqr()) return false;
    }
    while (!pending.empty()) {
      std::vector<int> sg = pending.back();
      pending.pop_back();
      {
        double by = 0.0;
        const double fl = filter_work(sg, &by);
        StageScope st(c, "eigh_chfsi_filter", fl, by);
        filter(sg);
      }
      StageScope st(c, "eigh_chfsi_cholqr");
      if (!cholqr()) return false;
    }
    {
      StageScope st(c, "eigh_chfsi_rr");
      rayleigh_ritz();
    }
    double rmaxK = 0.0;
    for (int j = 0; j < K; ++j) rmaxK = std::max(rmaxK, rs[j]);
    if (g_cf_debug)
      fprintf(stderr, "[chfsi] it %d: cut %.5f theta[0] %.6f theta[K-1] %.6f theta[nb-1] %.6f max res(K) %.2e cost %.1f\n",
              it, f.c, th[0], th[K - 1], th[nb - 1], rmaxK, cost);
    if (!std::isfinite(rmaxK)) return false;
    if (rmaxK <= tol && last_ratio >= 0.05) {  // accepted only on a basis that one Cholesky-QR pass made orthonormal
      converged = true;
      break;
    }
    // ---- next filter: cut, per-column need, per-group degrees, conditioning cap, segments
    f.c = (th[K - 1] > f.c) ? std::max(th[nb - 1], f.c) : th[nb - 1];
    f.top = std::max(th[0], f.c + 1e-3 * gnorm);
    if (!(f.c > f.a)) return false;
    const double e = 0.5 * (f.c - f.a), cen = 0.5 * (f.c + f.a);
    std::vector<double> rate(nb), need(nb, 0.0);
    for (int j = 0; j < nb; ++j) rate[j] = std::acosh(std::max(1.0, (th[j] - cen) / e));
    double mw = 0.0, cap = 1e9;
    for (int j = 0; j < K; ++j) {
      if (rs[j] > tol) need[j] = std::log(rs[j] / tol) / std::max(rate[j], 1e-2) * 1.06 + 1.0;
      mw = std::max(mw, need[j]);
    }
    for (int j = 0; j < nb; ++j) {
      const double rr = std::min(1.0, std::max(rs[j] / gnorm, 1e-16));
      cap = std::min(cap, std::log(1e6 / rr) / std::max(rate[j] - rate[nb - 1], 1e-3));
    }
    const int mcap = (int)std::max(8.0, std::min(60.0, cap));
    std::vector<int> want(ng);
    for (int g = 0; g < ng; ++g) {
      double dg = 0.0;
      for (int j = g * CF_BN; j < std::min((g + 1) * CF_BN, K); ++j) dg = std::max(dg, need[j]);
      if ((g + 1) * CF_BN > K) dg = mw;  // guards follow the slowest wanted column
      want[g] = (int)std::ceil(std::min(dg, 120.0));
    }
    for (int g = 1; g < ng; ++g) want[g] = std::max(want[g], want[g - 1]);
    // segments of at most mcap degrees each (first one now, the rest after re-orthonormalisation)
    std::vector<std::vector<int>> segs;
    std::vector<int> rem = want;
    while (rem[ng - 1] > 0 && segs.size() < 4) {
      std::vector<int> sg(ng);
      for (int g = 0; g < ng; ++g) {
        sg[g] = std::min(rem[g], mcap);
        rem[g] -= sg[g];
      }
      segs.push_back(sg);
    }
    if (segs.empty()) segs.push_back(std::vector<int>(ng, 0));  // residuals are there: only re-orthonormalise + RR
    deg = segs[0];
    for (size_t q = segs.size(); q-- > 1;) pending.push_back(segs[q]);
    if (g_cf_debug) {
      fprintf(stderr, "[chfsi]   next: cut %.5f mcap %d want", f.c, mcap);
      for (int g = 0; g < ng; ++g) fprintf(stderr, " %d", want[g]);
      fprintf(stderr, "\n");
    }
  }
  if (!converged) return false;
  FLGP_CUDA(cudaMemcpyAsync(lam, theta.p, sizeof(double) * K, cudaMemcpyDeviceToDevice, c->stream));
  FLGP_CUDA(cudaMemcpyAsync(Y, Xb[cur].p, sizeof(double) * (size_t)s * K, cudaMemcpyDeviceToDevice, c->stream));
  sync(c);
  return true;
}

}  // namespace flgp
