// lae.cu — Local Anchor Embedding weights and the fixed-r CSR cross-similarity matrix
// (replaces LAE_cpp / local_anchor_embedding_cpp / v_to_z_cpp, /root/reference/src/lae.cpp:15-153,
//  and the sparse output of KNN_cpp, src/Utils.cpp:145-189).
//
// One thread owns one point: gathers its r anchors into registers, runs the Nesterov projected
// gradient with doubling back-tracking in the oracle's exact operation order (core_math.cuh
// lae_solve; no FMA contraction, sequential reductions), sorts its r (column, weight) pairs and
// writes its CSR row.  p[i] = i*r is implicit; explicit zeros are kept (src/lae.cpp:61-67).
//
// Roofline: HBM (8d + 4r bytes read, 12r written per point); the solver itself is latency /
// divergence bound (data-dependent 1..100 outer iterations) and is reported with its iteration
// count, as SURVEY.md §8d asks.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "kernels.cuh"

namespace flgp {

namespace {

// momentum ratios alpha_t of the Nesterov iteration (core_math.cuh lae_alpha_table), filled once per process
__device__ double g_lae_alpha[LAE_T];

template <int R, int D>
struct RegU {
  double u[R][D];
  __device__ __forceinline__ double operator()(int a, int k) const { return u[a][k]; }
};
template <int D>
struct RegX {
  double x[D];
  __device__ __forceinline__ double operator()(int k) const { return x[k]; }
};
struct GlobU {  // anchors read in place (large d): U(c_a, k)
  const double* U;
  int64_t ldu;
  int c[LAE_RMAX];
  __device__ __forceinline__ double operator()(int a, int k) const { return U[c[a] + ldu * k]; }
};
struct GlobURow {  // the same from a row-major copy (k contiguous): a thread streams its r anchor rows, so every
                   // 32-byte sector it fetches is used four times instead of once
  const double* U;
  int64_t off[LAE_RMAX];
  int c[LAE_RMAX];
  __device__ __forceinline__ double operator()(int a, int k) const { return U[off[a] + k]; }
};
struct GlobX {
  const double* X;
  int64_t ldx;
  __device__ __forceinline__ double operator()(int k) const { return X[ldx * k]; }
};

template <int RMAXT>
__device__ __forceinline__ void write_row(int r, const int* col, const double* z, int64_t i, int64_t n,
                                          int32_t* Zj, double* Zx, double* Wd) {
  int cj[RMAXT];
  double cz[RMAXT];
#pragma unroll
  for (int a = 0; a < RMAXT; ++a)
    if (a < r) {
      cj[a] = col[a];
      cz[a] = z[a];
      if (Wd) Wd[i + n * a] = z[a];
    }
  // insertion sort by column (columns are distinct)
#pragma unroll
  for (int a = 1; a < RMAXT; ++a)
    if (a < r) {
      int kj = cj[a];
      double kz = cz[a];
      int b = a - 1;
      while (b >= 0 && cj[b] > kj) {
        cj[b + 1] = cj[b];
        cz[b + 1] = cz[b];
        --b;
      }
      cj[b + 1] = kj;
      cz[b + 1] = kz;
    }
#pragma unroll
  for (int a = 0; a < RMAXT; ++a)
    if (a < r) {
      Zj[i * r + a] = cj[a];
      Zx[i * r + a] = cz[a];
    }
}

__device__ __forceinline__ void add_stats(long long* stats, int it, int bt) {
  if (!stats) return;
  for (int o = 16; o; o >>= 1) {
    it += __shfl_xor_sync(0xffffffffu, it, o);
    bt += __shfl_xor_sync(0xffffffffu, bt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(stats), (unsigned long long)it);
    atomicAdd(reinterpret_cast<unsigned long long*>(stats) + 1, (unsigned long long)bt);
  }
}

template <int R, int D>
__global__ void __launch_bounds__(128)
lae_small_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, const double* __restrict__ U, int64_t ldu,
                 const int32_t* __restrict__ ind, int32_t* __restrict__ Zj, double* __restrict__ Zx,
                 double* __restrict__ Wd, long long* stats, const int32_t* __restrict__ perm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int it = 0, bt = 0;
  if (i < n) {
    RegX<D> x;
    RegU<R, D> Ur;
    int col[R];
#pragma unroll
    for (int k = 0; k < D; ++k) x.x[k] = X[i + ldx * k];
#pragma unroll
    for (int a = 0; a < R; ++a) {
      col[a] = ind[i + n * a];
#pragma unroll
      for (int k = 0; k < D; ++k) Ur.u[a][k] = U[col[a] + ldu * k];
    }
    double z[R];
    const LaeStats ls = lae_solve<R, D>(R, D, x, Ur, z, g_lae_alpha);
    it = ls.iters;
    bt = ls.backtracks;
    write_row<R>(R, col, z, perm ? (int64_t)perm[i] : i, n, Zj, Zx, Wd);  // perm: input row i is output row perm[i]
  }
  add_stats(stats, it, bt);
}

// RT > 0: compile-time r (the solver's r x r state then lives in registers instead of LAE_RMAX-sized local arrays)
template <bool ROWMAJOR, int RT>
__global__ void __launch_bounds__(128)
lae_generic_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, const double* __restrict__ U,
                   int64_t ldu, int r, const int32_t* __restrict__ ind, int32_t* __restrict__ Zj,
                   double* __restrict__ Zx, double* __restrict__ Wd, long long* stats,
                   const int32_t* __restrict__ perm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int it = 0, bt = 0;
  if (i < n) {
    GlobX x{X + i, ldx};
    typename std::conditional<ROWMAJOR, GlobURow, GlobU>::type Ur;
    Ur.U = U;
    if constexpr (ROWMAJOR) {
      for (int a = 0; a < r; ++a) {
        Ur.c[a] = ind[i + n * a];
        Ur.off[a] = (int64_t)Ur.c[a] * ldu;  // ldu = row pitch of the row-major copy
      }
    } else {
      Ur.ldu = ldu;
      for (int a = 0; a < r; ++a) Ur.c[a] = ind[i + n * a];
    }
    double z[RT ? RT : LAE_RMAX];
    const LaeStats ls = lae_solve<RT, 0>(r, d, x, Ur, z, g_lae_alpha);
    it = ls.iters;
    bt = ls.backtracks;
    write_row<(RT ? RT : LAE_RMAX)>(r, Ur.c, z, perm ? (int64_t)perm[i] : i, n, Zj, Zx, Wd);
  }
  add_stats(stats, it, bt);
}

// ---- large d: one WARP per point ------------------------------------------------------------------------------
// With d in the hundreds the per-thread solver is a chain of d-long dependent sums per objective evaluation and the
// kernel's time is that latency (58 ms at n = 70000, d = 784).  Here a warp owns the point: the d products of a sum
// are formed by the 32 lanes in parallel (coalesced reads of the anchors' row-major rows) into shared memory, then
// every lane adds them up in the oracle's order (k = 0, 1, 2, ...: an LDS broadcast + DADD per term, the same value
// in all lanes, so control flow stays warp-uniform).  The r x r and r chains of the set-up run one chain per lane.
// Dynamic shared memory per warp: x (d) + products (d) + UUt / xUt (LAE_RMAX^2 + LAE_RMAX).
constexpr int LW_EXTRA = LAE_RMAX * LAE_RMAX + LAE_RMAX;

template <int RT>
__global__ void __launch_bounds__(256)
lae_warp_kernel(const double* __restrict__ X, int64_t n, int64_t ldx, int d, const double* __restrict__ Ur, int64_t ldu,
                int r, const int32_t* __restrict__ ind, int32_t* __restrict__ Zj, double* __restrict__ Zx,
                double* __restrict__ Wd, long long* stats, const int32_t* __restrict__ perm) {
  extern __shared__ double lw_sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* xs = lw_sm + (size_t)wid * (2 * (size_t)d + LW_EXTRA);
  double* ps = xs + d;
  double* uu = ps + d;          // UUt, row stride LAE_RMAX
  double* xu = uu + LAE_RMAX * LAE_RMAX;
  int it_tot = 0, bt_tot = 0;
  for (int64_t i = (int64_t)blockIdx.x * nw + wid; i < n; i += (int64_t)gridDim.x * nw) {
    int col[LAE_RMAX];
    const double* row[LAE_RMAX];
    for (int a = 0; a < r; ++a) {
      col[a] = ind[i + n * a];
      row[a] = Ur + (int64_t)col[a] * ldu;
    }
    __syncwarp();
    for (int k = lane; k < d; k += 32) xs[k] = X[i + ldx * k];
    __syncwarp();
    // set-up chains: (a, b) with a <= b (the product commutes: UUt is exactly symmetric), then x.U_a
    const int npair = r * (r + 1) / 2, nchain = npair + r;
    for (int c0 = 0; c0 < nchain; c0 += 32) {
      const int ch = c0 + lane;
      if (ch < nchain) {
        int a = 0, b = 0;
        const double* pa;
        const double* pb;
        if (ch < npair) {
          int q = ch;
          while (q >= r - a) {
            q -= r - a;
            ++a;
          }
          b = a + q;
          pa = row[a];
          pb = row[b];
        } else {
          a = ch - npair;
          pa = xs;
          pb = row[a];
        }
        double sacc = 0.0;
        int k = 0;
        for (; k + 4 <= d; k += 4) {
          const double p0 = pa[k] * pb[k], p1 = pa[k + 1] * pb[k + 1], p2 = pa[k + 2] * pb[k + 2],
                       p3 = pa[k + 3] * pb[k + 3];
          sacc = sacc + p0;
          sacc = sacc + p1;
          sacc = sacc + p2;
          sacc = sacc + p3;
        }
        for (; k < d; ++k) sacc = sacc + pa[k] * pb[k];
        if (ch < npair) {
          uu[a * LAE_RMAX + b] = sacc;
          uu[b * LAE_RMAX + a] = sacc;
        } else {
          xu[a] = sacc;
        }
      }
    }
    __syncwarp();
    constexpr int RA = RT ? RT : LAE_RMAX;  // row stride lae_iterate<RT> expects
    double UUt[RA * RA], xUt[RA];
#pragma unroll
    for (int a = 0; a < RA; ++a)
      if (a < r) {
        xUt[a] = xu[a];
#pragma unroll
        for (int b = 0; b < RA; ++b)
          if (b < r) UUt[a * RA + b] = uu[a * LAE_RMAX + b];
      }
    auto objective = [&](const double* w) {
      __syncwarp();
      for (int k = lane; k < d; k += 32) {
        double wu = 0.0;
        for (int a = 0; a < r; ++a) wu = wu + w[a] * row[a][k];
        const double df = xs[k] - wu;
        ps[k] = df * df;
      }
      __syncwarp();
      double sq = 0.0;
      int k = 0;
      for (; k + 4 <= d; k += 4) {
        const double p0 = ps[k], p1 = ps[k + 1], p2 = ps[k + 2], p3 = ps[k + 3];
        sq = sq + p0;
        sq = sq + p1;
        sq = sq + p2;
        sq = sq + p3;
      }
      for (; k < d; ++k) sq = sq + ps[k];
      return sq / 2.0;
    };
    double z[RA];
    const LaeStats ls = lae_iterate<RT>(r, UUt, xUt, objective, z, g_lae_alpha);
    if (lane == 0) {
      it_tot += ls.iters;
      bt_tot += ls.backtracks;
      write_row<RA>(r, col, z, perm ? (int64_t)perm[i] : i, n, Zj, Zx, Wd);
    }
  }
  add_stats(stats, it_tot, bt_tot);
}

__global__ void knn_to_csr_kernel(int64_t n, int r, const int32_t* __restrict__ ind, const double* __restrict__ dist,
                                  int32_t* __restrict__ Zj, double* __restrict__ Zx, const int32_t* __restrict__ perm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int col[32];
  double v[32];
  for (int a = 0; a < r; ++a) {
    col[a] = ind[i + n * a];
    v[a] = dist[i + n * a];
  }
  write_row<32>(r, col, v, perm ? (int64_t)perm[i] : i, n, Zj, Zx, nullptr);
}

__global__ void se_weights_kernel(const double* __restrict__ dist, int64_t len, double denom, double* out) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < len) out[e] = exp(-dist[e] / denom);
}

__global__ void lae_point_kernel(const double* x, int d, const double* Ur, int r, double* z) {
  if (threadIdx.x || blockIdx.x) return;
  GlobX xa{x, 1};
  GlobU ua;
  ua.U = Ur;  // r x d column-major, ld r
  ua.ldu = r;
  for (int a = 0; a < r; ++a) ua.c[a] = a;
  double zz[LAE_RMAX];
  lae_solve<0, 0>(r, d, xa, ua, zz, g_lae_alpha);
  for (int a = 0; a < r; ++a) z[a] = zz[a];
}

__global__ void simplex_project_kernel(const double* v, int r, double* z, double* scratch) {
  if (threadIdx.x || blockIdx.x) return;
  simplex_project<0>(v, r, z, scratch);
}

void lae_tables_init(Ctx* c) {
  static bool done_dev[64] = {false};  // the table is a per-device symbol
  bool& done = done_dev[c->device & 63];
  if (done) return;
  double tab[LAE_T];
  lae_alpha_table(tab);
  FLGP_CUDA(cudaMemcpyToSymbolAsync(g_lae_alpha, tab, sizeof tab, 0, cudaMemcpyHostToDevice, c->stream));
  sync(c);  // tab is a stack array
  done = true;
}

}  // namespace

void lae_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, const double* U, int s, int64_t ldu,
             int r, const int32_t* ind, int32_t* Zj, double* Zx, double* Wd, long long* stats, const int32_t* perm) {
  if (r < 1 || r > LAE_RMAX) fail(2, "LAE: r=%d outside the supported range 1..%d", r, LAE_RMAX);
  if (n <= 0) return;
  lae_tables_init(c);
  const int grid = ceil_div(n, 128);
#define LAE_CASE(R_, D_)                                                                                  \
  if (r == R_ && d == D_) {                                                                               \
    FLGP_LAUNCH(c, (lae_small_kernel<R_, D_>), grid, 128, 0, X, n, ldx, U, ldu, ind, Zj, Zx, Wd, stats,   \
                perm);                                                                                    \
    return;                                                                                               \
  }
  LAE_CASE(2, 2) LAE_CASE(3, 2) LAE_CASE(4, 2) LAE_CASE(5, 2)
  LAE_CASE(2, 3) LAE_CASE(3, 3) LAE_CASE(4, 3) LAE_CASE(5, 3)
#undef LAE_CASE
  if (d >= 32) {  // long rows: one warp per point over a row-major copy of the anchors
    const size_t per_warp = (2 * (size_t)d + LW_EXTRA) * sizeof(double);
    int warps = (int)std::min<size_t>(8, (100 * 1024) / per_warp);
    if (warps >= 1) {
      DevBuf<double> Ur((size_t)s * d);
      to_rowmajor_run(c, U, s, ldu, d, d, Ur.p);
      const size_t smem = per_warp * warps;
      const int wgrid = (int)std::min<int64_t>(ceil_div(n, warps), (int64_t)c->sm_count * 16);
#define LAE_WARP(R_)                                                                                               \
  do {                                                                                                             \
    if (smem > 40 * 1024)                                                                                          \
      FLGP_CUDA(cudaFuncSetAttribute(lae_warp_kernel<R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    FLGP_LAUNCH(c, lae_warp_kernel<R_>, wgrid, warps * 32, smem, X, n, ldx, d, Ur.p, (int64_t)d, r, ind, Zj, Zx, Wd, \
                stats, perm);                                                                                      \
  } while (0)
      switch (r) {
        case 2: LAE_WARP(2); break;
        case 3: LAE_WARP(3); break;
        case 4: LAE_WARP(4); break;
        case 5: LAE_WARP(5); break;
        default: LAE_WARP(0); break;
      }
#undef LAE_WARP
      sync(c);  // Ur is released on return
      return;
    }
  }
  if (d >= 8) {  // medium rows: read the anchors from a row-major copy (sector reuse through L1)
    DevBuf<double> Ur((size_t)s * d);
    to_rowmajor_run(c, U, s, ldu, d, d, Ur.p);
#define LAE_GEN(R_) \
  FLGP_LAUNCH(c, (lae_generic_kernel<true, R_>), grid, 128, 0, X, n, ldx, d, Ur.p, (int64_t)d, r, ind, Zj, Zx, Wd, stats, perm)
    switch (r) {
      case 2: LAE_GEN(2); break;
      case 3: LAE_GEN(3); break;
      case 4: LAE_GEN(4); break;
      case 5: LAE_GEN(5); break;
      default: LAE_GEN(0); break;
    }
#undef LAE_GEN
    sync(c);  // Ur is released on return
    return;
  }
  FLGP_LAUNCH(c, (lae_generic_kernel<false, 0>), grid, 128, 0, X, n, ldx, d, U, ldu, r, ind, Zj, Zx, Wd, stats, perm);
}

void knn_to_csr_run(Ctx* c, int64_t n, int r, const int32_t* ind, const double* dist, int32_t* Zj, double* Zx,
                    const int32_t* perm) {
  if (r < 1 || r > 32) fail(2, "r=%d outside 1..32", r);
  if (n <= 0) return;
  FLGP_LAUNCH(c, knn_to_csr_kernel, ceil_div(n, 128), 128, 0, n, r, ind, dist, Zj, Zx, perm);
}

void se_weights_run(Ctx* c, const double* dist, int64_t len, double denom, double* out) {
  if (len <= 0) return;
  FLGP_LAUNCH(c, se_weights_kernel, ceil_div(len, 256), 256, 0, dist, len, denom, out);
}

void lae_point_run(Ctx* c, const double* x, int d, const double* Ur, int r, double* z) {
  if (r < 1 || r > LAE_RMAX) fail(2, "LAE: r=%d outside the supported range 1..%d", r, LAE_RMAX);
  lae_tables_init(c);
  FLGP_LAUNCH(c, lae_point_kernel, 1, 32, 0, x, d, Ur, r, z);
}

void simplex_project_run(Ctx* c, const double* v, int r, double* z) {
  DevBuf<double> scratch(r);
  FLGP_LAUNCH(c, simplex_project_kernel, 1, 32, 0, v, r, z, scratch.p);
  sync(c);
}

}  // namespace flgp
