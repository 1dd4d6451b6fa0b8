// core_math.cuh — per-thread arithmetic of the FLGP hot path, written once for device code.
//
// Everything here is `FLGP_HD` so that a test-only host shim (tests/hostcheck/) can compile the
// very same source with g++ and compare it against the oracle on the CPU box before any GPU
// time is spent.  The product only ever calls these from CUDA kernels.
//
// Arithmetic contract: fp64, round-to-nearest; multiplies and adds are NEVER
// contracted (the library is compiled with -fmad=false); fused multiply-adds appear only where
// `fma()` is written out (k-means scores, GEMMs, eigensolver).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define FLGP_HD __host__ __device__ __forceinline__
#else
#define FLGP_HD inline
#endif

namespace flgp {

// ---------------------------------------------------------------------------------------------
// Two-limb fixed point: order-independent (hence shard- and schedule-independent) summation.
// |x| < 2^E, at most `count` addends; L = ceil(log2 count), B = 62 - L bits per limb.
//   x ~ hi*q1 + lo*q2,  q1 = 2^(E-B), q2 = 2^(E-2B);  sums of hi / lo never overflow int64.
// ---------------------------------------------------------------------------------------------
struct Fx {
  double q1, q2, iq1, iq2;
  int E, B;
};

inline int fx_make(double maxabs, int64_t count, Fx* fx) {  // host only
  if (!std::isfinite(maxabs) || count < 1) return 1;
  int E = (maxabs > 0.0) ? std::ilogb(maxabs) + 1 : 0;
  if (E < -800) E = -800;
  if (E > 800) return 1;
  int L = 0;
  while (((int64_t)1 << L) < count) ++L;
  int B = 62 - L;
  fx->E = E;
  fx->B = B;
  fx->q1 = std::ldexp(1.0, E - B);
  fx->q2 = std::ldexp(1.0, E - 2 * B);
  fx->iq1 = std::ldexp(1.0, B - E);
  fx->iq2 = std::ldexp(1.0, 2 * B - E);
  return 0;
}

FLGP_HD long long fx_trunc(double v) {
#if defined(__CUDA_ARCH__)
  return __double2ll_rz(v);
#else
  return (long long)v;
#endif
}
FLGP_HD double fx_todouble(long long v) {
#if defined(__CUDA_ARCH__)
  return __ll2double_rn(v);
#else
  return (double)v;
#endif
}
FLGP_HD void fx_encode(const Fx& fx, double x, long long* hi, long long* lo) {
  long long h = fx_trunc(x * fx.iq1);
  double rem = x - fx_todouble(h) * fx.q1;
  *hi = h;
  *lo = fx_trunc(rem * fx.iq2);
}
FLGP_HD double fx_decode(const Fx& fx, long long hi, long long lo) {
  return fx_todouble(hi) * fx.q1 + fx_todouble(lo) * fx.q2;
}

// ---------------------------------------------------------------------------------------------
// libstdc++ heap primitives on (key, id) pairs, comparator key[a] < key[b].  These follow the
// ALGORITHM of std::partial_sort = __heap_select + __sort_heap (bits/stl_algo.h, stl_heap.h)
// step for step, so that the selected set AND order agree with the reference's literal
// std::partial_sort call (src/Utils.cpp:93) even on exact ties.
// ---------------------------------------------------------------------------------------------
FLGP_HD void heap_adjust(double* hk, int* hi, int hole, int len, double vk, int vi) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (hk[child] < hk[child - 1]) child--;
    hk[hole] = hk[child];
    hi[hole] = hi[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    hk[hole] = hk[child - 1];
    hi[hole] = hi[child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && hk[parent] < vk) {
    hk[hole] = hk[parent];
    hi[hole] = hi[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  hk[hole] = vk;
  hi[hole] = vi;
}
FLGP_HD void heap_make(double* hk, int* hi, int len) {
  if (len < 2) return;
  int parent = (len - 2) / 2;
  while (true) {
    double vk = hk[parent];
    int vi = hi[parent];
    heap_adjust(hk, hi, parent, len, vk, vi);
    if (parent == 0) return;
    parent--;
  }
}
// element (k,i) of the tail [middle,last): replaces the heap top iff k < top key
FLGP_HD void heap_select_step(double* hk, int* hi, int r, double k, int i) {
  if (k < hk[0]) heap_adjust(hk, hi, 0, r, k, i);
}
FLGP_HD void heap_sort(double* hk, int* hi, int r) {
  int last = r;
  while (last > 1) {
    --last;
    double vk = hk[last];
    int vi = hi[last];
    hk[last] = hk[0];
    hi[last] = hi[0];
    heap_adjust(hk, hi, 0, last, vk, vi);
  }
}

// The same three primitives over an accessor (getk/geti/set by index) instead of arrays, so that a heap
// of compile-time size can live entirely in registers (RegHeap below): every index is resolved by an
// unrolled compare-select chain, never by an address.
template <class H>
FLGP_HD void heap_adjust_acc(H& h, int hole, int len, double vk, int vi) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (h.getk(child) < h.getk(child - 1)) child--;
    h.set(hole, h.getk(child), h.geti(child));
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h.set(hole, h.getk(child - 1), h.geti(child - 1));
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && h.getk(parent) < vk) {
    h.set(hole, h.getk(parent), h.geti(parent));
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h.set(hole, vk, vi);
}
template <class H>
FLGP_HD void heap_make_acc(H& h, int len) {
  if (len < 2) return;
  int parent = (len - 2) / 2;
  while (true) {
    double vk = h.getk(parent);
    int vi = h.geti(parent);
    heap_adjust_acc(h, parent, len, vk, vi);
    if (parent == 0) return;
    parent--;
  }
}
template <class H>
FLGP_HD void heap_sort_acc(H& h, int r) {
  int last = r;
  while (last > 1) {
    --last;
    double vk = h.getk(last);
    int vi = h.geti(last);
    h.set(last, h.getk(0), h.geti(0));
    heap_adjust_acc(h, 0, last, vk, vi);
  }
}
template <int R>
struct RegHeap {
  double k[R];
  int id[R];
  FLGP_HD double getk(int i) const {
    double v = k[0];
#pragma unroll
    for (int u = 1; u < R; ++u) v = (i == u) ? k[u] : v;
    return v;
  }
  FLGP_HD int geti(int i) const {
    int v = id[0];
#pragma unroll
    for (int u = 1; u < R; ++u) v = (i == u) ? id[u] : v;
    return v;
  }
  FLGP_HD void set(int i, double kv, int iv) {
#pragma unroll
    for (int u = 0; u < R; ++u)
      if (i == u) {
        k[u] = kv;
        id[u] = iv;
      }
  }
};

// ---------------------------------------------------------------------------------------------
// v_to_z_cpp (src/lae.cpp:137-153): projection of v (length r) onto the simplex.
// `vd` is scratch of length r.  Sequential cumulative sum, as std::partial_sum.
// ---------------------------------------------------------------------------------------------
template <int RT>
FLGP_HD void simplex_project(const double* v, int r_in, double* z, double* vd) {
  const int r = RT ? RT : r_in;
#pragma unroll
  for (int a = 0; a < r; ++a) vd[a] = v[a];
  // insertion sort, descending (values only: the order of equal keys is irrelevant)
#pragma unroll
  for (int a = 1; a < r; ++a) {
    double key = vd[a];
    int b = a - 1;
    while (b >= 0 && vd[b] < key) {
      vd[b + 1] = vd[b];
      --b;
    }
    vd[b + 1] = key;
  }
  double cs = 0.0;
  // theta = (cs_rho - 1) / rho is the very quotient formed at a = rho - 1 (same operands): it is kept instead of
  // divided again (an IEEE fp64 division is ~30 instructions on the fp64 pipe); rho == 0 -> (0 - 1) / 0 = -inf
  double theta = -INFINITY;
#pragma unroll
  for (int a = 0; a < r; ++a) {
    cs = (a == 0) ? vd[0] : cs + vd[a];
    const double qv = (cs - 1.0) / (double)(a + 1);
    double vstar = vd[a] - qv;
    if (vstar > 0) theta = qv;  // the reference scans from r down and stops at the LAST index with v* > 0
  }
#pragma unroll
  for (int a = 0; a < r; ++a) {
    double t = v[a] - theta;
    z[a] = (t > 0.0) ? t : 0.0;  // std::max(t, 0.0)
  }
}

// exact 2^j as a double (inf for j >= 1024), replacing pow(2, j) at src/lae.cpp:110
FLGP_HD double pow2i(int j) {
  if (j >= 1024) return INFINITY;
#if defined(__CUDA_ARCH__)
  return __hiloint2double((1023 + j) << 20, 0);
#else
  return std::ldexp(1.0, j);
#endif
}

// ---------------------------------------------------------------------------------------------
// local_anchor_embedding_cpp (src/lae.cpp:76-133).
//   RT, DT : compile-time r, d (0 = run-time, arrays sized RMAX).
//   UAcc   : functor (a,k) -> U(a,k), the r gathered anchors.
//   x      : the point (length d), readable by index.
// Operation order is the oracle's (oracle/flgp_oracle.cpp orc_lae_point): every dot product and
// norm is a sequential left-to-right sum of separately rounded products.
// ---------------------------------------------------------------------------------------------
constexpr int LAE_RMAX = 16;
constexpr int LAE_T = 100;
constexpr int LAE_JCAP = 1100;

struct LaeStats {
  int iters, backtracks;
};

// 1 / beta for the back-tracking step size.  beta = 2^j * beta_curr with beta_curr = 1 initially, so beta is always
// an exact power of two >= 1 (or +inf): its reciprocal is an exponent flip, bit-identical to the IEEE division.
// Anything else (unreachable) takes the division.
FLGP_HD double lae_recip(double beta) {
#if defined(__CUDA_ARCH__)
  const int hi = __double2hiint(beta), lo = __double2loint(beta);
  const int e = (hi >> 20) & 0x7ff;
  if (lo == 0 && (hi & 0x800fffff) == 0 && e >= 1023 && e <= 2045) return __hiloint2double((2046 - e) << 20, 0);
#endif
  return 1.0 / beta;
}

// The momentum ratio alpha_t = (delta_{t-1} - 1) / delta_t of iteration t depends on t only
// (delta_0 = 0, delta_1 = 1, delta_{t+1} = (1 + sqrt(1 + 4 delta_t^2)) / 2): a table of LAE_T correctly rounded
// values replaces one division and one square root per iteration and point.  IEEE sqrt and division are correctly
// rounded on the host and on sm_100a alike, so the table equals the on-the-fly values bit for bit.
inline void lae_alpha_table(double* tab) {
  double delta_prev = 0.0, delta_curr = 1.0;
  for (int t = 0; t < 100; ++t) {
    tab[t] = (delta_prev - 1.0) / delta_curr;
    delta_prev = delta_curr;
    delta_curr = (1.0 + std::sqrt(1.0 + 4.0 * delta_curr * delta_curr)) / 2.0;
  }
}

// The iteration proper, given UUt (r x r, row stride RA), xUt (r) and the objective w -> |x - w U|^2 / 2 as a callable
// (sequential per thread in lae_solve; warp-cooperative in lae.cu's large-d kernel: same values, same order).
template <int RT, class Obj>
FLGP_HD LaeStats lae_iterate(int r_in, const double* UUt, const double* xUt, Obj&& objective, double* z_out,
                             const double* alpha_tab = nullptr) {
  constexpr int RA = RT ? RT : LAE_RMAX;
  const int r = RT ? RT : r_in;
  double zp[RA], v[RA], g[RA], vt[RA], z[RA], scratch[RA];
  double* zc = z_out;  // the caller's array IS the current iterate (no copy at exit; see DESIGN.md §7 note on nvcc)
#pragma unroll
  for (int a = 0; a < r; ++a) {
    zp[a] = 1.0 / (double)r;
    zc[a] = 1.0 / (double)r;
  }
  double delta_prev = 0.0, delta_curr = 1.0, beta_curr = 1.0;
  int t = 0, nbt = 0;
  for (t = 0; t < LAE_T; ++t) {
    const double alpha = alpha_tab ? alpha_tab[t] : (delta_prev - 1.0) / delta_curr;
#pragma unroll
    for (int a = 0; a < r; ++a) v[a] = zc[a] + alpha * (zc[a] - zp[a]);
    const double g_v = objective(v);
#pragma unroll
    for (int b = 0; b < r; ++b) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < r; ++a) s = s + v[a] * UUt[a * RA + b];
      g[b] = s - xUt[b];
    }
    int j = 0;
    while (true) {
      const double beta = pow2i(j) * beta_curr;
      const double ib = lae_recip(beta);
#pragma unroll
      for (int a = 0; a < r; ++a) vt[a] = v[a] - ib * g[a];
      simplex_project<RT>(vt, r, z, scratch);
      const double g_z = objective(z);
      double dot = 0.0, sq = 0.0;
#pragma unroll
      for (int a = 0; a < r; ++a) {
        double dz = z[a] - v[a];
        dot = dot + g[a] * dz;
        sq = sq + dz * dz;
      }
      const double g_tilde = g_v + dot + beta * sq / 2.0;
      if (g_z <= g_tilde || j >= LAE_JCAP) {
        beta_curr = beta;
#pragma unroll
        for (int a = 0; a < r; ++a) {
          zp[a] = zc[a];
          zc[a] = z[a];
        }
        break;
      }
      ++j;
      ++nbt;
    }
    if (!alpha_tab) {
      delta_prev = delta_curr;
      delta_curr = (1.0 + sqrt(1.0 + 4.0 * delta_curr * delta_curr)) / 2.0;
    }
    double sq = 0.0;
#pragma unroll
    for (int a = 0; a < r; ++a) {
      double dz = zc[a] - zp[a];
      sq = sq + dz * dz;
    }
    if (sq < 1e-5) {
      ++t;
      break;
    }
  }
  LaeStats st;
  st.iters = t;
  st.backtracks = nbt;
  return st;
}

template <int RT, int DT, class XAcc, class UAcc>
FLGP_HD LaeStats lae_solve(int r_in, int d_in, const XAcc& x, const UAcc& U, double* z_out,
                           const double* alpha_tab = nullptr) {
  constexpr int RA = RT ? RT : LAE_RMAX;
  const int r = RT ? RT : r_in;
  const int d = DT ? DT : d_in;
  double UUt[RA * RA], xUt[RA];
#pragma unroll
  for (int a = 0; a < r; ++a) {
#pragma unroll
    for (int b = 0; b < r; ++b) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < d; ++k) s = s + U(a, k) * U(b, k);
      UUt[a * RA + b] = s;
    }
  }
#pragma unroll
  for (int a = 0; a < r; ++a) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < d; ++k) s = s + x(k) * U(a, k);
    xUt[a] = s;
  }
  auto objective = [&](const double* w) {
    double sq = 0.0;
#pragma unroll
    for (int k = 0; k < d; ++k) {
      double wu = 0.0;
#pragma unroll
      for (int a = 0; a < r; ++a) wu = wu + w[a] * U(a, k);
      double df = x(k) - wu;
      sq = sq + df * df;
    }
    return sq / 2.0;
  };
  return lae_iterate<RT>(r, UUt, xUt, objective, z_out, alpha_tab);
}

// ---------------------------------------------------------------------------------------------
// Symmetric tridiagonal helpers (eigensolver).  d[0..n), e[0..n-1) sub-diagonal, e2 = e*e.
// ---------------------------------------------------------------------------------------------
// number of eigenvalues strictly below x (Sturm count, LAPACK dlaebz/dstebz recurrence)
FLGP_HD int sturm_count(const double* d, const double* e2, int n, double x, double pivmin) {
  int cnt = 0;
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  if (q < 0.0) ++cnt;
  for (int i = 1; i < n; ++i) {
    q = d[i] - x - e2[i - 1] / q;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) ++cnt;
  }
  return cnt;
}

// ---------------------------------------------------------------------------------------------
// Mini-batch k-means (subsample_cpp "minibatchkmeans", src/Utils.cpp:49-62; contract in DESIGN.md §2).
// ---------------------------------------------------------------------------------------------
FLGP_HD uint64_t mb_mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// key of batch `it` (0-based)
FLGP_HD uint64_t mb_batch_key(uint64_t seed, int it) { return mb_mix64(seed ^ ((uint64_t)(it + 1) * 0xD1B54A32D192ED03ull)); }
// k -> perm(k): a bijection of [0, n) (4-round Feistel network on the smallest even number of bits covering n,
// cycle-walked back into range), so that the first b values are b DISTINCT rows: a batch drawn without replacement.
FLGP_HD int64_t mb_perm(int64_t k, int64_t n, uint64_t key) {
  int bits = 2;
  while (((int64_t)1 << bits) < n) bits += 2;
  const int half = bits / 2;
  const uint64_t mask = ((uint64_t)1 << half) - 1;
  uint64_t x = (uint64_t)k;
  do {
    uint64_t L = x >> half, R = x & mask;
    for (int rd = 0; rd < 4; ++rd) {
      const uint64_t f = mb_mix64(R ^ (key + (uint64_t)rd * 0xA24BAED4963EE407ull)) & mask;
      const uint64_t t = L ^ f;
      L = R;
      R = t;
    }
    x = (L << half) | R;
  } while (x >= (uint64_t)n);
  return (int64_t)x;
}
// The Lloyd score rule on a group of G centres at once (mb_assign_kernel's per-thread work): e[g] starts at
// cn[g] = |c_g|^2 + 2 d maxabs^2 and takes fma(x_q, -2 c_gq, .) for q ascending; rec = G x dp records of -2 c, zero
// padded from d to dp (a multiple of QC), x(q) = coordinate q of the point (read only for q < d): the padding adds
// fma(0, 0, e) = e, so the chain equals the d-term chain of the oracle bit for bit.
template <int G, int QC, class XAcc>
FLGP_HD void mb_score_group(const XAcc& x, int d, int dp, const double* rec, const double* cn, double* e) {
#pragma unroll
  for (int g = 0; g < G; ++g) e[g] = cn[g];
  for (int q0 = 0; q0 < dp; q0 += QC) {
    double xr[QC];
#pragma unroll
    for (int qq = 0; qq < QC; ++qq) xr[qq] = (q0 + qq < d) ? x(q0 + qq) : 0.0;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int qq = 0; qq < QC; ++qq) e[g] = fma(xr[qq], rec[g * dp + q0 + qq], e[g]);
  }
}
// arg-min over the group's live centres j0 .. min(j0 + G, s) - 1, lowest index on ties (centres come in index order)
template <int G>
FLGP_HD void mb_argmin_group(const double* e, int j0, int s, double* best, int* bj) {
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int j = j0 + g;
    if (j < s && (j == 0 || e[g] < *best)) {
      *best = e[g];
      *bj = j;
    }
  }
}
// record of centre j for the group staging: -2 c (q < d), 0 beyond; cn = fma chain of c^2 plus m2
FLGP_HD double mb_centre_norm(const double* C, int64_t ldc, int j, int d, double m2) {
  double a = 0.0;
  for (int q = 0; q < d; ++q) {
    const double c = C[j + ldc * q];
    a = fma(c, c, a);
  }
  return a + m2;
}
// Sculley's per-sample update with the per-centre learning rate eta = 1 / count: c <- (1 - eta) c + eta x, every
// operation rounded separately
FLGP_HD double mb_update_coord(double c, double x, double eta) { return (1.0 - eta) * c + eta * x; }

}  // namespace flgp
