// gemm.cu — small/medium dense fp64 products of the path: the heat-kernel covariance
// H = V0 diag(exp(-t(1-lambda))) V1^T  (HK_from_spectrum_cpp, /root/reference/src/Spectrum.cpp:83-94),
// the K x K blocks of the GPR tail and the folds W M W^T.  One strided, shared-memory tiled FMA
// kernel (64 x 64 tile, 4 x 4 per thread, ascending-k accumulation => deterministic).
// The diagonal scaling is fused into the A-operand load, as the reference scales V0's columns first.
#include <algorithm>

#include "kernels.cuh"
#include "tma_dmma.cuh"

namespace flgp {

namespace {

// ---- FP64 tensor-core GEMM, operands staged by TMA ---------------------------------------------------------------
// C(i, j) = sum_k A(i,k) sc[k] B(j,k), A: M x K and B: N x K, both row-major (k contiguous) = the "row.col" operand
// layout of mma.sync.m8n8k4.f64 (tcgen05 has no fp64 kind; DMMA is the fp64 tensor path of sm_100a).
// CTA tile 64 x 64, K step 16: one TMA box of 64 rows x 128 bytes per operand and stage, SWIZZLE_128B so that the
// fragment loads (8 rows x 4 k) are bank-conflict free; DG_STAGES-deep mbarrier pipeline, one producer thread.
// 4 warps, each a 32 x 32 sub-tile = 4 x 4 DMMA tiles (32 accumulator registers per thread).
__global__ void __launch_bounds__(128)
dmma_gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const double* __restrict__ sc, int64_t M, int64_t N, int K, double* __restrict__ C, int64_t ldc,
                    int kchunk, int64_t csplit) {
  // blockIdx.z = K slab [kbeg, kend) (kchunk is a multiple of 16); partial results csplit apart
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  C += (int64_t)blockIdx.z * csplit;
  extern __shared__ __align__(1024) unsigned char dsm_raw[];
  // 1024-byte alignment of every box is what SWIZZLE_128B requires
  double* boxes = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[DG_STAGES];
  __shared__ double scs[1024];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wm = (wid >> 1) * 32, wn = (wid & 1) * 32;  // this warp's corner inside the CTA tile
  const int i0 = blockIdx.y * DG_T, j0 = blockIdx.x * DG_T;
  const int nk = (kend - kbeg + DG_K - 1) / DG_K;
  if (tid == 0) {
    for (int st = 0; st < DG_STAGES; ++st) mbar_init(&full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int kt) {
    const int st = kt % DG_STAGES;
    double* a = boxes + (size_t)st * 2 * DG_T * DG_K;
    mbar_expect_tx(&full[st], DG_STAGE_BYTES);
    tma_load_2d(a, &mapA, kbeg + kt * DG_K, i0, &full[st]);           // rows i0.., columns kbeg+kt*16.. (zero filled outside)
    tma_load_2d(a + DG_T * DG_K, &mapB, kbeg + kt * DG_K, j0, &full[st]);
  };
  if (tid == 0)
    for (int kt = 0; kt < DG_STAGES - 1 && kt < nk; ++kt) issue(kt);
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int fr = lane >> 2, fk = lane & 3;  // fragment coordinates: row (or column) within the 8, k within the 4
  for (int kt = 0; kt < nk; ++kt) {
    const int st = kt % DG_STAGES;
    // the scale factors of this K slab (block-uniform, tiny)
    if (sc) {
      __syncthreads();
      if (tid < DG_K) scs[tid] = (kbeg + kt * DG_K + tid < kend) ? sc[kbeg + kt * DG_K + tid] : 0.0;
    }
    mbar_wait(&full[st], (kt / DG_STAGES) & 1);
    if (sc) __syncthreads();
    const double* As = boxes + (size_t)st * 2 * DG_T * DG_K;
    const double* Bs = As + DG_T * DG_K;
#pragma unroll
    for (int ks = 0; ks < DG_K; ks += 4) {
      double af[4], bf[4];
      const double sk = sc ? scs[ks + fk] : 1.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) af[a] = box_at(As, wm + a * 8 + fr, ks + fk) * sk;
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = box_at(Bs, wn + b * 8 + fr, ks + fk);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    __syncthreads();  // everyone is done with stage st-of-(kt): the producer may refill the slot freed one step ago
    if (tid == 0 && kt + DG_STAGES - 1 < nk) issue(kt + DG_STAGES - 1);
  }
  // epilogue: thread holds C(row = lane/4, cols = 2 (lane%4) + {0,1}) of every 8 x 8 tile
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t i = i0 + wm + a * 8 + fr, j = j0 + wn + b * 8 + 2 * fk;
      if (i < M && j < N) C[i + ldc * j] = acc[a][b][0];
      if (i < M && j + 1 < N) C[i + ldc * (j + 1)] = acc[a][b][1];
    }
}

constexpr int GT = 64, GK = 16, GPAD = 66;

// C(i,j) = sum_k (A(i,k) * sc[k]) * B(k,j);  element strides for every operand.
__global__ void __launch_bounds__(256)
gemm_strided_kernel(const double* __restrict__ A, int64_t ars, int64_t acs, const double* __restrict__ B,
                    int64_t brs, int64_t bcs, const double* __restrict__ sc, int64_t M, int64_t N, int K,
                    double* __restrict__ C, int64_t crs, int64_t ccs, int kchunk, int64_t csplit) {
  // blockIdx.z = K slab (split-K for short-and-wide products: partial results csplit apart, summed in slab order by
  // splitk_reduce_kernel, so the result does not depend on the schedule)
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  C += (int64_t)blockIdx.z * csplit;
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
  for (int k0 = kbeg; k0 < kend; k0 += GK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + q * 256;
      int kk = a_kfast ? (e & 15) : (e >> 6), m = a_kfast ? (e >> 4) : (e & 63);
      int k = k0 + kk;
      int64_t i = i0 + m;
      double v = 0.0;
      if (k < kend && i < M) {
        v = A[i * ars + k * acs];
        if (sc) v = __dmul_rn(v, sc[k]);
      }
      As[kk][m] = v;
      kk = b_kfast ? (e & 15) : (e >> 6);
      m = b_kfast ? (e >> 4) : (e & 63);
      k = k0 + kk;
      int64_t j = j0 + m;
      Bs[kk][m] = (k < kend && j < N) ? B[k * brs + j * bcs] : 0.0;
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

__global__ void splitk_reduce_kernel(const double* __restrict__ part, int64_t len, int nsplit, double* __restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= len) return;
  double a = 0.0;
  for (int z = 0; z < nsplit; ++z) a += part[(int64_t)z * len + e];
  out[e] = a;
}

void gemm_strided(Ctx* c, const double* A, int64_t ars, int64_t acs, const double* B, int64_t brs, int64_t bcs,
                  const double* sc, int64_t M, int64_t N, int K, double* C, int64_t crs, int64_t ccs) {
  if (M <= 0 || N <= 0) return;
  dim3 grid(ceil_div(N, GT), ceil_div(M, GT));
  FLGP_LAUNCH(c, gemm_strided_kernel, grid, 256, 0, A, ars, acs, B, brs, bcs, sc, M, N, K, C, crs, ccs, K > 0 ? K : 1,
              (int64_t)0);
}
// the same with the contraction split over up to 64 slabs when the output is too small to fill the GPU
// (dense column-major M x N output, ld M)
void gemm_strided_splitk(Ctx* c, const double* A, int64_t ars, int64_t acs, const double* B, int64_t brs, int64_t bcs,
                         int64_t M, int64_t N, int K, double* C) {
  if (M <= 0 || N <= 0) return;
  const int tiles = ceil_div(N, GT) * ceil_div(M, GT);
  int nsplit = std::min(64, std::max(1, std::min(K / 128, (2 * c->sm_count) / tiles)));
  if (nsplit <= 1) {
    gemm_strided(c, A, ars, acs, B, brs, bcs, nullptr, M, N, K, C, 1, M);
    return;
  }
  const int kchunk = ceil_div(ceil_div(K, nsplit), GK) * GK;
  nsplit = ceil_div(K, kchunk);
  DevBuf<double> part((size_t)nsplit * M * N);
  dim3 grid(ceil_div(N, GT), ceil_div(M, GT), nsplit);
  FLGP_LAUNCH(c, gemm_strided_kernel, grid, 256, 0, A, ars, acs, B, brs, bcs, (const double*)nullptr, M, N, K, part.p,
              (int64_t)1, M, kchunk, M * N);
  FLGP_LAUNCH(c, splitk_reduce_kernel, ceil_div(M * N, 256), 256, 0, part.p, M * N, nsplit, C);
  sync(c);  // part is released on return
}

}  // namespace

void gemm_nt_run(Ctx* c, const double* A, const double* B, const double* sc, int64_t M, int64_t N, int K,
                 double* C, int64_t ldc) {
  gemm_nt_ld_run(c, A, K, B, K, sc, M, N, K, C, ldc);
}

void gemm_nt_ld_run(Ctx* c, const double* A, int64_t lda, const double* B, int64_t ldb, const double* sc, int64_t M,
                    int64_t N, int K, double* C, int64_t ldc) {
  if (M <= 0 || N <= 0) return;
  // TMA needs 16-byte aligned bases and row pitches; anything else (odd leading dimensions) takes the FMA kernel
  const bool tma_ok = (lda % 2 == 0) && (ldb % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(B) % 16 == 0) && K >= 1 && M * (int64_t)N >= 64 * 64;
  if (!tma_ok) {
    gemm_strided(c, A, lda, 1, B, 1, ldb, sc, M, N, K, C, 1, ldc);
    return;
  }
  const CUtensorMap mapA = make_map(A, M, K, lda), mapB = make_map(B, N, K, ldb);
  const size_t smem = (size_t)DG_STAGES * DG_STAGE_BYTES + 1024;
  FLGP_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(N, DG_T), ceil_div(M, DG_T));
  FLGP_LAUNCH(c, dmma_gemm_nt_kernel, grid, 128, smem, mapA, mapB, sc, M, N, K, C, ldc, K > 0 ? ((K + 15) / 16) * 16 : 16,
              (int64_t)0);
}

void gemm_general_run(Ctx* c, const double* A, int64_t ars, int64_t acs, const double* B, int64_t brs, int64_t bcs,
                      int64_t M, int64_t N, int K, double* C, int64_t crs, int64_t ccs) {
  gemm_strided(c, A, ars, acs, B, brs, bcs, nullptr, M, N, K, C, crs, ccs);
}
void gemm_general_splitk_run(Ctx* c, const double* A, int64_t ars, int64_t acs, const double* B, int64_t brs,
                             int64_t bcs, int64_t M, int64_t N, int K, double* C) {
  gemm_strided_splitk(c, A, ars, acs, B, brs, bcs, M, N, K, C);
}

void gemm_nn_run(Ctx* c, const double* A, const double* B, int64_t M, int64_t N, int K, double* C) {
  gemm_strided(c, A, K, 1, B, N, 1, nullptr, M, N, K, C, N, 1);
}

void gemm_tn_splitk_run(Ctx* c, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t M, int64_t N,
                        int Kd, double* C) {
  if (M <= 0 || N <= 0) return;
  const bool tma_ok = (lda % 2 == 0) && (ldb % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(B) % 16 == 0) && M * N >= 64 * 64 && Kd >= 256;
  if (!tma_ok) {
    gemm_strided_splitk(c, A, lda, 1, B, 1, ldb, M, N, Kd, C);
    return;
  }
  // A^T B with both operands column-major = the NT product of the row-major views (M x Kd) and (N x Kd): FP64 tensor
  // cores, the contraction split into slabs so that the few output tiles still fill the GPU, slabs summed in order
  const int tiles = ceil_div(N, DG_T) * ceil_div(M, DG_T);
  int nsplit = std::max(1, std::min(Kd / 64, (2 * c->sm_count) / tiles));
  const int kchunk = ceil_div(ceil_div(Kd, nsplit), DG_K) * DG_K;
  nsplit = ceil_div(Kd, kchunk);
  const CUtensorMap mapA = make_map(A, M, Kd, lda), mapB = make_map(B, N, Kd, ldb);
  const size_t smem = (size_t)DG_STAGES * DG_STAGE_BYTES + 1024;
  FLGP_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(N, DG_T), ceil_div(M, DG_T), nsplit);
  if (nsplit == 1) {
    FLGP_LAUNCH(c, dmma_gemm_nt_kernel, grid, 128, smem, mapA, mapB, (const double*)nullptr, M, N, Kd, C, M, kchunk,
                (int64_t)0);
    return;
  }
  DevBuf<double> part((size_t)nsplit * M * N);  // stream-ordered pool: released blocks are only reused on this stream
  FLGP_LAUNCH(c, dmma_gemm_nt_kernel, grid, 128, smem, mapA, mapB, (const double*)nullptr, M, N, Kd, part.p, M, kchunk,
              M * N);
  FLGP_LAUNCH(c, splitk_reduce_kernel, ceil_div(M * N, 256), 256, 0, part.p, M * N, nsplit, C);
}

void gemv_run(Ctx* c, const double* A, const double* x, int64_t M, int K, double* y) {
  gemm_strided(c, A, K, 1, x, 1, 0, nullptr, M, 1, K, y, 1, 0);
}

void gram_small_run(Ctx* c, const double* V, const double* y, int64_t n_rows, int K, double* G, double* g) {
  // G = V^T V : A(i,k) = V[k*K + i], B(k,j) = V[k*K + j]
  if (n_rows > INT32_MAX) fail(2, "too many training rows");
  gemm_strided_splitk(c, V, 1, K, V, K, 1, K, K, (int)n_rows, G);
  if (y && g) gemm_strided_splitk(c, V, 1, K, y, 1, 0, K, 1, (int)n_rows, g);
}

}  // namespace flgp
