// kernels.cuh — host-callable stage launchers.  All pointers are DEVICE pointers unless named *_h.
// Matrices are column-major with explicit leading dimensions; CSR has exactly r entries per row.
#pragma once
#include "common.cuh"

namespace flgp {

// ---- kmeans.cu -------------------------------------------------------------------------------
// Lloyd iterations on this rank's rows; centroid sums are all-reduced (int64 limbs).
// U: s x (d+1) column-major (centres, sizes).  assign: n_local (scratch/out).
// Cluster-sorted layout left behind by the pruned passes (d <= 4): position p holds original row perm[p];
// as[p] is a centre close to that row (its final assignment), Rbits[a] + move[a] bounds |x - U_a| over the
// rows with as == a.  The KNN stage reuses it for exact candidate pruning.
struct KMeansSorted {
  bool valid = false;
  DevBuf<double> Xs;                 // n_local x d, column-major, ld n_local
  DevBuf<int32_t> perm, as;
  DevBuf<unsigned long long> Rbits;  // bit patterns of non-negative doubles
  DevBuf<double> move;
  DevBuf<int> seg_start;             // s + 1: positions [seg_start[a], seg_start[a+1]) were in cluster a at the last sort
  double eta = 0.0;
  double maxabs = 0.0;               // max |X| over all ranks
};
void kmeans_run(Ctx* c, const double* X, int64_t n_local, int64_t ldx, int d, int s, int64_t n_total,
                int64_t row_offset, const int32_t* init_idx_h, int iter_max, double* U, int32_t* assign,
                int* iters_out, KMeansSorted* sorted_out = nullptr);
double maxabs_run(Ctx* c, const double* X, int64_t n_local, int64_t ldx, int d);  // all-reduced, host value
// tot.withinss of an assignment (exact fixed point, all-reduced): the criterion stats::kmeans ranks its nstart runs by
double kmeans_withinss_run(Ctx* c, const double* X, int64_t n_local, int64_t ldx, int d, const double* U, int s,
                           const int32_t* assign, int64_t n_total);

// ---- minibatch.cu ----------------------------------------------------------------------------
// subsample_cpp "minibatchkmeans" (src/Utils.cpp:49-62): U = s x (d+1), centroids by mini-batch k-means from the
// start rows init_idx_h (batches keyed by seed, at most max_iters of them), sizes = rows per 1-NN label.  Single GPU.
void minibatch_kmeans_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, int s, const int32_t* init_idx_h,
                          int max_iters, uint64_t seed, double* U, int* iters_out);

// ---- knn.cu ----------------------------------------------------------------------------------
// ind: n x r (ld n), ascending distance, libstdc++ partial_sort tie behaviour.  dist: optional.
// sorted (optional): the cluster-sorted layout of the SAME rows left by kmeans_run; with it (d <= 4, r <= 5) every
// point scans only the anchors that can reach its top r (exact: see knn.cu), everything else is unchanged.
// out_sorted (optional, with sorted): when the pruned path runs, row p of ind/dist describes sorted position p
// (= original row sorted->perm[p]) and *out_sorted is set; the consumer un-permutes when it writes its own output.
void knn_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, const double* U, int s, int64_t ldu,
             int r, int32_t* ind, double* dist, const KMeansSorted* sorted = nullptr, bool* out_sorted = nullptr);

// ---- distsel.cu ------------------------------------------------------------------------------
// Large-d distance stages on the FP64 tensor cores (DMMA) with exact, certified selection.
// to_rowmajor: column-major n x d (ld ldx) -> row-major n x dp, zero padded (dp even: 16-byte row pitch for TMA).
void to_rowmajor_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, int dp, double* Xr);
bool dist_select_supported(int64_t n, int s, int r);
// For every row i of Xr (n x dp): the r+1 smallest of  add[j] - 2 <Xr_i, Cr_j>  over the s rows of Cr (s x dp).
// out_idx(i, 0..r-1) (ld ldo) = their indices in ascending order of value; rows whose r+1 smallest values are not
// pairwise more than thr (= thr_row[i] if given, else thr0) apart are appended to und_list (*und_count entries;
// the caller zeroes the counter) and must be re-done in the oracle's summation order.
void dist_select_run(Ctx* c, const double* Xr, int64_t n, const double* Cr, int s, int dp, const double* add, int r,
                     double thr0, const double* thr_row, int32_t* out_idx, int64_t ldo, int* und_count,
                     int32_t* und_list, const int* n_rows_dev = nullptr, double* out_val = nullptr);
// n_rows_dev (optional, device): only the first *n_rows_dev rows are live.  out_val (optional): 2 doubles per row, the
// smallest and second smallest value (r + 1 >= 2 always).

// ---- lae.cu ----------------------------------------------------------------------------------
// Zj/Zx: n*r CSR (row i at i*r), rows sorted by column.  Wd: optional dense n x r weights (ld n)
// in KNN order.  stats: optional 2 x int64 on device (iterations, back-tracks), accumulated.
// perm (optional): input row i (of X and ind) is written as output row perm[i].
void lae_run(Ctx* c, const double* X, int64_t n, int64_t ldx, int d, const double* U, int s, int64_t ldu,
             int r, const int32_t* ind, int32_t* Zj, double* Zx, double* Wd, long long* stats,
             const int32_t* perm = nullptr);
void knn_to_csr_run(Ctx* c, int64_t n, int r, const int32_t* ind, const double* dist, int32_t* Zj, double* Zx,
                    const int32_t* perm = nullptr);
void se_weights_run(Ctx* c, const double* dist, int64_t len, double denom, double* out);
// one point / one vector, for the exported helpers
void lae_point_run(Ctx* c, const double* x, int d, const double* Ur, int r, double* z);
void simplex_project_run(Ctx* c, const double* v, int r, double* z);

// ---- sparse.cu -------------------------------------------------------------------------------
// vmax bounds |Zx| (fixed-point scale); the pipeline's own Z has entries in [0, 1]
void colsum_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx, int64_t n_total,
                double* colsum, double vmax = 1.0);  // fixed-point, all-reduced
// caller-supplied CSR: throws unless 0 <= column < s, columns strictly ascending per row, values finite; max |value|
double csr_validate_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx);
void gl_apply_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, double* Zx, int mode,
                  const double* colsum, const double* num_class);
void spectrum_scale_run(Ctx* c, int s, const double* colsum, double* w);
// pmax bounds |Z(i,p) w(p) Z(i,q) w(q)| (fixed-point scale); <= 1 for the pipeline's own non-negative Z
void gram_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx, const double* w,
              int64_t n_total, double* G, double pmax = 1.0);  // s x s, fixed-point, all-reduced, symmetric
// rows of the lifted eigenvectors: out(a, k) = sum_p Z(i_a, c_p) w(c_p) Wm(c_p, k),  i_a = idx ? idx[a] : a
// Wm: s x K row-major.  out: n_rows x K, column-major (ld = ldo) if colmajor else row-major (ld = K).
void lift_rows_run(Ctx* c, int r, const int32_t* Zj, const double* Zx, const double* w, const double* Wm,
                   int K, const int32_t* idx, int64_t n_rows, double* out, int64_t ldo, bool colmajor);
// y_i = sum_p a_ip v(c_p)   (folded prediction)
void sparse_rowdot_run(Ctx* c, int64_t n, int r, const int32_t* Zj, const double* Zx, const double* w,
                       const double* v, double* y);
// q_i = add + sum_{p,q} a_ip a_iq B(c_p, c_q)   (folded posterior variance), B: s x s
void sparse_quadform_run(Ctx* c, int64_t n, int s, int r, const int32_t* Zj, const double* Zx,
                         const double* w, const double* B, double add, double* q);

// ---- gemm.cu ---------------------------------------------------------------------------------
// C(M x N, col-major ldc) = sum_k A(i,k) * sc[k] * B(j,k);  A: M x K and B: N x K, both ROW-major (ld K).
// sc may be null.  fp64 FMA accumulation in ascending k.
void gemm_nt_run(Ctx* c, const double* A, const double* B, const double* sc, int64_t M, int64_t N, int K,
                 double* C, int64_t ldc);
// the same with explicit leading dimensions (elements) of the row-major operands; FP64 tensor cores (DMMA) fed by TMA
void gemm_nt_ld_run(Ctx* c, const double* A, int64_t lda, const double* B, int64_t ldb, const double* sc, int64_t M,
                    int64_t N, int K, double* C, int64_t ldc);
// C(M x N row-major) = A(M x K row-major) * B(K x N row-major)
void gemm_nn_run(Ctx* c, const double* A, const double* B, int64_t M, int64_t N, int K, double* C);
// generic strides (FMA kernel): C(i,j) [at i*crs + j*ccs] = sum_k A(i,k) [i*ars + k*acs] B(k,j) [k*brs + j*bcs]
void gemm_general_run(Ctx* c, const double* A, int64_t ars, int64_t acs, const double* B, int64_t brs, int64_t bcs,
                      int64_t M, int64_t N, int K, double* C, int64_t crs, int64_t ccs);
// the same with the contraction split into slabs (few output tiles, long K); C dense column-major M x N
void gemm_general_splitk_run(Ctx* c, const double* A, int64_t ars, int64_t acs, const double* B, int64_t brs,
                             int64_t bcs, int64_t M, int64_t N, int K, double* C);
// small: y(M) = A(M x K row-major) x(K)
void gemv_run(Ctx* c, const double* A, const double* x, int64_t M, int K, double* y);
// G(K x K, col-major) = V^T V and g = V^T y over n_rows rows of V (row-major n_rows x K); deterministic
void gram_small_run(Ctx* c, const double* V, const double* y, int64_t n_rows, int K, double* G, double* g);
// C(M x N, col-major ld M) = A^T B;  A: Kd x M and B: Kd x N column-major (lda, ldb); deterministic split over Kd
void gemm_tn_splitk_run(Ctx* c, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t M, int64_t N,
                        int Kd, double* C);

// ---- nystrom.cu ------------------------------------------------------------------------------
// Nystrom extension (fit_nystrom_regression_gp_cpp, src/Fit.cpp:222-357).  U: s x d column-major (ld ldu).
// un (s): |u|^2;  D (s x s): squared distances between anchors;  *mean_h: their mean (host).
void nys_anchor_distances_run(Ctx* c, const double* U, int s, int64_t ldu, int d, double* un, double* D, double* mean_h);
// for one bandwidth (denom = a2 * mean): rs (s) = rowsums of Z_UU + 1e-9, lam (K) = top eigenvalues of W_UU,
// Bt (K x s row-major) = rescaled anchor eigenvectors divided by (|lam| + 1e-9)
void nys_anchor_operator_run(Ctx* c, const double* D, int s, int K, double denom, double* rs, double* lam, double* Bt);
// rows row0 .. row0+nb-1 of X (column-major, ld ldx): Wx (nb x s row-major scratch), V (nb x K row-major) = extension
void nys_extend_rows_run(Ctx* c, const double* X, int64_t ldx, int64_t row0, int64_t nb, int d, const double* U, int s,
                         int64_t ldu, const double* un, const double* rs, double denom, const double* Bt, int K,
                         double* Wx, double* V);
void nys_rowdot_run(Ctx* c, const double* T, const double* V, int64_t n, int K, double add, double* out);

// ---- tail.cu ---------------------------------------------------------------------------------
// m > K branch of the GPR tail: from Gg = [G1 (K_ld x K_ld col-major) | g1 (K_ld)] to coef (K_ld) and M (K_ld x K_ld),
// both zero padded beyond K; *flag = 1 if Q is not positive definite.  All pointers on the device.
void tail_woodbury_run(Ctx* c, const double* Gg, int K_ld, int K, const double* ls, const double* lam, double ns,
                       double* coef, double* M, int* flag);

// ---- eigh.cu ---------------------------------------------------------------------------------
// Top-K eigenpairs (descending) of the symmetric s x s matrix G (full storage; destroyed).
// lam: K.  Y: s x K column-major, orthonormal columns.
// psd: the caller knows G is positive semi-definite (a Gram): lower spectral bound 0 for the iterative route.
void eigh_topk_run(Ctx* c, double* G, int s, int K, double* lam, double* Y, bool psd = false);
// the direct route: Householder tridiagonalisation, multisection, inverse iteration, blocked back-transformation
void eigh_direct_run(Ctx* c, double* G, int s, int K, double* lam, double* Y);
// chfsi.cu: Chebyshev-filtered subspace iteration (K << s, s even); G is only read.  false = not applicable or not
// converged (outputs undefined).
bool chfsi_topk_run(Ctx* c, const double* G, int s, int K, double* lam, double* Y, bool psd);
// Cholesky factor and its inverse of an SPD matrix S (nb x nb column-major, nb <= 512) on one thread-block cluster:
// Linv (nb x nb ROW-major, lower) = L^-1, S = L L^T.  false: a pivot was not positive.  *ratio = min / max diag(L).
bool chol_inv_run(Ctx* c, const double* S, int nb, double* Linv, double* ratio);

// ---- misc ------------------------------------------------------------------------------------
double dfma_peak_run(Ctx* c, int iters);  // measured fp64 FMA TFLOP/s (roofline denominator)

}  // namespace flgp
