/* flgp.h — C ABI of libflgp_b200.so: the B200-native spectral core of FLGP.
 *
 * Drop-in boundary for the path BASELINE.json:north_star names.  Each entry point replaces one seam
 * of the reference (citations relative to the reference repository junhuihe2000/FLGP):
 * an Rcpp export registered in src/RcppExports.cpp:471-504, or an internal C++ seam the fit drivers
 * call (src/Spectrum.h:53-114, src/Utils.h:39-41).  INTEGRATION.md shows the Rcpp shim that binds them.
 *
 * Conventions (same as the R/Eigen side):
 *   - HOST pointers in and out unless the name ends in _dev; the caller owns every buffer.
 *   - matrices are COLUMN-MAJOR fp64 (R's layout); indices are int32 and 0-BASED
 *     (the reference returns 0-based indices to R as well, src/Utils.cpp:144).
 *   - sparse matrices are CSR with exactly r entries per row, rows sorted by column
 *     (dgRMatrix slots j, x; p[i] = i*r is implicit), explicit zeros kept (src/lae.cpp:61-67).
 *   - every function returns 0 on success; otherwise flgp_last_error() holds the message that the
 *     R shim passes to Rcpp::stop.  2 = invalid argument, 3 = CUDA error, 4 = NCCL error.
 *   - there is NO CPU fallback: without a CUDA device flgp_ctx_create fails.  Entries that take neither a context nor a
 *     spectrum handle (the *_rows training entries, the small exported helpers, the optimisers) are m- or K-sized dense
 *     host algebra — the part the reference's drivers also run on the host (Eigen LLT, NLopt) — and need no device.
 *   - k-means: the reference calls R's stats::kmeans (Hartigan-Wong, R RNG).  This library runs
 *     Lloyd's algorithm from explicit initial row indices `init_idx` (s distinct rows of X), or, when
 *     init_idx is NULL, from flgp_default_init(n, s, seed); at most `iter_max` (reference: 100)
 *     iterations; stops when no assignment changes.  See DESIGN.md §2.
 */
#ifndef FLGP_H
#define FLGP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flgp_ctx flgp_ctx;            /* device, stream, communicator; one per process/GPU */
typedef struct flgp_spectrum flgp_spectrum;  /* device-resident EigenPair (src/Spectrum.h:117-124) */

/* graph-Laplacian types of graphLaplacian_cpp (src/Utils.cpp:198-208) */
enum { FLGP_GL_RW = 0, FLGP_GL_NORMALIZED = 1, FLGP_GL_CLUSTER_NORMALIZED = 2 };

/* ---- context ------------------------------------------------------------------------------- */
int flgp_version(void);
const char* flgp_last_error(void);
int flgp_ctx_create(int device, flgp_ctx** out);
void flgp_ctx_destroy(flgp_ctx* ctx);
/* run on the caller's CUDA stream (cudaStream_t as void*); NULL restores the library's own stream */
int flgp_ctx_set_stream(flgp_ctx* ctx, void* cuda_stream);
int flgp_ctx_synchronize(flgp_ctx* ctx);
/* kernels launched by this context so far */
uint64_t flgp_ctx_launch_count(const flgp_ctx* ctx);
/* per-stage CUDA-event timing on the library's stream */
int flgp_ctx_set_timing(flgp_ctx* ctx, int on);
int flgp_ctx_stage_reset(flgp_ctx* ctx);
int flgp_ctx_stage_count(flgp_ctx* ctx);
int flgp_ctx_stage_get(flgp_ctx* ctx, int i, char* name, int name_len, double* ms, uint64_t* launches,
                       double* flops, double* bytes);
/* measured fp64 FMA throughput of this GPU in TFLOP/s (the FP64 roofline denominator) */
int flgp_dfma_peak(flgp_ctx* ctx, int iters, double* tflops);
/* Host -> device -> host through the copy path every entry point uses for its arguments (self-test).  Every `const
 * double*` / `int32_t*` argument of this header may be ordinary pageable memory (R vectors, numpy arrays): copies of
 * 8 MB and more from / to such memory are staged through pinned bounce buffers by several host threads (csrc/hostcopy.cu,
 * ~37 GB/s instead of the driver's ~13 GB/s); pinned or registered memory is copied directly. */
int flgp_copy_roundtrip(flgp_ctx* ctx, const void* in, void* out, size_t bytes);

/* ---- multi-GPU: one process per GPU, rows of X_all sharded in contiguous blocks -------------- */
/* rank 0 calls flgp_comm_unique_id and ships the 128 bytes to the other ranks (torch.distributed);
 * every rank then calls flgp_ctx_comm_init.  Collectives used: int64 all-reduce of fixed-point limbs
 * (k-means sums, column sums, Gram) and fp64 all-reduce of K x K blocks. */
int flgp_comm_unique_id(void* out128);
int flgp_ctx_comm_init(flgp_ctx* ctx, const void* id128, int rank, int nranks);

/* ---- stage entry points (one per Rcpp export / C++ seam) ------------------------------------ */
/* deterministic default initial rows for k-means: s distinct sorted indices in [0, n) */
int flgp_default_init(int64_t n, int s, uint64_t seed, int32_t* init_idx);

/* subsample_cpp (src/Utils.cpp:32-68).  method "kmeans": U is s x (d+1) = [centres, cluster sizes];
 * "random": U is s x d = X[init_idx,] (no size column, as the reference).  assign (n, optional) and
 * iters (optional) report the final assignment and the Lloyd iterations run.
 * "minibatchkmeans" (src/Utils.cpp:49-62): U is s x (d+1) = [centroids, rows per nearest centroid].  The reference's
 * centroids come from ClusterR::MiniBatchKmeans (un-vendored, R-RNG start: parity unpinned); here Sculley's mini-batch
 * k-means with ClusterR's defaults (early_stop_iter 10, tol 1e-4) from the start rows init_idx, at most iter_max
 * batches of min(10 s, n) distinct rows drawn by a bijection keyed by (seed, batch); the sizes column is the
 * reference's own 1-NN count (:57-62).  iters <- batches run; assign is not written; nstart must be 1; single GPU.
 * nstart > 1 (stats::kmeans's restarts): nstart Lloyd runs, start 0 from init_idx (or the seed's default rows),
 * start q from the default rows of a seed derived from (seed, q); the run with the smallest total within-cluster
 * sum of squares (exact fixed-point sum: the same choice on every rank) is returned.
 * Limits: r <= 16 for the LAE kernel ("lae"), r <= 32 for the SE kernel and KNN; sparse matrices handed in by the
 * caller (flgp_graph_laplacian, flgp_spectrum_from_z) need 0 <= column < s and strictly ascending columns in every
 * row (what dgRMatrix stores) -- anything else is rejected with status 2. */
int flgp_subsample(flgp_ctx* ctx, const double* X, int64_t n, int d, int s, const char* method, int iter_max,
                   int nstart, const int32_t* init_idx, uint64_t seed, double* U, int32_t* assign, int* iters);

/* KNN_cpp (src/Utils.cpp:102-192), distance "Euclidean".  ind: n x r, ascending distance, ties as
 * libstdc++ std::partial_sort.  dist (optional): n x r squared distances in the same order.
 * Zj/Zx (optional, both or neither): distances_sp as CSR. */
int flgp_knn(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int r, int32_t* ind,
             double* dist, int32_t* Zj, double* Zx);

/* v_to_z_cpp (src/lae.cpp:137-153) */
int flgp_simplex_project(flgp_ctx* ctx, const double* v, int r, double* z);
/* local_anchor_embedding_cpp (src/lae.cpp:76-133); Ur is r x d */
int flgp_lae_point(flgp_ctx* ctx, const double* x, int d, const double* Ur, int r, double* z);
/* LAE_cpp (src/lae.cpp:48-70); U is s x d.  stats (optional, 2): solver iterations, back-tracks */
int flgp_lae(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int r, int32_t* Zj,
             double* Zx, int64_t* stats);
/* graphLaplacian_cpp (src/Utils.cpp:195-212), in place on Zx; num_class (s) only for cluster-normalized */
int flgp_graph_laplacian(flgp_ctx* ctx, int64_t n, int s, int r, const int32_t* Zj, double* Zx, int gl,
                         const double* num_class);
/* cross_similarity_lae_cpp / cross_similarity_se_cpp (src/Spectrum.cpp:101-142); U is s x ucols,
 * ucols = d or d+1 (cluster sizes in the last column; required for cluster-normalized) */
int flgp_cross_similarity_lae(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int ucols,
                              int r, int gl, int32_t* Zj, double* Zx);
int flgp_cross_similarity_se(flgp_ctx* ctx, const double* X, int64_t n, int d, const double* U, int s, int ucols,
                             int r, int gl, double epsilon, int32_t* Zj, double* Zx);
/* spectrum_from_Z_cpp + truncated_SVD_cpp (src/Spectrum.cpp:146-161, src/TruncatedSVD.cpp:9-34).
 * K < 0 means K = s.  values (K): sigma (root) or sigma^2, descending.  vectors (optional): n x K
 * = sqrt(n) U.  handle (optional) keeps the result on the device. */
int flgp_spectrum_from_z(flgp_ctx* ctx, int64_t n, int s, int r, const int32_t* Zj, const double* Zx, int K,
                         int root, double* values, double* vectors, flgp_spectrum** handle);

/* Dense symmetric top-K eigensolver: the RSpectra::eigs_sym(A, k = K) call sites of the Nystrom / GLGP drivers
 * (src/Fit.cpp:262-278, 410-428) and the engine behind flgp_spectrum_from_z.  A is s x s symmetric (full storage,
 * column- or row-major alike); values (K) are the K algebraically largest eigenvalues, descending; vectors is
 * s x K column-major with orthonormal columns (signs arbitrary).  No host LAPACK, no cuSOLVER. */
int flgp_eigs_sym(flgp_ctx* ctx, const double* A, int s, int K, double* values, double* vectors);

/* heat_kernel_spectrum_cpp (src/Spectrum.cpp:48-76): X (m x d) and X_new (m_new x d, may be NULL with
 * m_new = 0) are concatenated, training rows first.  models = {subsample, kernel, gl, root}. */
int flgp_heat_kernel_spectrum(flgp_ctx* ctx, const double* X, int64_t m, const double* X_new, int64_t m_new, int d,
                              int s, int r, int K, const char* subsample, const char* kernel, int gl, int root,
                              int nstart, double epsilon, int iter_max, const int32_t* init_idx, uint64_t seed,
                              flgp_spectrum** out);
/* the same on one shard of X_all (rows [row_offset, row_offset + n_local) of n_total), for multi-GPU runs;
 * X_local is n_local x d with leading dimension n_local. */
int flgp_heat_kernel_spectrum_sharded(flgp_ctx* ctx, const double* X_local, int64_t n_local, int64_t n_total,
                                      int64_t row_offset, int d, int s, int r, int K, const char* subsample,
                                      const char* kernel, int gl, int root, int nstart, double epsilon,
                                      int iter_max, const int32_t* init_idx, uint64_t seed, flgp_spectrum** out);
/* the same with X_local already in device memory (column-major, ld = n_local) */
int flgp_heat_kernel_spectrum_dev(flgp_ctx* ctx, const double* X_local_dev, int64_t n_local, int64_t n_total,
                                  int64_t row_offset, int d, int s, int r, int K, const char* subsample,
                                  const char* kernel, int gl, int root, int nstart, double epsilon, int iter_max,
                                  const int32_t* init_idx, uint64_t seed, flgp_spectrum** out);

/* ---- EigenPair handle ------------------------------------------------------------------------ */
void flgp_spectrum_free(flgp_spectrum* h);
/* info[0..7] = n_local, n_total, row_offset, d, s, r, K, kmeans iterations; info[8..9] = LAE iterations, back-tracks */
int flgp_spectrum_info(const flgp_spectrum* h, int64_t* info10);
int flgp_spectrum_values(const flgp_spectrum* h, double* values);  /* K */
int flgp_spectrum_anchors(const flgp_spectrum* h, double* U);      /* s x (d+1), or s x d for "random" */
int flgp_spectrum_z(const flgp_spectrum* h, int32_t* Zj, double* Zx); /* local rows */
int flgp_spectrum_vectors(flgp_spectrum* h, double* vectors);      /* n_local x K */
/* mat_indexing(eigenvectors, idx, 0..K-1) (src/Utils.h:130-137); idx are LOCAL row numbers */
int flgp_spectrum_gather_rows(flgp_spectrum* h, const int32_t* idx, int64_t n_idx, double* V);
/* HK_from_spectrum_cpp (src/Spectrum.cpp:83-94): H (n0 x n1) = V[idx0,:K] diag(exp(-t(1-values))) V[idx1,:K]^T */
int flgp_hk_from_spectrum(flgp_spectrum* h, int K, double t, const int32_t* idx0, int64_t n0, const int32_t* idx1,
                          int64_t n1, double* H);

/* lae_eigenmap (src/Spectrum.cpp:17-25): eigenvalues (ndim) = 1 - sigma, eigenvectors n x ndim */
int flgp_lae_eigenmap(flgp_ctx* ctx, const double* X, int64_t n, int d, int s, int r, int ndim, const char* subsample,
                      int gl, int nstart, int iter_max, const int32_t* init_idx, uint64_t seed, double* eigenvalues,
                      double* eigenvectors);
/* heat_kernel_covariance_cpp (src/Spectrum.cpp:28-43): H is (m + m_new) x m */
int flgp_heat_kernel_covariance(flgp_ctx* ctx, const double* X, int64_t m, const double* X_new, int64_t m_new, int d,
                                int s, int r, double t, int K, const char* subsample, const char* kernel, int gl,
                                int root, int nstart, double epsilon, int iter_max, const int32_t* init_idx,
                                uint64_t seed, double* H);

/* ---- GPR tail at FIXED hyper-parameters --------------------------------- */
/* predict_regression_cpp, noise = "same" (src/Predict.cpp:40-75) + posterior_covariance_regression
 * (src/Utils.cpp:215-249) as called by fit_lae_regression_gp_cpp (src/Fit.cpp:64-77), with
 * pars = (t, noise) supplied instead of optimised.  The first m_total rows of X_all are the training
 * rows; Y_local holds the labels of the training rows THIS rank owns.  Outputs are per local row:
 * y_pred (n_local): posterior mean at every local row (training rows first), cov (n_local): posterior
 * variance (meaningful for the test rows, as the reference evaluates it there). */
int flgp_regression_fixed(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double t, double noise,
                          double sigma, double* y_pred, double* cov);
int flgp_regression_fixed_dev(flgp_spectrum* h, const double* Y_local_dev, int64_t m_total, int K, double t,
                              double noise, double sigma, double* y_pred_dev, double* cov_dev);
/* fit_lae_regression_gp_cpp (src/Fit.cpp:20-99) with fixed pars, single process: Y_train (m),
 * outputs train (m), test (m_new), cov (m_new). */
int flgp_fit_lae_regression_fixed(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                  int64_t m_new, int d, int s, int r, int K, double sigma, double t, double noise,
                                  const char* subsample, const char* kernel, int gl, int root, int nstart,
                                  int iter_max, const int32_t* init_idx, uint64_t seed, double* train, double* test,
                                  double* cov);

/* ---- hyper-parameter training of the GP regression ---------------- */
/* Objective of train_regression_gp_cpp with noise = "same": negative_marginal_likelihood_regression_cpp
 * (approach "marginal") or negative_log_posterior_regression_cpp ("posterior"), src/train.cpp:333-436, at
 * pars = (t, noise variance); grad (2, may be NULL) carries the reference's clipping of grad[1] to +-10.
 * Y_local: labels of the training rows this rank owns (first m_total rows of X_all are the training rows). */
int flgp_regression_objective(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double sigma,
                              const char* approach, const double* pars, double* obj, double* grad);
/* train_regression_gp_cpp, noise = "same" (src/train.cpp:557-671): NLopt's LD_MMA from x0 = (10, 1) (or pars_io when
 * not NaN), lb = (1e-3, 1e-4), ub = +inf, xtol_rel = 1e-5.  pars_io <- optimum, *obj <- -(minimum).  NLopt is an
 * un-vendored dependency of the reference: the optimiser is a restatement of the published CCSA/MMA algorithm, so
 * parity with the reference is to optimiser tolerance. */
int flgp_train_regression(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double sigma,
                          const char* approach, double* pars_io, double* obj, int* nevals);
/* The optimiser itself behind an nlopt-style callback (value = f(n, x, grad, data)); host only, no GPU needed. */
typedef double (*flgp_objective_fn)(unsigned n, const double* x, double* grad, void* data);
int flgp_mma_minimize(int n, flgp_objective_fn f, void* data, const double* lb, const double* ub, double* x,
                      double* minf, double xtol_rel, int maxeval, int* nevals);
/* fit_lae_regression_gp_cpp (src/Fit.cpp:20-99) including the training: pars_io = (NaN, NaN) trains, finite values
 * are used as given; *obj (may be NULL) <- the objective (sign as the reference prints it: larger is better). */
int flgp_fit_lae_regression(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                            int64_t m_new, int d, int s, int r, int K, double sigma, const char* approach,
                            const char* subsample, const char* kernel, int gl, int root, int nstart, int iter_max,
                            const int32_t* init_idx, uint64_t seed, double* pars_io, double* train, double* test,
                            double* cov, double* obj);
/* fit_se_regression_gp_cpp (src/Fit.cpp:102-219): one k-means + KNN, then per a2 of the grid
 * Z = exp(-dist / (a2 * mean dist)) -> graph Laplacian -> spectrum -> training; the a2 with the largest objective
 * wins.  fixed_pars (may be NULL): evaluate the objective at these (t, noise) instead of training.
 * out (may be NULL): the winning spectrum handle (caller frees). */
int flgp_fit_se_regression(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                           int64_t m_new, int d, int s, int r, int K, double sigma, const double* a2s, int n_a2,
                           const char* approach, const char* subsample, int gl, int root, int nstart, int iter_max,
                           const int32_t* init_idx, uint64_t seed, const double* fixed_pars, double* train,
                           double* test, double* cov, double* pars_out, double* best_a2, double* best_obj,
                           flgp_spectrum** out);
/* fit_nystrom_regression_gp_cpp (src/Fit.cpp:222-357; R wrapper R/Fit.R:177-195), SURVEY.md §8f row 4: anchors by
 * subsample_cpp, dense SE kernel on the anchors (bandwidth a2 * mean anchor distance), doubly normalised; its top-K
 * eigenpairs (the RSpectra::eigs_sym callback) are extended to every row by the Nystrom formula; training and GPR
 * tail as in the other drivers.  Single process (the sharded form follows).  fixed_pars (may be NULL): skip the training. */
int flgp_fit_nystrom_regression(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                int64_t m_new, int d, int s, int K, double sigma, const double* a2s, int n_a2,
                                const char* approach, const char* subsample, int nstart, int iter_max,
                                const int32_t* init_idx, uint64_t seed, const double* fixed_pars, double* train,
                                double* test, double* cov, double* pars_out, double* best_a2, double* best_obj);
/* The same on one contiguous block of rows of X_all per rank (rows row_offset .. row_offset + n_local - 1; the first
 * m_total rows of X_all are the training rows; Y_local = labels of the training rows this rank owns): anchors and the
 * s x s anchor operator replicated, the extension row-parallel, the K x K statistics all-reduced.  mean_local /
 * cov_local: posterior mean and variance of the local rows (BASELINE config 5's "Nystrom variant" over 8 GPUs). */
int flgp_fit_nystrom_regression_sharded(flgp_ctx* ctx, const double* X_local, int64_t n_local, int64_t n_total,
                                        int64_t row_offset, int d, const double* Y_local, int64_t m_total, int s, int K,
                                        double sigma, const double* a2s, int n_a2, const char* approach,
                                        const char* subsample, int nstart, int iter_max, const int32_t* init_idx,
                                        uint64_t seed, const double* fixed_pars, double* mean_local, double* cov_local,
                                        double* pars_out, double* best_a2, double* best_obj);
/* posterior_distribution_classification (src/Utils.cpp:252-299; exported, src/Utils.h:77-80) as the binary logit
 * drivers call it (fit_lae_logit_gp_cpp, src/Fit.cpp:563-582) at a FIXED diffusion time t: Laplace approximation of
 * the GP classifier — Newton iterations on the m labelled rows (labels 0/1, f = 0 start, |df|_1 < tol), then the
 * predictive mean C21 (Y - pi) and variance C22 - rowsum((C21 beta) o C21) of EVERY local row, folded through the
 * factored eigenvectors (C21 = V2 Lam V1^T is never formed).  Cvv carries + sigma on its diagonal and C22 + sigma, as
 * in the driver.  The labels' Polya-Gamma sampler (R RNG) is not part of this path (training of t: flgp_train_logit).
 * The Newton mode is m x m dense algebra on the host (as in the reference: Eigen LLT): O(m^3) per Newton step, meant
 * for the m of the reference's use (hundreds to a few thousand labelled rows; m <= 8192 is accepted). */
int flgp_classification_posterior_fixed(flgp_spectrum* h, const double* Y_local, int64_t m_total, int K, double t,
                                        double sigma, double tol, int max_iter, double* mean, double* cov);
/* The export itself, posterior_distribution_classification(C11, C21, C22, Y, tol, max_iter) (src/Utils.h:77-80) on
 * explicit covariance blocks: C11 m x m, C21 m_new x m (column-major), C22 m_new, labels 0/1.  Newton mode on the
 * host (m x m), the m_new-sized products on the device. */
int flgp_posterior_distribution_classification(flgp_ctx* ctx, const double* C11, const double* C21, const double* C22,
                                               const double* Y, int m, int64_t m_new, double tol, int max_iter,
                                               double* mean, double* cov);

/* ---- training of the binary GP classifier (fit_lae_logit_gp_cpp, src/Fit.cpp:521-600) -------- */
/* negative_marginal_likelihood_logit_cpp ("marginal") / negative_log_posterior_logit_cpp ("posterior"),
 * src/train.cpp:14-36: minus the Laplace-approximate marginal log-likelihood of the m labelled rows at diffusion time t
 * (marginal_log_likelihood_logit_la_cpp, src/train.cpp:716-760; N = trials per row, NULL = all 1), plus the prior
 * p log(t + 1e-9) + (t / tau)^(-q) with the defaults of PostOFData (p = 1e-2, q = 10, tau = 2).  Single process or
 * every rank with the same Y / N (the labelled rows are gathered). */
int flgp_logit_objective(flgp_spectrum* h, const double* Y, const double* N, int64_t m_total, int K, double sigma,
                         const char* approach, double t, double* obj);
/* train_lae_logit_gp_cpp (src/train.cpp:38-71): NLopt's LN_COBYLA on t from t0 = 10 (or *t_io when finite and >= 0),
 * lb = 1e-3, ub = +inf, xtol_rel = 1e-4.  *t_io <- optimum, *obj <- -(minimum).  NLopt is an un-vendored dependency:
 * the optimiser is Powell's COBYLA iteration restated for one variable (flgp_cobyla_minimize_1d), so parity with the
 * reference is to optimiser tolerance. */
int flgp_train_logit(flgp_spectrum* h, const double* Y, const double* N, int64_t m_total, int K, double sigma,
                     const char* approach, double* t_io, double* obj, int* nevals);
/* train_logit_mult_gp_cpp (src/MultiClassification.cpp:30-53), the training half of fit_lae_logit_mult_gp_cpp
 * (src/Fit.cpp:603-662; BASELINE config 3): labels Y in {0 .. J-1}, J = max(Y) + 1 one-vs-rest binary trainings with
 * N = 1, each from t0 = 10.  *J_out <- J; t_out / obj_out (J_cap entries; obj_out may be NULL) <- (t_j, -minimum_j).
 * The trainings are independent host loops and run on as many threads.  predict_logit_mult_gp_cpp draws its labels
 * with the Polya-Gamma sampler (R RNG): not part of this path; per class, flgp_classification_posterior_fixed at t_j
 * gives the deterministic Laplace posterior. */
int flgp_train_logit_mult(flgp_spectrum* h, const double* Y, int64_t m_total, int K, double sigma,
                          const char* approach, int J_cap, int* J_out, double* t_out, double* obj_out);
/* The optimiser itself behind an nlopt-style callback (n = 1); host only, no GPU needed. */
int flgp_cobyla_minimize_1d(flgp_objective_fn f, void* data, double lb, double ub, double* x, double* minf,
                            double xtol_rel, int maxeval, int* nevals);
/* fit_lae_logit_gp_cpp (src/Fit.cpp:521-600) without the Polya-Gamma label sampler (R RNG): spectrum, training of t
 * (*t_io = NaN trains; a finite value is used as given), then posterior_distribution_classification on the m_new test
 * rows: post_mean / post_cov (m_new each, may be NULL).  C_out (may be NULL): the n x m matrix [Cvv + sigma I; Cnv] of
 * output_cov = TRUE.  *obj (may be NULL) <- the objective as the reference prints it (larger is better). */
int flgp_fit_lae_logit(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m, int64_t m_new,
                       int d, int s, int r, int K, const double* N, double sigma, const char* approach,
                       const char* subsample, const char* kernel, int gl, int root, int nstart, int iter_max,
                       const int32_t* init_idx, uint64_t seed, double* t_io, double* post_mean, double* post_cov,
                       double* C_out, double* obj);
/* fit_se_logit_gp_cpp (src/Fit.cpp:668-794; R wrapper fit_se_logit_gp_rcpp, the README's GPC example) without the
 * Polya-Gamma label sampler: one k-means + KNN, then per a2 of the grid Z = exp(-dist / (a2 * mean dist)) -> graph
 * Laplacian -> spectrum -> COBYLA training of t (src/Fit.cpp:712-743); the a2 with the largest objective wins.
 * *t_io = NaN trains; a finite value is used at every grid point as given (the objective is then evaluated there).
 * post_mean / post_cov / C_out as flgp_fit_lae_logit; *best_a2, *best_obj, out (the winning handle; caller frees) may
 * be NULL.  Single process. */
int flgp_fit_se_logit(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m, int64_t m_new,
                      int d, int s, int r, int K, const double* N, double sigma, const double* a2s, int n_a2,
                      const char* approach, const char* subsample, int gl, int root, int nstart, int iter_max,
                      const int32_t* init_idx, uint64_t seed, double* t_io, double* post_mean, double* post_cov,
                      double* C_out, double* best_a2, double* best_obj, flgp_spectrum** out);
/* fit_se_logit_mult_gp_cpp (src/Fit.cpp:797-895) without the label sampler: the same grid with the J one-vs-rest
 * trainings of flgp_train_logit_mult per a2; a grid point's objective is the sum of its class objectives (:855-859).
 * t_out / obj_out (J_cap entries; obj_out may be NULL) <- (t_j, objective_j) of the winning a2. */
int flgp_fit_se_logit_mult(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                           int64_t m_new, int d, int s, int r, int K, double sigma, const double* a2s, int n_a2,
                           const char* approach, const char* subsample, int gl, int root, int nstart, int iter_max,
                           const int32_t* init_idx, uint64_t seed, int J_cap, int* J_out, double* t_out,
                           double* obj_out, double* best_a2, double* best_obj, flgp_spectrum** out);

/* fit_nystrom_logit_gp_cpp (src/Fit.cpp:896-1038) without the label sampler: the Nystrom grid of
 * flgp_fit_nystrom_regression (anchors, dense SE anchor kernel, eigs_sym, extension of the labelled rows) with the COBYLA
 * training of t per bandwidth; then the winning extension of every row, posterior_distribution_classification on the
 * test rows (post_mean / post_cov, m_new each, may be NULL) and, if C_out != NULL, the n x m block [Cvv + sigma I; Cnv].
 * *t_io = NaN trains, a finite value is used at every bandwidth as given.  Single process. */
int flgp_fit_nystrom_logit(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                           int64_t m_new, int d, int s, int K, const double* N, double sigma, const double* a2s,
                           int n_a2, const char* approach, const char* subsample, int nstart, int iter_max,
                           const int32_t* init_idx, uint64_t seed, double* t_io, double* post_mean, double* post_cov,
                           double* C_out, double* best_a2, double* best_obj);
/* The training half of fit_nystrom_logit_mult_gp_cpp (src/Fit.cpp:1045-1162): per bandwidth the J one-vs-rest
 * trainings on the extended labelled rows, summed objective selects.  t_out / obj_out (J_cap entries) of the winner;
 * values_out (K) and vectors_out (n x K column-major), both optional: the winning extended eigenpair, which is what
 * predict_logit_mult_gp_cpp (Polya-Gamma sampler, stays in R) consumes. */
int flgp_fit_nystrom_logit_mult(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                int64_t m_new, int d, int s, int K, double sigma, const double* a2s, int n_a2,
                                const char* approach, const char* subsample, int nstart, int iter_max,
                                const int32_t* init_idx, uint64_t seed, int J_cap, int* J_out, double* t_out,
                                double* obj_out, double* values_out, double* vectors_out, double* best_a2,
                                double* best_obj);

/* ---- the reference's small exported helpers: m-sized dense host algebra, no device work (usable without a context) -- */
/* marginal_log_likelihood_logit_la_cpp(C, Y, N, tol, max_iter) (src/train.cpp:716-760; export src/RcppExports.cpp:457-469):
 * Laplace-approximate marginal log-likelihood of the m labelled rows for the covariance C (m x m, column-major);
 * N = trials per row (NULL = all 1); tol <= 0 / max_iter <= 0 take the reference's defaults 1e-5 / 100. */
int flgp_marginal_log_likelihood_logit_la(const double* C, const double* Y, const double* N, int m, double tol,
                                          int max_iter, double* out);
/* flgp_logit_objective / flgp_train_logit on explicit labelled rows V1 (m x K ROW-major) of the eigenvectors and
 * `values` (K): the same host code the handle-based entries run after fetching those rows (src/train.cpp:14-71, 716-760). */
int flgp_logit_objective_rows(const double* V1, const double* values, const double* Y, const double* N, int m, int K,
                              double sigma, const char* approach, double t, double* obj);
int flgp_train_logit_rows(const double* V1, const double* values, const double* Y, const double* N, int m, int K,
                          double sigma, const char* approach, double* t_io, double* obj, int* nevals);
/* posterior_distribution_classification (src/Utils.cpp:252-299) as the logit drivers call it (src/Fit.cpp:563-582),
 * folded onto the eigenvector rows: from the m labelled rows V1 (m x K ROW-major) and `values` (K), at diffusion time t,
 * the Newton mode of the Laplace approximation and the two operators
 *   coef (K) = Lam V1^T (Y - pi),   Mq (K x K column-major) = Lam - Lam V1^T beta V1 Lam,
 * so that any row v of the eigenvectors has posterior mean v . coef and variance v Mq v^T + sigma.  This is the m-sized
 * half of flgp_classification_posterior_fixed / flgp_fit_*_logit (their n-sized half runs on the device). */
int flgp_classification_fold_rows(const double* V1, const double* values, const double* Y, int m, int K, double t,
                                  double sigma, double tol, int max_iter, double* coef, double* Mq);
/* multi_train_split (src/MultiClassification.cpp:14-27): J = max(Y) + 1, aug_y (m x J column-major, may be NULL to
 * query J) = one-vs-rest indicator columns. */
int flgp_multi_train_split(const double* Y, int64_t m, int J_cap, int* J_out, double* aug_y);
/* negative_log_likelihood(mean, cov, target, type) (src/Utils.cpp:302-318), type "regression": the mean over the rows of
 * ((target - mean)^2 / cov + log(cov + 1e-9)), plus log(2 * 3.1415926), halved.  "binary" / "multinomial" average over
 * rnorm draws from R's RNG (nll_classification, :321-336): not deterministic, rejected with status 2. */
int flgp_negative_log_likelihood(const double* mean, const double* cov, const double* target, int64_t n,
                                 const char* type, double* out);
/* test_regression_cpp(C, Y, Cnv) (src/Predict.cpp:29-37): Y_pred = Cnv C^{-1} Y by Cholesky; C m x m, Cnv m_new x m. */
int flgp_test_regression(const double* C, const double* Y, const double* Cnv, int m, int64_t m_new, double* Y_pred);

/* flgp_regression_objective / flgp_train_regression (noise = "same") on explicit training rows V1 (m x K ROW-major) of
 * the eigenvectors and `values` (K): host only; the objective and optimiser code of the handle-based entries, with the
 * K x K statistics formed by plain host loops instead of the device reduction (src/train.cpp:333-436, 557-671). */
int flgp_regression_objective_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                                   const char* approach, const double* pars, double* obj, double* grad);
int flgp_train_regression_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                               const char* approach, double* pars_io, double* obj, int* nevals);

/* ---- noise = "different": one noise variance per training row (src/train.cpp:438-556, src/Predict.cpp:76-113) -------
 * x = (t, noise_1 .. noise_m).  The *_rows entries are host algebra on explicit training rows V1 (m x K ROW-major) of the
 * eigenvectors and `values` (K, as exported: the Laplacian spectrum is 1 - values); no device work, no context.
 * negative_marginal_likelihood_diff_noise_regression_cpp ("marginal") / negative_log_posterior_diff_noise_regression_cpp
 * ("posterior"): *obj <- the value to be minimised, grad (m + 1, may be NULL) its gradient, the noise components clipped
 * to [-1, 1] in the m > K branch as in the reference. */
int flgp_regression_objective_diff_rows(const double* V1, const double* values, const double* Y, int m, int K,
                                        double sigma, const char* approach, const double* x, double* obj,
                                        double* grad);
/* train_regression_gp_cpp, noise = "different" (src/train.cpp:588-611, 614-671): MMA on m + 1 variables from
 * (10, 1 .. 1), lb = (1e-3, 1e-4 ..), ub = +inf, xtol_rel = 1e-5.  x_io: NaN entries take the start values.  The
 * reference sizes these vectors with an m read through a pointer of the wrong struct type (undefined behaviour, SURVEY.md
 * appendix A.10); m here is the number of training rows.  *obj <- -(minimum). */
int flgp_train_regression_diff_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                                    const char* approach, double* x_io, double* obj, int* nevals);
/* predict_regression_cpp, noisepar = "different": coef (K) = Lam V1^T alpha, so that Y_pred = V_new coef. */
int flgp_predict_coef_diff_rows(const double* V1, const double* values, const double* Y, int m, int K, double sigma,
                                const double* x, double* coef);
/* fit_lae_regression_gp_cpp with noise = "different": spectrum, training (pars_io: m + 1 values, any NaN trains), mean of
 * every row; cov = the reference's posterior_covariance_regression, which takes pars[1] (the first row's noise) as the
 * common variance (src/Utils.cpp:218-220).  Single process. */
int flgp_fit_lae_regression_diff_noise(flgp_ctx* ctx, const double* X, const double* Y, const double* X_new, int64_t m,
                                       int64_t m_new, int d, int s, int r, int K, double sigma, const char* approach,
                                       const char* subsample, const char* kernel, int gl, int root, int nstart,
                                       int iter_max, const int32_t* init_idx, uint64_t seed, double* pars_io,
                                       double* train, double* test, double* cov, double* obj);

#ifdef __cplusplus
}
#endif
#endif /* FLGP_H */
