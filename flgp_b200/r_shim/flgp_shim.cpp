// flgp_shim.cpp — Rcpp glue that a maintainer of junhuihe2000/FLGP drops into the R package's src/
// directory IN PLACE OF the bodies of the spectral-core functions, keeping every [[Rcpp::export]] signature
// (and therefore R/RcppExports.R, NAMESPACE, man/, R/Fit.R) unchanged.  Each function marshals R's
// column-major REALSXP / INTSXP buffers straight into the C ABI of include/flgp.h (no copies on the way in:
// Eigen::Map aliases R memory exactly as the reference does at src/Fit.cpp:29-32) and turns a non-zero
// status into Rcpp::stop, the reference's own error convention (src/Utils.cpp:64,123,207).
//
// NOT BUILT IN THIS REPOSITORY: the build environment has no R, Rcpp or RcppEigen.  It is
// parsed and type-checked by g++ against the reference's own headers and include/flgp.h with a type-level
// stand-in for <RcppEigen.h> (tests/r_mock/, tests/test_abi_cpu.py: every function defined here is the definition
// of a reference prototype, every C-ABI call has the declared argument types); the same entry points are
// exercised through ctypes by flgp_b200/api.py and the -m gpu tests.  Link with: PKG_LIBS += -L<dir> -lflgp_b200  (see INTEGRATION.md).
//
// Replaces (file:line in the reference):
//   subsample_cpp                 src/Utils.cpp:32-68          -> flgp_subsample
//   KNN_cpp                       src/Utils.cpp:102-192        -> flgp_knn
//   graphLaplacian_cpp            src/Utils.cpp:195-212        -> flgp_graph_laplacian
//   LAE_cpp / local_anchor_embedding_cpp / v_to_z_cpp  src/lae.cpp:48-153 -> flgp_lae / flgp_lae_point / flgp_simplex_project
//   cross_similarity_{lae,se}_cpp src/Spectrum.cpp:101-142     -> flgp_cross_similarity_{lae,se}
//   spectrum_from_Z_cpp           src/Spectrum.cpp:146-161     -> flgp_spectrum_from_z
//   heat_kernel_spectrum_cpp      src/Spectrum.cpp:48-76       -> flgp_heat_kernel_spectrum
//   HK_from_spectrum_cpp          src/Spectrum.cpp:83-94       -> flgp_hk_from_spectrum (handle) or Eigen as before
//   lae_eigenmap                  src/Spectrum.cpp:17-25       -> flgp_lae_eigenmap
//   heat_kernel_covariance_cpp    src/Spectrum.cpp:28-43       -> flgp_heat_kernel_covariance
//   fit_lae_regression_gp_cpp     src/Fit.cpp:20-99            -> flgp_fit_lae_regression (spectrum + MMA training + GPR tail)
//   fit_se_regression_gp_cpp      src/Fit.cpp:102-219          -> flgp_fit_se_regression  (one k-means/KNN, bandwidth grid)
//   train_regression_gp_cpp       src/train.cpp:557-671        -> flgp_train_regression   (noise = "same")
//   fit_nystrom_regression_gp_cpp src/Fit.cpp:222-357          -> flgp_fit_nystrom_regression
//   posterior_distribution_classification src/Utils.cpp:252-299 -> flgp_posterior_distribution_classification
// [[Rcpp::depends(RcppEigen)]]
#include <RcppEigen.h>

#include "Fit.h"      // the fit_*_gp_cpp prototypes ([[Rcpp::export]]) re-defined at the end of this file
#include "Predict.h"  // test_pgbinary_cpp, test_regression_cpp (and train.h: ReturnValue, the exported objectives)
#include "Spectrum.h"
#include "Utils.h"    // also MultiClassification.h: MultiClassifier, multi_train_split
#include "flgp.h"
#include "lae.h"

namespace {

flgp_ctx* ctx() {  // one context per R session, created on first use
  static flgp_ctx* c = nullptr;
  if (!c && flgp_ctx_create(0, &c)) Rcpp::stop(flgp_last_error());
  return c;
}
inline void ok(int rc) {
  if (rc) Rcpp::stop(flgp_last_error());
}
int gl_code(const std::string& gl) {
  if (gl == "rw") return FLGP_GL_RW;
  if (gl == "normalized") return FLGP_GL_NORMALIZED;
  if (gl == "cluster-normalized") return FLGP_GL_CLUSTER_NORMALIZED;
  Rcpp::stop("Error: the type of graph Laplacian is not supported!");
}
// k-means needs explicit start rows (the reference lets stats::kmeans draw them from R's RNG):
// draw them from R's RNG here too, so set.seed() keeps governing reproducibility.
std::vector<int32_t> r_init(int n, int s) {
  Rcpp::IntegerVector idx = Rcpp::sample(n, s);  // 1-based, without replacement
  std::vector<int32_t> out(idx.begin(), idx.end());
  for (auto& v : out) v -= 1;
  std::sort(out.begin(), out.end());
  return out;
}
// fixed-r CSR (Zj, Zx) -> Eigen row-major sparse matrix (the dgRMatrix the reference returns)
Eigen::SparseMatrix<double, Eigen::RowMajor> to_sparse(int n, int s, int r, const std::vector<int32_t>& Zj,
                                                       const std::vector<double>& Zx) {
  Eigen::SparseMatrix<double, Eigen::RowMajor> Z(n, s);
  Z.reserve(Eigen::VectorXi::Constant(n, r));
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < r; ++a) Z.insert(i, Zj[(size_t)i * r + a]) = Zx[(size_t)i * r + a];
  return Z;
}

}  // namespace

Eigen::MatrixXd subsample_cpp(const Eigen::MatrixXd& X, int s, std::string method, int nstart) {
  const int n = X.rows(), d = X.cols();
  std::vector<int32_t> init = r_init(n, s);
  Eigen::MatrixXd U(s, method == "random" ? d : d + 1);  // "kmeans" / "minibatchkmeans": centres + sizes (src/Utils.cpp:43-62)
  ok(flgp_subsample(ctx(), X.data(), n, d, s, method.c_str(), 100, nstart, init.data(), 0, U.data(), nullptr, nullptr));
  return U;
}

Rcpp::List KNN_cpp(const Eigen::MatrixXd& X, const Eigen::MatrixXd& U, int r, std::string distance, bool output,
                   int /*batch: bounded the reference's working set only*/) {
  if (distance != "Euclidean") Rcpp::stop("The distance method of KNN is not supported!\n");
  const int n = X.rows(), d = X.cols(), s = U.rows();
  Eigen::MatrixXi ind(n, r);
  if (!output) {
    ok(flgp_knn(ctx(), X.data(), n, d, U.data(), s, r, ind.data(), nullptr, nullptr, nullptr));
    return Rcpp::List::create(Rcpp::Named("ind_knn") = ind);
  }
  std::vector<int32_t> Zj((size_t)n * r);
  std::vector<double> Zx((size_t)n * r);
  ok(flgp_knn(ctx(), X.data(), n, d, U.data(), s, r, ind.data(), nullptr, Zj.data(), Zx.data()));
  return Rcpp::List::create(Rcpp::Named("ind_knn") = ind, Rcpp::Named("distances_sp") = to_sparse(n, s, r, Zj, Zx));
}

Eigen::RowVectorXd v_to_z_cpp(const Eigen::RowVectorXd& v) {
  Eigen::RowVectorXd z(v.size());
  ok(flgp_simplex_project(ctx(), v.data(), (int)v.size(), z.data()));
  return z;
}

Eigen::RowVectorXd local_anchor_embedding_cpp(const Eigen::RowVectorXd& x, const Eigen::MatrixXd& U) {
  Eigen::RowVectorXd z(U.rows());
  ok(flgp_lae_point(ctx(), x.data(), (int)U.cols(), U.data(), (int)U.rows(), z.data()));
  return z;
}

Eigen::SparseMatrix<double, Eigen::RowMajor> LAE_cpp(const Eigen::MatrixXd& X, const Eigen::MatrixXd& U, int r) {
  const int n = X.rows(), s = U.rows();
  std::vector<int32_t> Zj((size_t)n * r);
  std::vector<double> Zx((size_t)n * r);
  ok(flgp_lae(ctx(), X.data(), n, (int)X.cols(), U.data(), s, r, Zj.data(), Zx.data(), nullptr));
  return to_sparse(n, s, r, Zj, Zx);
}

void graphLaplacian_cpp(Eigen::SparseMatrix<double, Eigen::RowMajor>& Z, std::string gl, const Eigen::VectorXd& num_class) {
  Z.makeCompressed();
  const int n = Z.rows(), s = Z.cols(), r = n ? Z.nonZeros() / n : 0;
  ok(flgp_graph_laplacian(ctx(), n, s, r, Z.innerIndexPtr(), Z.valuePtr(), gl_code(gl),
                          num_class.size() ? num_class.data() : nullptr));
}

static Eigen::SparseMatrix<double, Eigen::RowMajor> cross_similarity(const Eigen::MatrixXd& X, const Eigen::MatrixXd& U, int r,
                                                                    const std::string& gl, bool se, double epsilon) {
  const int n = X.rows(), d = X.cols(), s = U.rows();
  std::vector<int32_t> Zj((size_t)n * r);
  std::vector<double> Zx((size_t)n * r);
  ok(se ? flgp_cross_similarity_se(ctx(), X.data(), n, d, U.data(), s, (int)U.cols(), r, gl_code(gl), epsilon, Zj.data(), Zx.data())
        : flgp_cross_similarity_lae(ctx(), X.data(), n, d, U.data(), s, (int)U.cols(), r, gl_code(gl), Zj.data(), Zx.data()));
  return to_sparse(n, s, r, Zj, Zx);
}
Eigen::SparseMatrix<double, Eigen::RowMajor> cross_similarity_lae_cpp(const Eigen::MatrixXd& X, const Eigen::MatrixXd& U, int r,
                                                                     Rcpp::String gl) {
  return cross_similarity(X, U, r, gl, false, 0.1);
}
Eigen::SparseMatrix<double, Eigen::RowMajor> cross_similarity_se_cpp(const Eigen::MatrixXd& X, const Eigen::MatrixXd& U, int r,
                                                                    Rcpp::String gl, double epsilon) {
  return cross_similarity(X, U, r, gl, true, epsilon);
}

// EigenPair keeps its reference layout {values, vectors}; the n x K block is materialised once so that
// train.cpp / Predict.cpp (out of scope) keep working unchanged.  Large-n callers use the handle API.
static EigenPair from_handle(flgp_spectrum* h) {
  int64_t info[10];
  ok(flgp_spectrum_info(h, info));
  Eigen::VectorXd values(info[6]);
  Eigen::MatrixXd vectors(info[0], info[6]);
  int rc = flgp_spectrum_values(h, values.data());
  if (!rc) rc = flgp_spectrum_vectors(h, vectors.data());
  flgp_spectrum_free(h);
  ok(rc);
  return EigenPair(values, vectors);
}

EigenPair spectrum_from_Z_cpp(const Eigen::SparseMatrix<double, Eigen::RowMajor>& Zin, int K, bool root) {
  Eigen::SparseMatrix<double, Eigen::RowMajor> Z = Zin;
  Z.makeCompressed();
  const int n = Z.rows(), s = Z.cols(), r = n ? Z.nonZeros() / n : 0;
  flgp_spectrum* h = nullptr;
  ok(flgp_spectrum_from_z(ctx(), n, s, r, Z.innerIndexPtr(), Z.valuePtr(), K, root, nullptr, nullptr, &h));
  return from_handle(h);
}

EigenPair heat_kernel_spectrum_cpp(const Eigen::MatrixXd& X, const Eigen::MatrixXd& X_new, int s, int r, int K,
                                   const Rcpp::List& models, int nstart, double epsilon) {
  const std::string sub = Rcpp::as<std::string>(models["subsample"]), ker = Rcpp::as<std::string>(models["kernel"]);
  std::vector<int32_t> init = r_init(X.rows() + X_new.rows(), s);
  flgp_spectrum* h = nullptr;
  ok(flgp_heat_kernel_spectrum(ctx(), X.data(), X.rows(), X_new.data(), X_new.rows(), (int)X.cols(), s, r, K, sub.c_str(),
                               ker.c_str(), gl_code(Rcpp::as<std::string>(models["gl"])), Rcpp::as<bool>(models["root"]),
                               nstart, epsilon, 100, init.data(), 0, &h));
  return from_handle(h);
}

Rcpp::List lae_eigenmap(const Eigen::MatrixXd& X, int s, int r, int ndim, std::string subsample, std::string norm, int nstart) {
  std::vector<int32_t> init = r_init(X.rows(), s);
  Eigen::VectorXd ev(ndim);
  Eigen::MatrixXd V(X.rows(), ndim);
  ok(flgp_lae_eigenmap(ctx(), X.data(), X.rows(), (int)X.cols(), s, r, ndim, subsample.c_str(), gl_code(norm), nstart, 100,
                       init.data(), 0, ev.data(), V.data()));
  return Rcpp::List::create(Rcpp::Named("eigenvalues") = ev, Rcpp::Named("eigenvectors") = V);
}

Eigen::MatrixXd heat_kernel_covariance_cpp(const Eigen::MatrixXd& X, const Eigen::MatrixXd& X_new, int s, int r, double t, int K,
                                           Rcpp::List models, int nstart, double epsilon) {
  const std::string sub = Rcpp::as<std::string>(models["subsample"]), ker = Rcpp::as<std::string>(models["kernel"]);
  std::vector<int32_t> init = r_init(X.rows() + X_new.rows(), s);
  Eigen::MatrixXd H(X.rows() + X_new.rows(), X.rows());
  ok(flgp_heat_kernel_covariance(ctx(), X.data(), X.rows(), X_new.data(), X_new.rows(), (int)X.cols(), s, r, t, K, sub.c_str(),
                                 ker.c_str(), gl_code(Rcpp::as<std::string>(models["gl"])), Rcpp::as<bool>(models["root"]),
                                 nstart, epsilon, 100, init.data(), 0, H.data()));
  return H;
}

// ---- regression fit drivers (noise = "same"; noise = "different" keeps the reference's nlopt path) -------------
namespace {
Rcpp::List pack_fit(const Eigen::VectorXd& train, const Eigen::VectorXd& test, const Eigen::VectorXd& cov,
                    const std::vector<double>& pars) {
  Rcpp::List Y_pred = Rcpp::List::create(Rcpp::Named("train") = train, Rcpp::Named("test") = test);
  Rcpp::List post = Rcpp::List::create(Rcpp::Named("mean") = test, Rcpp::Named("cov") = cov);
  return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("posterior") = post, Rcpp::Named("pars") = pars);
}
}  // namespace

Rcpp::List fit_lae_regression_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                                     int s, int r, int K, double sigma, std::string approach, std::string noise,
                                     Rcpp::List models, bool output_cov, int nstart) {
  if (noise != "same" && noise != "different") Rcpp::stop("The noise setting is illegal!");
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::Map<Eigen::VectorXd> Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const int m = X.rows(), m_new = X_new.rows();
  const std::string sub = Rcpp::as<std::string>(models["subsample"]), ker = Rcpp::as<std::string>(models["kernel"]);
  std::vector<int32_t> init = r_init(m + m_new, s);
  std::vector<double> pars = {NA_REAL, NA_REAL};  // NaN: train
  Eigen::VectorXd train(m), test(m_new), cov(m_new);
  double obj = 0.0;
  if (noise == "different") {  // (t, noise_1 .. noise_m): src/train.cpp:438-556, src/Predict.cpp:76-113
    pars.assign(m + 1, NA_REAL);
    ok(flgp_fit_lae_regression_diff_noise(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, r, K, sigma,
                                          approach.c_str(), sub.c_str(), ker.c_str(),
                                          gl_code(Rcpp::as<std::string>(models["gl"])), Rcpp::as<bool>(models["root"]),
                                          nstart, 100, init.data(), 0, pars.data(), train.data(), test.data(), cov.data(),
                                          &obj));
  } else
  ok(flgp_fit_lae_regression(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, r, K, sigma,
                             approach.c_str(), sub.c_str(), ker.c_str(), gl_code(Rcpp::as<std::string>(models["gl"])),
                             Rcpp::as<bool>(models["root"]), nstart, 100, init.data(), 0, pars.data(), train.data(),
                             test.data(), cov.data(), &obj));
  Rcpp::List res = pack_fit(train, test, cov, pars);
  if (output_cov) {
    Eigen::MatrixXd C(m + m_new, m);
    ok(flgp_heat_kernel_covariance(ctx(), X.data(), m, X_new.data(), m_new, (int)X.cols(), s, r, pars[0], K, sub.c_str(),
                                   ker.c_str(), gl_code(Rcpp::as<std::string>(models["gl"])),
                                   Rcpp::as<bool>(models["root"]), nstart, 0.1, 100, init.data(), 0, C.data()));
    res["C"] = C;
  }
  return res;
}

Rcpp::List fit_se_regression_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                                    int s, int r, int K, double sigma, std::vector<double> a2s, std::string approach,
                                    std::string noise, Rcpp::List models, bool output_cov, int nstart) {
  if (noise != "same") Rcpp::stop("noise=\"different\" is not offloaded; call the reference path");
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::Map<Eigen::VectorXd> Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const int m = X.rows(), m_new = X_new.rows(), n = m + m_new;
  const std::string sub = Rcpp::as<std::string>(models["subsample"]);
  std::vector<int32_t> init = r_init(n, s);
  std::vector<double> pars(2);
  Eigen::VectorXd train(m), test(m_new), cov(m_new);
  double a2 = 0.0, obj = 0.0;
  flgp_spectrum* h = nullptr;
  ok(flgp_fit_se_regression(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, r, K, sigma, a2s.data(),
                            (int)a2s.size(), approach.c_str(), sub.c_str(), gl_code(Rcpp::as<std::string>(models["gl"])),
                            Rcpp::as<bool>(models["root"]), nstart, 100, init.data(), 0, nullptr, train.data(),
                            test.data(), cov.data(), pars.data(), &a2, &obj, output_cov ? &h : nullptr));
  Rcpp::Rcout << "By " << approach << " method, optimal epsilon = " << std::sqrt(a2) << ", t = " << pars[0]
              << ", sigma = " << std::sqrt(pars[1]) << ", the objective function is " << obj << "\n";
  Rcpp::List res = pack_fit(train, test, cov, pars);
  if (output_cov) {
    std::vector<int32_t> i0(n), i1(m);
    for (int i = 0; i < n; ++i) i0[i] = i;
    for (int i = 0; i < m; ++i) i1[i] = i;
    Eigen::MatrixXd C(n, m);
    int rc = flgp_hk_from_spectrum(h, K < 0 ? s : K, pars[0], i0.data(), n, i1.data(), m, C.data());
    flgp_spectrum_free(h);
    ok(rc);
    res["C"] = C;
  }
  return res;
}

Rcpp::List fit_nystrom_regression_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                                         int s, int K, double sigma, std::vector<double> a2s, std::string approach,
                                         std::string noise, std::string subsample, bool output_cov, int nstart) {
  if (noise != "same" || output_cov) Rcpp::stop("noise=\"different\" / output_cov are not offloaded; call the reference path");
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::Map<Eigen::VectorXd> Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const int m = X.rows(), m_new = X_new.rows();
  std::vector<int32_t> init = r_init(m + m_new, s);
  std::vector<double> pars(2);
  Eigen::VectorXd train(m), test(m_new), cov(m_new);
  double a2 = 0.0, obj = 0.0;
  ok(flgp_fit_nystrom_regression(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, K, sigma, a2s.data(),
                                 (int)a2s.size(), approach.c_str(), subsample.c_str(), nstart, 100, init.data(), 0, nullptr,
                                 train.data(), test.data(), cov.data(), pars.data(), &a2, &obj));
  Rcpp::Rcout << "By " << approach << " method, optimal epsilon = " << std::sqrt(a2) << ", t = " << pars[0]
              << ", sigma = " << std::sqrt(pars[1]) << ", the objective function is " << obj << "\n";
  return pack_fit(train, test, cov, pars);
}

// fit_lae_logit_gp_cpp (src/Fit.cpp:521-600), signature unchanged (src/Fit.h).  Spectrum, COBYLA training of t and the
// Laplace posterior of the test rows run behind the C ABI; the labels still come from the reference's Polya-Gamma Gibbs
// sampler (test_pgbinary_cpp, src/Predict.cpp:11-26: R's RNG) on the covariance block C the library returns.
Rcpp::List fit_lae_logit_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                                int s, int r, int K, Rcpp::NumericVector N_train, double sigma, std::string approach,
                                Rcpp::List models, bool output_cov, int nstart) {
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::Map<Eigen::VectorXd> Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const Eigen::Map<Eigen::VectorXd> N(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(N_train));
  const int m = X.rows(), m_new = X_new.rows();
  const std::string sub = Rcpp::as<std::string>(models["subsample"]), ker = Rcpp::as<std::string>(models["kernel"]);
  std::vector<int32_t> init = r_init(m + m_new, s);
  double t = NA_REAL, obj = 0.0;  // NaN: train
  Eigen::VectorXd mean(m_new), cov(m_new);
  Eigen::MatrixXd C(m + m_new, m);
  ok(flgp_fit_lae_logit(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, r, K, N.data(), sigma,
                        approach.c_str(), sub.c_str(), ker.c_str(), gl_code(Rcpp::as<std::string>(models["gl"])),
                        Rcpp::as<bool>(models["root"]), nstart, 100, init.data(), 0, &t, mean.data(), cov.data(),
                        C.data(), &obj));
  Eigen::VectorXd label = Rcpp::as<Eigen::VectorXd>(test_pgbinary_cpp(C.topRows(m), Y, C)["Y_pred"]);
  Rcpp::List Y_pred = Rcpp::List::create(Rcpp::Named("train") = label.head(m), Rcpp::Named("test") = label.tail(m_new));
  Rcpp::List post = Rcpp::List::create(Rcpp::Named("mean") = mean, Rcpp::Named("cov") = cov);
  if (output_cov)
    return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("C") = C, Rcpp::Named("posterior") = post,
                              Rcpp::Named("pars") = t);
  return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("posterior") = post, Rcpp::Named("pars") = t);
}

// fit_se_logit_gp_cpp (src/Fit.cpp:668-794), signature unchanged: the bandwidth grid with the COBYLA training of t per
// grid point behind the C ABI; labels from the reference's sampler on the returned covariance block, as above.
Rcpp::List fit_se_logit_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                               int s, int r, int K, Rcpp::NumericVector N_train, double sigma, std::vector<double> a2s,
                               std::string approach, Rcpp::List models, bool output_cov, int nstart) {
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::Map<Eigen::VectorXd> Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const Eigen::Map<Eigen::VectorXd> N(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(N_train));
  const int m = X.rows(), m_new = X_new.rows();
  const std::string sub = Rcpp::as<std::string>(models["subsample"]);
  std::vector<int32_t> init = r_init(m + m_new, s);
  double t = NA_REAL, a2 = 0.0, obj = 0.0;  // NaN: train
  Eigen::VectorXd mean(m_new), cov(m_new);
  Eigen::MatrixXd C(m + m_new, m);
  ok(flgp_fit_se_logit(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, r, K, N.data(), sigma,
                       a2s.data(), (int)a2s.size(), approach.c_str(), sub.c_str(),
                       gl_code(Rcpp::as<std::string>(models["gl"])), Rcpp::as<bool>(models["root"]), nstart, 100,
                       init.data(), 0, &t, mean.data(), cov.data(), C.data(), &a2, &obj, nullptr));
  Rcpp::Rcout << "By " << approach << " method, optimal epsilon = " << std::sqrt(a2) << ", t = " << t
              << ", the objective function is " << obj << "\n";
  Eigen::VectorXd label = Rcpp::as<Eigen::VectorXd>(test_pgbinary_cpp(C.topRows(m), Y, C)["Y_pred"]);
  Rcpp::List Y_pred = Rcpp::List::create(Rcpp::Named("train") = label.head(m), Rcpp::Named("test") = label.tail(m_new));
  Rcpp::List post = Rcpp::List::create(Rcpp::Named("mean") = mean, Rcpp::Named("cov") = cov);
  if (output_cov)
    return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("C") = C, Rcpp::Named("posterior") = post,
                              Rcpp::Named("pars") = t);
  return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("posterior") = post, Rcpp::Named("pars") = t);
}

// fit_nystrom_logit_gp_cpp (src/Fit.cpp:896-1038), signature unchanged: Nystrom grid, COBYLA training of t per bandwidth
// and the Laplace posterior behind the C ABI; labels from the reference's sampler on the returned covariance block.
Rcpp::List fit_nystrom_logit_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                                    int s, int K, Rcpp::NumericVector N_train, double sigma, std::vector<double> a2s,
                                    std::string approach, std::string subsample, bool output_cov, int nstart) {
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::Map<Eigen::VectorXd> Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const Eigen::Map<Eigen::VectorXd> N(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(N_train));
  const int m = X.rows(), m_new = X_new.rows();
  std::vector<int32_t> init = r_init(m + m_new, s);
  double t = NA_REAL, a2 = 0.0, obj = 0.0;  // NaN: train
  Eigen::VectorXd mean(m_new), cov(m_new);
  Eigen::MatrixXd C(m + m_new, m);
  ok(flgp_fit_nystrom_logit(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, K, N.data(), sigma,
                            a2s.data(), (int)a2s.size(), approach.c_str(), subsample.c_str(), nstart, 100, init.data(), 0,
                            &t, mean.data(), cov.data(), C.data(), &a2, &obj));
  Rcpp::Rcout << "By " << approach << " method, optimal epsilon = " << std::sqrt(a2) << ", t = " << t
              << ", the objective function is " << obj << "\n";
  Eigen::VectorXd label = Rcpp::as<Eigen::VectorXd>(test_pgbinary_cpp(C.topRows(m), Y, C)["Y_pred"]);
  Rcpp::List Y_pred = Rcpp::List::create(Rcpp::Named("train") = label.head(m), Rcpp::Named("test") = label.tail(m_new));
  Rcpp::List post = Rcpp::List::create(Rcpp::Named("mean") = mean, Rcpp::Named("cov") = cov);
  if (output_cov)
    return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("C") = C, Rcpp::Named("posterior") = post,
                              Rcpp::Named("pars") = t);
  return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("posterior") = post, Rcpp::Named("pars") = t);
}

// train_logit_mult_gp_cpp (src/MultiClassification.cpp:30-53) for fit_lae_logit_mult_gp_cpp (src/Fit.cpp:603-662): the J
// one-vs-rest trainings run behind the C ABI on the spectrum handle; the MultiClassifier keeps the reference's layout
// (aug_y + one ReturnValue(t, obj) per class), so predict_logit_mult_gp_cpp (Polya-Gamma sampler, R RNG) is unchanged.
// (ReturnValue: src/train.h:173-181)
std::vector<ReturnValue> train_logit_mult_on_handle(flgp_spectrum* h, const Eigen::VectorXd& Y, int K, double sigma,
                                                    const std::string& approach) {
  int J = 0;
  std::vector<double> t(256), obj(256);
  ok(flgp_train_logit_mult(h, Y.data(), Y.size(), K, sigma, approach.c_str(), 256, &J, t.data(), obj.data()));
  std::vector<ReturnValue> res(J);
  for (int j = 0; j < J; ++j) res[j] = ReturnValue(t[j], obj[j]);
  return res;
}

// The grid loop of fit_se_logit_mult_gp_cpp (src/Fit.cpp:839-867) as one call: the winning handle and the per-class
// ReturnValues; the caller builds the MultiClassifier from them and goes on with predict_logit_mult_gp_cpp unchanged.
std::vector<ReturnValue> se_logit_mult_grid(const Eigen::MatrixXd& X, const Eigen::VectorXd& Y, const Eigen::MatrixXd& X_new,
                                            int s, int r, int K, double sigma, const std::vector<double>& a2s,
                                            const std::string& approach, Rcpp::List models, int nstart,
                                            flgp_spectrum** best, double* best_a2, double* max_obj) {
  const int m = X.rows(), m_new = X_new.rows();
  const std::string sub = Rcpp::as<std::string>(models["subsample"]);
  std::vector<int32_t> init = r_init(m + m_new, s);
  int J = 0;
  std::vector<double> t(256), obj(256);
  ok(flgp_fit_se_logit_mult(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, r, K, sigma, a2s.data(),
                            (int)a2s.size(), approach.c_str(), sub.c_str(), gl_code(Rcpp::as<std::string>(models["gl"])),
                            Rcpp::as<bool>(models["root"]), nstart, 100, init.data(), 0, 256, &J, t.data(), obj.data(),
                            best_a2, max_obj, best));
  std::vector<ReturnValue> res(J);
  for (int j = 0; j < J; ++j) res[j] = ReturnValue(t[j], obj[j]);
  return res;
}

// The grid loop of fit_nystrom_logit_mult_gp_cpp (src/Fit.cpp:1087-1132) as one call: per-class ReturnValues and the
// winning extended EigenPair (values K, vectors n x K); predict_logit_mult_gp_cpp goes on unchanged from there.
std::vector<ReturnValue> nystrom_logit_mult_grid(const Eigen::MatrixXd& X, const Eigen::VectorXd& Y,
                                                 const Eigen::MatrixXd& X_new, int s, int K, double sigma,
                                                 const std::vector<double>& a2s, const std::string& approach,
                                                 const std::string& subsample, int nstart, Eigen::VectorXd& values,
                                                 Eigen::MatrixXd& vectors, double* best_a2, double* max_obj) {
  const int m = X.rows(), m_new = X_new.rows();
  std::vector<int32_t> init = r_init(m + m_new, s);
  values.resize(K);
  vectors.resize(m + m_new, K);
  int J = 0;
  std::vector<double> t(256), obj(256);
  ok(flgp_fit_nystrom_logit_mult(ctx(), X.data(), Y.data(), X_new.data(), m, m_new, (int)X.cols(), s, K, sigma,
                                 a2s.data(), (int)a2s.size(), approach.c_str(), subsample.c_str(), nstart, 100,
                                 init.data(), 0, 256, &J, t.data(), obj.data(), values.data(), vectors.data(), best_a2,
                                 max_obj));
  std::vector<ReturnValue> res(J);
  for (int j = 0; j < J; ++j) res[j] = ReturnValue(t[j], obj[j]);
  return res;
}

// ---- the three multi-class drivers, signatures unchanged (src/Fit.h): spectrum / grids and the J one-vs-rest trainings
// behind the C ABI; labels and the `posterior` entry through the reference's own predict_logit_mult_gp_cpp (Polya-Gamma
// sampler on R's RNG) and posterior_distribution_multiclassification, on the EigenPair the library hands back --------------
namespace {
Rcpp::List mult_tail(const EigenPair& eigenpair, const Eigen::VectorXd& Y, const std::vector<ReturnValue>& res_vec, int m,
                     int m_new, int K, double sigma) {
  const MultiClassifier multiclassifier(multi_train_split(Y), res_vec);
  const int n = m + m_new;
  const Eigen::VectorXi idx = Eigen::VectorXi::LinSpaced(m, 0, m - 1);
  const Eigen::VectorXi idx_pred = Eigen::VectorXi::LinSpaced(n, 0, n - 1);
  const Eigen::VectorXi idx_new = Eigen::VectorXi::LinSpaced(m_new, m, n - 1);
  const Eigen::VectorXd label_pred = predict_logit_mult_gp_cpp(multiclassifier, eigenpair, idx, idx_pred, K, sigma);
  const Eigen::VectorXd train_pred = label_pred.head(m), test_pred = label_pred.tail(m_new);
  Rcpp::List Y_pred = Rcpp::List::create(Rcpp::Named("train") = train_pred, Rcpp::Named("test") = test_pred);
  Rcpp::List post = posterior_distribution_multiclassification(eigenpair, multiclassifier, idx, idx_new, K, sigma);
  return Rcpp::List::create(Rcpp::Named("Y_pred") = Y_pred, Rcpp::Named("posterior") = post);
}
}  // namespace

Rcpp::List fit_lae_logit_mult_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                                     int s, int r, int K, double sigma, std::string approach, Rcpp::List models,
                                     int nstart) {
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::VectorXd Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const int m = X.rows(), m_new = X_new.rows();
  if (K < 0) K = s;
  const std::string sub = Rcpp::as<std::string>(models["subsample"]), ker = Rcpp::as<std::string>(models["kernel"]);
  std::vector<int32_t> init = r_init(m + m_new, s);
  flgp_spectrum* h = nullptr;
  ok(flgp_heat_kernel_spectrum(ctx(), X.data(), m, X_new.data(), m_new, (int)X.cols(), s, r, K, sub.c_str(), ker.c_str(),
                               gl_code(Rcpp::as<std::string>(models["gl"])), Rcpp::as<bool>(models["root"]), nstart, 0.1,
                               100, init.data(), 0, &h));
  const std::vector<ReturnValue> res_vec = train_logit_mult_on_handle(h, Y, K, sigma, approach);
  const EigenPair eigenpair = from_handle(h);  // materialises {values, vectors} and frees the handle
  return mult_tail(eigenpair, Y, res_vec, m, m_new, K, sigma);
}

Rcpp::List fit_se_logit_mult_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train, Rcpp::NumericMatrix X_test,
                                    int s, int r, int K, double sigma, std::vector<double> a2s, std::string approach,
                                    Rcpp::List models, int nstart) {
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::VectorXd Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const int m = X.rows(), m_new = X_new.rows();
  if (K < 0) K = s;
  flgp_spectrum* h = nullptr;
  double a2 = 0.0, obj = 0.0;
  const std::vector<ReturnValue> res_vec = se_logit_mult_grid(X, Y, X_new, s, r, K, sigma, a2s, approach, models, nstart, &h,
                                                              &a2, &obj);
  Rcpp::Rcout << "By " << approach << " method, optimal epsilon = " << std::sqrt(a2) << ", the objective function is "
              << obj << "\n";
  const EigenPair eigenpair = from_handle(h);
  return mult_tail(eigenpair, Y, res_vec, m, m_new, K, sigma);
}

Rcpp::List fit_nystrom_logit_mult_gp_cpp(Rcpp::NumericMatrix X_train, Rcpp::NumericVector Y_train,
                                         Rcpp::NumericMatrix X_test, int s, int K, double sigma, std::vector<double> a2s,
                                         std::string approach, std::string subsample, int nstart) {
  const Eigen::Map<Eigen::MatrixXd> X(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_train));
  const Eigen::VectorXd Y(Rcpp::as<Eigen::Map<Eigen::VectorXd>>(Y_train));
  const Eigen::Map<Eigen::MatrixXd> X_new(Rcpp::as<Eigen::Map<Eigen::MatrixXd>>(X_test));
  const int m = X.rows(), m_new = X_new.rows();
  if (K < 0) K = s;
  Eigen::VectorXd values;
  Eigen::MatrixXd vectors;
  double a2 = 0.0, obj = 0.0;
  const std::vector<ReturnValue> res_vec = nystrom_logit_mult_grid(X, Y, X_new, s, K, sigma, a2s, approach, subsample, nstart,
                                                                   values, vectors, &a2, &obj);
  Rcpp::Rcout << "By " << approach << " method, optimal epsilon = " << std::sqrt(a2) << ", the objective function is "
              << obj << "\n";
  const EigenPair eigenpair(values, vectors);
  return mult_tail(eigenpair, Y, res_vec, m, m_new, K, sigma);
}

// The small exported helpers (marginal_log_likelihood_logit_la_cpp, multi_train_split, test_regression_cpp,
// negative_log_likelihood) are m-sized host algebra in the reference as well: their bodies stay where they are
// (src/train.cpp, src/MultiClassification.cpp, src/Predict.cpp, src/Utils.cpp).  The library offers the same functions
// to callers without R (flgp_marginal_log_likelihood_logit_la, flgp_multi_train_split, flgp_test_regression,
// flgp_negative_log_likelihood); re-defining them here would only duplicate symbols the package already links.

// [[Rcpp::export(posterior_distribution_classification)]]  -- signature unchanged (src/Utils.h:77-80)
Rcpp::List posterior_distribution_classification(const Eigen::MatrixXd& C11, const Eigen::MatrixXd& C21,
                                                 const Eigen::VectorXd& C22, const Eigen::VectorXd& Y, double tol,
                                                 int max_iter) {
  Eigen::VectorXd mean(C21.rows()), cov(C21.rows());
  ok(flgp_posterior_distribution_classification(ctx(), C11.data(), C21.data(), C22.data(), Y.data(), (int)C11.rows(),
                                                C21.rows(), tol, max_iter, mean.data(), cov.data()));
  return Rcpp::List::create(Rcpp::Named("mean") = mean, Rcpp::Named("cov") = cov);
}
