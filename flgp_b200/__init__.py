"""flgp_b200 — B200-native spectral core of FLGP behind the reference's Rcpp export names.

The package holds only what the hot path needs: csrc/ (hand-written sm_100a CUDA kernels + the C ABI of
include/flgp.h) and a ctypes mirror of the reference interface.  Importing the API loads
libflgp_b200.so and fails loudly when it is absent; there is no CPU or PyTorch fallback.
"""
from .api import (  # noqa: F401
    Context, EigenPair, FlgpError, HK_from_spectrum_cpp, KNN_cpp, LAE_cpp, cross_similarity_lae_cpp,
    cross_similarity_se_cpp, default_ctx, default_init, eigs_sym, fit_lae_regression_gp_rcpp, fit_nystrom_regression_gp_rcpp,
    fit_se_regression_gp_rcpp,
    graphLaplacian_cpp,
    heat_kernel_covariance_rcpp, heat_kernel_spectrum_cpp, heat_kernel_spectrum_sharded, knn_distances,
    lae_eigenmap, local_anchor_embedding_cpp, mma_minimize, posterior_distribution_classification,
    posterior_distribution_classification_rcpp, regression_fixed, regression_objective, spectrum_from_Z_cpp, subsample_cpp, train_regression_gp,
    v_to_z_cpp, cobyla_minimize_1d, fit_lae_logit_gp_rcpp, fit_nystrom_regression_sharded, logit_objective, train_lae_logit_gp, train_logit_mult_gp, fit_lae_logit_mult_gp_rcpp,
    fit_se_logit_gp_rcpp, fit_se_logit_mult_gp_rcpp, fit_nystrom_logit_gp_rcpp, fit_nystrom_logit_mult_gp_rcpp,
    regression_objective_rows, train_regression_rows, regression_objective_diff_rows, train_regression_diff_rows, predict_coef_diff_rows,
    marginal_log_likelihood_logit_la_cpp, classification_fold_rows, posterior_distribution_multiclassification, logit_objective_rows, train_logit_rows, multi_train_split, negative_log_likelihood, test_regression_cpp,
)

__version__ = "0.1.0"
