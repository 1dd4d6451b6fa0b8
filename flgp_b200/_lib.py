"""ctypes loader for libflgp_b200.so (the C ABI declared in include/flgp.h).

There is no fallback of any kind: if the CUDA library is missing or cannot be loaded the import
of the product API fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflgp_b200.so")

c_i64 = C.c_int64
c_u64 = C.c_uint64
p_f64 = C.POINTER(C.c_double)
p_i32 = C.POINTER(C.c_int32)
p_i64 = C.POINTER(C.c_int64)
p_void = C.c_void_p

OBJECTIVE_FN = C.CFUNCTYPE(C.c_double, C.c_uint, p_f64, p_f64, p_void)  # nlopt-style: f(n, x, grad, data)

_lib = None


class FlgpError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the R shim turns the same message into Rcpp::stop)."""


def _declare(lib):
    H = p_void  # opaque handles
    sig = {
        "flgp_version": (C.c_int, []),
        "flgp_last_error": (C.c_char_p, []),
        "flgp_ctx_create": (C.c_int, [C.c_int, C.POINTER(H)]),
        "flgp_ctx_destroy": (None, [H]),
        "flgp_ctx_set_stream": (C.c_int, [H, p_void]),
        "flgp_ctx_synchronize": (C.c_int, [H]),
        "flgp_ctx_launch_count": (c_u64, [H]),
        "flgp_ctx_set_timing": (C.c_int, [H, C.c_int]),
        "flgp_ctx_stage_reset": (C.c_int, [H]),
        "flgp_ctx_stage_count": (C.c_int, [H]),
        "flgp_ctx_stage_get": (C.c_int, [H, C.c_int, C.c_char_p, C.c_int, p_f64, C.POINTER(c_u64), p_f64, p_f64]),
        "flgp_dfma_peak": (C.c_int, [H, C.c_int, p_f64]),
        "flgp_copy_roundtrip": (C.c_int, [H, p_void, p_void, C.c_size_t]),
        "flgp_comm_unique_id": (C.c_int, [p_void]),
        "flgp_ctx_comm_init": (C.c_int, [H, p_void, C.c_int, C.c_int]),
        "flgp_default_init": (C.c_int, [c_i64, C.c_int, c_u64, p_i32]),
        "flgp_subsample": (C.c_int, [H, p_f64, c_i64, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, p_i32, c_u64,
                                     p_f64, p_i32, C.POINTER(C.c_int)]),
        "flgp_knn": (C.c_int, [H, p_f64, c_i64, C.c_int, p_f64, C.c_int, C.c_int, p_i32, p_f64, p_i32, p_f64]),
        "flgp_simplex_project": (C.c_int, [H, p_f64, C.c_int, p_f64]),
        "flgp_lae_point": (C.c_int, [H, p_f64, C.c_int, p_f64, C.c_int, p_f64]),
        "flgp_lae": (C.c_int, [H, p_f64, c_i64, C.c_int, p_f64, C.c_int, C.c_int, p_i32, p_f64, p_i64]),
        "flgp_graph_laplacian": (C.c_int, [H, c_i64, C.c_int, C.c_int, p_i32, p_f64, C.c_int, p_f64]),
        "flgp_cross_similarity_lae": (C.c_int, [H, p_f64, c_i64, C.c_int, p_f64, C.c_int, C.c_int, C.c_int, C.c_int,
                                                p_i32, p_f64]),
        "flgp_cross_similarity_se": (C.c_int, [H, p_f64, c_i64, C.c_int, p_f64, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_double, p_i32, p_f64]),
        "flgp_spectrum_from_z": (C.c_int, [H, c_i64, C.c_int, C.c_int, p_i32, p_f64, C.c_int, C.c_int, p_f64, p_f64,
                                           C.POINTER(H)]),
        "flgp_eigs_sym": (C.c_int, [H, p_f64, C.c_int, C.c_int, p_f64, p_f64]),
        "flgp_heat_kernel_spectrum": (C.c_int, [H, p_f64, c_i64, p_f64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                                p_i32, c_u64, C.POINTER(H)]),
        "flgp_heat_kernel_spectrum_sharded": (C.c_int, [H, p_f64, c_i64, c_i64, c_i64, C.c_int, C.c_int, C.c_int,
                                                        C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                                        C.c_double, C.c_int, p_i32, c_u64, C.POINTER(H)]),
        "flgp_heat_kernel_spectrum_dev": (C.c_int, [H, p_void, c_i64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                                    C.c_int, p_i32, c_u64, C.POINTER(H)]),
        "flgp_spectrum_free": (None, [H]),
        "flgp_spectrum_info": (C.c_int, [H, p_i64]),
        "flgp_spectrum_values": (C.c_int, [H, p_f64]),
        "flgp_spectrum_anchors": (C.c_int, [H, p_f64]),
        "flgp_spectrum_z": (C.c_int, [H, p_i32, p_f64]),
        "flgp_spectrum_vectors": (C.c_int, [H, p_f64]),
        "flgp_spectrum_gather_rows": (C.c_int, [H, p_i32, c_i64, p_f64]),
        "flgp_hk_from_spectrum": (C.c_int, [H, C.c_int, C.c_double, p_i32, c_i64, p_i32, c_i64, p_f64]),
        "flgp_lae_eigenmap": (C.c_int, [H, p_f64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                        C.c_int, C.c_int, p_i32, c_u64, p_f64, p_f64]),
        "flgp_heat_kernel_covariance": (C.c_int, [H, p_f64, c_i64, p_f64, c_i64, C.c_int, C.c_int, C.c_int, C.c_double,
                                                  C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                                  C.c_double, C.c_int, p_i32, c_u64, p_f64]),
        "flgp_regression_fixed": (C.c_int, [H, p_f64, c_i64, C.c_int, C.c_double, C.c_double, C.c_double, p_f64,
                                            p_f64]),
        "flgp_regression_fixed_dev": (C.c_int, [H, p_void, c_i64, C.c_int, C.c_double, C.c_double, C.c_double, p_void,
                                                p_void]),
        "flgp_fit_lae_regression_fixed": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int,
                                                    C.c_int, C.c_double, C.c_double, C.c_double, C.c_char_p,
                                                    C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, p_i32, c_u64,
                                                    p_f64, p_f64, p_f64]),
        "flgp_regression_objective": (C.c_int, [H, p_f64, c_i64, C.c_int, C.c_double, C.c_char_p, p_f64, p_f64, p_f64]),
        "flgp_train_regression": (C.c_int, [H, p_f64, c_i64, C.c_int, C.c_double, C.c_char_p, p_f64, p_f64,
                                            C.POINTER(C.c_int)]),
        "flgp_mma_minimize": (C.c_int, [C.c_int, OBJECTIVE_FN, p_void, p_f64, p_f64, p_f64, p_f64, C.c_double, C.c_int,
                                        C.POINTER(C.c_int)]),
        "flgp_fit_lae_regression": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_double, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                              C.c_int, C.c_int, p_i32, c_u64, p_f64, p_f64, p_f64, p_f64, p_f64]),
        "flgp_fit_se_regression": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_double, p_f64, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                             C.c_int, C.c_int, p_i32, c_u64, p_f64, p_f64, p_f64, p_f64, p_f64, p_f64,
                                             p_f64, C.POINTER(H)]),
        "flgp_fit_nystrom_regression": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int,
                                                  C.c_double, p_f64, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                                  p_i32, c_u64, p_f64, p_f64, p_f64, p_f64, p_f64, p_f64, p_f64]),
        "flgp_fit_nystrom_regression_sharded": (C.c_int, [H, p_f64, c_i64, c_i64, c_i64, C.c_int, p_f64, c_i64, C.c_int,
                                                          C.c_int, C.c_double, p_f64, C.c_int, C.c_char_p, C.c_char_p,
                                                          C.c_int, C.c_int, p_i32, c_u64, p_f64, p_f64, p_f64, p_f64,
                                                          p_f64, p_f64]),
        "flgp_logit_objective": (C.c_int, [H, p_f64, p_f64, c_i64, C.c_int, C.c_double, C.c_char_p, C.c_double, p_f64]),
        "flgp_train_logit": (C.c_int, [H, p_f64, p_f64, c_i64, C.c_int, C.c_double, C.c_char_p, p_f64, p_f64,
                                       C.POINTER(C.c_int)]),
        "flgp_train_logit_mult": (C.c_int, [H, p_f64, c_i64, C.c_int, C.c_double, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                            p_f64, p_f64]),
        "flgp_cobyla_minimize_1d": (C.c_int, [OBJECTIVE_FN, p_void, C.c_double, C.c_double, p_f64, p_f64, C.c_double,
                                              C.c_int, C.POINTER(C.c_int)]),
        "flgp_fit_lae_logit": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, p_f64,
                                         C.c_double, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, p_i32, c_u64, p_f64, p_f64, p_f64, p_f64, p_f64]),
        "flgp_fit_se_logit": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, p_f64,
                                        C.c_double, p_f64, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, p_i32, c_u64, p_f64, p_f64, p_f64, p_f64, p_f64, p_f64, C.POINTER(H)]),
        "flgp_fit_se_logit_mult": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_double, p_f64, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                             C.c_int, C.c_int, p_i32, c_u64, C.c_int, C.POINTER(C.c_int), p_f64, p_f64,
                                             p_f64, p_f64, C.POINTER(H)]),
        "flgp_fit_nystrom_logit": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int, p_f64,
                                             C.c_double, p_f64, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, p_i32,
                                             c_u64, p_f64, p_f64, p_f64, p_f64, p_f64, p_f64]),
        "flgp_fit_nystrom_logit_mult": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int,
                                                  C.c_double, p_f64, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                                  p_i32, c_u64, C.c_int, C.POINTER(C.c_int), p_f64, p_f64, p_f64, p_f64,
                                                  p_f64, p_f64]),
        "flgp_marginal_log_likelihood_logit_la": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, C.c_double, C.c_int, p_f64]),
        "flgp_logit_objective_rows": (C.c_int, [p_f64, p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, C.c_char_p,
                                                C.c_double, p_f64]),
        "flgp_train_logit_rows": (C.c_int, [p_f64, p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, C.c_char_p, p_f64,
                                            p_f64, C.POINTER(C.c_int)]),
        "flgp_classification_fold_rows": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, C.c_double,
                                                    C.c_double, C.c_int, p_f64, p_f64]),
        "flgp_multi_train_split": (C.c_int, [p_f64, c_i64, C.c_int, C.POINTER(C.c_int), p_f64]),
        "flgp_negative_log_likelihood": (C.c_int, [p_f64, p_f64, p_f64, c_i64, C.c_char_p, p_f64]),
        "flgp_test_regression": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, c_i64, p_f64]),
        "flgp_regression_objective_rows": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, C.c_char_p, p_f64,
                                                     p_f64, p_f64]),
        "flgp_train_regression_rows": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, C.c_char_p, p_f64,
                                                 p_f64, C.POINTER(C.c_int)]),
        "flgp_regression_objective_diff_rows": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, C.c_char_p,
                                                          p_f64, p_f64, p_f64]),
        "flgp_train_regression_diff_rows": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, C.c_char_p, p_f64,
                                                      p_f64, C.POINTER(C.c_int)]),
        "flgp_predict_coef_diff_rows": (C.c_int, [p_f64, p_f64, p_f64, C.c_int, C.c_int, C.c_double, p_f64, p_f64]),
        "flgp_fit_lae_regression_diff_noise": (C.c_int, [H, p_f64, p_f64, p_f64, c_i64, c_i64, C.c_int, C.c_int, C.c_int,
                                                         C.c_int, C.c_double, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                                                         C.c_int, C.c_int, C.c_int, p_i32, c_u64, p_f64, p_f64, p_f64,
                                                         p_f64, p_f64]),
        "flgp_classification_posterior_fixed": (C.c_int, [H, p_f64, c_i64, C.c_int, C.c_double, C.c_double, C.c_double,
                                                          C.c_int, p_f64, p_f64]),
        "flgp_posterior_distribution_classification": (C.c_int, [H, p_f64, p_f64, p_f64, p_f64, C.c_int, c_i64,
                                                                 C.c_double, C.c_int, p_f64, p_f64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return sig


SIGNATURES = None


def load():
    """Load the CUDA library; raises if it has not been built (python -m flgp_b200.build)."""
    global _lib, SIGNATURES
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libflgp_b200.so is missing (%s). Build it with `python -m flgp_b200.build`; "
                "flgp_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        SIGNATURES = _declare(lib)
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().flgp_last_error()
        raise FlgpError("%s (status %d)" % (msg.decode("utf-8", "replace") if msg else "unknown error", rc))
