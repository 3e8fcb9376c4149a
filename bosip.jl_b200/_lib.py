"""ctypes binding of libbosip_b200.so -- the Python stand-in for Julia's `ccall` (INTEGRATION.md shows the Julia stubs).

Loading fails loudly: there is no CPU fallback and no alternative implementation behind these symbols."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BOSIP_B200_LIB", os.path.join(HERE, "libbosip_b200.so"))  # env override: development A/B builds only

OK, EINVAL, ECUDA, ENOTPD, EDOMAIN, ENOMEM = 0, -1, -2, -3, -4, -5
SE, MATERN32, MATERN52 = 0, 1, 2
HOST, DEVICE = 0, 1
WANT_APPROX, WANT_MEAN, WANT_VAR = 1, 2, 4
ACQ_MAXVAR, ACQ_LOGMAXVAR = 0, 1
PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_TRUNCNORMAL = 0, 1, 2
CAND_POINTS, CAND_PHILOX, CAND_GRID = 0, 1, 2
VAR_FP64, VAR_INT8, VAR_AUTO = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class GpOpts(C.Structure):
    _fields_ = [("jitter", C.c_double), ("pred_noise", C.c_int32), ("clamp_var", C.c_int32)]


LIKE_NORMAL, LIKE_LOGNORMAL, LIKE_NORMAL_DIFF, LIKE_NORMAL_SUM, LIKE_EXP = 0, 1, 2, 3, 4


class LikeNormal(C.Structure):
    _fields_ = [("p", C.c_int32), ("kind", C.c_int32), ("z_obs", c_double_p), ("std_obs", c_double_p), ("sum_lengths", c_int32_p)]


class Prior(C.Structure):
    _fields_ = [("d", C.c_int32), ("kind", c_int32_p), ("params", c_double_p), ("logprior", C.c_void_p)]


class CandSrc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mem", C.c_int32), ("points", C.c_void_p), ("seed", C.c_uint64),
                ("lb", c_double_p), ("ub", c_double_p), ("n_per_dim", C.c_int64)]


# name -> (restype, argtypes); every symbol include/bosip_b200.h declares
PROTOTYPES = {
    "bosip_b200_version": (C.c_int, []),
    "bosip_b200_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "bosip_b200_shutdown": (C.c_int, [C.c_void_p]),
    "bosip_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "bosip_b200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bosip_b200_stream_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bosip_b200_i8_issuers": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "bosip_b200_i8_pair": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "bosip_b200_get_variance_engine": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bosip_b200_launch_count": (C.c_int, [C.c_void_p, c_int64_p]),
    "bosip_b200_set_chunk": (C.c_int, [C.c_void_p, C.c_int64]),
    "bosip_b200_set_variance_engine": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bosip_b200_int8_peak": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "bosip_b200_int8_peak_random": (C.c_int, [C.c_void_p, c_double_p]),
    "bosip_b200_gp_fit": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                    c_double_p, c_double_p, c_double_p, C.POINTER(GpOpts), C.POINTER(C.c_void_p)]),
    "bosip_b200_gp_load_factors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                             c_double_p, c_double_p, c_double_p, c_double_p, C.POINTER(GpOpts),
                                             C.POINTER(C.c_void_p)]),
    "bosip_b200_gp_get_factors": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "bosip_b200_gp_free": (C.c_int, [C.c_void_p]),
    "bosip_b200_gp_dims": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bosip_b200_init_multi": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "bosip_b200_shutdown_multi": (C.c_int, [C.c_void_p]),
    "bosip_b200_multi_size": (C.c_int, [C.c_void_p]),
    "bosip_b200_multi_context": (C.c_void_p, [C.c_void_p, C.c_int]),
    "bosip_b200_gp_multi_replica": (C.c_void_p, [C.c_void_p, C.c_int]),
    "bosip_b200_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "bosip_b200_gp_fit_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                          c_double_p, c_double_p, c_double_p, C.POINTER(GpOpts), C.POINTER(C.c_void_p)]),
    "bosip_b200_gp_free_multi": (C.c_int, [C.c_void_p]),
    "bosip_b200_posterior_eval_multi": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), C.c_void_p, C.c_int64, C.c_void_p,
                                                  C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, c_int64_p, C.c_int, C.c_int64, c_int64_p,
                                                  c_double_p]),
    "bosip_b200_acq_argmax_multi": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), C.POINTER(CandSrc), C.c_int64,
                                              C.c_int64, C.c_int, c_int64_p, c_double_p, c_double_p]),
    "bosip_b200_evidence_sum_multi": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), C.c_void_p, C.c_int64, C.c_uint,
                                                c_double_p]),
    "bosip_b200_mmd_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                       c_double_p]),
    "bosip_b200_tv_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, c_double_p]),
    "bosip_b200_gp_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bosip_b200_posterior_eval": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), C.c_void_p, C.c_int64,
                                            C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, c_int64_p]),
    "bosip_b200_posterior_eval_argmax": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), C.c_void_p, C.c_int64,
                                                   C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, c_int64_p,
                                                   C.c_int, C.c_int64, c_int64_p, c_double_p]),
    "bosip_b200_posterior_eval_bi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(LikeNormal), C.POINTER(Prior), C.c_void_p,
                                               C.c_int64, C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, c_int64_p,
                                               C.c_int, C.c_int64, c_int64_p, c_double_p]),
    "bosip_b200_normal_likelihood_eval": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.c_void_p, C.c_void_p, C.c_int64,
                                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, c_int64_p]),
    "bosip_b200_acq_argmax": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), C.POINTER(CandSrc), C.c_int64,
                                        C.c_int64, C.c_int, c_int64_p, c_double_p, c_double_p]),
    "bosip_b200_acq_argmax_bi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(LikeNormal), C.POINTER(Prior), C.POINTER(CandSrc),
                                           C.c_int64, C.c_int64, C.c_int, c_int64_p, c_double_p, c_double_p]),
    "bosip_b200_evidence_sum_bi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(LikeNormal), C.POINTER(Prior), C.c_void_p, C.c_int64,
                                             C.c_int, C.c_uint, c_double_p]),
    "bosip_b200_marginals_int": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), c_double_p, C.c_int64, C.c_int,
                                           c_double_p, c_double_p, C.c_uint, c_double_p, c_double_p]),
    "bosip_b200_candidates": (C.c_int, [C.c_void_p, C.POINTER(CandSrc), C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "bosip_b200_evidence_sum": (C.c_int, [C.c_void_p, C.POINTER(LikeNormal), C.POINTER(Prior), C.c_void_p, C.c_int64,
                                          C.c_int, C.c_uint, c_double_p]),
    "bosip_b200_mmd_sums": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_int64, C.c_int, C.c_int, C.c_int, c_double_p]),
    "bosip_b200_mmd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_int, c_double_p]),
    "bosip_b200_tv_stage1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "bosip_b200_tv_stage2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double,
                                       C.c_double, c_double_p]),
    "bosip_b200_tv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "bosip_b200_mvnormal_sample": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "bosip_b200_mixture_pdf_sum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_void_p, C.c_int64, C.c_int,
                                             C.c_int, C.c_void_p]),
    "bosip_b200_amis_log_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_int, C.c_void_p]),
    "bosip_b200_weighted_moments": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p, c_double_p,
                                              c_double_p, C.c_void_p]),
    "bosip_b200_resample": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "bosip_b200_weighted_quantile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, c_double_p]),
    "bosip_b200_cutoff_area": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, c_double_p]),
    "bosip_b200_set_iou": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "bosip_b200_confidence_iou": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, c_double_p,
                                            c_double_p, c_double_p]),
    "bosip_b200_fp64_peak": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "bosip_b200_math_selftest": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "bosip_b200_timing_reset": (C.c_int, [C.c_void_p, C.c_int]),
    "bosip_b200_timing_get": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree CUDA library; raises (never falls back) when it is missing or a symbol is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python bosip.jl_b200/build.py` (nvcc, sm_100a). "
            "bosip_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
