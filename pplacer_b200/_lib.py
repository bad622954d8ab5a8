"""ctypes binding of libpplacer_b200.so (the C ABI declared in include/pplacer_b200.h).

The library is the product: this module never falls back to any CPU implementation.  If the
shared object is missing it is built with nvcc (pplacer_b200/build.py); if no CUDA device is
present every compute entry point raises PplacerB200Error(PP_ENODEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u8_p = C.POINTER(C.c_uint8)

PP_OK, PP_EINVAL, PP_ENODEVICE, PP_ECUDA, PP_ENOMEM, PP_ESTATE = 0, 1, 2, 3, 4, 5

EXPORTED_SYMBOLS = [
    "pp_last_error", "pp_abi_version", "pp_default_prefs", "pp_create", "pp_destroy", "pp_precompute",
    "pp_export_glvs", "pp_get_model", "pp_place", "pp_place_device", "pp_prefilter_scores", "pp_prefilter_time",
    "pp_three_tax_log_like", "pp_evolve", "pp_discrete_gamma", "pp_diagd", "pp_emper_freq",
    "pp_kernel_launches", "pp_phase_times_ms", "pp_measure_fp64_peak", "pp_set_stream",
    "pp_set_prefilter_kernel", "pp_prefilter_kernel_used", "pp_set_three_tax_engine", "pp_three_tax_engine_used", "pp_three_tax_work", "pp_three_tax_profile", "pp_set_rank_mode", "pp_rank_fallbacks",
    "pp_create_multi", "pp_destroy_multi", "pp_multi_precompute", "pp_multi_place", "pp_multi_n_devices", "pp_multi_ctx",
    "pp_multi_last_counts", "pp_refine_site_categories", "pp_check_like", "pp_map_states", "pp_set_figs", "pp_n_figs",
]


class PplacerB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libpplacer_b200 error %d: %s" % (code, msg))
        self.code = code


class Model(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("model_class", C.c_int32), ("n_rates", C.c_int32), ("_pad", C.c_int32),
                ("exch", c_double_p), ("statd", c_double_p), ("statd_chars", c_u8_p), ("n_statd_chars", C.c_int64),
                ("rates", c_double_p), ("gamma_alpha", C.c_double), ("site_cat", c_i32_p)]


class Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("root_id", C.c_int32), ("child_ptr", c_i32_p), ("child_idx", c_i32_p),
                ("bl", c_double_p), ("leaf_row", c_i32_p)]


class Prefs(C.Structure):
    _fields_ = [("start_pend", C.c_double), ("max_pend", C.c_double), ("initial_tolerance", C.c_double),
                ("final_tolerance", C.c_double), ("max_iter", C.c_int32), ("max_strikes", C.c_int32),
                ("strike_box", C.c_double), ("max_pitches", C.c_int32), ("keep_at_most", C.c_int32),
                ("keep_factor", C.c_double), ("calc_pp", C.c_int32), ("prior_kind", C.c_int32),
                ("pp_rel_err", C.c_double), ("prior_mean", C.c_double), ("prior_means", c_double_p)]


class Debug(C.Structure):
    _fields_ = [("scores", c_double_p), ("n_pitches", c_i32_p), ("pitch_edges", c_i32_p), ("pitch_ll", c_double_p),
                ("pitch_pend", c_double_p), ("pitch_dist", c_double_p), ("n_like_calls", c_i64_p),
                ("timing", c_double_p)]


RESULT_DTYPE = np.dtype([("edge", "<i4"), ("_pad", "<i4"), ("log_like", "<f8"), ("pendant", "<f8"),
                         ("distal", "<f8"), ("ml_ratio", "<f8"), ("marginal_like", "<f8"), ("post_prob", "<f8")])

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building if stale/missing) libpplacer_b200.so and declare the ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.is_stale():
        _build.build()
    L = C.CDLL(_build.LIB)
    L.pp_last_error.restype = C.c_char_p
    L.pp_abi_version.restype = C.c_int
    L.pp_default_prefs.argtypes = [C.POINTER(Prefs)]
    L.pp_default_prefs.restype = None
    L.pp_create.argtypes = [C.POINTER(Model), C.POINTER(Tree), c_u8_p, C.c_int32, C.c_int32, C.c_int32,
                            C.POINTER(C.c_void_p)]
    L.pp_destroy.argtypes = [C.c_void_p]
    L.pp_destroy.restype = None
    L.pp_precompute.argtypes = [C.c_void_p]
    L.pp_export_glvs.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_i64_p]
    L.pp_get_model.argtypes = [C.c_void_p] + [c_double_p] * 5
    L.pp_place.argtypes = [C.c_void_p, C.POINTER(Prefs), C.c_int64, c_u8_p, C.c_void_p, c_i32_p, c_i32_p,
                           C.POINTER(Debug)]
    L.pp_place_device.argtypes = [C.c_void_p, C.POINTER(Prefs), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    L.pp_prefilter_scores.argtypes = [C.c_void_p, C.c_double, C.c_int64, c_u8_p, c_double_p, C.c_int32]
    L.pp_prefilter_time.argtypes = [C.c_void_p, C.c_double, C.c_int64, c_u8_p, C.c_int32, C.c_int32, c_double_p]
    L.pp_three_tax_log_like.argtypes = [C.c_void_p, C.c_int64, c_u8_p, c_i32_p, c_double_p, c_double_p, c_double_p]
    L.pp_evolve.argtypes = [C.c_void_p, c_double_p, c_i64_p, C.c_int32, C.c_double, c_double_p, c_i64_p]
    L.pp_discrete_gamma.argtypes = [C.c_int32, C.c_double, c_double_p]
    L.pp_diagd.argtypes = [C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
    L.pp_emper_freq.argtypes = [C.c_int32, c_u8_p, C.c_int64, c_double_p]
    L.pp_kernel_launches.argtypes = [C.c_void_p, C.c_int32]
    L.pp_kernel_launches.restype = C.c_int64
    L.pp_phase_times_ms.argtypes = [C.c_void_p, c_double_p, C.c_int32]
    L.pp_measure_fp64_peak.argtypes = [C.c_void_p, c_double_p]
    L.pp_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.pp_set_prefilter_kernel.argtypes = [C.c_void_p, C.c_int32]
    L.pp_prefilter_kernel_used.argtypes = [C.c_void_p]
    L.pp_set_three_tax_engine.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.pp_three_tax_engine_used.argtypes = [C.c_void_p]
    L.pp_three_tax_work.argtypes = [C.c_void_p, c_double_p, C.c_int32]
    L.pp_three_tax_profile.argtypes = [C.c_void_p, c_double_p, C.c_int32]
    L.pp_set_rank_mode.argtypes = [C.c_void_p, C.c_int32]
    L.pp_rank_fallbacks.argtypes = [C.c_void_p]
    L.pp_set_figs.argtypes = [C.c_void_p, C.c_int32, c_i32_p, c_i32_p, c_i32_p]
    L.pp_n_figs.argtypes = [C.c_void_p]
    L.pp_refine_site_categories.argtypes = [C.c_void_p, c_i32_p, c_double_p]
    L.pp_check_like.argtypes = [C.c_void_p, c_double_p]
    L.pp_map_states.argtypes = [C.c_void_p, C.c_int32, c_i32_p, C.c_double, C.POINTER(C.c_int8)]
    L.pp_create_multi.argtypes = [C.POINTER(Model), C.POINTER(Tree), c_u8_p, C.c_int32, C.c_int32, C.c_int32, c_i32_p,
                                  C.POINTER(C.c_void_p)]
    L.pp_destroy_multi.argtypes = [C.c_void_p]
    L.pp_destroy_multi.restype = None
    L.pp_multi_precompute.argtypes = [C.c_void_p]
    L.pp_multi_place.argtypes = [C.c_void_p, C.POINTER(Prefs), C.c_int64, c_u8_p, C.c_void_p, c_i32_p, c_i32_p]
    L.pp_multi_n_devices.argtypes = [C.c_void_p]
    L.pp_multi_ctx.argtypes = [C.c_void_p, C.c_int32]
    L.pp_multi_ctx.restype = C.c_void_p
    L.pp_multi_last_counts.argtypes = [C.c_void_p, c_i64_p]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != PP_OK:
        raise PplacerB200Error(rc, load().pp_last_error().decode("utf-8", "replace"))


def dp(a):
    return a.ctypes.data_as(c_double_p)


def ip(a):
    return a.ctypes.data_as(c_i32_p)


def up(a):
    return a.ctypes.data_as(c_u8_p)


def default_prefs(**kw) -> Prefs:
    p = Prefs()
    load().pp_default_prefs(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def pplacer_p_prefs(**kw) -> Prefs:
    """`-p` (prefs.ml:198-203): calc_pp, keep_at_most 20, keep_factor 0.001, max_strikes 20."""
    base = dict(calc_pp=1, keep_at_most=20, keep_factor=0.001, max_strikes=20)
    base.update(kw)
    return default_prefs(**base)


def discrete_gamma(n_cat: int, alpha: float) -> np.ndarray:
    out = np.zeros(n_cat)
    check(load().pp_discrete_gamma(n_cat, alpha, dp(out)))
    return out


def diagd(exch: np.ndarray, pi: np.ndarray):
    S = len(pi)
    exch = np.ascontiguousarray(exch, dtype=np.float64)
    pi = np.ascontiguousarray(pi, dtype=np.float64)
    x, lam, xit = np.zeros((S, S)), np.zeros(S), np.zeros((S, S))
    check(load().pp_diagd(S, dp(exch), dp(pi), dp(x), dp(lam), dp(xit)))
    return x, lam, xit


def emper_freq(S: int, chars: np.ndarray) -> np.ndarray:
    flat = np.ascontiguousarray(chars, dtype=np.uint8).reshape(-1)
    out = np.zeros(S)
    check(load().pp_emper_freq(S, up(flat), flat.size, dp(out)))
    return out
