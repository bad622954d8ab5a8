"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/pplacer_oracle.c) and of
the reference's own linear_c.c compiled into oracle/_ref.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by the
product package pplacer_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "liblinear_ref.so")

c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u8_p = C.POINTER(C.c_uint8)
c_u16_p = C.POINTER(C.c_uint16)


def build(force: bool = False) -> None:
    srcs = [os.path.join(HERE, f) for f in ("pplacer_oracle.c", "pplacer_oracle.h", "qng_tables.h")]
    stale = (not os.path.exists(ORACLE_SO)
             or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs))
    if force or stale:
        subprocess.run(["make", "-C", HERE, "_build/liboracle.so"], check=True, capture_output=True)
    if os.path.exists("/root/reference/pplacer_src/linear_c.c") and (force or not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class Prefs(C.Structure):
    _fields_ = [("start_pend", C.c_double), ("max_pend", C.c_double), ("initial_tolerance", C.c_double),
                ("final_tolerance", C.c_double), ("max_iter", C.c_int), ("strike_box", C.c_double),
                ("max_strikes", C.c_int), ("max_pitches", C.c_int), ("keep_at_most", C.c_int),
                ("keep_factor", C.c_double), ("calc_pp", C.c_int), ("pp_rel_err", C.c_double),
                ("prior_kind", C.c_int), ("prior_mean", C.c_double), ("prior_means", c_double_p)]


class Result(C.Structure):
    _fields_ = [("edge", C.c_int32), ("_pad", C.c_int32), ("log_like", C.c_double), ("pendant", C.c_double),
                ("distal", C.c_double), ("ml_ratio", C.c_double), ("marginal_like", C.c_double),
                ("post_prob", C.c_double)]


RESULT_DTYPE = np.dtype([("edge", "<i4"), ("_pad", "<i4"), ("log_like", "<f8"), ("pendant", "<f8"),
                         ("distal", "<f8"), ("ml_ratio", "<f8"), ("marginal_like", "<f8"),
                         ("post_prob", "<f8")])


class Debug(C.Structure):
    _fields_ = [("scores", c_double_p), ("n_pitches", c_i32_p), ("pitch_edges", c_i32_p),
                ("pitch_ll", c_double_p), ("pitch_pend", c_double_p), ("pitch_dist", c_double_p),
                ("n_like_calls", c_i64_p), ("timing", c_double_p)]


_lib = None
_ref = None
FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        L.orc_default_prefs.argtypes = [C.POINTER(Prefs)]
        L.orc_discrete_gamma.argtypes = [C.c_int, C.c_double, c_double_p]
        L.orc_gamma_inc_P.restype = C.c_double
        L.orc_gamma_inc_P.argtypes = [C.c_double, C.c_double]
        L.orc_gamma_Pinv.restype = C.c_double
        L.orc_gamma_Pinv.argtypes = [C.c_double, C.c_double]
        L.orc_diagd_normed_of_exchangeable_pair.argtypes = [C.c_int] + [c_double_p] * 5
        L.orc_symm_eigs.argtypes = [C.c_int] + [c_double_p] * 3
        L.orc_multi_exp.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                                    c_u8_p, C.c_double, c_double_p]
        L.orc_emper_freq.argtypes = [C.c_int, c_u8_p, C.c_int64, c_double_p]
        L.orc_lv_of_char.argtypes = [C.c_int, C.c_uint8, c_double_p]
        L.orc_brent.argtypes = [FN, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                c_double_p, C.POINTER(C.c_int)]
        L.orc_qng.argtypes = [FN, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, c_double_p,
                              c_double_p, C.POINTER(C.c_int)]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                                 c_double_p, c_i32_p, C.c_int, C.c_int, c_i32_p, c_i32_p, c_double_p,
                                 c_i32_p, c_u8_p, C.c_int, C.c_int, C.c_char_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_precompute.argtypes = [C.c_void_p]
        L.orc_glv_a.restype = c_double_p
        L.orc_glv_a.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_glv_e.restype = c_i64_p
        L.orc_glv_e.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_glv_a_len.restype = C.c_int64
        L.orc_glv_a_len.argtypes = [C.c_void_p]
        L.orc_node_tree_likes.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
        L.orc_place.argtypes = [C.c_void_p, C.POINTER(Prefs), C.c_int64, c_u8_p, C.c_void_p, c_i32_p,
                                c_i32_p, C.POINTER(Debug), C.c_int]
        L.orc_three_tax_ll.argtypes = [C.c_void_p, c_u8_p, C.c_int, C.c_double, C.c_double, c_double_p]
        L.orc_set_figs.argtypes = [C.c_void_p, C.c_int, c_i32_p, c_i32_p, c_i32_p]
        L.orc_refine_site_categories.argtypes = [C.c_void_p, c_i32_p, c_double_p]
        i, d = C.c_int, C.c_double
        L.orc_k_gemmish.argtypes = [c_double_p, c_double_p, c_double_p, i, i]
        L.orc_k_dediagonalize.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, i]
        L.orc_k_pairwise_prod.argtypes = [c_double_p, c_double_p, c_double_p, i, i, i, i]
        L.orc_k_statd_pairwise_prod.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, i, i, i, i]
        L.orc_k_log_like3.restype = d
        L.orc_k_log_like3.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, i, i, i, i]
        L.orc_k_masked_logdot.restype = d
        L.orc_k_masked_logdot.argtypes = [c_double_p, c_double_p, c_u16_p, i, i, i, i]
        L.orc_k_logdot.restype = d
        L.orc_k_logdot.argtypes = [c_double_p, c_double_p, i, i, i, i]
        _lib = L
    return _lib


def ref_lib():
    """The reference's own linear_c.c (oracle/_ref/liblinear_ref.so); None when not built."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(REF_SO):
            return None
        R = C.CDLL(REF_SO)
        d, i = C.c_double, C.c_int
        R.ref_gemmish.argtypes = [c_double_p, c_double_p, c_double_p, i, i]
        R.ref_dediagonalize.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, i]
        R.ref_mat_log_like3.restype = d
        R.ref_mat_log_like3.argtypes = [c_double_p] * 4 + [i, i]
        R.ref_mat_pairwise_prod.argtypes = [c_double_p] * 3 + [i, i]
        R.ref_mat_statd_pairwise_prod.argtypes = [c_double_p] * 4 + [i, i]
        R.ref_mat_masked_logdot.restype = d
        R.ref_mat_masked_logdot.argtypes = [c_double_p, c_double_p, c_u16_p, i, i]
        R.ref_mat_bounded_logdot.restype = d
        R.ref_mat_bounded_logdot.argtypes = [c_double_p, c_double_p, i, i, i, i]
        R.ref_ten_log_like3.restype = d
        R.ref_ten_log_like3.argtypes = [c_double_p] * 5 + [i, i, i]
        R.ref_ten_pairwise_prod.argtypes = [c_double_p] * 3 + [i, i, i]
        R.ref_ten_statd_pairwise_prod.argtypes = [c_double_p] * 4 + [i, i, i]
        R.ref_ten_masked_logdot.restype = d
        R.ref_ten_masked_logdot.argtypes = [c_double_p, c_double_p, c_u16_p, c_double_p, i, i, i]
        R.ref_ten_bounded_logdot.restype = d
        R.ref_ten_bounded_logdot.argtypes = [c_double_p, c_double_p, i, i, c_double_p, i, i, i]
        R.ref_vec_pairwise_prod.argtypes = [c_double_p] * 3 + [i]
        R.ref_int_vec_tot.restype = i
        R.ref_int_vec_tot.argtypes = [c_u16_p, i]
        R.ref_int_vec_pairwise_prod.argtypes = [c_u16_p] * 3 + [i]
        R.ref_float_mat_int_vec_mul.restype = i
        R.ref_float_mat_int_vec_mul.argtypes = [c_double_p, i, c_double_p, i, i, c_u16_p, i]
        _ref = R
    return _ref


def dp(a):
    return a.ctypes.data_as(c_double_p)


def ip(a):
    return a.ctypes.data_as(c_i32_p)


def up(a):
    return a.ctypes.data_as(c_u8_p)


def default_prefs(**kw) -> Prefs:
    p = Prefs()
    lib().orc_default_prefs(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def discrete_gamma(n_cat: int, alpha: float) -> np.ndarray:
    out = np.zeros(n_cat)
    lib().orc_discrete_gamma(n_cat, alpha, dp(out))
    return out


def diagd(exch: np.ndarray, pi: np.ndarray):
    S = len(pi)
    exch = np.ascontiguousarray(exch, dtype=np.float64)
    pi = np.ascontiguousarray(pi, dtype=np.float64)
    x, lam, xit = np.zeros((S, S)), np.zeros(S), np.zeros((S, S))
    lib().orc_diagd_normed_of_exchangeable_pair(S, dp(exch), dp(pi), dp(x), dp(lam), dp(xit))
    return x, lam, xit


def emper_freq(S: int, chars: np.ndarray) -> np.ndarray:
    flat = np.ascontiguousarray(chars, dtype=np.uint8).reshape(-1)
    out = np.zeros(S)
    st = lib().orc_emper_freq(S, up(flat), flat.size, dp(out))
    if st:
        raise ValueError("unknown base in alignment")
    return out


def multi_exp(x, lam, xit, rates, bl, mask=None) -> np.ndarray:
    S, R = len(lam), len(rates)
    out = np.zeros((R, S, S))
    rates = np.ascontiguousarray(rates, dtype=np.float64)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    lib().orc_multi_exp(S, R, dp(np.ascontiguousarray(x)), dp(np.ascontiguousarray(lam)),
                        dp(np.ascontiguousarray(xit)), dp(rates), None if m is None else up(m), bl, dp(out))
    return out


def brent(pyf, raw_start, left, right, tol, max_iters=100):
    calls = C.c_int(0)
    out = C.c_double(0.0)
    cb = FN(lambda x, _ud: float(pyf(x)))
    st = lib().orc_brent(cb, None, raw_start, left, right, tol, max_iters, C.byref(out), C.byref(calls))
    return st, out.value, calls.value


def qng(pyf, a, b, epsabs, epsrel):
    res, err, ne = C.c_double(0), C.c_double(0), C.c_int(0)
    cb = FN(lambda x, _ud: float(pyf(x)))
    st = lib().orc_qng(cb, None, a, b, epsabs, epsrel, C.byref(res), C.byref(err), C.byref(ne))
    return st, res.value, err.value, ne.value


class Oracle:
    """CPU oracle over a pplacer_b200.reference.Problem."""

    def __init__(self, prob, mmap_dir: str | None = None):
        L = lib()
        m = prob.model
        self.prob = prob
        self.S, self.gcat = m.n_states, m.model_class
        if m.model_statd is None or m.empirical_freqs:
            self.statd = emper_freq(m.n_states, prob.ref_chars_raw)   # un-premasked alignment
        else:
            self.statd = np.ascontiguousarray(m.model_statd, dtype=np.float64)
        self.x, self.lam, self.xit = diagd(m.exch, self.statd)
        if m.model_class == 0:
            self.rates = discrete_gamma(m.n_rates, m.gamma_alpha) if m.rates is None else np.asarray(m.rates, float)
        else:
            self.rates = np.ascontiguousarray(m.rates, dtype=np.float64)
        self.R = len(self.rates)
        t = prob.tree
        ptr, idx = t.csr()
        bl = np.ascontiguousarray(np.nan_to_num(t.bl, nan=0.0), dtype=np.float64)
        sc = None if prob.site_cat is None else np.ascontiguousarray(prob.site_cat, dtype=np.int32)
        self._keep = (ptr, idx, bl, sc)
        self.n_nodes, self.L = t.n_nodes, prob.L
        self.ctx = L.orc_create(self.S, self.R, self.gcat, dp(self.statd), dp(self.x), dp(self.lam),
                                dp(self.xit), dp(self.rates), None if sc is None else ip(sc), t.n_nodes,
                                t.root, ip(ptr), ip(idx), dp(bl), ip(prob.leaf_row), up(prob.ref_chars),
                                prob.ref_chars.shape[0], prob.L,
                                None if mmap_dir is None else mmap_dir.encode())
        if not self.ctx:
            raise RuntimeError("orc_create failed")
        self.precomputed = False

    def close(self):
        if self.ctx:
            lib().orc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def precompute(self):
        lib().orc_precompute(self.ctx)
        self.precomputed = True

    def set_figs(self, figs):
        """figs: list of (rep, edges) from pplacer_b200.fig.figs_of_tree, or None for the plain ranking"""
        if not figs:
            lib().orc_set_figs(self.ctx, 0, None, None, None)
            return
        from pplacer_b200.fig import figs_csr
        rep, ptr, idx = figs_csr(figs)
        lib().orc_set_figs(self.ctx, len(rep), ip(rep), ip(ptr), ip(idx))

    def refine_site_categories(self, want_site_ll: bool = False):
        """Gcat_model.refine (gcat_model.ml:387-419); the context keeps the new categories."""
        cats = np.zeros(self.L, dtype=np.int32)
        sll = np.zeros((self.R, self.L)) if want_site_ll else None
        st = lib().orc_refine_site_categories(self.ctx, ip(cats), None if sll is None else dp(sll))
        if st:
            raise RuntimeError("a site has zero likelihood")
        self.precomputed = False
        return (cats, sll) if want_site_ll else cats

    def glv(self, which: int, node: int):
        """(a, e) in the reference layout: Gmix a[R][L][S], Gcat a[L][S]; copies."""
        n = lib().orc_glv_a_len(self.ctx)
        a = np.ctypeslib.as_array(lib().orc_glv_a(self.ctx, which, node), shape=(n,)).copy()
        e = np.ctypeslib.as_array(lib().orc_glv_e(self.ctx, which, node), shape=(self.L,)).copy()
        shape = (self.L, self.S) if self.gcat else (self.R, self.L, self.S)
        return a.reshape(shape), e

    def node_tree_likes(self, node: int):
        a, b = C.c_double(0), C.c_double(0)
        lib().orc_node_tree_likes(self.ctx, node, C.byref(a), C.byref(b))
        return a.value, b.value

    def three_tax_ll(self, query_chars: np.ndarray, edge: int, pendant: float, distal: float) -> float:
        q = np.ascontiguousarray(query_chars, dtype=np.uint8)
        out = C.c_double(0)
        st = lib().orc_three_tax_ll(self.ctx, up(q), edge, pendant, distal, C.byref(out))
        if st:
            raise RuntimeError("orc_three_tax_ll status %d" % st)
        return out.value

    def place(self, query_chars: np.ndarray, prefs: Prefs | None = None, n_threads: int = 1, debug: bool = False):
        if not self.precomputed:
            self.precompute()
        prefs = prefs or default_prefs()
        q = np.ascontiguousarray(query_chars, dtype=np.uint8)
        Q = q.shape[0]
        K = prefs.keep_at_most
        out = np.zeros((Q, K), dtype=RESULT_DTYPE)
        n_out = np.zeros(Q, dtype=np.int32)
        status = np.zeros(Q, dtype=np.int32)
        dbg = None
        info = {}
        if debug:
            P = prefs.max_pitches + 1
            info = dict(scores=np.zeros((Q, self.n_nodes)), n_pitches=np.zeros(Q, np.int32),
                        pitch_edges=np.full((Q, P), -1, np.int32), pitch_ll=np.full((Q, P), np.nan),
                        pitch_pend=np.full((Q, P), np.nan), pitch_dist=np.full((Q, P), np.nan),
                        n_like_calls=np.zeros(Q, np.int64), timing=np.zeros(4))
            dbg = Debug(dp(info["scores"]), ip(info["n_pitches"]), ip(info["pitch_edges"]), dp(info["pitch_ll"]),
                        dp(info["pitch_pend"]), dp(info["pitch_dist"]),
                        info["n_like_calls"].ctypes.data_as(c_i64_p), dp(info["timing"]))
        lib().orc_place(self.ctx, C.byref(prefs), Q, up(q), out.ctypes.data_as(C.c_void_p), ip(n_out),
                        ip(status), None if dbg is None else C.byref(dbg), n_threads)
        return out, n_out, status, info
