"""Host-side driver over the C ABI: the part of Pplacer_run.run_placements (pplacer_run.ml:167-410)
and Core.pplacer_core (core.ml:76-409) that the reference's OCaml host would keep, written in
Python because no OCaml toolchain exists in the build image (INTEGRATION.md shows the OCaml stub).
All numerics happen inside libpplacer_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .reference import Problem, flat_exp_prior_mean, informative_prior_means


def abi_inputs(prob: Problem):
    """The pp_model / pp_tree structs and the reference alignment of a Problem, plus the arrays they point into."""
    m = prob.model
    t = prob.tree
    exch = np.ascontiguousarray(m.exch, dtype=np.float64)
    statd = None
    statd_chars = None
    if m.model_statd is not None and not m.empirical_freqs:
        statd = np.ascontiguousarray(m.model_statd, dtype=np.float64)
    else:
        statd_chars = np.ascontiguousarray(prob.ref_chars_raw, dtype=np.uint8).reshape(-1)
    rates = None if m.rates is None else np.ascontiguousarray(m.rates, dtype=np.float64)
    site_cat = None if prob.site_cat is None else np.ascontiguousarray(prob.site_cat, dtype=np.int32)
    ptr, idx = t.csr()
    bl = np.array(t.bl, dtype=np.float64)
    missing = [i for i in range(t.n_nodes) if i != t.root and not np.isfinite(bl[i])]
    if missing:   # Gtree.get_bl raises for a missing branch length (gtree.ml); only the root edge may lack one
        raise ValueError("tree edges without a branch length: %s" % missing[:8])
    bl[t.root] = 0.0 if not np.isfinite(bl[t.root]) else bl[t.root]
    bl = np.ascontiguousarray(bl)
    leaf_row = np.ascontiguousarray(prob.leaf_row, dtype=np.int32)
    ref_chars = np.ascontiguousarray(prob.ref_chars, dtype=np.uint8)
    keep = (exch, statd, statd_chars, rates, site_cat, ptr, idx, bl, leaf_row, ref_chars)
    cm = _lib.Model(m.n_states, m.model_class, m.n_rates, 0, _lib.dp(exch),
                    None if statd is None else _lib.dp(statd),
                    None if statd_chars is None else _lib.up(statd_chars),
                    0 if statd_chars is None else statd_chars.size,
                    None if rates is None else _lib.dp(rates), float(m.gamma_alpha),
                    None if site_cat is None else _lib.ip(site_cat))
    ct = _lib.Tree(t.n_nodes, t.root, _lib.ip(ptr), _lib.ip(idx), _lib.dp(bl), _lib.ip(leaf_row))
    return cm, ct, ref_chars, keep


class MultiPlacer:
    """One reference replicated on several GPUs of the box behind ONE handle (pp_create_multi): what replaces
    Multiprocessing's fork-per-job (pplacer_run.ml:618-649).  place() shards the queries inside the library - one host
    thread per device - and returns the records in query order."""

    def __init__(self, prob: Problem, devices):
        L = _lib.load()
        self.prob = prob
        cm, ct, ref_chars, self._keep = abi_inputs(prob)
        ids = np.ascontiguousarray(list(devices), dtype=np.int32)
        h = C.c_void_p()
        _lib.check(L.pp_create_multi(C.byref(cm), C.byref(ct), _lib.up(ref_chars), ref_chars.shape[0], prob.L, len(ids),
                                     _lib.ip(ids), C.byref(h)))
        self.h = h
        self.n_devices = len(ids)
        self.L = prob.L
        self.precomputed = False

    def close(self):
        if getattr(self, "h", None):
            _lib.load().pp_destroy_multi(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def precompute(self):
        _lib.check(_lib.load().pp_multi_precompute(self.h))
        self.precomputed = True

    def prefs(self, pp: bool = False, **kw) -> _lib.Prefs:
        p = _lib.pplacer_p_prefs(**kw) if pp else _lib.default_prefs(**kw)
        p.prior_kind = 1
        p.prior_mean = flat_exp_prior_mean(self.prob.tree)
        return p

    def place(self, query_chars: np.ndarray, prefs: Optional[_lib.Prefs] = None):
        if not self.precomputed:
            self.precompute()
        prefs = prefs or self.prefs()
        q = np.ascontiguousarray(query_chars, dtype=np.uint8)
        if q.ndim != 2 or q.shape[1] != self.L:
            raise ValueError("query is not the same length as the ref alignment")
        Q, K = q.shape[0], prefs.keep_at_most
        out = np.zeros((Q, K), dtype=_lib.RESULT_DTYPE)
        n_out = np.zeros(Q, dtype=np.int32)
        status = np.zeros(Q, dtype=np.int32)
        _lib.check(_lib.load().pp_multi_place(self.h, C.byref(prefs), Q, _lib.up(q), out.ctypes.data_as(C.c_void_p),
                                              _lib.ip(n_out), _lib.ip(status)))
        return out, n_out, status

    def last_counts(self) -> np.ndarray:
        c = np.zeros(self.n_devices, dtype=np.int64)
        _lib.check(_lib.load().pp_multi_last_counts(self.h, c.ctypes.data_as(_lib.c_i64_p)))
        return c


class Placer:
    """One reference (tree + alignment + model) resident on one GPU."""

    def __init__(self, prob: Problem, device: int = 0):
        L = _lib.load()
        self.prob = prob
        m = prob.model
        t = prob.tree
        cm, ct, ref_chars, self._keep = abi_inputs(prob)
        self.S, self.gcat = m.n_states, m.model_class
        self.R = m.n_rates
        self.n_nodes, self.L = t.n_nodes, prob.L
        ctx = C.c_void_p()
        _lib.check(L.pp_create(C.byref(cm), C.byref(ct), _lib.up(ref_chars), ref_chars.shape[0], prob.L, device,
                               C.byref(ctx)))
        self.ctx = ctx
        self.device = device
        self.precomputed = False

    # ---- lifecycle ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            _lib.load().pp_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- precompute -----------------------------------------------------------------------
    def precompute(self):
        _lib.check(_lib.load().pp_precompute(self.ctx))
        self.precomputed = True

    def set_figs(self, figs=None, cutoff: Optional[float] = None):
        """--fig-cutoff (fig.ml): figs = list of (rep, edges) from pplacer_b200.fig.figs_of_tree, or a cutoff to build them
        from this reference tree; None / 0 restores the plain ranking."""
        from . import fig as figmod
        if cutoff is not None:
            figs = figmod.figs_of_tree(cutoff, self.prob.tree)
        if not figs:
            _lib.check(_lib.load().pp_set_figs(self.ctx, 0, None, None, None))
            return 0
        rep, ptr, idx = figmod.figs_csr(figs)
        _lib.check(_lib.load().pp_set_figs(self.ctx, len(rep), _lib.ip(rep), _lib.ip(ptr), _lib.ip(idx)))
        return len(rep)

    def refine_site_categories(self, want_site_ll: bool = False):
        """Gcat_model.refine on the device (gcat_model.ml:387-419): returns the new zero-based category per site (and
        the R x L per-category site log-likelihoods).  The context keeps the new categories; precompute afterwards."""
        cats = np.zeros(self.L, dtype=np.int32)
        sll = np.zeros((self.R, self.L)) if want_site_ll else None
        _lib.check(_lib.load().pp_refine_site_categories(self.ctx, _lib.ip(cats), None if sll is None else _lib.dp(sll)))
        self.precomputed = False
        return (cats, sll) if want_site_ll else cats

    def check_like(self) -> np.ndarray:
        """--check-like (pplacer_run.ml:344-377): [n_nodes, 2] = tree_likelihood, supernode_like per edge."""
        out = np.zeros((self.n_nodes, 2))
        _lib.check(_lib.load().pp_check_like(self.ctx, _lib.dp(out)))
        return out

    def map_states(self, node_ids, cutoff: float = 0.8) -> np.ndarray:
        """Map_seq.of_map (map_seq.ml:17-37): int8[len(node_ids), L], per site the MAP state on the proximal side of
        each node or -1 below `cutoff` (Prefs.map_cutoff, 0.8) - what --map-mrca / --map-identity read."""
        ids = np.ascontiguousarray(node_ids, dtype=np.int32)
        out = np.zeros((len(ids), self.L), dtype=np.int8)
        _lib.check(_lib.load().pp_map_states(self.ctx, len(ids), ids.ctypes.data_as(_lib.c_i32_p), float(cutoff),
                                             out.ctypes.data_as(_lib.C.POINTER(_lib.C.c_int8))))
        return out

    def model_numerics(self):
        S, R = self.S, self.R
        statd, x, lam, xit, rates = np.zeros(S), np.zeros((S, S)), np.zeros(S), np.zeros((S, S)), np.zeros(R)
        _lib.check(_lib.load().pp_get_model(self.ctx, _lib.dp(statd), _lib.dp(x), _lib.dp(lam), _lib.dp(xit),
                                            _lib.dp(rates)))
        return dict(statd=statd, x=x, lam=lam, xit=xit, rates=rates)

    def glv(self, which: int, node: int):
        """(a, e) of one node in the reference layout: Gmix a[R][L][S], Gcat a[L][S]."""
        Rg = 1 if self.gcat else self.R
        a = np.zeros(Rg * self.L * self.S)
        e = np.zeros(self.L, dtype=np.int64)
        _lib.check(_lib.load().pp_export_glvs(self.ctx, which, node, node + 1, _lib.dp(a),
                                              e.ctypes.data_as(_lib.c_i64_p)))
        return a.reshape((self.L, self.S) if self.gcat else (self.R, self.L, self.S)), e

    # ---- stage-level entry points ---------------------------------------------------------
    def prefilter_scores(self, query_chars: np.ndarray, start_pend: float = 0.1, kernel: int = 0) -> np.ndarray:
        q = np.ascontiguousarray(query_chars, dtype=np.uint8)
        out = np.zeros((q.shape[0], self.n_nodes))
        _lib.check(_lib.load().pp_prefilter_scores(self.ctx, start_pend, q.shape[0], _lib.up(q), _lib.dp(out), kernel))
        return out

    def prefilter_time_ms(self, query_chars: np.ndarray, kernel: int = 0, reps: int = 3, start_pend: float = 0.1) -> float:
        """pp_prefilter_time: mean device time of one prefilter launch (kernel: 0 default, 1 direct FMA form, 2 table form)."""
        q = np.ascontiguousarray(query_chars, dtype=np.uint8)
        ms = C.c_double(0)
        _lib.check(_lib.load().pp_prefilter_time(self.ctx, start_pend, q.shape[0], _lib.up(q), kernel, reps, C.byref(ms)))
        return ms.value

    def three_tax_log_like(self, query_chars, edge, pendant, distal) -> np.ndarray:
        q = np.ascontiguousarray(query_chars, dtype=np.uint8)
        edge = np.ascontiguousarray(edge, dtype=np.int32)
        pend = np.ascontiguousarray(pendant, dtype=np.float64)
        dist = np.ascontiguousarray(distal, dtype=np.float64)
        out = np.zeros(q.shape[0])
        _lib.check(_lib.load().pp_three_tax_log_like(self.ctx, q.shape[0], _lib.up(q), _lib.ip(edge), _lib.dp(pend),
                                                     _lib.dp(dist), _lib.dp(out)))
        return out

    def evolve(self, a: np.ndarray, e: np.ndarray, bl: float):
        a = np.ascontiguousarray(a, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.int64)
        n = a.shape[-2]
        out_a = np.zeros_like(a)
        out_e = np.zeros_like(e)
        _lib.check(_lib.load().pp_evolve(self.ctx, _lib.dp(a), e.ctypes.data_as(_lib.c_i64_p), n, bl, _lib.dp(out_a),
                                         out_e.ctypes.data_as(_lib.c_i64_p)))
        return out_a, out_e

    # ---- the path -------------------------------------------------------------------------
    def prefs(self, pp: bool = False, informative_prior: bool = False, uniform_prior: bool = False,
              prior_lower: float = 0.0, **kw) -> _lib.Prefs:
        """Defaults of prefs.ml:65-120 with the prior chosen as pplacer_run.ml:380-390."""
        p = _lib.pplacer_p_prefs(**kw) if pp else _lib.default_prefs(**kw)
        if uniform_prior:
            p.prior_kind = 0
        elif informative_prior:
            p.prior_kind = 2
            self._prior_means = np.ascontiguousarray(
                np.nan_to_num(informative_prior_means(self.prob.tree, prior_lower), nan=1.0))
            p.prior_means = _lib.dp(self._prior_means)
        else:
            p.prior_kind = 1
            p.prior_mean = flat_exp_prior_mean(self.prob.tree)
        return p

    def place(self, query_chars: np.ndarray, prefs: Optional[_lib.Prefs] = None, debug: bool = False):
        """Core.pplacer_core over a batch: returns (results[Q, keep_at_most], n_out[Q], status[Q], info)."""
        if not self.precomputed:
            self.precompute()
        prefs = prefs or self.prefs()
        q = np.ascontiguousarray(query_chars, dtype=np.uint8)
        if q.ndim != 2 or q.shape[1] != self.L:
            raise ValueError("query is not the same length as the ref alignment")
        Q = q.shape[0]
        K = prefs.keep_at_most
        out = np.zeros((Q, K), dtype=_lib.RESULT_DTYPE)
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
            dbg = _lib.Debug(_lib.dp(info["scores"]), _lib.ip(info["n_pitches"]), _lib.ip(info["pitch_edges"]),
                             _lib.dp(info["pitch_ll"]), _lib.dp(info["pitch_pend"]), _lib.dp(info["pitch_dist"]),
                             info["n_like_calls"].ctypes.data_as(_lib.c_i64_p), _lib.dp(info["timing"]))
        _lib.check(_lib.load().pp_place(self.ctx, C.byref(prefs), Q, _lib.up(q), out.ctypes.data_as(C.c_void_p),
                                        _lib.ip(n_out), _lib.ip(status), None if dbg is None else C.byref(dbg)))
        return out, n_out, status, info

    def place_device(self, d_chars_ptr: int, Q: int, d_out_ptr: int, d_n_out_ptr: int, d_status_ptr: int,
                     prefs: Optional[_lib.Prefs] = None):
        """pp_place_device: everything already resident in HBM (device pointers as ints)."""
        if not self.precomputed:
            self.precompute()
        prefs = prefs or self.prefs()
        _lib.check(_lib.load().pp_place_device(self.ctx, C.byref(prefs), Q, C.c_void_p(d_chars_ptr),
                                               C.c_void_p(d_out_ptr), C.c_void_p(d_n_out_ptr),
                                               C.c_void_p(d_status_ptr)))

    def set_stream(self, cuda_stream: Optional[int]):
        """Run on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        _lib.check(_lib.load().pp_set_stream(self.ctx, C.c_void_p(cuda_stream) if cuda_stream else None))

    def set_prefilter_kernel(self, kernel: int):
        """0 default, 1 direct FMA form, 2 four-base table form (DNA Gamma-4 only)."""
        _lib.check(_lib.load().pp_set_prefilter_kernel(self.ctx, int(kernel)))

    def prefilter_kernel_used(self) -> int:
        return int(_lib.load().pp_prefilter_kernel_used(self.ctx))

    def set_three_tax_engine(self, engine: int, blocks_per_sm: int = 0):
        """0 default, 1 block-per-candidate, 2 warp-per-candidate, 3 block + shared-memory w (2, 3: DNA Gamma-4)."""
        _lib.check(_lib.load().pp_set_three_tax_engine(self.ctx, int(engine), int(blocks_per_sm)))

    def three_tax_engine_used(self) -> int:
        return int(_lib.load().pp_three_tax_engine_used(self.ctx))

    # ---- instrumentation ------------------------------------------------------------------
    def kernel_launches(self, reset: bool = False) -> int:
        return int(_lib.load().pp_kernel_launches(self.ctx, 1 if reset else 0))

    def phase_times_ms(self, reset: bool = False) -> np.ndarray:
        out = np.zeros(4)
        _lib.check(_lib.load().pp_phase_times_ms(self.ctx, _lib.dp(out), 1 if reset else 0))
        return out

    def set_rank_mode(self, mode: int):
        """0 truncated ranking on large trees (default), 1 always the full sort."""
        _lib.check(_lib.load().pp_set_rank_mode(self.ctx, int(mode)))

    def rank_fallbacks(self) -> int:
        return int(_lib.load().pp_rank_fallbacks(self.ctx))

    def three_tax_work(self, reset: bool = False) -> np.ndarray:
        """[pendant calls, distal calls, pendant call-sites, distal call-sites] of the warp engine."""
        out = np.zeros(4)
        _lib.check(_lib.load().pp_three_tax_work(self.ctx, _lib.dp(out), 1 if reset else 0))
        return out

    def three_tax_profile(self, reset: bool = False) -> np.ndarray:
        """pp_three_tax_profile: per-phase cycle counters (library built with -DPPB_TT_PROFILE)."""
        out = np.zeros(32)
        _lib.check(_lib.load().pp_three_tax_profile(self.ctx, _lib.dp(out), 1 if reset else 0))
        return out

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0)
        _lib.check(_lib.load().pp_measure_fp64_peak(self.ctx, C.byref(v)))
        return v.value
