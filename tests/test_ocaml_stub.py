"""The OCaml C stub of INTEGRATION.md (integration/pplacer_b200_stubs.c) is real code: it is compiled here with
-Wall -Wextra against the mock caml/ headers of oracle/caml_shim (no OCaml toolchain in the image) and driven through
a small mock runtime (integration/shim_driver.c: builds the bigarray / option / boxed-float values, catches Failure,
runs the custom block's finaliser).  CPU: the Failure path (no CUDA device).  GPU: create -> precompute -> place ->
export_glv end to end, equal to the ctypes binding."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pplacer_b200 import _lib, configs
from pplacer_b200.placer import Placer, abi_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    _lib.load()
    so = str(tmp_path_factory.mktemp("stub") / "libppb_stub_test.so")
    libdir = os.path.join(ROOT, "pplacer_b200")
    cmd = ["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-std=gnu11", "-fPIC", "-shared",
           "-I", os.path.join(ROOT, "oracle", "caml_shim"), "-I", os.path.join(ROOT, "include"), "-o", so,
           os.path.join(ROOT, "integration", "pplacer_b200_stubs.c"), os.path.join(ROOT, "integration", "shim_driver.c"),
           "-L", libdir, "-lpplacer_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(so)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.drv_last_failure.restype = C.c_char_p
    L.drv_create.restype = C.c_ssize_t
    L.drv_create.argtypes = [i, i, vp, vp, vp, C.c_int64, vp, d, i, vp, i, i, i, vp, vp, i, vp, vp, vp, i, vp, i]
    L.drv_precompute.argtypes = [C.c_ssize_t]
    L.drv_place.argtypes = [C.c_ssize_t, vp, vp, i, vp, C.c_int64, i, i, vp, vp, vp, vp]
    L.drv_export_glv.argtypes = [C.c_ssize_t, i, i, vp, C.c_int64, vp, i]
    L.drv_release.argtypes = [C.c_ssize_t]
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _create(L, prob, devices):
    m, t = prob.model, prob.tree
    cm, ct, ref_chars, keep = abi_inputs(prob)
    exch, statd, statd_chars, rates, site_cat, ptr, idx, bl, leaf_row, _ = keep
    dev = np.ascontiguousarray(devices, dtype=np.int32)
    h = L.drv_create(m.n_states, m.model_class, _ptr(exch), _ptr(statd), _ptr(statd_chars),
                     0 if statd_chars is None else statd_chars.size, _ptr(rates), float(m.gamma_alpha), m.n_rates,
                     _ptr(site_cat), prob.L, t.root, t.n_nodes, _ptr(ptr), _ptr(idx), len(idx), _ptr(bl), _ptr(leaf_row),
                     _ptr(ref_chars), ref_chars.shape[0], _ptr(dev), len(dev))
    return h, keep


def test_stub_compiles_and_raises_failure_without_a_gpu(stub):
    import torch
    if torch.cuda.is_available():
        pytest.skip("the Failure path of this test needs a box without a CUDA device")
    h, _ = _create(stub, configs.c1(), [0])
    assert h == 0
    assert b"no CUDA device" in stub.drv_last_failure()


@pytest.mark.gpu
def test_stub_end_to_end_equals_ctypes_binding(stub):
    prob = configs.c1()
    h, keep = _create(stub, prob, [0])
    assert h != 0, stub.drv_last_failure()
    try:
        assert stub.drv_precompute(h) == 0, stub.drv_last_failure()
        with Placer(prob) as p:
            p.precompute()
            for pp in (False, True):
                pf = p.prefs(pp=pp)
                want, n_want, st_want, _ = p.place(prob.query_chars, pf)
                vec = np.array([pf.start_pend, pf.max_pend, pf.initial_tolerance, pf.final_tolerance, pf.max_iter, pf.max_strikes,
                                pf.strike_box, pf.max_pitches, pf.keep_at_most, pf.keep_factor, pf.calc_pp, pf.prior_kind,
                                pf.pp_rel_err, pf.prior_mean], dtype=np.float64)
                Q, K = prob.query_chars.shape[0], pf.keep_at_most
                q = np.ascontiguousarray(prob.query_chars)
                edges = np.zeros(Q * K, np.int32); out = np.zeros(Q * K * 6); n_out = np.zeros(Q, np.int32); st = np.zeros(Q, np.int32)
                rc = stub.drv_place(h, _ptr(vec), None, prob.tree.n_nodes, _ptr(q), Q, prob.L, K, _ptr(edges), _ptr(out), _ptr(n_out), _ptr(st))
                assert rc == 0, stub.drv_last_failure()
                assert np.array_equal(n_out, n_want) and np.array_equal(st, st_want)
                assert np.array_equal(edges.reshape(Q, K), want["edge"])
                o = out.reshape(Q, K, 6)
                for j, f in enumerate(("log_like", "pendant", "distal", "ml_ratio", "marginal_like", "post_prob")):
                    assert np.array_equal(o[:, :, j], want[f], equal_nan=True), f
            a, e = p.glv(2, 3)
            ga, ge = np.zeros(a.size), np.zeros(prob.L, np.int64)
            assert stub.drv_export_glv(h, 2, 3, _ptr(ga), ga.size, _ptr(ge), prob.L) == 0
            assert np.array_equal(ga.reshape(a.shape), a) and np.array_equal(ge, e)
        # a bad argument surfaces as Failure with the library's message
        bad = np.zeros(14)
        rc = stub.drv_place(h, _ptr(bad), None, prob.tree.n_nodes, _ptr(q), Q, prob.L, 1, _ptr(edges), _ptr(out), _ptr(n_out), _ptr(st))
        assert rc == 1 and b"bad prefs" in stub.drv_last_failure()
    finally:
        stub.drv_release(h)
