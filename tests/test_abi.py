"""The C-ABI shared library loads without a GPU, exports every symbol include/*.h declares, and
refuses to compute without a CUDA device (no CPU fallback).  CPU only."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from pplacer_b200 import _lib, configs
from pplacer_b200.placer import Placer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported():
    L = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "libpplacer_b200.so does not export " + s
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms
    assert L.pp_abi_version() == 1


def test_struct_layouts_match_header():
    """ctypes mirrors of the header structs have the sizes the C compiler gives them."""
    import subprocess, tempfile
    src = ('#include <stdio.h>\n#include "pplacer_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
           'sizeof(pp_prefs),sizeof(pp_model),sizeof(pp_tree),sizeof(pp_result),sizeof(pp_debug));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")],
                       check=True)
        sizes = [int(v) for v in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.Prefs), C.sizeof(_lib.Model), C.sizeof(_lib.Tree), _lib.RESULT_DTYPE.itemsize,
                     C.sizeof(_lib.Debug)]
    p = _lib.default_prefs()
    assert (p.start_pend, p.max_pend, p.initial_tolerance, p.final_tolerance) == (0.1, 2.0, 0.01, 1e-5)
    assert (p.max_iter, p.max_strikes, p.max_pitches, p.keep_at_most) == (200, 6, 40, 7)
    assert (p.strike_box, p.keep_factor, p.calc_pp, p.pp_rel_err) == (3.0, 0.01, 0, 0.01)
    pp = _lib.pplacer_p_prefs()     # prefs.ml:198-203
    assert (pp.calc_pp, pp.keep_at_most, pp.keep_factor, pp.max_strikes) == (1, 20, 0.001, 20)


def test_host_numerics_match_oracle(oracle_binding):
    """The library's host-side model numerics (a2, a3) against the oracle."""
    assert np.allclose(_lib.discrete_gamma(4, 0.5), oracle_binding.discrete_gamma(4, 0.5), rtol=1e-12)
    assert np.allclose(_lib.discrete_gamma(8, 1.7), oracle_binding.discrete_gamma(8, 1.7), rtol=1e-12)
    prob = configs.c1()
    pi = _lib.emper_freq(4, prob.ref_chars_raw)
    assert np.allclose(pi, oracle_binding.emper_freq(4, prob.ref_chars_raw), rtol=0, atol=1e-15)
    x, lam, xit = _lib.diagd(prob.model.exch, pi)
    ox, olam, oxit = oracle_binding.diagd(prob.model.exch, pi)
    # eigenvectors are defined up to sign/order: compare P(t)
    for t in (0.05, 1.0):
        P = (x * np.exp(t * lam)[None, :]) @ xit.T
        OP = (ox * np.exp(t * olam)[None, :]) @ oxit.T
        assert np.allclose(P, OP, atol=1e-13)


def test_no_cpu_fallback():
    """Without a CUDA device pp_create fails with PP_ENODEVICE; with one it succeeds."""
    import torch
    prob = configs.c1()
    if torch.cuda.is_available():
        with Placer(prob) as p:
            assert p.ctx
        return
    with pytest.raises(_lib.PplacerB200Error) as ei:
        Placer(prob)
    assert ei.value.code == _lib.PP_ENODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_oracle():
    """Nothing under pplacer_b200/ imports, includes, links or calls anything in oracle/."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include[^\n]*oracle|\borc_[a-z]|liboracle|liblinear_ref", re.M)
    for path in glob.glob(os.path.join(ROOT, "pplacer_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
            assert not pat.search(open(path, errors="replace").read()), path


def test_library_is_sm100a_with_tensor_core_and_tma_instructions():
    """The shipped binary is sm_100a SASS and holds what DESIGN.md 3.2 claims for the staged prefilter forms and the
    query-lane table: FP64 tensor-core MMAs (DMMA), TMA bulk copies (UBLKCP = cp.async.bulk) and mbarrier waits (SYNCS)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    lib = _lib.lib_path()
    arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in arch, arch
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout   # ~10 s
    assert "prefilter_staged_kernel" in sass and "prefilter_qlane_kernel" in sass and "tt_eigen_kernel" in sass
    assert "DMMA.8x8x4" in sass and "UBLKCP" in sass and "SYNCS" in sass
