"""The oracle's inner loops against the reference's OWN linear_c.c (oracle/_ref/liblinear_ref.so, compiled
unmodified from /root/reference by oracle/Makefile), BIT FOR BIT, on random inputs of every shape the path uses
(DNA and protein, Gamma mixtures "ten_*" and Price-CAT "mat_*").  This is what pins the restatement's arithmetic:
same operation order, no FMA contraction, same libm log.

  gemmish_c              linear_c.c:29     dediagonalize           linear_c.c:92
  mat_log_like3_c        linear_c.c:170    mat_pairwise_prod_c     linear_c.c:218
  mat_statd_pairwise..   linear_c.c:232    mat_masked_logdot_c     linear_c.c:251
  mat_bounded_logdot_c   linear_c.c:298    ten_log_like3_c         linear_c.c:367
  ten_pairwise_prod_c    linear_c.c:426    ten_statd_pairwise..    linear_c.c:443
  ten_masked_logdot_c    linear_c.c:465    ten_bounded_logdot_c    linear_c.c:541
"""
import numpy as np
import pytest

SHAPES = [(4, 4, 0, 333), (4, 4, 0, 1), (20, 4, 0, 57), (20, 1, 0, 64), (4, 1, 1, 129), (20, 1, 1, 75), (4, 3, 0, 16), (20, 6, 0, 9)]


@pytest.fixture(scope="module")
def libs(oracle_binding):
    R = oracle_binding.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt .so)")
    return oracle_binding, oracle_binding.lib(), R


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _same(a, b):
    """bit-identical (NaN payloads included)"""
    return np.array_equal(_bits(np.atleast_1d(a)), _bits(np.atleast_1d(b)))


def _glv(rng, R, gcat, n, S, scale=1.0):
    shape = (n, S) if gcat else (R, n, S)
    a = rng.random(shape) * scale
    a[rng.random(shape) < 0.05] = 0.0          # zeros as in leaf vectors
    return np.ascontiguousarray(a)


@pytest.mark.parametrize("S,n", [(4, 1), (4, 257), (20, 1), (20, 123)])
def test_gemmish_bit_exact(libs, S, n):
    ob, O, R = libs
    rng = np.random.default_rng(S * 1000 + n)
    a, b = rng.random((S, S)), rng.random((n, S))
    d1, d2 = np.zeros((n, S)), np.zeros((n, S))
    O.orc_k_gemmish(ob.dp(d1), ob.dp(a), ob.dp(b), S, n)
    R.ref_gemmish(ob.dp(d2), ob.dp(a), ob.dp(b), S, n)
    assert _same(d1, d2)
    assert np.allclose(d1, b @ a.T, rtol=1e-13)


@pytest.mark.parametrize("S", [4, 20])
def test_dediagonalize_bit_exact(libs, S):
    ob, O, R = libs
    rng = np.random.default_rng(S)
    for _ in range(8):
        u, lam, uit = rng.standard_normal((S, S)), np.exp(-rng.random(S) * 3), rng.standard_normal((S, S))
        d1, d2 = np.zeros((S, S)), np.zeros((S, S))
        O.orc_k_dediagonalize(ob.dp(d1), ob.dp(u), ob.dp(lam), ob.dp(uit), S)
        R.ref_dediagonalize(ob.dp(d2), ob.dp(u), ob.dp(lam), ob.dp(uit), S)
        assert _same(d1, d2)
        assert np.allclose(d1, (u * lam) @ uit.T, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("S,Rn,gcat,n", SHAPES)
def test_pairwise_products_bit_exact(libs, S, Rn, gcat, n):
    ob, O, R = libs
    rng = np.random.default_rng(S + 7 * Rn + 13 * n)
    x, y = _glv(rng, Rn, gcat, n, S), _glv(rng, Rn, gcat, n, S)
    statd = rng.random(S); statd /= statd.sum()
    d1, d2 = np.zeros_like(x), np.zeros_like(x)
    O.orc_k_pairwise_prod(ob.dp(d1), ob.dp(x), ob.dp(y), Rn, gcat, n, S)
    if gcat:
        R.ref_mat_pairwise_prod(ob.dp(d2), ob.dp(x), ob.dp(y), n, S)
    else:
        R.ref_ten_pairwise_prod(ob.dp(d2), ob.dp(x), ob.dp(y), Rn, n, S)
    assert _same(d1, d2)
    d1[:] = 0; d2[:] = 0
    O.orc_k_statd_pairwise_prod(ob.dp(statd), ob.dp(d1), ob.dp(x), ob.dp(y), Rn, gcat, n, S)
    if gcat:
        R.ref_mat_statd_pairwise_prod(ob.dp(statd), ob.dp(d2), ob.dp(x), ob.dp(y), n, S)
    else:
        R.ref_ten_statd_pairwise_prod(ob.dp(statd), ob.dp(d2), ob.dp(x), ob.dp(y), Rn, n, S)
    assert _same(d1, d2)


@pytest.mark.parametrize("S,Rn,gcat,n", SHAPES)
def test_log_like3_bit_exact(libs, S, Rn, gcat, n):
    ob, O, R = libs
    rng = np.random.default_rng(3 * S + Rn + n)
    for scale in (1.0, 1e-40, 1e-150):          # small mantissas as after many exponent pulls
        x, y, z = (_glv(rng, Rn, gcat, n, S, scale) + 1e-3 * scale for _ in range(3))
        statd = rng.random(S); statd /= statd.sum()
        a = O.orc_k_log_like3(ob.dp(statd), ob.dp(x), ob.dp(y), ob.dp(z), Rn, gcat, n, S)
        if gcat:
            b = R.ref_mat_log_like3(ob.dp(statd), ob.dp(x), ob.dp(y), ob.dp(z), n, S)
        else:
            util = np.zeros(n)
            b = R.ref_ten_log_like3(ob.dp(statd), ob.dp(x), ob.dp(y), ob.dp(z), ob.dp(util), Rn, n, S)
        assert np.isfinite(b) or scale < 1e-100
        assert _same(a, b), (a, b)
    # a site whose product is exactly zero: log 0 = -inf propagates identically (linear_c.c:418)
    x[..., 0, :] = 0.0
    a = O.orc_k_log_like3(ob.dp(statd), ob.dp(x), ob.dp(y), ob.dp(z), Rn, gcat, n, S)
    if gcat:
        b = R.ref_mat_log_like3(ob.dp(statd), ob.dp(x), ob.dp(y), ob.dp(z), n, S)
    else:
        util = np.zeros(n)
        b = R.ref_ten_log_like3(ob.dp(statd), ob.dp(x), ob.dp(y), ob.dp(z), ob.dp(util), Rn, n, S)
    assert a == -np.inf and _same(a, b)


@pytest.mark.parametrize("S,Rn,gcat,n", SHAPES)
def test_logdots_bit_exact(libs, S, Rn, gcat, n):
    ob, O, R = libs
    rng = np.random.default_rng(5 * S + Rn + 3 * n)
    x, y = _glv(rng, Rn, gcat, n, S) + 1e-6, _glv(rng, Rn, gcat, n, S) + 1e-6
    for frac in (1.0, 0.6, 0.0):
        mask = (rng.random(n) < frac).astype(np.uint16)
        a = O.orc_k_masked_logdot(ob.dp(x), ob.dp(y), mask.ctypes.data_as(ob.c_u16_p), Rn, gcat, n, S)
        if gcat:
            b = R.ref_mat_masked_logdot(ob.dp(x), ob.dp(y), mask.ctypes.data_as(ob.c_u16_p), n, S)
        else:
            util = np.zeros(n)
            b = R.ref_ten_masked_logdot(ob.dp(x), ob.dp(y), mask.ctypes.data_as(ob.c_u16_p), ob.dp(util), Rn, n, S)
        assert _same(a, b), (frac, a, b)
    # the unmasked log-dot (Glv.logdot -> bounded_logdot over every site, glv.ml / linear_c.c:298,541)
    a = O.orc_k_logdot(ob.dp(x), ob.dp(y), Rn, gcat, n, S)
    if gcat:
        b = R.ref_mat_bounded_logdot(ob.dp(x), ob.dp(y), 0, n - 1, n, S)
    else:
        util = np.zeros(n)
        b = R.ref_ten_bounded_logdot(ob.dp(x), ob.dp(y), 0, n - 1, ob.dp(util), Rn, n, S)
    assert _same(a, b), (a, b)


# ---------------------------------------------------------------------------------------------
# stage level: the oracle's prefilter score and three-taxon likelihood, on its own precomputed GLVs, against the
# same quantities assembled from the reference's C functions only (tests/refc.py) - to rounding of the final sums
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_gtr", "wag", "gcat_dna", "gcat_aa"])
def test_oracle_stages_match_reference_c(libs, name):
    from pplacer_b200 import configs, reference, seqio
    from tests.refc import LN2, RefC, indicator, sites
    ob, O, R = libs
    prob = {"c1_gtr": configs.c1, "wag": lambda: configs.like_problem("wag", seed=11),
            "gcat_dna": lambda: configs.gcat_variant("gtr"), "gcat_aa": lambda: configs.gcat_variant("wag")}[name]()
    gcat = prob.model.model_class == 1
    o = ob.Oracle(prob)
    o.precompute()
    num = dict(statd=o.statd, x=o.x, lam=o.lam, xit=o.xit, rates=o.rates)
    rc = RefC(ob, R, num, gcat, prob.site_cat)
    S = rc.S
    rng = np.random.default_rng(17)
    nodes = rng.choice(np.array(prob.tree.nonroot_ids()), size=6, replace=False)
    queries = prob.query_chars[:4]
    _, _, _, info = o.place(queries, ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree)), debug=True)
    for q in range(len(queries)):
        mask = seqio.informative(queries[q])
        lv = indicator(queries[q], S)
        qev = rc.evolve(lv if gcat else np.repeat(lv[None], len(num["rates"]), axis=0), 0.1)   # start_pend 0.1 (prefs.ml)
        for node in nodes:
            sa, se = o.glv(2, int(node))
            want = rc.masked_logdot(np.ascontiguousarray(qev), np.ascontiguousarray(sa), mask) + LN2 * float(se[mask].sum())
            assert abs(info["scores"][q, node] - want) <= 4e-16 * abs(want), (q, node)
    for i in range(16):
        q = int(rng.integers(0, len(queries)))
        edge = int(rng.choice(nodes))
        pend, dist = float(rng.uniform(1e-8, 2.0)), float(rng.uniform(0, 1) * prob.tree.bl[edge])
        idx = np.nonzero(seqio.informative(queries[q]))[0]
        da, de = o.glv(0, edge)
        pa, pe = o.glv(1, edge)
        sub = RefC(ob, R, num, True, prob.site_cat[idx]) if gcat else rc
        lv = indicator(queries[q], S)
        lv = lv[idx] if gcat else np.repeat(lv[None], len(num["rates"]), axis=0)[:, idx]
        x = sub.evolve(sites(pa, idx, gcat), float(prob.tree.bl[edge] - dist))
        y = sub.evolve(sites(da, idx, gcat), dist)
        z = sub.evolve(np.ascontiguousarray(lv), pend)
        want = sub.log_like3(x, y, z) + LN2 * float(de[idx].sum() + pe[idx].sum())
        got = o.three_tax_ll(queries[q], edge, pend, dist)
        assert abs(got - want) <= 4e-16 * abs(want), (i, got, want)
    o.close()
