"""The stage-level C-ABI entry points against the reference's OWN linear_c.c (oracle/_ref/liblinear_ref.so, the
unmodified file compiled from /root/reference; the prebuilt .so travels to the GPU box), with NO oracle code in between:
the device GLVs are exported in the reference layout (pp_export_glvs) and fed to the reference's C functions.

  pp_prefilter_scores   == Glv.masked_logdot  -> ten_masked_logdot_c (linear_c.c:465) / mat_masked_logdot_c (:251)
  pp_three_tax_log_like == Three_tax.log_like -> ten_log_like3_c (linear_c.c:367) / mat_log_like3_c (:170)
  pp_evolve             == Model.evolve_into  -> dediagonalize (linear_c.c:92) + gemmish_c (linear_c.c:29)
"""
import numpy as np
import pytest

from pplacer_b200 import configs, seqio
from pplacer_b200.placer import Placer

pytestmark = pytest.mark.gpu

from tests.refc import LN2, RefC, indicator, sites as _sites


PROBLEMS = {
    "c1_gtr": lambda: configs.c1(),
    "wag": lambda: configs.like_problem("wag", seed=11),
    "gcat_dna": lambda: configs.gcat_variant("gtr"),
    "gcat_aa": lambda: configs.gcat_variant("wag"),
}


@pytest.fixture(scope="module")
def ref(oracle_binding):
    R = oracle_binding.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref/liblinear_ref.so not present")
    return oracle_binding, R


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_stage_entry_points_match_reference_c(ref, name):
    ob, R = ref
    prob = PROBLEMS[name]()
    gcat = prob.model.model_class == 1
    with Placer(prob) as p:
        p.precompute()
        rc = RefC(ob, R, p.model_numerics(), gcat, prob.site_cat)
        S = rc.S
        rng = np.random.default_rng(17)
        nodes = rng.choice(np.array(prob.tree.nonroot_ids()), size=6, replace=False)
        Q = min(4, prob.query_chars.shape[0])
        queries = prob.query_chars[:Q]
        start_pend = 0.1
        scores = p.prefilter_scores(queries, start_pend=start_pend)
        for q in range(Q):
            mask = seqio.informative(queries[q])
            qev = rc.evolve(indicator(queries[q], S) if gcat else np.repeat(indicator(queries[q], S)[None], len(rc.num["rates"]), axis=0),
                            start_pend)
            for node in nodes:
                # ---- Glv.masked_logdot (glv.ml) of the evolved query against the supernode -------------------
                sa, se = p.glv(2, int(node))
                want = rc.masked_logdot(np.ascontiguousarray(qev), np.ascontiguousarray(sa), mask) + LN2 * float(se[mask].sum())
                got = scores[q, node]
                assert abs(got - want) <= 1e-10 * abs(want), (name, q, node, got, want)
        # ---- Three_tax.log_like at random branch lengths: the reference gathers the masked sites (core.ml:225-239) ----
        n = 24
        qi = rng.integers(0, Q, size=n)
        edges = rng.choice(nodes, size=n).astype(np.int32)
        pend = rng.uniform(1e-8, 2.0, size=n)
        dist = rng.uniform(0.0, 1.0, size=n) * prob.tree.bl[edges]
        got = p.three_tax_log_like(queries[qi], edges, pend, dist)
        for i in range(n):
            idx = np.nonzero(seqio.informative(queries[qi[i]]))[0]
            da, de = p.glv(0, int(edges[i]))
            pa, pe = p.glv(1, int(edges[i]))
            if gcat:
                sub = RefC(ob, R, rc.num, True, prob.site_cat[idx])
                lv = indicator(queries[qi[i]], S)[idx]
            else:
                sub = rc
                lv = np.repeat(indicator(queries[qi[i]], S)[None], len(rc.num["rates"]), axis=0)[:, idx]
            x = sub.evolve(_sites(pa, idx, gcat), float(prob.tree.bl[edges[i]] - dist[i]))
            y = sub.evolve(_sites(da, idx, gcat), float(dist[i]))
            z = sub.evolve(np.ascontiguousarray(lv), float(pend[i]))
            want = sub.log_like3(x, y, z) + LN2 * float(de[idx].sum() + pe[idx].sum())
            assert abs(got[i] - want) <= 1e-10 * abs(want), (name, i, got[i], want)
        # ---- Model.evolve_into -------------------------------------------------------------------------
        a, e = p.glv(0, int(nodes[0]))
        got_a, got_e = p.evolve(a, e, 0.2371)
        want_a = rc.evolve(np.ascontiguousarray(a), 0.2371)
        assert np.array_equal(got_e, e)
        # P(t) entries carry ~1e-16 ABSOLUTE rounding (device exp / fused multiply-adds vs libm and separate roundings)
        assert np.allclose(got_a, want_a, rtol=1e-10, atol=1e-15 * np.abs(want_a).max())
