"""Parity tests proper: the CUDA path, called through the C ABI (libpplacer_b200.so), against the
CPU oracle on the same seeded inputs, against the reference's own known answers, and through
size-independent properties.  Tolerances are the north_star's: identical best edge and kept-edge
sets except at documented near-ties, log-likelihood 1e-6 relative, branch lengths 1e-4,
like_weight_ratio / post_prob 1e-5 absolute."""
import json
import os

import numpy as np
import pytest

from pplacer_b200 import _lib, configs, reference, seqio
from pplacer_b200.placer import Placer
from tests.util import assert_placements_match, near_queries

pytestmark = pytest.mark.gpu

PROBLEMS = {
    "c1_gtr": lambda: configs.c1(),
    "wag": lambda: configs.like_problem("wag", seed=11),
    "gcat_dna": lambda: configs.gcat_variant("gtr"),
    "gcat_aa": lambda: configs.gcat_variant("wag"),
    "synth": lambda: configs.synthetic(60, 300, 64, seed=7, min_len=40),
    # a long tree under Gamma(0.5): the slowest rate category underflows to exactly 0 at many sites, and
    # the reference's exponent pull then stops for good (frexp 0. = (0., 0) is the maximum over the
    # rates, gmix_model.ml:172-176) - 12 % of the (edge, site, rate) maxima are 0.  The device must
    # reproduce that, including the prefilter's exact fallback for values outside its fast range.
    "zero_cat": lambda: configs.synthetic(1500, 48, 6, seed=9, alpha=0.5, mean_bl=0.12, min_len=30),
}
_cache = {}
DEFAULT_BIG = 7 if os.environ.get("PPB_PF_PAIRS", "1") == "0" else 8   # prefilter form of DNA Gamma-4 batches of >= 256 queries
PP_KEEP = dict(keep_factor=0.001, keep_at_most=20)     # what -p sets (prefs.ml:198-203)


def setup(name, oracle_binding):
    if name not in _cache:
        prob = PROBLEMS[name]()
        o = oracle_binding.Oracle(prob)
        o.precompute()
        p = Placer(prob)
        p.precompute()
        _cache[name] = (prob, o, p)
    return _cache[name]


def oracle_prefs(ob, prob, **kw):
    return ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree), **kw)


# ---------------------------------------------------------------------------------------------
# known answers of the reference, through the GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gtr", "wag", "jtt"])
def test_phyml_tree_likelihood_at_every_edge(name):
    """tests/pplacer/test_like.ml:57-66 on device: with an all-ones query (N / X everywhere)
    masked_logdot(snode_i) and log_like3(evolve(d_i,bl/2), evolve(p_i,bl/2), 1) are the tree
    log-likelihood PhyML reports, at EVERY edge, to 1e-7 relative."""
    tree, aln, model = configs.like_inputs(name)
    any_char = "N" if model.n_states == 4 else "X"
    L = len(aln[0][1])
    prob = reference.build_problem(tree, model, aln, [("ones", any_char * L)], pre_mask=False)
    want = configs.PHYML_LL[name]
    with Placer(prob) as p:
        p.precompute()
        scores = p.prefilter_scores(prob.query_chars)[0]
        for e in range(tree.n_nodes):
            if e == tree.root:
                assert np.isnan(scores[e])
            else:
                assert abs(scores[e] - want) <= 1e-7 * abs(want), (e, scores[e], want)
        edges = np.array(tree.nonroot_ids(), dtype=np.int32)
        q = np.repeat(prob.query_chars, len(edges), axis=0)
        ll = p.three_tax_log_like(q, edges, np.full(len(edges), 0.37), tree.bl[edges] / 2.0)
        assert np.all(np.abs(ll - want) <= 1e-7 * abs(want)), (ll, want)


# ---------------------------------------------------------------------------------------------
# stage by stage against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_gtr", "wag", "gcat_dna", "gcat_aa", "zero_cat"])
def test_precompute_glvs_match_oracle(oracle_binding, name):
    prob, o, p = setup(name, oracle_binding)
    exp_equal = exp_total = 0
    for which in (0, 1, 2):
        for node in range(prob.tree.n_nodes):
            a, e = p.glv(which, node)
            oa, oe = o.glv(which, node)
            sh = (1, -1, 1) if a.ndim == 3 else (-1, 1)
            va = a * np.exp2(e.reshape(sh).astype(float))
            vo = oa * np.exp2(oe.reshape(sh).astype(float))
            # P(t) = X exp(t Lambda) Xit^T carries ~1e-16 ABSOLUTE rounding per entry (fused vs unfused
            # multiply-adds differ there), i.e. up to ~1e-10 relative on the small off-diagonal entries
            # of short branches - the conditioning of the reference's own arithmetic
            site_max = np.abs(vo).max(axis=(0, 2), keepdims=True) if a.ndim == 3 else np.abs(vo).max(axis=1, keepdims=True)
            assert np.all(np.abs(va - vo) <= 1e-9 * np.abs(vo) + 1e-12 * site_max), (which, node)
            exp_equal += int((e == oe).sum())
            exp_total += e.size
    # exponent pulls happen at the same sites (a boundary flip needs max|a| within an ulp of 2^-k)
    assert exp_equal == exp_total


@pytest.mark.parametrize("name", ["c1_gtr", "wag", "gcat_dna", "gcat_aa", "synth", "zero_cat"])
def test_prefilter_scores_and_ranking_match_oracle(oracle_binding, name):
    prob, o, p = setup(name, oracle_binding)
    _, _, st, info = o.place(prob.query_chars, oracle_prefs(oracle_binding, prob), n_threads=4, debug=True)
    want = info["scores"]
    got = p.prefilter_scores(prob.query_chars)
    root = prob.tree.root
    assert np.isnan(got[:, root]).all()
    keep = [i for i in range(prob.tree.n_nodes) if i != root]
    g, w = got[:, keep], want[:, keep]
    assert np.all(np.abs(g - w) <= 1e-10 * np.abs(w)), np.abs((g - w) / w).max()
    # ranking (fig.ml:143-147): identical order wherever the oracle's scores are not tied to 1e-9
    for q in range(g.shape[0]):
        og, ow = np.argsort(-g[q], kind="stable"), np.argsort(-w[q], kind="stable")
        for a, b in zip(og[:41], ow[:41]):
            assert a == b or abs(w[q][a] - w[q][b]) <= 1e-9 * abs(w[q][a]), (q, a, b)


@pytest.mark.parametrize("name,kernel", [(n, k) for n in ("c1_gtr", "synth", "zero_cat") for k in (1, 2, 3, 4, 5, 6, 7, 8)])
def test_prefilter_kernel_variants_match_oracle(oracle_binding, name, kernel):
    """The direct FMA form (1: evolved-query table in the constant bank, 3: in shared memory), the
    four-base table form (2) and the staged forms (TMA bulk copies + mbarrier ring; 4: FMA, 5: FP64 tensor
    cores, DMMA) of the DNA Gamma-4 prefilter all reproduce ten_masked_logdot_c."""
    prob, o, p = setup(name, oracle_binding)
    _, _, st, info = o.place(prob.query_chars, oracle_prefs(oracle_binding, prob), n_threads=4, debug=True)
    want = info["scores"]
    got = p.prefilter_scores(prob.query_chars, kernel=kernel)
    assert p.prefilter_kernel_used() == {1: 1, 2: 2, 3: 1, 4: 4, 5: 5, 6: 6, 7: 7, 8: 8}[kernel]
    root = prob.tree.root
    keep = [i for i in range(prob.tree.n_nodes) if i != root]
    g, w = got[:, keep], want[:, keep]
    assert np.isnan(got[:, root]).all()
    assert np.all(np.abs(g - w) <= 1e-10 * np.abs(w)), np.abs((g - w) / w).max()


@pytest.mark.parametrize("name", ["wag", "gcat_aa", "gcat_dna"])
def test_generic_query_lane_prefilter_matches_oracle(oracle_binding, name):
    """The generic query-lane log table (S + 2 columns: 20-state Gamma mixture and Price-CAT, DNA Price-CAT), with
    two-state ambiguity codes (exact terms added by the fix kernel), the any-state code X / N (its own column) and gaps
    (uninformative), on more queries than one warp holds: same scores as mat_masked_logdot_c / ten_masked_logdot_c."""
    prob, o, p = setup(name, oracle_binding)
    q = np.concatenate([prob.query_chars] * 5, axis=0)[:70].copy()
    two_a, two_b, any_c = ("R", "Y", "N") if name == "gcat_dna" else ("B", "Z", "X")
    q[1, ::7] = ord(two_a); q[2, 3::11] = ord(two_b); q[3, ::5] = ord(any_c); q[4, 1::9] = ord("-")
    _, _, st, info = o.place(q, oracle_prefs(oracle_binding, prob), n_threads=4, debug=True)
    want = info["scores"]
    got = p.prefilter_scores(q, kernel=7)
    assert p.prefilter_kernel_used() == 7
    direct = p.prefilter_scores(q, kernel=1)
    root = prob.tree.root
    keep = [i for i in range(prob.tree.n_nodes) if i != root]
    assert np.isnan(got[:, root]).all()
    for g in (got[:, keep], direct[:, keep]):
        w = want[:, keep]
        assert np.all(np.abs(g - w) <= 1e-10 * np.abs(w)), np.abs((g - w) / w).max()


def test_prefilter_many_queries_matches_few_query_kernels(oracle_binding):
    """Q > 64 takes the multi-warp tiles (W = 4/8, QT = 16), Q <= 64 the one-warp blocks with the site
    range split over blockIdx.z: same scores for the same queries either way, for every kernel form."""
    prob, o, p = setup("synth", oracle_binding)
    q = prob.query_chars
    big = np.concatenate([q, q, q[:17]], axis=0)
    assert big.shape[0] > 64
    for kernel in (0, 1, 2, 3, 4, 5, 6, 7, 8):
        a = p.prefilter_scores(q, kernel=kernel)
        b = p.prefilter_scores(big, kernel=kernel)
        for lo in (0, q.shape[0]):
            d = np.abs(b[lo:lo + q.shape[0]] - a)
            assert np.all((d <= 1e-12 * np.abs(a)) | np.isnan(a)), (kernel, np.nanmax(d))
        d = np.abs(b[2 * q.shape[0]:] - a[:17])
        assert np.all((d <= 1e-12 * np.abs(a[:17])) | np.isnan(a[:17]))


@pytest.mark.parametrize("name", ["c1_gtr", "wag", "gcat_dna", "gcat_aa"])
def test_three_tax_log_like_matches_oracle(oracle_binding, name):
    prob, o, p = setup(name, oracle_binding)
    rng = np.random.default_rng(3)
    Q = prob.query_chars.shape[0]
    n = 96
    qi = rng.integers(0, Q, size=n)
    edges = rng.choice(np.array(prob.tree.nonroot_ids()), size=n).astype(np.int32)
    pend = rng.uniform(1e-8, 2.0, size=n)
    dist = rng.uniform(0.0, 1.0, size=n) * prob.tree.bl[edges]
    got = p.three_tax_log_like(prob.query_chars[qi], edges, pend, dist)
    want = np.array([o.three_tax_ll(prob.query_chars[qi[i]], int(edges[i]), float(pend[i]), float(dist[i])) for i in range(n)])
    assert np.all(np.abs(got - want) <= 1e-11 * np.abs(want)), np.abs((got - want) / want).max()


@pytest.mark.parametrize("name", ["c1_gtr", "gcat_aa"])
def test_evolve_matches_oracle(oracle_binding, name):
    """pp_evolve (Model.evolve_into) against the oracle's multi_exp + the reference's own gemmish_c."""
    prob, o, p = setup(name, oracle_binding)
    a, e = o.glv(0, prob.tree.children[prob.tree.root][0])
    bl = 0.2371
    got_a, got_e = p.evolve(a, e, bl)
    assert np.array_equal(got_e, e)
    P = oracle_binding.multi_exp(o.x, o.lam, o.xit, o.rates, bl)
    if prob.model.model_class == 0:
        want = np.einsum("rij,rlj->rli", P, a)
    else:
        want = np.einsum("lij,lj->li", P[prob.site_cat], a)
    assert np.allclose(got_a, want, rtol=1e-12, atol=1e-300)


# ---------------------------------------------------------------------------------------------
# the path end to end
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_gtr", "wag", "gcat_dna", "gcat_aa", "synth", "zero_cat"])
def test_place_matches_oracle(oracle_binding, name):
    prob, o, p = setup(name, oracle_binding)
    exp, n_exp, st_exp, oinfo = o.place(prob.query_chars, oracle_prefs(oracle_binding, prob), n_threads=4, debug=True)
    got, n_got, st_got, info = p.place(prob.query_chars, p.prefs(), debug=True)
    near = assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)
    assert len(near_queries(near)) <= max(1, len(n_exp) // 50), near
    # baseball visits the same number of locations (same stop rules on the same ranking)
    same = (info["n_pitches"] == oinfo["n_pitches"]).mean()
    assert same >= 0.98, (info["n_pitches"], oinfo["n_pitches"])
    # objective-evaluation counts agree closely (identical Brent state machines; FMA-level
    # differences in f can change a trajectory only at exact ties)
    rel = np.abs(info["n_like_calls"] - oinfo["n_like_calls"]) / np.maximum(oinfo["n_like_calls"], 1)
    assert np.median(rel) <= 0.02, (info["n_like_calls"][:8], oinfo["n_like_calls"][:8])


@pytest.mark.parametrize("name,pp", [("c1_gtr", False), ("synth", False), ("c1_gtr", True)])
def test_three_tax_engines_agree_and_match_oracle(oracle_binding, name, pp):
    """The DNA Gamma-4 second stage has four engines: 1 block-per-candidate (one objective value per
    pass), 2 warp-per-candidate (values at points known in advance computed several per pass and
    memoised; resumable Brent state machine), 3 block-per-candidate with the cached product in
    shared memory, 4 (the default) block-per-candidate with the pendant objective in the eigenbasis
    of the rate matrix.  All restate the same Brent / QNG call sequence, so they must give the same
    placements, the same number of objective calls, and each must match the oracle."""
    prob, o, p = setup(name, oracle_binding)
    sub = prob.query_chars[:12] if pp else prob.query_chars
    kw = dict(calc_pp=1, keep_at_most=20, keep_factor=0.001, max_strikes=20) if pp else {}
    exp, n_exp, st_exp, oinfo = o.place(sub, oracle_prefs(oracle_binding, prob, **kw), n_threads=4, debug=True)
    res = {}
    try:
        assert p.three_tax_engine_used() == 4
        for eng in (1, 2, 3, 4):
            p.set_three_tax_engine(eng)
            assert p.three_tax_engine_used() == eng
            got, n_got, st_got, info = p.place(sub, p.prefs(pp=pp), debug=True)
            near = assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, pp=pp, near_tie_ll=1e-3, **PP_KEEP if pp else {})
            assert len(near_queries(near)) <= max(1, len(n_exp) // 50), near
            res[eng] = (got, n_got, st_got, info)
    finally:
        p.set_three_tax_engine(0)
    g1, n1, s1, i1 = res[1]
    for eng in (2, 3, 4):
        g2, n2, s2, i2 = res[eng]
        assert np.array_equal(n1, n2) and np.array_equal(s1, s2)
        assert_placements_match(g2, n2, s2, g1, n1, s1, pp=pp, **PP_KEEP if pp else {})
        assert np.array_equal(i1["n_pitches"], i2["n_pitches"])
        rel = np.abs(i1["n_like_calls"] - i2["n_like_calls"]) / np.maximum(i1["n_like_calls"], 1)
        assert np.median(rel) <= 0.02 and rel.max() <= 0.25, (eng, i1["n_like_calls"][:8], i2["n_like_calls"][:8])


@pytest.mark.parametrize("name", ["c1_gtr", "gcat_aa"])
def test_place_posterior_probability_matches_oracle(oracle_binding, name):
    """-p: marginal likelihood by nested QNG (three_tax.ml:169-215), keep_at_most 20,
    keep_factor 0.001, max_strikes 20 (prefs.ml:198-203).  Parity UNPINNED in the reference - the
    oracle restatement of GSL's qng is the arbiter."""
    prob, o, p = setup(name, oracle_binding)
    op = oracle_prefs(oracle_binding, prob, calc_pp=1, keep_at_most=20, keep_factor=0.001, max_strikes=20)
    sub = prob.query_chars[:12]
    exp, n_exp, st_exp, _ = o.place(sub, op, n_threads=4)
    got, n_got, st_got, _ = p.place(sub, p.prefs(pp=True))
    near = assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, pp=True, near_tie_ll=1e-3, **PP_KEEP)
    assert len(near_queries(near)) <= 1
    for q in range(len(n_got)):
        r = got[q][:n_got[q]]
        assert abs(r["post_prob"].sum() - 1.0) < 1e-9


def test_place_informative_and_uniform_priors(oracle_binding):
    prob, o, p = setup("c1_gtr", oracle_binding)
    sub = prob.query_chars[:6]
    means = np.ascontiguousarray(np.nan_to_num(reference.informative_prior_means(prob.tree), nan=1.0))
    for kind in (0, 2):
        op = oracle_prefs(oracle_binding, prob, calc_pp=1, keep_at_most=20, keep_factor=0.001, max_strikes=20, prior_kind=kind)
        if kind == 2:
            op.prior_means = oracle_binding.dp(means)
        exp, n_exp, st_exp, _ = o.place(sub, op, n_threads=2)
        got, n_got, st_got, _ = p.place(sub, p.prefs(pp=True, uniform_prior=(kind == 0), informative_prior=(kind == 2)))
        assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, pp=True, near_tie_ll=1e-3, **PP_KEEP)


def test_edge_cases(oracle_binding):
    """Ragged and degenerate inputs: one informative site, all-N query, no
    informative site (core.ml:157 -> status 103), unknown character (-> 105), keep_at_most 1,
    max_strikes 0 (= optimise every edge, core.ml:100-102), Q = 1."""
    prob, o, p = setup("c1_gtr", oracle_binding)
    L = prob.L
    base = prob.query_chars[0].copy()
    one = np.full(L, ord("-"), np.uint8); one[L // 2] = base[L // 2] if base[L // 2] != ord("-") else ord("A")
    alln = np.full(L, ord("N"), np.uint8)
    gaps = np.full(L, ord("-"), np.uint8)
    bad = base.copy(); bad[3] = ord("J")
    amb = base.copy(); amb[::5] = ord("R"); amb[1::7] = ord("?")
    qs = np.stack([one, alln, gaps, bad, amb, base])
    op = oracle_prefs(oracle_binding, prob)
    exp, n_exp, st_exp, _ = o.place(qs, op)
    got, n_got, st_got, _ = p.place(qs, p.prefs())
    assert list(st_exp) == [0, 0, 103, 105, 0, 0]
    assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)
    # keep_at_most 1 and a single query
    exp, n_exp, st_exp, _ = o.place(qs[5:6], oracle_prefs(oracle_binding, prob, keep_at_most=1))
    got, n_got, st_got, _ = p.place(qs[5:6], p.prefs(keep_at_most=1))
    assert n_got[0] == 1 and got[0][0]["ml_ratio"] == 1.0
    assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, keep_at_most=1)
    # max_strikes 0: every edge is optimised
    sub = prob.query_chars[:4]
    exp, n_exp, st_exp, oi = o.place(sub, oracle_prefs(oracle_binding, prob, max_strikes=0), debug=True)
    got, n_got, st_got, gi = p.place(sub, p.prefs(max_strikes=0), debug=True)
    assert np.all(gi["n_pitches"] == prob.tree.n_nodes - 1) and np.array_equal(gi["n_pitches"], oi["n_pitches"])
    assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)
    # Q = 0 is a no-op
    got, n_got, st_got, _ = p.place(np.zeros((0, L), np.uint8), p.prefs())
    assert got.shape[0] == 0


def test_api_errors():
    prob = configs.c1()
    with Placer(prob) as p:
        with pytest.raises(_lib.PplacerB200Error) as ei:
            p.prefilter_scores(prob.query_chars[:1])      # before precompute
        assert ei.value.code == _lib.PP_ESTATE
        with pytest.raises(ValueError, match="not the same length"):
            p.place(prob.query_chars[:, :-1])
        p.precompute()
        with pytest.raises(_lib.PplacerB200Error) as ei:
            p.place(prob.query_chars[:1], p.prefs(keep_at_most=0))
        assert ei.value.code == _lib.PP_EINVAL


# ---------------------------------------------------------------------------------------------
# size-independent properties at the bench shape (C2: 800 taxa x 1500 columns)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2_small():
    prob = configs.c2(Q=1536)
    p = Placer(prob)
    p.precompute()
    yield prob, p
    p.close()


def test_c2_properties(c2_small, oracle_binding):
    prob, p = c2_small
    Q = prob.query_chars.shape[0]
    out, n_out, st, info = p.place(prob.query_chars, p.prefs(), debug=True)
    assert (st == 0).all() and (n_out >= 1).all() and (n_out <= 7).all()
    for q in range(Q):
        r = out[q][:n_out[q]]
        assert abs(r["ml_ratio"].sum() - 1.0) < 1e-9                     # weights normalised
        assert np.all(np.diff(r["ml_ratio"]) <= 0)                        # sorted descending
        assert r["ml_ratio"][0] >= 1.0 / 7 - 1e-12
        assert len(set(r["edge"].tolist())) == n_out[q]
        assert np.all((r["pendant"] >= 1e-8) & (r["pendant"] <= 2.0))
        assert np.all((r["distal"] >= 0) & (r["distal"] <= prob.tree.bl[r["edge"]] + 1e-12))
        # the best placement's likelihood is at least the best prefilter-ranked start
        assert np.isfinite(r["log_like"]).all()
    # idempotent / deterministic: a second pass gives the same bytes
    out2, n2, st2, _ = p.place(prob.query_chars, p.prefs())
    assert out.tobytes() == out2.tobytes() and np.array_equal(n_out, n2)
    # order- and batch-independent: a permuted subset gives the same per-query records
    rng = np.random.default_rng(1)
    perm = rng.permutation(Q)[:300]
    out3, n3, st3, _ = p.place(prob.query_chars[perm], p.prefs())
    assert out3.tobytes() == out[perm].tobytes() and np.array_equal(n3, n_out[perm])
    # oracle spot check on a sample
    o = oracle_binding.Oracle(prob)
    o.precompute()
    idx = perm[:256]
    exp, n_exp, st_exp, _ = o.place(prob.query_chars[idx], oracle_prefs(oracle_binding, prob), n_threads=8)
    near = assert_placements_match(out[idx], n_out[idx], st[idx], exp, n_exp, st_exp, near_tie_ll=1e-3)
    assert len(near_queries(near)) <= max(2, len(idx) // 50), near
    o.close()


def test_c5_shape_posterior_probability_matches_oracle(c2_small, oracle_binding):
    """BASELINE config 5 at its own shape: the C2 tree (800 taxa x 1500 columns) with -p --keep-at-most 40
    (calc_pp, keep_factor 0.001, max_strikes 20; prefs.ml:198-203): nested QNG integrals over up to 40 kept
    locations per query, against the oracle."""
    prob, p = c2_small
    q = prob.query_chars[:32]
    kw = dict(calc_pp=1, keep_at_most=40, keep_factor=0.001, max_strikes=20)
    o = oracle_binding.Oracle(prob)
    o.precompute()
    exp, n_exp, st_exp, oi = o.place(q, oracle_prefs(oracle_binding, prob, **kw), n_threads=os.cpu_count(), debug=True)
    o.close()
    got, n_got, st_got, gi = p.place(q, p.prefs(pp=True, keep_at_most=40), debug=True)
    assert n_exp.max() >= 4 and (n_exp > 1).mean() > 0.4   # several kept locations per query: the integrals run for each
    near = assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, pp=True, near_tie_ll=1e-3, keep_factor=0.001, keep_at_most=40)
    assert len(near_queries(near)) <= 1, near
    assert (gi["n_pitches"] == oi["n_pitches"]).mean() >= 0.95
    for i in range(len(n_got)):
        r = got[i][:n_got[i]]
        assert abs(r["post_prob"].sum() - 1.0) < 1e-9 and abs(r["ml_ratio"].sum() - 1.0) < 1e-9


@pytest.mark.slow
def test_c4_shape_matches_oracle(oracle_binding):
    """BASELINE config 4 at its own shape: the 10k-taxon tree (19 998 nodes, mean branch length 0.002), where
    the four-base prefilter table (960 MB) streams from HBM, the edge dimension is tiled (Np) and the ranking is
    the truncated one.  Prefilter scores over EVERY edge, the head of the ranking and the placements of 32
    queries against the oracle (its precompute is 11.5 GB of host memory and ~25 s)."""
    prob = configs.c4(Q=32)
    assert prob.tree.n_nodes > 19_000
    o = oracle_binding.Oracle(prob)
    o.precompute()
    exp, n_exp, st_exp, oi = o.place(prob.query_chars, oracle_prefs(oracle_binding, prob), n_threads=os.cpu_count(), debug=True)
    o.close()
    with Placer(prob) as p:
        p.precompute()
        got, n_got, st_got, gi = p.place(prob.query_chars, p.prefs(), debug=True)
        assert p.rank_fallbacks() == 0
        direct = p.prefilter_scores(prob.query_chars[:8], kernel=1)
        staged = [p.prefilter_scores(prob.query_chars[:8], kernel=k) for k in (6, 4, 5)]
        qlane = p.prefilter_scores(prob.query_chars, kernel=7)   # the forms large batches use (313 edge tiles, 720 MB / 2.2 GB tables)
        assert p.prefilter_kernel_used() == 7
        qpair = p.prefilter_scores(prob.query_chars, kernel=8)
        assert p.prefilter_kernel_used() == 8
    root = prob.tree.root
    keep = np.array([i for i in range(prob.tree.n_nodes) if i != root])
    g, w = gi["scores"][:, keep], oi["scores"][:, keep]
    assert np.isfinite(w).all()
    assert np.all(np.abs(g - w) <= 1e-10 * np.abs(w)), np.abs((g - w) / w).max()
    for t in (qlane, qpair):
        assert np.isnan(t[:, root]).all()
        assert np.all(np.abs(t[:, keep] - w) <= 1e-10 * np.abs(w)), np.abs((t[:, keep] - w) / w).max()
    for d in [direct] + staged:
        d = d[:, keep]
        assert np.all(np.abs(d - w[:8]) <= 1e-10 * np.abs(w[:8])), np.abs((d - w[:8]) / w[:8]).max()
    for q in range(g.shape[0]):
        og, ow = np.argsort(-g[q], kind="stable"), np.argsort(-w[q], kind="stable")
        for a, b in zip(og[:41], ow[:41]):
            assert a == b or abs(w[q][a] - w[q][b]) <= 1e-9 * abs(w[q][a]), (q, a, b)
    near = assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)
    assert len(near_queries(near)) <= 1, near
    assert (gi["n_pitches"] == oi["n_pitches"]).mean() >= 0.95


def test_c2_prefilter_forms_agree(c2_small):
    """At the bench shape (1598 nodes = 25 edge tiles, windows of 250-1500 sites, 300 queries = multi-warp blocks) every
    prefilter form gives the same scores: product table (default), log table, direct FMA, staged FMA (TMA ring), staged DMMA."""
    prob, p = c2_small
    q = prob.query_chars[:300]
    ref = p.prefilter_scores(q, kernel=1)
    keep = np.array([i for i in range(prob.tree.n_nodes) if i != prob.tree.root])
    assert np.isfinite(ref[:, keep]).all() and np.isnan(ref[:, prob.tree.root]).all()
    assert p.prefilter_kernel_used() == 1
    for kernel in (0, 2, 6, 7, 8, 4, 5):
        got = p.prefilter_scores(q, kernel=kernel)
        assert p.prefilter_kernel_used() == (DEFAULT_BIG if kernel == 0 else kernel)   # >= 256 queries: a query-lane log table is the default
        assert np.isnan(got[:, prob.tree.root]).all()
        d = np.abs(got[:, keep] - ref[:, keep])
        assert np.all(d <= 1e-11 * np.abs(ref[:, keep])), (kernel, float((d / np.abs(ref[:, keep])).max()))


def test_device_resident_entry_point_matches_host_entry_point(c2_small):
    import torch
    prob, p = c2_small
    q = prob.query_chars[:256]
    out, n_out, st, _ = p.place(q, p.prefs())
    d_chars = torch.from_numpy(q).cuda()
    d_out = torch.zeros(q.shape[0] * 7 * 56, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(q.shape[0], dtype=torch.int32, device="cuda")
    d_st = torch.zeros(q.shape[0], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    p.place_device(d_chars.data_ptr(), q.shape[0], d_out.data_ptr(), d_n.data_ptr(), d_st.data_ptr(), p.prefs())
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == out.tobytes()
    assert np.array_equal(d_n.cpu().numpy(), n_out) and np.array_equal(d_st.cpu().numpy(), st)


# ---------------------------------------------------------------------------------------------
# C3: the reference's only end-to-end golden (donovan-test ar53, WAG + Price-CAT 20)
# ---------------------------------------------------------------------------------------------
@pytest.mark.slow
def test_ar53_golden_placements():
    import torch
    if not configs.ar53_available():
        pytest.skip("tests/golden/_ar53 fixture not present (tests/golden/make_fixtures.py --ar53)")
    if torch.cuda.get_device_properties(0).total_memory < 80e9:
        pytest.skip("needs ~58 GB of HBM")
    prob = configs.ar53()
    assert prob.L == 8035 and prob.tree.n_nodes == 13934
    gold = configs.ar53_golden()["GOLDEN"]
    fields = gold["fields"]
    want = {pq["nm"][0][0]: dict(zip(fields, pq["p"][0])) for pq in gold["placements"]}
    with Placer(prob) as p:
        p.precompute()
        out, n_out, st, info = p.place(prob.query_chars, p.prefs(), debug=True)
    assert (st == 0).all()
    for q, names in enumerate(prob.query_names):
        g = want[names[0]]
        assert n_out[q] == 1
        r = out[q][0]
        assert int(r["edge"]) == g["edge_num"], (names, r, g)
        assert abs(r["log_like"] - g["likelihood"]) <= 1e-6 * abs(g["likelihood"])
        assert abs(r["pendant"] - g["pendant_length"]) <= 1e-4
        assert abs(r["distal"] - g["distal_length"]) <= 1e-4
        assert abs(r["ml_ratio"] - g["like_weight_ratio"]) <= 1e-5


# ---------------------------------------------------------------------------------------------
# C1 end to end: refpkg + query FASTA -> jplace (run_file), compared with the oracle under the
# north_star tolerances through the jplace comparator
# ---------------------------------------------------------------------------------------------
def test_c1_refpkg_to_jplace_matches_oracle(oracle_binding, tmp_path):
    import shutil
    from pplacer_b200 import jplace, run, synth, models
    src = os.path.join(configs.GOLDEN, "like", "gtr")
    rp = tmp_path / "hiv.refpkg"
    rp.mkdir()
    for f in ("hiv.fasta", "hiv.phy_phyml_tree.txt", "phylo_model.jplace"):
        shutil.copyfile(os.path.join(src, f), rp / f)
    (rp / "CONTENTS.json").write_text(json.dumps({"files": {"aln_fasta": "hiv.fasta", "tree": "hiv.phy_phyml_tree.txt",
                                                             "phylo_model": "phylo_model.jplace"}}))
    tree, aln, model = configs.like_inputs("gtr")
    queries = synth.perturbed_queries(aln, models.NUC_CODE, 1)
    queries.append(("dup_of_first", queries[0][1].lower()))          # de-duplicated and upper-cased
    qf = tmp_path / "q.fasta"
    qf.write_text("".join(">%s\n%s\n" % (n, s) for n, s in queries))
    for pp in (False, True):
        out_path = str(tmp_path / ("out_pp.jplace" if pp else "out.jplace"))
        path, prob, (res, n_out, st) = run.run_file(str(rp), str(qf), out_path, pp=pp)
        assert (st == 0).all()
        tree_str, fields, got = jplace.read_jplace(path)
        assert tree_str == prob.tree.to_newick(with_node_numbers=True)
        assert fields == (jplace.PP_FIELDS if pp else jplace.BASE_FIELDS)
        assert ["q_%s_0" % aln[0][0], "dup_of_first"] in [pq.names for pq in got]
        o = oracle_binding.Oracle(prob)
        kw = dict(calc_pp=1, keep_at_most=20, keep_factor=0.001, max_strikes=20) if pp else {}
        exp, n_exp, st_exp, _ = o.place(prob.query_chars, oracle_prefs(oracle_binding, prob, **kw), n_threads=4)
        want = jplace.pqueries_from_results(prob.query_names, exp, n_exp, st_exp, pp=pp)
        rep = jplace.compare(got, want, keep_factor=0.001 if pp else 0.01)
        assert rep.ok and rep.n_queries == len(prob.query_names), rep.summary() + " " + "; ".join(rep.failures[:5])
        o.close()


def test_second_stage_work_counters(c2_small):
    """pp_three_tax_work: the objective calls the second stage made, split by kind - the unit of its
    roofline in bench.py - must add up to the per-query call counts pp_place reports."""
    prob, p = c2_small
    q = prob.query_chars[:200]
    p.three_tax_work(reset=True)
    out, n_out, st, info = p.place(q, p.prefs(), debug=True)
    w = p.three_tax_work()
    assert w[0] > 0 and w[1] > 0 and w[0] + w[1] == info["n_like_calls"].sum()
    nq = seqio.informative(q).sum(axis=1)
    assert nq.min() * (w[0] + w[1]) <= w[2] + w[3] <= nq.max() * (w[0] + w[1])


def test_truncated_ranking_equals_full_sort_and_falls_back(c2_small):
    """Large trees rank with topk_rank_kernel (first 256 entries of Fig.enum_by_score's order); the
    placements must be byte-identical to the full segmented sort, also when a query runs through more
    than 256 locations and the batch is redone on the full order."""
    prob, p = c2_small                                  # 1598 nodes > 1024: the truncated ranking is used
    q = prob.query_chars[:96]
    try:
        p.set_rank_mode(1)
        full, n_full, st_full, _ = p.place(q, p.prefs())
        wide_full, nw_full, _, iw_full = p.place(q[:6], p.prefs(max_pitches=300, max_strikes=300, strike_box=1e9), debug=True)
        p.set_rank_mode(0)
        before = p.rank_fallbacks()
        got, n_got, st_got, _ = p.place(q, p.prefs())
        assert p.rank_fallbacks() == before             # defaults never leave the first 256 ranks
        assert got.tobytes() == full.tobytes() and np.array_equal(n_got, n_full) and np.array_equal(st_got, st_full)
        wide, nw, _, iw = p.place(q[:6], p.prefs(max_pitches=300, max_strikes=300, strike_box=1e9), debug=True)
        assert p.rank_fallbacks() == before + 1         # 301 pitches per query: fallback to the full order
        assert (iw["n_pitches"] == 301).all() and np.array_equal(iw["n_pitches"], iw_full["n_pitches"])
        assert wide.tobytes() == wide_full.tobytes() and np.array_equal(nw, nw_full)
    finally:
        p.set_rank_mode(0)


# ---------------------------------------------------------------------------------------------
# all GPUs of the box behind one handle (pp_create_multi / pp_multi_place)
# ---------------------------------------------------------------------------------------------
def test_multi_device_handle_equals_single_device(c2_small):
    """pp_multi_place shards the queries over its contexts from a shared counter, one host thread per device; the records
    must equal pp_place on one device byte for byte.  On a single-GPU box both contexts live on device 0 (still two
    contexts, two streams, two host threads); with >= 2 GPUs they are on different devices."""
    import torch
    from pplacer_b200.placer import MultiPlacer
    prob, p = c2_small
    q = prob.query_chars[:700]
    want, n_want, st_want, _ = p.place(q, p.prefs())
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    with MultiPlacer(prob, devs) as mp:
        mp.precompute()
        got, n_got, st_got = mp.place(q, mp.prefs())
        counts = mp.last_counts()
        assert counts.sum() == len(q) and (counts > 0).all(), counts
        assert got.tobytes() == want.tobytes() and np.array_equal(n_got, n_want) and np.array_equal(st_got, st_want)
        # ragged: fewer queries than devices, and an empty batch
        g1, n1, s1 = mp.place(q[:1], mp.prefs())
        assert g1.tobytes() == want[:1].tobytes()
        g0, n0, s0 = mp.place(q[:0], mp.prefs())
        assert g0.shape[0] == 0
    with pytest.raises(_lib.PplacerB200Error):
        MultiPlacer(prob, [0, 99])


# ---------------------------------------------------------------------------------------------
# Gcat_model.refine and --check-like on the device
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gcat_dna", "gcat_aa"])
def test_gcat_refine_site_categories_matches_oracle(oracle_binding, name):
    """Gcat_model.refine (gcat_model.ml:387-419): per-site tree log-likelihood under every rate category (R passes of
    Like_stree.site_log_like_arr on the device), best category under FastTree's prior; then the whole path with the
    refined categories."""
    prob = PROBLEMS[name]()
    o = oracle_binding.Oracle(prob)
    want_cats, want_ll = o.refine_site_categories(True)
    with Placer(prob) as p:
        got_cats, got_ll = p.refine_site_categories(True)
        assert np.all(np.abs(got_ll - want_ll) <= 1e-10 * np.abs(want_ll)), np.abs((got_ll - want_ll) / want_ll).max()
        differ = np.nonzero(got_cats != want_cats)[0]
        rates = p.model_numerics()["rates"]
        for s in differ:      # only at a tie of the two categories' penalised log-likelihoods
            pen = want_ll[:, s] + 2 * np.log(rates) - 3 * rates
            assert abs(pen[got_cats[s]] - pen[want_cats[s]]) <= 1e-9 * abs(pen[want_cats[s]]), (s, pen)
        assert len(differ) <= 1
        assert (got_cats != prob.site_cat).any()           # the refinement changed something
        if len(differ) == 0:
            o.precompute()
            p.precompute()
            exp, n_exp, st_exp, _ = o.place(prob.query_chars[:10], oracle_prefs(oracle_binding, prob), n_threads=4)
            got, n_got, st_got, _ = p.place(prob.query_chars[:10], p.prefs())
            assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)
    o.close()
    with Placer(configs.c1()) as p:
        with pytest.raises(_lib.PplacerB200Error):
            p.refine_site_categories()                     # Gmix models have no site categories


@pytest.mark.parametrize("name", ["gtr", "wag"])
def test_check_like_table(name):
    """--check-like (pplacer_run.ml:344-377): tree_likelihood and supernode_like agree at every edge and equal PhyML's."""
    tree, aln, model = configs.like_inputs(name)
    prob = reference.build_problem(tree, model, aln, [("q", aln[0][1])], pre_mask=False)
    with Placer(prob) as p:
        p.precompute()
        t = p.check_like()
    want = configs.PHYML_LL[name]
    assert t.shape == (tree.n_nodes, 2)
    assert np.all(np.abs(t[:, 0] - want) <= 1e-7 * abs(want))
    nr = np.array(tree.nonroot_ids())
    assert np.all(np.abs(t[nr, 1] - want) <= 1e-7 * abs(want)) and np.isnan(t[tree.root, 1])


# ---------------------------------------------------------------------------------------------
# fig ranking (--fig-cutoff)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["c1_gtr", "wag", "gcat_dna", "gcat_aa"])
def test_map_states_match_the_posterior_of_the_oracle_glvs(oracle_binding, name):
    """pp_map_states (Map_seq.of_map, map_seq.ml:17-37; Seq_post.get_summary Proximal, seq_post.ml:14-31) against a numpy
    restatement on the ORACLE's darr / parr: statd .* evolve(darr[id], bl[id]) .* parr[id], summed over the rates,
    arg-max per site, -1 below the cutoff."""
    prob, o, p = setup(name, oracle_binding)
    mn = p.model_numerics()
    t = prob.tree
    ids = [i for i in t.nonroot_ids()][:: max(1, len(t.nonroot_ids()) // 6)][:6]
    cutoff = 0.8
    got = p.map_states(ids, cutoff)
    gcat = prob.model.model_class == 1
    for row, nid in zip(got, ids):
        ad, _ = o.glv(0, nid)
        ap, _ = o.glv(1, nid)
        P = oracle_binding.multi_exp(mn["x"], mn["lam"], mn["xit"], mn["rates"], float(t.bl[nid]))
        if gcat:
            ev = np.einsum("lsj,lj->ls", P[prob.site_cat], ad)
            u = mn["statd"][None, :] * ev * ap
        else:
            ev = np.einsum("rsj,rlj->rls", P, ad)
            u = (mn["statd"][None, None, :] * ev * ap).sum(axis=0)
        best = u.argmax(axis=1)
        share = u.max(axis=1) / np.abs(u).sum(axis=1)
        want = np.where(share < cutoff, -1, best)
        srt = np.sort(u, axis=1)
        clear = (np.abs(share - cutoff) > 1e-9) & ((srt[:, -1] - srt[:, -2]) > 1e-12 * np.abs(srt[:, -1]))
        assert clear.mean() > 0.9
        assert np.array_equal(row[clear], want[clear]), (name, nid, int((row[clear] != want[clear]).sum()))
        assert (row >= -1).all() and (row < p.S).all()


@pytest.mark.parametrize("name,cutoffs", [("c1_gtr", (0.05, 0.3)), ("synth", (0.05, 0.2, 1.0)), ("gcat_aa", (0.2,)), ("wag", (0.3,))])
def test_fig_ranking_matches_oracle(oracle_binding, name, cutoffs):
    """Fig._enum_by_score on the device (fig.cuh): representatives scored for every query, groups of figs expanded and
    their members scored only when baseball needs more locations.  Same placements, the same number of locations
    visited AND THE SAME EDGES SCORED as the oracle's lazy enumeration (scores of untouched edges stay NaN)."""
    from pplacer_b200 import fig
    prob, o, p = setup(name, oracle_binding)
    try:
        for cutoff in cutoffs:
            figs = fig.figs_of_tree(cutoff, prob.tree)
            assert figs and p.set_figs(figs) == len(figs)
            o.set_figs(figs)
            exp, n_exp, st_exp, oi = o.place(prob.query_chars, oracle_prefs(oracle_binding, prob), n_threads=4, debug=True)
            got, n_got, st_got, gi = p.place(prob.query_chars, p.prefs(), debug=True)
            near = assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)
            assert len(near_queries(near)) <= max(1, len(n_exp) // 50), near
            assert (gi["n_pitches"] == oi["n_pitches"]).mean() >= 0.97
            scored_g, scored_o = np.isfinite(gi["scores"]), np.isfinite(oi["scores"])
            assert (scored_g == scored_o).all(axis=1).mean() >= 0.97          # the lazy expansion stops where the reference's does
            both = scored_g & scored_o
            assert np.all(np.abs(gi["scores"][both] - oi["scores"][both]) <= 1e-10 * np.abs(oi["scores"][both]))
            assert scored_o.sum(axis=1).mean() < prob.tree.n_nodes - 1         # fewer edges scored than the plain ranking
    finally:
        p.set_figs(None)
        o.set_figs(None)
    # back to the plain ranking
    exp, n_exp, st_exp, _ = o.place(prob.query_chars[:8], oracle_prefs(oracle_binding, prob), n_threads=4)
    got, n_got, st_got, _ = p.place(prob.query_chars[:8], p.prefs())
    assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)


def test_fig_ranking_at_the_c2_shape(c2_small, oracle_binding):
    from pplacer_b200 import fig
    prob, p = c2_small
    figs = fig.figs_of_tree(0.1, prob.tree)
    q = prob.query_chars[:288]   # >= 256 queries: the representatives go through the query-lane table on the fig view
    o = oracle_binding.Oracle(prob)
    o.precompute()
    o.set_figs(figs)
    exp, n_exp, st_exp, oi = o.place(q, oracle_prefs(oracle_binding, prob), n_threads=os.cpu_count(), debug=True)
    o.close()
    try:
        p.set_figs(figs)
        got, n_got, st_got, gi = p.place(q, p.prefs(), debug=True)
    finally:
        p.set_figs(None)
    assert p.prefilter_kernel_used() == DEFAULT_BIG
    near = assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, near_tie_ll=1e-3)
    assert len(near_queries(near)) <= 5, near
    assert (np.isfinite(gi["scores"]) == np.isfinite(oi["scores"])).all(axis=1).mean() >= 0.97
    assert np.isfinite(oi["scores"]).sum(axis=1).mean() < 0.5 * prob.tree.n_nodes
