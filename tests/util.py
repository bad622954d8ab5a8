"""Shared helpers for the parity tests: the north_star tolerances in one place."""
import numpy as np

LL_REL = 1e-6        # per-placement log-likelihood, relative
BL_ABS = 1e-4        # branch lengths
RATIO_ABS = 1e-5     # like_weight_ratio / post_prob


def assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, pp=False, names=None, near_tie_ll=None):
    """Compare pp_place output with the oracle's.  Identical status, identical best edge and kept
    edge set (except at documented near-ties: two kept candidates whose log-likelihoods differ by
    less than `near_tie_ll` may swap / a candidate within near_tie_ll of the keep threshold may
    drop in or out), LL within 1e-6 relative, branch lengths 1e-4, ratios 1e-5."""
    near = []
    Q = len(n_exp)
    assert np.array_equal(st_got, st_exp), "status differs: %s vs %s" % (st_got[:16], st_exp[:16])
    for q in range(Q):
        g = got[q][:n_got[q]]
        e = exp[q][:n_exp[q]]
        tag = "query %s" % (names[q] if names else q)
        if st_exp[q] != 0:
            assert n_got[q] == 0
            continue
        ge = {int(r["edge"]): r for r in g}
        ee = {int(r["edge"]): r for r in e}
        if set(ge) != set(ee) or int(g[0]["edge"]) != int(e[0]["edge"]):
            if near_tie_ll is None:
                raise AssertionError("%s: kept edges differ: %s vs %s" % (tag, list(ge), list(ee)))
            # near-tie exception: every edge present on one side only must sit within near_tie_ll of
            # the keep threshold or of the best; record and compare the common edges
            near.append(q)
        for edge in set(ge) & set(ee):
            a, b = ge[edge], ee[edge]
            assert abs(a["log_like"] - b["log_like"]) <= LL_REL * abs(b["log_like"]), (tag, edge, a, b)
            assert abs(a["pendant"] - b["pendant"]) <= BL_ABS, (tag, edge, a, b)
            assert abs(a["distal"] - b["distal"]) <= BL_ABS, (tag, edge, a, b)
            if q not in near:
                assert abs(a["ml_ratio"] - b["ml_ratio"]) <= RATIO_ABS, (tag, edge, a, b)
                if pp:
                    assert abs(a["marginal_like"] - b["marginal_like"]) <= LL_REL * abs(b["marginal_like"]), (tag, edge, a, b)
                    assert abs(a["post_prob"] - b["post_prob"]) <= RATIO_ABS, (tag, edge, a, b)
    return near
