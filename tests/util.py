"""Shared helpers for the parity tests: the north_star tolerances in one place."""
import math

import numpy as np

LL_REL = 1e-6        # per-placement log-likelihood, relative
BL_ABS = 1e-4        # branch lengths
RATIO_ABS = 1e-5     # like_weight_ratio / post_prob


def _one_sided_reason(rec, own, other, log_keep_factor, keep_at_most, near_tie_ll):
    """Why may `rec` (kept on the `own` side only) legitimately be missing from `other`?  Returns the
    documented near-tie category or None.  core.ml:336-352: a result is kept when its rank is below
    keep_at_most and its log-likelihood is at least best + log keep_factor."""
    ll = float(rec["log_like"])
    for side in (own, other):                       # the threshold of either side (their bests agree to LL_REL)
        thr = float(max(r["log_like"] for r in side)) + log_keep_factor
        if abs(ll - thr) <= near_tie_ll:
            return "threshold"
    if len(other) >= keep_at_most and abs(float(min(r["log_like"] for r in other)) - ll) <= near_tie_ll:
        return "truncation"                          # rank keep_at_most vs keep_at_most + 1 swapped
    return None


def assert_placements_match(got, n_got, st_got, exp, n_exp, st_exp, pp=False, names=None, near_tie_ll=None,
                            keep_factor=0.01, keep_at_most=7):
    """Compare pp_place output with the oracle's.  Identical status, identical best edge and kept
    edge set except at VERIFIED near-ties, LL within 1e-6 relative, branch lengths 1e-4, ratios 1e-5.

    Near-tie rule (checked, not assumed): with `near_tie_ll` given,
      * an edge kept on one side only must have its log-likelihood within near_tie_ll of the keep
        threshold best + log keep_factor ("threshold"), or the other side must be full (keep_at_most
        entries) with its last entry within near_tie_ll of this edge ("truncation");
      * a differing best edge needs both sides' best log-likelihoods within near_tie_ll of each
        other (two different edges tie for the maximum, e.g. every edge for an all-N query).
      * a like_weight_ratio / post_prob off by more than 1e-5 must be fully explained by log-likelihood differences
        that are themselves within tolerance ("stop-point", see _stop_point_explains).
    Anything else raises.  Returns the list of (query, category) near-ties found."""
    near = []
    Q = len(n_exp)
    lkf = math.log(keep_factor)
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
        is_near = False
        if set(ge) != set(ee):
            if near_tie_ll is None:
                raise AssertionError("%s: kept edges differ: %s vs %s" % (tag, list(ge), list(ee)))
            for edge in set(ge) ^ set(ee):
                own, other = (g, e) if edge in ge else (e, g)
                rec = ge[edge] if edge in ge else ee[edge]
                why = _one_sided_reason(rec, own, other, lkf, keep_at_most, near_tie_ll)
                assert why is not None, ("%s: edge %d kept on one side only and NOT at a near-tie (ll %.9g; got %s / exp %s)"
                                         % (tag, edge, rec["log_like"], [(int(r["edge"]), float(r["log_like"])) for r in g],
                                            [(int(r["edge"]), float(r["log_like"])) for r in e]))
                near.append((q, why))
            is_near = True
        bg = max(g, key=lambda r: r["log_like"])
        be = max(e, key=lambda r: r["log_like"])
        if int(bg["edge"]) != int(be["edge"]):
            ok = near_tie_ll is not None and abs(float(bg["log_like"]) - float(be["log_like"])) <= near_tie_ll
            assert ok, "%s: best edge differs and is not a near-tie: %s vs %s" % (tag, bg, be)
            near.append((q, "best"))
            is_near = True
        for edge in set(ge) & set(ee):
            a, b = ge[edge], ee[edge]
            assert abs(a["log_like"] - b["log_like"]) <= LL_REL * abs(b["log_like"]), (tag, edge, a, b)
            assert abs(a["pendant"] - b["pendant"]) <= BL_ABS, (tag, edge, a, b)
            assert abs(a["distal"] - b["distal"]) <= BL_ABS, (tag, edge, a, b)
            if not is_near:
                if abs(a["ml_ratio"] - b["ml_ratio"]) > RATIO_ABS:
                    # like_weight_ratio is exp(ll - best) normalised over the kept set: a log-likelihood that moved
                    # by d (inside its own 1e-6 relative tolerance) moves the ratios by up to d.  Accept ONLY when that
                    # is verifiably the whole story (see _stop_point_explains); counted as a near-tie ("stop-point").
                    assert near_tie_ll is not None and _stop_point_explains(g, e, "ml_ratio", "log_like", near_tie_ll), \
                        ("%s: ml_ratio differs beyond 1e-5 and the log-likelihood differences do not explain it" % tag, _dump(g), _dump(e))
                    near.append((q, "stop-point"))
                    is_near = True
                    continue
                if pp:
                    assert abs(a["marginal_like"] - b["marginal_like"]) <= LL_REL * abs(b["marginal_like"]), (tag, edge, a, b)
                    if abs(a["post_prob"] - b["post_prob"]) > RATIO_ABS:
                        assert near_tie_ll is not None and _stop_point_explains(g, e, "post_prob", "marginal_like", near_tie_ll), \
                            ("%s: post_prob differs beyond 1e-5 and the marginal likelihoods do not explain it" % tag, _dump(g), _dump(e))
                        near.append((q, "stop-point"))
                        is_near = True
    return near


def _dump(recs):
    return [(int(r["edge"]), float(r["log_like"]), float(r["pendant"]), float(r["distal"]), float(r["ml_ratio"])) for r in recs]


def _stop_point_explains(g, e, ratio_field, ll_field, near_tie_ll):
    """Brent stops when its bracket is narrower than the tolerance (minimization.ml:88-96), so two runs whose objective
    values differ in the last bits may stop a step apart: branch lengths within 1e-4 of each other, log-likelihoods
    within a few 1e-5 ABSOLUTE - inside the 1e-6 relative tolerance, but enough to move a like_weight_ratio by more
    than 1e-5.  True when (checked, not assumed)
      * both sides kept the same edges, every log-likelihood within min(near_tie_ll, 1e-6 relative) of the other side's,
      * each side's ratios are exactly the normalised exp(ll - max ll) of ITS OWN log-likelihoods (ppatteries.ml:175-183)
        to 1e-9, i.e. nothing but the log-likelihoods differs, and
      * every ratio difference is bounded by the largest log-likelihood difference (d ratio <= ratio (1 - ratio) d ll)."""
    ge = {int(r["edge"]): r for r in g}
    ee = {int(r["edge"]): r for r in e}
    if set(ge) != set(ee):
        return False
    dmax = 0.0
    for edge in ge:
        d = abs(float(ge[edge][ll_field]) - float(ee[edge][ll_field]))
        if not d <= min(near_tie_ll, LL_REL * abs(float(ee[edge][ll_field]))):
            return False
        dmax = max(dmax, d)
    for side in (g, e):
        ll = np.array([float(r[ll_field]) for r in side])
        w = np.exp(ll - ll.max())
        w /= w.sum()
        if not np.allclose(w, [float(r[ratio_field]) for r in side], rtol=0, atol=1e-9):
            return False
    return all(abs(float(ge[k][ratio_field]) - float(ee[k][ratio_field])) <= 2.0 * dmax + 1e-12 for k in ge)


def near_queries(near):
    return sorted({q for q, _ in near})
