"""Figs (pplacer --fig-cutoff): groups of edges whose leaves lie within `cutoff` of each other, used to rank locations
without scoring every edge - score one representative per fig, then expand the best figs (pplacer_src/fig.ml).

This module restates the HOST part: Fig._figs_of_gtree (fig.ml:70-99) - the list of (representative edge, edge set) -
including its list orders, which decide ties in the later stable sorts.  The enumeration itself (fig.ml:120-148) runs on
the device (csrc/fig.cuh) and in the oracle.

What a fig contains, as the reference builds it (restated, not "fixed"): the leaf edges of the group, plus - added when
the group is closed at a node or passed on upwards - the node where it was closed and every ancestor edge above it
except the root (add_edge_to_figs, fig.ml:46-47).  Internal edges strictly below the closing node are in no fig.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

Fig = Tuple[int, List[int]]          # (representative edge, sorted edge ids)


def figs_of_tree(cutoff: float, tree) -> Optional[List[Fig]]:
    """Fig.figs_of_gtree (fig.ml:103-107): None (= Dummy: plain ranking of every non-root edge) when cutoff == 0."""
    if cutoff == 0.0:
        return None
    bl = tree.bl
    top = tree.root

    def aux(i: int):
        # returns (max_bl | None, edge set, rep_edge, accum list)
        kids = tree.children[i]
        if not kids:
            return float(bl[i]), {i}, i, []
        # fold_figs (fig.ml:52-66): fold_left over the subtrees in order
        bls: List[float] = []
        all_edges = set()
        rep: Optional[Tuple[int, int]] = None
        maybe_figs: List[Tuple[int, set]] = []
        tot_accum: List[Tuple[int, set]] = []
        for j in kids:
            max_bl, edges, rep_edge, accum = aux(j)
            if max_bl is not None:
                bls.insert(0, max_bl)                       # maybe_cons: prepended
            all_edges |= edges
            card = len(edges)
            if not (rep is not None and rep[1] >= card):
                rep = (j, card)
            maybe_figs.insert(0, (rep_edge, edges))         # consed
            tot_accum = accum + tot_accum                   # List.append cur.accum accum.tot_accum
        acc = tot_accum if i == top else [(r, e | {i}) for r, e in tot_accum]
        rep_edge = rep[0]
        srt = sorted(bls, reverse=True)
        if not srt:
            return None, all_edges, rep_edge, acc
        if i == top:
            return None, all_edges, rep_edge, acc + maybe_figs
        if len(srt) == 1:
            return srt[0] + float(bl[i]), all_edges, rep_edge, acc
        if srt[0] + srt[1] <= cutoff:
            return srt[0] + float(bl[i]), all_edges, rep_edge, acc
        return None, set(), rep_edge, acc + [(r, e | {i}) for r, e in maybe_figs]

    import sys
    lim = sys.getrecursionlimit()
    sys.setrecursionlimit(max(lim, 4 * tree.n_nodes + 1000))
    try:
        accum = aux(top)[3]
    finally:
        sys.setrecursionlimit(lim)
    return [(r, sorted(e)) for r, e in accum]


def figs_csr(figs: List[Fig]):
    """-> (rep int32[m], ptr int32[m+1], idx int32[...]) for pp_set_figs / the oracle."""
    rep = np.array([r for r, _ in figs], dtype=np.int32)
    ptr = np.zeros(len(figs) + 1, dtype=np.int32)
    for k, (_, e) in enumerate(figs):
        ptr[k + 1] = ptr[k] + len(e)
    idx = np.array([x for _, e in figs for x in e], dtype=np.int32) if figs else np.zeros(0, np.int32)
    return rep, ptr, idx


def enum_by_score(figs: Optional[List[Fig]], score, strike_box: float, nonroot_ids: List[int]):
    """Fig.enum_by_score (fig.ml:120-148) as a Python generator of (score, edge): the reference semantics in the plainest
    form, used by the host tests as a third opinion next to the oracle and the device."""
    import math

    def key_desc(s):       # descending with OCaml's compare on floats: nan below everything
        return (0, -s) if not math.isnan(s) else (1, 0.0)

    if figs is None:
        for s, e in sorted(((score(e), e) for e in nonroot_ids), key=lambda t: key_desc(t[0])):
            yield s, e
        return
    scored = sorted(((score(r), es) for r, es in figs), key=lambda t: key_desc(t[0]))    # stable
    yielded = set()
    k = 0
    while k < len(scored):
        s0, edges = scored[k]
        k += 1
        group = set(edges)
        while k < len(scored) and abs(s0 - scored[k][0]) < strike_box:      # approx_equal ~epsilon:strike_box
            group |= set(scored[k][1])
            k += 1
        fresh = sorted(group - yielded)
        yielded |= group
        for s, e in sorted(((score(e), e) for e in fresh), key=lambda t: key_desc(t[0])):
            yield s, e
