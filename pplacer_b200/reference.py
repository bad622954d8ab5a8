"""Assemble the inputs of the placement path from a tree, a reference alignment, a model and a
query list — the host-side steps of Pplacer_run.run_placements that precede the numerics
(pplacer_src/pplacer_run.ml:167-264, 380-410): length check, statd source (UN-premasked reference
alignment, :205 precedes :207-218), pre-masking (:59-144), Gcat mask_sites
(gcat_model.ml:429-439), de-duplication (:230-253), leaf -> alignment-row map
(like_stree.ml:17-24, alignment.ml:154-169), upper-casing (:409), prior choice (:380-390) and the
informative prior means (core.ml:34-52).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .models import ModelSpec
from .newick import Tree
from .seqio import informative, to_char_matrix


@dataclass
class Problem:
    tree: Tree
    model: ModelSpec
    ref_names: List[str]
    ref_chars_raw: np.ndarray        # uint8[N, L_raw]  (statd source)
    ref_chars: np.ndarray            # uint8[N, L] pre-masked
    site_cat: Optional[np.ndarray]   # int32[L] (Gcat, masked) or None
    leaf_row: np.ndarray             # int32[n_nodes]
    query_names: List[List[str]]     # per unique query: the names sharing that sequence
    query_chars: np.ndarray          # uint8[Q, L] unique, pre-masked, upper-cased
    mask: Optional[np.ndarray]       # bool[L_raw] or None when --no-pre-mask
    L_raw: int = 0

    @property
    def L(self) -> int:
        return int(self.ref_chars.shape[1])


def leaf_row_map(tree: Tree, ref_names: List[str]) -> np.ndarray:
    leaves = tree.leaves()
    if len(leaves) != len(ref_names):
        raise ValueError("tree has %d taxa, and ref align has %d." % (len(leaves), len(ref_names)))
    index: Dict[str, int] = {}
    for i, n in enumerate(ref_names):
        if n in index:
            raise ValueError("taxon in alignment repeatedly: '%s'" % n)
        index[n] = i
    out = np.full(tree.n_nodes, -1, dtype=np.int32)
    for lf in leaves:
        name = tree.node_label[lf]
        if name not in index:
            raise ValueError("taxon not found in alignment: '%s'" % name)
        out[lf] = index[name]
    return out


def _check_alphabet(model: ModelSpec, names: List[str], chars: np.ndarray) -> None:
    ok = np.zeros(256, dtype=bool)
    for ch in model.alphabet():
        ok[ord(ch)] = True
    bad = ~ok[chars]
    if bad.any():
        r, c = np.argwhere(bad)[0]
        raise ValueError("%s is not a known base in %s" % (chr(chars[r, c]), names[r]))


def premask(model: ModelSpec, ref_names, ref_chars, q_names, q_chars, discard_nonoverlapped=False):
    """pplacer_run.ml:59-144.  Returns (q_names, q_chars_cut, ref_chars_cut, mask)."""
    _check_alphabet(model, ref_names, ref_chars)
    _check_alphabet(model, q_names, q_chars)
    ref_mask = informative(ref_chars).any(axis=0)
    overlaps = (informative(q_chars) & ref_mask[None, :]).any(axis=1)
    if not overlaps.all():
        if not discard_nonoverlapped:
            bad = q_names[int(np.argmin(overlaps))]
            raise ValueError("Sequence %s doesn't overlap any reference sequences." % bad)
        keep = np.nonzero(overlaps)[0]
        q_names = [q_names[i] for i in keep]
        q_chars = q_chars[keep]
    q_mask = informative(q_chars).any(axis=0) if len(q_chars) else np.zeros(ref_chars.shape[1], bool)
    mask = ref_mask & q_mask
    if not mask.any():
        raise ValueError("Sequence length cut to 0 by pre-masking; can't proceed with no information.")
    return q_names, np.ascontiguousarray(q_chars[:, mask]), np.ascontiguousarray(ref_chars[:, mask]), mask


def build_problem(tree: Tree, model: ModelSpec, ref_align: List[Tuple[str, str]],
                  queries: List[Tuple[str, str]], pre_mask: bool = True,
                  discard_nonoverlapped: bool = False) -> Problem:
    ref_names = [n for n, _ in ref_align]
    ref_raw = to_char_matrix([s.upper() for _, s in ref_align])   # refpkg.ml:37
    L_raw = ref_raw.shape[1]
    q_names = [n for n, _ in queries]
    for n, s in queries:
        if len(s) != L_raw:
            raise ValueError("query %s is not the same length as the reference alignment (got %d; expected %d)"
                             % (n, len(s), L_raw))
    q_chars = to_char_matrix([s.upper() for _, s in queries]) if queries else np.zeros((0, L_raw), np.uint8)
    site_cat = None
    if model.model_class == 1:
        if len(model.site_cat) != L_raw:   # Model.check gcat_model.ml:421-427
            raise ValueError("mismatch: %d sites in site_categories and %d sites in reference alignment"
                             % (len(model.site_cat), L_raw))
        site_cat = model.site_cat.astype(np.int32)
    mask = None
    ref_cut = ref_raw
    if pre_mask:
        q_names, q_chars, ref_cut, mask = premask(model, ref_names, ref_raw, q_names, q_chars,
                                                  discard_nonoverlapped)
        if site_cat is not None:
            site_cat = np.ascontiguousarray(site_cat[mask])
    # de-duplicate identical (pre-masked) query strings; first-seen order (the reference's order is
    # Hashtbl-fold order, compare by name)
    seen: Dict[bytes, int] = {}
    names: List[List[str]] = []
    rows: List[int] = []
    for i, n in enumerate(q_names):
        key = q_chars[i].tobytes()
        j = seen.get(key)
        if j is None:
            seen[key] = len(rows)
            rows.append(i)
            names.append([n])
        else:
            names[j].append(n)
    uq = np.ascontiguousarray(q_chars[rows]) if rows else np.zeros((0, ref_cut.shape[1]), np.uint8)
    return Problem(tree=tree, model=model, ref_names=ref_names, ref_chars_raw=ref_raw, ref_chars=ref_cut,
                   site_cat=site_cat, leaf_row=leaf_row_map(tree, ref_names), query_names=names,
                   query_chars=uq, mask=mask, L_raw=L_raw)


def flat_exp_prior_mean(tree: Tree) -> float:
    """pplacer_run.ml:387-390: exponential with mean = tree length / number of edges."""
    return tree.tree_length() / float(tree.n_edges())


def informative_prior_means(tree: Tree, prior_lower: float = 0.0, top_segment_f=None) -> np.ndarray:
    """Core.prior_mean_map top_segment_f (core.ml:34-52): for edge i, the average over leaves below
    i of (distance from the leaf up to the bottom of edge i) + top_segment_f(bl_i).  pplacer passes
    top_segment_f = (+.) prior_lower (pplacer_run.ml:380-386), the default here."""
    if top_segment_f is None:
        top_segment_f = lambda b: prior_lower + b
    out = np.full(tree.n_nodes, np.nan)
    dists: Dict[int, List[float]] = {}
    for node in tree.postorder():
        top = top_segment_f(float(tree.bl[node]))
        if not tree.children[node]:
            out[node] = top
            dists[node] = [float(tree.bl[node])]
        else:
            below: List[float] = []
            for c in tree.children[node]:
                below.extend(dists.pop(c))
            out[node] = sum(d + top for d in below) / float(len(below))
            dists[node] = [d + tree.bl[node] for d in below]
    out[tree.root] = np.nan
    return out
