"""Seeded synthetic inputs for the configs of BASELINE.json / SURVEY.md section 8(d).

  C1  hiv (tests/data/like/gtr: 20 taxa x 216 sites, GTR+G4 alpha 0.771) + 40 synthetic queries
  C2  16S shape: 800 taxa, 1500 columns, 10k reads, GTR+G4 (seed 2)
  C4  10k taxa, 1500 columns, 1M reads (seed 4)
  C5  C2 with -p

Trees: random stepwise addition onto a trifurcating root, branch lengths ~ Exp(mean 0.05) floored
at 1e-6.  Alignment: sites evolved down the tree under GTR+G4 from the stationary distribution,
~10 % of leaf cells gapped in geometric runs.  Queries: a new tip hung off a random edge with
pendant ~ Exp(0.05), restricted to a random window of 250-1500 columns, 0.5 % 'N'.
Only numpy + the library's host-side model numerics (no GPU) are used.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from . import _lib
from .models import ModelSpec, gtr_exch
from .newick import Tree, parse
from .reference import Problem, build_problem

HIV_GTR = dict(ac=1.17728, ag=3.08642, at=0.60698, cg=0.92672, ct=4.74366, gt=1.0)
HIV_PI = np.array([0.42, 0.15, 0.19, 0.24])   # roughly the hiv empirical frequencies; only used to simulate


def random_tree(n_taxa: int, rng: np.random.Generator, mean_bl: float = 0.05, min_bl: float = 1e-6) -> Tree:
    """Random stepwise addition; returns a Tree numbered like the reference parser would."""
    # node records: children lists; start with a trifurcating root over 3 leaves
    children: List[List[int]] = [[1, 2, 3], [], [], []]
    parent = [-1, 0, 0, 0]
    for _ in range(3, n_taxa):
        # pick a random existing non-root node = edge above it; split that edge
        v = int(rng.integers(1, len(children)))
        p = parent[v]
        mid = len(children)
        leaf = mid + 1
        children.append([v, leaf])      # new internal node
        children.append([])             # new leaf
        parent.append(p)
        parent.append(mid)
        children[p][children[p].index(v)] = mid
        parent[v] = mid
    n = len(children)
    bl = np.maximum(rng.exponential(mean_bl, size=n), min_bl)
    # Newick string written iteratively, leaves named t<k> in order of appearance in the string
    out: List[str] = []
    leaf_no = 0
    stack: List[Tuple[int, int]] = [(0, 0)]
    while stack:
        node, k = stack.pop()
        ch = children[node]
        if not ch:
            out.append("t%d:%.10g" % (leaf_no, bl[node]))
            leaf_no += 1
            continue
        if k == 0:
            out.append("(")
        if k < len(ch):
            if k > 0:
                out.append(",")
            stack.append((node, k + 1))
            stack.append((ch[k], 0))
        else:
            out.append(")" if node == 0 else "):%.10g" % bl[node])
    return parse("".join(out) + ";")


def _pmats(x, lam, xit, rates, t):
    # P_r(t) = X diag(exp(rate t lam)) Xit^T   (diagd.ml:79-98)
    return np.stack([(x * np.exp(r * t * lam)[None, :]) @ xit.T for r in rates])


def _sample_children(states: np.ndarray, cats: np.ndarray, P: np.ndarray, rng) -> np.ndarray:
    probs = P[cats, states]                      # [L, S]
    probs = np.maximum(probs, 0.0)
    cdf = np.cumsum(probs, axis=1)
    u = rng.random(len(states)) * cdf[:, -1]
    return np.minimum((u[:, None] > cdf).sum(axis=1), P.shape[-1] - 1).astype(np.int8)


@dataclass
class Synth:
    tree: Tree
    model: ModelSpec
    ref_align: List[Tuple[str, str]]
    node_states: np.ndarray      # int8[n_nodes, L] true states at every node
    site_cats: np.ndarray        # rate category used per site in the simulation
    rates: np.ndarray
    numerics: tuple              # (x, lam, xit)
    pi: np.ndarray


def simulate_reference(n_taxa: int, L: int, seed: int, alpha: float = 0.5, n_cats: int = 4,
                       gap_frac: float = 0.10, mean_bl: float = 0.05) -> Synth:
    rng = np.random.default_rng(seed)
    tree = random_tree(n_taxa, rng, mean_bl=mean_bl)
    exch = gtr_exch(**HIV_GTR)
    pi = HIV_PI / HIV_PI.sum()
    x, lam, xit = _lib.diagd(exch, pi)
    rates = _lib.discrete_gamma(n_cats, alpha)
    cats = rng.integers(0, n_cats, size=L)
    states = np.zeros((tree.n_nodes, L), dtype=np.int8)
    states[tree.root] = rng.choice(4, size=L, p=pi)
    order = list(reversed(tree.postorder()))     # parents before children
    for node in order:
        for c in tree.children[node]:
            states[c] = _sample_children(states[node], cats, _pmats(x, lam, xit, rates, float(tree.bl[c])), rng)
    code = np.frombuffer(b"ACGT", dtype=np.uint8)
    ref = []
    for lf in tree.leaves():
        row = code[states[lf]].copy()
        # gaps in geometric runs
        n_gap = int(gap_frac * L)
        placed = 0
        while placed < n_gap:
            start = int(rng.integers(0, L))
            run = int(min(rng.geometric(1.0 / 12.0), n_gap - placed, L - start))
            row[start:start + run] = ord("-")
            placed += run
        ref.append((tree.node_label[lf], row.tobytes().decode("ascii")))
    model = ModelSpec(4, 0, "GTR", exch, None, True, n_cats, alpha, None, None)
    return Synth(tree, model, ref, states, cats, rates, (x, lam, xit), pi)


def simulate_queries(s: Synth, Q: int, seed: int, min_len: int = 250, max_len: int = None,
                     pend_mean: float = 0.05, n_frac: float = 0.005) -> List[Tuple[str, str]]:
    rng = np.random.default_rng(seed + 7919)
    L = s.node_states.shape[1]
    max_len = L if max_len is None else min(max_len, L)
    min_len = min(min_len, max_len)
    x, lam, xit = s.numerics
    code = np.frombuffer(b"ACGT", dtype=np.uint8)
    nonroot = np.array(s.tree.nonroot_ids())
    out = []
    # quantise branch lengths so the P matrices can be cached
    cache = {}
    for k in range(Q):
        edge = int(rng.choice(nonroot))
        t = float(rng.exponential(pend_mean)) + float(rng.random() * s.tree.bl[edge])
        tq = max(round(t, 3), 1e-3)
        P = cache.get(tq)
        if P is None:
            P = cache[tq] = _pmats(x, lam, xit, s.rates, tq)
        st = _sample_children(s.node_states[edge], s.site_cats, P, rng)
        row = code[st].copy()
        nmask = rng.random(L) < n_frac
        row[nmask] = ord("N")
        n_keep = int(rng.integers(min_len, max_len + 1))
        start = int(rng.integers(0, L - n_keep + 1))
        seq = np.full(L, ord("-"), dtype=np.uint8)
        seq[start:start + n_keep] = row[start:start + n_keep]
        out.append(("q%d" % k, seq.tobytes().decode("ascii")))
    return out


def make_problem(n_taxa: int, L: int, Q: int, seed: int, **kw) -> Problem:
    s = simulate_reference(n_taxa, L, seed, **{k: v for k, v in kw.items() if k in ("alpha", "n_cats", "gap_frac", "mean_bl")})
    queries = simulate_queries(s, Q, seed, **{k: v for k, v in kw.items() if k in ("min_len", "max_len", "pend_mean")})
    return build_problem(s.tree, s.model, s.ref_align, queries)


def perturbed_queries(ref_align: List[Tuple[str, str]], alphabet: str, seed: int, sub_frac: float = 0.05,
                      window: int = 100) -> List[Tuple[str, str]]:
    """C1-style queries: each reference row twice - full length with i.i.d. substitutions, and a
    random window with the rest gapped."""
    rng = np.random.default_rng(seed)
    out = []
    letters = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    for name, seq in ref_align:
        row = np.frombuffer(seq.upper().encode(), dtype=np.uint8).copy()
        L = len(row)
        a = row.copy()
        hit = (rng.random(L) < sub_frac) & (a != ord("-"))
        a[hit] = rng.choice(letters, size=int(hit.sum()))
        out.append(("q_%s_0" % name, a.tobytes().decode()))
        b = np.full(L, ord("-"), dtype=np.uint8)
        w = min(window, L)
        st = int(rng.integers(0, L - w + 1))
        b[st:st + w] = row[st:st + w]
        if not ((b != ord("-")) & (b != ord("?"))).any():
            b[st:st + w] = a[st:st + w]
        out.append(("q_%s_1" % name, b.tobytes().decode()))
    return out
