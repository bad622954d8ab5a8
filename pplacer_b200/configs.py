"""The five workloads of BASELINE.json `configs` plus small test problems.

  C1  hiv GTR+G4 (the reference's own tests/data/like/gtr: 20 taxa x 216 columns) + 40 queries
  C2  16S shape: 800 taxa x 1500 columns, 10k reads, GTR+G4            <- the bench workload
  C3  donovan-test ar53 (WAG + Price-CAT 20, 6968 taxa x 8035 columns, 3 bundled genomes)
  C4  10k taxa x 1500 columns, GTR+G4, reads sharded over the GPUs
  C5  C2 with -p (posterior probability; keep_at_most raised)

All inputs are synthetic and seeded, except the reference's bundled data files copied as DATA
into tests/golden/ by tests/golden/make_fixtures.py.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Tuple

import numpy as np

from . import models, newick, seqio, synth
from .reference import Problem, build_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# PhyML tree log-likelihoods asserted by the reference (tests/pplacer/test_like.ml:71-90)
PHYML_LL = {"gtr": -1752.42340, "wag": -2334.79151, "jtt": -2326.22075}
_LIKE_FILES = {"gtr": "hiv", "wag": "actin", "jtt": "actin"}


def like_inputs(name: str):
    d = os.path.join(GOLDEN, "like", name)
    stem = _LIKE_FILES[name]
    tree = newick.parse_file(os.path.join(d, stem + ".phy_phyml_tree.txt"))
    aln = seqio.read_fasta(os.path.join(d, stem + ".fasta"))
    model = models.model_from_json_file(os.path.join(d, "phylo_model.jplace"))
    return tree, aln, model


def like_problem(name: str, seed: int = 1, pre_mask: bool = True, queries=None) -> Problem:
    """C1-style problem on one of the reference's likelihood fixtures (gtr | wag | jtt)."""
    tree, aln, model = like_inputs(name)
    if queries is None:
        queries = synth.perturbed_queries(aln, models.NUC_CODE if model.n_states == 4 else models.PROT_CODE, seed)
    return build_problem(tree, model, aln, queries, pre_mask=pre_mask)


def c1() -> Problem:
    return like_problem("gtr", seed=1)


def gcat_variant(name: str, n_cats: int = 6, seed: int = 5) -> Problem:
    """The same data under a Price-CAT model (Gcat_model) with seeded rates / site categories:
    exercises the mat_* path on small inputs (the only bundled Gcat data is the 54 GB ar53)."""
    tree, aln, model = like_inputs(name)
    rng = np.random.default_rng(seed)
    L = len(aln[0][1])
    rates = np.sort(rng.gamma(2.0, 0.5, size=n_cats))
    cats = rng.integers(0, n_cats, size=L).astype(np.int32)
    m = models.ModelSpec(model.n_states, 1, model.subs_model, model.exch, model.model_statd, model.empirical_freqs,
                         n_cats, 0.0, rates, cats)
    queries = synth.perturbed_queries(aln, models.NUC_CODE if model.n_states == 4 else models.PROT_CODE, seed)
    return build_problem(tree, m, aln, queries)


def synthetic(n_taxa: int, L: int, Q: int, seed: int, **kw) -> Problem:
    return synth.make_problem(n_taxa, L, Q, seed, **kw)


def c2(Q: int = 10_000) -> Problem:
    return synth.make_problem(800, 1500, Q, seed=2, alpha=0.5)


# C4 branch lengths: with the C2 mean (0.05) a 10k-taxon tree is ~1000 substitutions/site long and the
# REFERENCE ITSELF degenerates under Gamma(0.5): the slowest rate category underflows to 0, after which
# perhaps_pull_exponent never rescales that site again (frexp 0. = (0., 0), gmix_model.ml:172-176) and
# whole sites decay to 0 => log-likelihood -inf at every edge.  Mean 0.002 (tree
# length ~40, dense species-level sampling) keeps every site maximum above 1e-8.
C4_MEAN_BL = 0.002


def c4(Q: int, n_taxa: int = 10_000) -> Problem:
    return synth.make_problem(n_taxa, 1500, Q, seed=4, alpha=0.5, mean_bl=C4_MEAN_BL)


def ar53_available() -> bool:
    return os.path.exists(os.path.join(GOLDEN, "_ar53", "msa.npz"))


def ar53(extra_queries: Optional[List[Tuple[str, str]]] = None) -> Problem:
    """C3 from the git-ignored fixture tests/golden/_ar53 (make_fixtures.py --ar53)."""
    d = os.path.join(GOLDEN, "_ar53")
    tree = newick.parse_file(os.path.join(d, "tree.nwk"))
    model = models.model_from_json_file(os.path.join(d, "phylo_model.json"))
    z = np.load(os.path.join(d, "msa.npz"))
    aln = [(str(n), row.tobytes().decode("ascii")) for n, row in zip(z["names"], z["chars"])]
    queries = seqio.read_fasta(os.path.join(d, "queries.fasta.gz"))
    if extra_queries:
        queries = queries + extra_queries
    return build_problem(tree, model, aln, queries)


def ar53_golden():
    with open(os.path.join(GOLDEN, "ar53_golden.json")) as fh:
        return json.load(fh)
