"""Reference packages: the three inputs the placement path reads from a refpkg directory
(pplacer_src/refpkg_parse.ml:86-104 CONTENTS.json -> file map; refpkg.ml:26-63 get_ref_tree /
get_aln_fasta / get_model).  Only `tree`, `aln_fasta` and `phylo_model` are consumed; a refpkg that
carries nothing but `tree_stats` (deprecated in the reference, refpkg.ml:51-59) is rejected with a
clear message - PhyML / RAxML stats parsing (parse_stats.ml) is not part of this path."""
from __future__ import annotations

import json
import os
from typing import List, Tuple

from . import models, newick, seqio


def contents(path: str) -> dict:
    fn = os.path.join(path, "CONTENTS.json")
    if not os.path.isfile(fn):
        raise ValueError("%s is not a reference package (no CONTENTS.json)" % path)
    with open(fn) as fh:
        return json.load(fh)["files"]


def load(path: str):
    """-> (Tree, [(name, SEQ)], ModelSpec).  The alignment is upper-cased (refpkg.ml:37)."""
    files = contents(path)
    if "tree" not in files:
        raise ValueError("refpkg has no tree")
    tree = newick.parse_file(os.path.join(path, files["tree"]))
    if "aln_fasta" not in files:
        raise ValueError("refpkg has no aln_fasta (Stockholm-only refpkgs are not supported on this path)")
    aln: List[Tuple[str, str]] = seqio.read_fasta(os.path.join(path, files["aln_fasta"]), upper=True)
    if "phylo_model" not in files:
        raise ValueError("refpkg has no phylo_model JSON (statistics-file refpkgs are deprecated in the reference "
                         "and not supported here; rebuild the refpkg with a current taxtastic)")
    model = models.model_from_json_file(os.path.join(path, files["phylo_model"]))
    return tree, aln, model
