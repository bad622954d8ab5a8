"""Reference packages: the inputs the placement path reads from a refpkg directory
(pplacer_src/refpkg_parse.ml:86-104 CONTENTS.json -> file map; refpkg.ml:26-63 get_ref_tree /
get_aln_fasta / get_model): `tree`; `aln_fasta`, else `aln_sto` (Stockholm, refpkg.ml:33-36); `phylo_model` JSON, else
the deprecated `tree_stats` statistics file of PhyML / RAxML (refpkg.ml:51-59, parse_stats.ml)."""
from __future__ import annotations

import json
import os
import sys
from typing import List, Tuple

from . import models, newick, seqio, stats


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
    aln: List[Tuple[str, str]]
    if "aln_fasta" in files:
        aln = seqio.read_fasta(os.path.join(path, files["aln_fasta"]), upper=True)
    elif "aln_sto" in files:
        aln = seqio.read_stockholm(os.path.join(path, files["aln_sto"]), upper=True)
    else:
        raise ValueError("refpkg has neither aln_fasta nor aln_sto")
    if "phylo_model" in files:
        model = models.model_from_json_file(os.path.join(path, files["phylo_model"]))
    elif "tree_stats" in files:
        sys.stderr.write("Warning: using a statistics file directly is now deprecated. We suggest using a reference "
                         "package. If you already are, then please use the latest version of taxtastic.\n")
        model = stats.model_from_stats_file(os.path.join(path, files["tree_stats"]))
    else:
        raise ValueError("refpkg has neither phylo_model nor tree_stats")
    return tree, aln, model
