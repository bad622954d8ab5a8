"""`pplacer -c REFPKG queries.fasta` for the placement path, without the OCaml host:
refpkg + query FASTA -> Problem (pre-mask, de-duplicate; pplacer_run.ml:59-253) -> pp_place on one
GPU or sharded over the ranks of a torch.distributed job -> jplace (placerun_io.ml:58-89).
This is the thin host front-end of SURVEY.md section 8(f) rows 1-2; flag names follow prefs.ml:185-290.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from typing import List, Optional, Tuple

import numpy as np

from . import jplace, refpkg, seqio
from .placer import Placer
from .reference import build_problem


def place_problem(prob, device: int = 0, pp: bool = False, always_refine: bool = False, check_like: bool = False,
                  devices: Optional[List[int]] = None, fig_cutoff: float = 0.0, **prefs_kw):
    """devices: several GPUs behind one handle (pp_create_multi) - what -j does in the reference."""
    if devices is not None and len(devices) > 1 and not always_refine and not check_like and fig_cutoff == 0.0:
        from .placer import MultiPlacer
        with MultiPlacer(prob, devices) as mp:
            mp.precompute()
            return mp.place(prob.query_chars, mp.prefs(pp=pp, **prefs_kw))
    with Placer(prob, device=device) as p:
        if always_refine and prob.model.model_class == 1:     # Model.refine, pplacer_run.ml:316-317 (Gmix: a no-op)
            sys.stderr.write("Optimizing site categories... ")
            p.refine_site_categories()
            sys.stderr.write("done.\n")
        p.precompute()
        if check_like:                                         # pplacer_run.ml:344-377
            t = p.check_like()
            sys.stderr.write("node  tree_likelihood  supernode_like\n")
            for i in range(t.shape[0]):
                sys.stderr.write("%d  %g  %g\n" % (i, t[i, 0], t[i, 1]))
        if fig_cutoff != 0.0:                                  # pplacer_run.ml:274-278
            sys.stderr.write("Determining figs... %d figs.\n" % p.set_figs(cutoff=fig_cutoff))
        prefs = p.prefs(pp=pp, **prefs_kw)
        out, n_out, st, _ = p.place(prob.query_chars, prefs)
    return out, n_out, st


def run_file(refpkg_path: str, query_fasta: str, out_path: Optional[str] = None, pp: bool = False,
             device: int = 0, pre_mask: bool = True, discard_nonoverlapped: bool = False, invocation: str = "",
             always_refine: bool = False, check_like: bool = False, devices: Optional[List[int]] = None,
             fig_cutoff: float = 0.0, **prefs_kw):
    """Pplacer_run.run_file (pplacer_run.ml:682-783) for a refpkg and one query file."""
    tree, aln, model = refpkg.load(refpkg_path)
    queries = seqio.read_fasta(query_fasta, upper=True)          # pplacer_run.ml:409
    prob = build_problem(tree, model, aln, queries, pre_mask=pre_mask, discard_nonoverlapped=discard_nonoverlapped)
    out, n_out, st = place_problem(prob, device=device, pp=pp, always_refine=always_refine, check_like=check_like,
                                   devices=devices, fig_cutoff=fig_cutoff, **prefs_kw)
    bad = np.nonzero(st != 0)[0]
    for q in bad:
        sys.stderr.write("Warning: %s: no placement (status %d)\n" % (prob.query_names[q][0], st[q]))
    if out_path is None:
        stem = os.path.basename(query_fasta)
        for ext in (".gz", ".fasta", ".fa", ".fna", ".faa"):
            if stem.endswith(ext):
                stem = stem[: -len(ext)]
        out_path = stem + ".jplace"
    jplace.write_jplace(out_path, prob.tree, prob.query_names, out, n_out, st, pp=pp,
                        invocation=invocation or "pplacer_b200 -c %s %s" % (refpkg_path, query_fasta))
    return out_path, prob, (out, n_out, st)


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="pplacer_b200", description=__doc__.split("\n")[0])
    ap.add_argument("-c", dest="refpkg", required=True, help="reference package path")
    ap.add_argument("-o", dest="out", default=None, help="output jplace file")
    ap.add_argument("-p", dest="pp", action="store_true", help="posterior probability mode (prefs.ml:198-203)")
    ap.add_argument("--keep-at-most", type=int)
    ap.add_argument("--keep-factor", type=float)
    ap.add_argument("--max-strikes", type=int)
    ap.add_argument("--strike-box", type=float)
    ap.add_argument("--max-pitches", type=int)
    ap.add_argument("--max-pend", type=float)
    ap.add_argument("--no-pre-mask", action="store_true")
    ap.add_argument("--discard-nonoverlapped", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--devices", type=str, default=None, help="comma-separated GPU ids: shard the queries over them (in place of -j)")
    ap.add_argument("--fig-cutoff", type=float, default=0.0, help="the cutoff for determining figs; 0 disables (prefs.ml:229-230)")
    ap.add_argument("--always-refine", action="store_true", help="re-derive the Price-CAT site categories (prefs.ml always_refine)")
    ap.add_argument("--check-like", action="store_true", help="print the tree likelihood seen from every edge (prefs.ml check_like)")
    ap.add_argument("queries")
    a = ap.parse_args(argv)
    kw = {k: getattr(a, k) for k in ("keep_at_most", "keep_factor", "max_strikes", "strike_box", "max_pitches", "max_pend")
          if getattr(a, k) is not None}
    t0 = time.perf_counter()
    path, prob, (out, n_out, st) = run_file(a.refpkg, a.queries, a.out, pp=a.pp, device=a.device, pre_mask=not a.no_pre_mask,
                                           discard_nonoverlapped=a.discard_nonoverlapped, always_refine=a.always_refine,
                                           check_like=a.check_like, fig_cutoff=a.fig_cutoff, devices=None if a.devices is None else [int(x) for x in a.devices.split(",")],
                                           invocation=" ".join(["pplacer_b200"] + (argv if argv is not None else sys.argv[1:])), **kw)
    sys.stderr.write("placed %d unique queries (%d names) in %.2f s -> %s\n"
                     % (int((st == 0).sum()), sum(len(n) for n in prob.query_names), time.perf_counter() - t0, path))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
