"""jplace (version 3) output and the parity comparator.

Writer: the JSON content of Placerun_io.to_json_file (pplacer_src/placerun_io.ml:58-89):
  * "fields" = the sorted union of the placements' keys (StringSet.elements, :80-83); a placement
    carries edge_num, likelihood, like_weight_ratio, distal_length, pendant_length and, with -p,
    post_prob and marginal_like (placement.ml:168-181);
  * every pquery is {"p": [[...fields...], ...], "nm": [[name, multiplicity], ...]}
    (pquery_io.ml:28-45) with placements in Pquery.make_ml_sorted order (pquery.ml:86-92);
  * floats print as %.12g, ints as %d, strings with '/' escaped as '\\/' (json_src/json.ml:25-44);
  * "tree" = Newick_gtree.to_string ~with_node_numbers:true (newick_gtree.ml:55-69; labels, ':%g'
    branch lengths and '{edge}' numbers, newick_bark.ml:69-74); "version": 3;
    "metadata": {"invocation": ...}.
Line breaking differs from OCaml's Format boxes (whitespace only); key order inside the top-level
object is Hashtbl order in the reference and fixed here.

Comparator: the north_star tolerances (identical best edge and kept-edge set except documented
near-ties, log-likelihood 1e-6 relative, branch lengths 1e-4, like_weight_ratio / post_prob 1e-5).
"""
from __future__ import annotations

import gzip
import json
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

BASE_FIELDS = ["distal_length", "edge_num", "like_weight_ratio", "likelihood", "pendant_length"]
PP_FIELDS = ["distal_length", "edge_num", "like_weight_ratio", "likelihood", "marginal_like", "pendant_length",
             "post_prob"]
_RESULT_KEY = {"distal_length": "distal", "edge_num": "edge", "like_weight_ratio": "ml_ratio",
               "likelihood": "log_like", "pendant_length": "pendant", "marginal_like": "marginal_like",
               "post_prob": "post_prob"}


def fmt_float(x: float) -> str:
    """Jsontype.Float -> '%.12g' (json.ml:41); OCaml prints nan / inf / -inf for non-finite values."""
    x = float(x)
    if x != x:
        return "nan"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    return "%.12g" % x


_ESC = {"\\": "\\\\", '"': '\\"', "/": "\\/", "\b": "\\b", "\f": "\\f", "\n": "\\n", "\r": "\\r", "\t": "\\t"}


def quote(s: str) -> str:
    """json.ml:25-36."""
    return '"' + "".join(_ESC.get(ch, ch) for ch in s) + '"'


def placements_to_rows(results: np.ndarray, n_out: int, fields: Sequence[str]) -> List[List[str]]:
    rows = []
    for r in results[:n_out]:
        row = []
        for f in fields:
            v = r[_RESULT_KEY[f]]
            row.append("%d" % int(v) if f == "edge_num" else fmt_float(v))
        rows.append(row)
    return rows


def to_jplace_string(tree, query_names: List[List[str]], results: np.ndarray, n_out: np.ndarray,
                     status: Optional[np.ndarray] = None, pp: bool = False, invocation: str = "pplacer_b200") -> str:
    """Render a placerun.  `query_names[q]` are the names sharing unique query q (the reference
    re-duplicates after placement, pplacer_run.ml:530-560: one pquery per unique sequence with all of
    its names, multiplicity 1 each).  Queries with non-zero status or no placements are omitted, as
    the reference omits sequences for which Core returned nothing."""
    fields = PP_FIELDS if pp else BASE_FIELDS
    out = ['{"tree":\n  ', quote(tree.to_newick(with_node_numbers=True)), ',\n  "placements":\n  [']
    first = True
    for q, names in enumerate(query_names):
        if (status is not None and status[q] != 0) or n_out[q] == 0:
            continue
        rows = placements_to_rows(results[q], int(n_out[q]), fields)
        p = ", ".join("[" + ", ".join(r) + "]" for r in rows)
        nm = ", ".join("[%s, 1]" % quote(n) for n in names)
        out.append(("" if first else ",") + '\n    {"p": [' + p + '],\n      "nm": [' + nm + "]\n    }")
        first = False
    out.append('\n  ], "metadata":\n  {"invocation":\n    ' + quote(invocation) + '\n  }, "version": 3, "fields":\n  [')
    out.append(", ".join(quote(f) for f in fields))
    out.append("\n  ]\n}\n")
    return "".join(out)


def write_jplace(path: str, *a, **kw) -> None:
    s = to_jplace_string(*a, **kw)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as fh:
        fh.write(s)


# ---------------------------------------------------------------------------------------------
# reading + comparison
# ---------------------------------------------------------------------------------------------
@dataclass
class Pquery:
    names: List[str]
    placements: List[Dict[str, float]]


def read_jplace(path: str) -> Tuple[str, List[str], List[Pquery]]:
    """Placerun_io.of_json_file for versions 1-3: 'nm' [[name, mult]], 'n' [name] or a single name."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        j = json.load(fh)
    return parse_jplace(j)


def parse_jplace(j: dict) -> Tuple[str, List[str], List[Pquery]]:
    fields = list(j["fields"])
    pqs: List[Pquery] = []
    for pq in j["placements"]:
        if "nm" in pq:
            names = [x[0] for x in pq["nm"]]
        else:
            n = pq["n"]
            names = [n] if isinstance(n, str) else list(n)
        pqs.append(Pquery(names, [dict(zip(fields, row)) for row in pq["p"]]))
    return j["tree"], fields, pqs


@dataclass
class CompareReport:
    n_queries: int = 0
    n_missing: int = 0
    n_best_edge_differs: int = 0
    n_kept_set_differs: int = 0
    near_ties: List[str] = field(default_factory=list)
    failures: List[str] = field(default_factory=list)
    max_ll_rel: float = 0.0
    max_bl_abs: float = 0.0
    max_ratio_abs: float = 0.0

    @property
    def ok(self) -> bool:
        return not self.failures

    def summary(self) -> str:
        return ("%d queries: %d missing, best edge differs in %d, kept set differs in %d (%d explained as near-ties), "
                "max |dLL|/|LL| %.3g, max |d branch| %.3g, max |d ratio| %.3g, %d failures"
                % (self.n_queries, self.n_missing, self.n_best_edge_differs, self.n_kept_set_differs,
                   len(self.near_ties), self.max_ll_rel, self.max_bl_abs, self.max_ratio_abs, len(self.failures)))


def compare(got: Iterable[Pquery], want: Iterable[Pquery], ll_rel: float = 1e-6, bl_abs: float = 1e-4,
            ratio_abs: float = 1e-5, near_tie_ll: float = 1e-3, keep_factor: float = 0.01) -> CompareReport:
    """Compare two placement sets query by query (matched on the first name).
    Near-tie rule (documented exception of the north_star): an edge kept on one side only is
    accepted when its log-likelihood is within `near_tie_ll` of the keep threshold
    (best + log keep_factor) or of a kept neighbour on the other side; a differing best edge is
    accepted when the two best log-likelihoods differ by less than `near_tie_ll`."""
    rep = CompareReport()
    gmap = {n: pq for pq in got for n in pq.names}
    for w in want:
        rep.n_queries += 1
        g = gmap.get(w.names[0])
        tag = w.names[0]
        if g is None:
            rep.n_missing += 1
            rep.failures.append("%s: missing" % tag)
            continue
        ge = {int(p["edge_num"]): p for p in g.placements}
        we = {int(p["edge_num"]): p for p in w.placements}
        gb, wb = g.placements[0], w.placements[0]
        near = False
        if int(gb["edge_num"]) != int(wb["edge_num"]):
            rep.n_best_edge_differs += 1
            if abs(gb["likelihood"] - wb["likelihood"]) <= near_tie_ll:
                near = True
            else:
                rep.failures.append("%s: best edge %d vs %d" % (tag, gb["edge_num"], wb["edge_num"]))
        if set(ge) != set(we):
            rep.n_kept_set_differs += 1
            thr = wb["likelihood"] + np.log(keep_factor)
            lows = [p["likelihood"] for p in w.placements] + [p["likelihood"] for p in g.placements]
            for e in set(ge) ^ set(we):
                p = ge.get(e) or we.get(e)
                cut = min(lows)
                if abs(p["likelihood"] - thr) <= near_tie_ll or abs(p["likelihood"] - cut) <= near_tie_ll:
                    near = True
                else:
                    rep.failures.append("%s: edge %d kept on one side only (LL %.6f, threshold %.6f)"
                                        % (tag, e, p["likelihood"], thr))
        if near:
            rep.near_ties.append(tag)
        for e in set(ge) & set(we):
            a, b = ge[e], we[e]
            d_ll = abs(a["likelihood"] - b["likelihood"]) / max(abs(b["likelihood"]), 1e-300)
            d_bl = max(abs(a["pendant_length"] - b["pendant_length"]), abs(a["distal_length"] - b["distal_length"]))
            rep.max_ll_rel = max(rep.max_ll_rel, d_ll)
            rep.max_bl_abs = max(rep.max_bl_abs, d_bl)
            if d_ll > ll_rel:
                rep.failures.append("%s edge %d: likelihood %.12g vs %.12g" % (tag, e, a["likelihood"], b["likelihood"]))
            if d_bl > bl_abs:
                rep.failures.append("%s edge %d: branch lengths differ by %.3g" % (tag, e, d_bl))
            if not near:
                keys = ["like_weight_ratio"] + (["post_prob"] if "post_prob" in a and "post_prob" in b
                                                and a["post_prob"] is not None and b["post_prob"] is not None else [])
                for k in keys:
                    d = abs(a[k] - b[k])
                    rep.max_ratio_abs = max(rep.max_ratio_abs, d)
                    if d > ratio_abs:
                        rep.failures.append("%s edge %d: %s %.8g vs %.8g" % (tag, e, k, a[k], b[k]))
    return rep


def pqueries_from_results(query_names: List[List[str]], results: np.ndarray, n_out: np.ndarray,
                          status: Optional[np.ndarray] = None, pp: bool = False) -> List[Pquery]:
    """The values a jplace reader would see (after the %.12g round trip)."""
    fields = PP_FIELDS if pp else BASE_FIELDS
    out = []
    for q, names in enumerate(query_names):
        if (status is not None and status[q] != 0) or n_out[q] == 0:
            continue
        rows = placements_to_rows(results[q], int(n_out[q]), fields)
        out.append(Pquery(list(names), [{f: (int(v) if f == "edge_num" else float(v)) for f, v in zip(fields, r)}
                                        for r in rows]))
    return out
