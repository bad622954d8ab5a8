"""jplace writer / reader / comparator against the reference's own
output files: the ar53 GOLDEN / TEST jplace of donovan-test, whose "tree" string pins Newick
numbering, '%g' branch lengths, label quoting and '{n}' edge numbers.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from pplacer_b200 import _lib, configs, jplace, newick


def test_float_and_string_formatting():
    # json.ml:41 '%.12g'; json.ml:25-36 escapes '/' as '\/'
    assert jplace.fmt_float(-23657343.98917654) == "-23657343.9892"
    assert jplace.fmt_float(9.372762011718751e-06) == "9.37276201172e-06"
    assert jplace.fmt_float(1.0) == "1" and jplace.fmt_float(0.5) == "0.5"
    assert jplace.quote('/srv/x "y"\n') == '"\\/srv\\/x \\"y\\"\\n"'
    gold = configs.ar53_golden()["GOLDEN"]
    # the invocation string of the GOLDEN file round-trips through quote + json.loads
    assert json.loads(jplace.quote(gold["invocation"])) == gold["invocation"]


def _records(rows):
    out = np.zeros((len(rows), 7), dtype=_lib.RESULT_DTYPE)
    n = np.zeros(len(rows), np.int32)
    for q, pl in enumerate(rows):
        n[q] = len(pl)
        for i, (e, ll, pend, dist, w) in enumerate(pl):
            out[q, i] = (e, 0, ll, pend, dist, w, np.nan, np.nan)
    return out, n


def test_write_read_compare_round_trip(tmp_path):
    tree = newick.parse("((A:0.1,'b c':0.25):0.05,C:0.3,D:1e-06):0;")
    names = [["q1", "q1_dup"], ["q/2"], ["q3"]]
    out, n = _records([[(0, -1234.5678901234, 0.0123456789, 0.05, 0.75), (3, -1236.1, 0.02, 0.1, 0.25)],
                       [(2, -99.5, 1e-8, 0.0, 1.0)], []])
    st = np.array([0, 0, 104], np.int32)
    s = jplace.to_jplace_string(tree, names, out, n, st, invocation="pplacer_b200 -c x.refpkg q.fasta")
    j = json.loads(s)
    assert j["version"] == 3 and j["fields"] == jplace.BASE_FIELDS and sorted(j["fields"]) == j["fields"]
    assert j["tree"] == "((A:0.1{0},'b c':0.25{1}):0.05{2},C:0.3{3},D:1e-06{4}){5};"
    assert j["placements"][0]["nm"] == [["q1", 1], ["q1_dup", 1]]
    assert j["placements"][0]["p"][0] == [0.05, 0, 0.75, -1234.56789012, 0.0123456789]
    assert len(j["placements"]) == 2                     # the query without placements is omitted
    assert "\\/" in s                                      # '/' escaped like the reference writer
    path = str(tmp_path / "o.jplace")
    jplace.write_jplace(path, tree, names, out, n, st)
    _, fields, pqs = jplace.read_jplace(path)
    rep = jplace.compare(pqs, jplace.pqueries_from_results(names, out, n, st))
    assert rep.ok and rep.n_queries == 2 and rep.max_ll_rel == 0.0, rep.summary()
    # tolerances: 1e-4 on branch lengths, 1e-6 relative on likelihood, 1e-5 on ratios
    out2 = out.copy(); out2[0, 0]["pendant"] += 5e-5; out2[1, 0]["log_like"] *= 1 + 5e-7
    assert jplace.compare(jplace.pqueries_from_results(names, out2, n, st), pqs).ok
    out3 = out.copy(); out3[0, 0]["pendant"] += 2e-4
    assert not jplace.compare(jplace.pqueries_from_results(names, out3, n, st), pqs).ok
    out4 = out.copy(); out4[0, 0]["edge"] = 1
    rep = jplace.compare(jplace.pqueries_from_results(names, out4, n, st), pqs)
    assert not rep.ok and rep.n_best_edge_differs == 1
    # -p adds marginal_like and post_prob, fields stay sorted
    outp = out.copy(); outp["marginal_like"] = -1240.0; outp["post_prob"] = 0.5
    jp = json.loads(jplace.to_jplace_string(tree, names, outp, n, st, pp=True))
    assert jp["fields"] == sorted(jp["fields"]) == jplace.PP_FIELDS and len(jp["placements"][0]["p"][0]) == 7


def test_reference_golden_and_test_jplace_agree_under_the_comparator():
    """The reference's own two runs of the ar53 test set (GOLDEN from 2025, TEST re-run) are
    'identical' under the north_star tolerances - the comparator must say so."""
    g = configs.ar53_golden()
    mk = lambda d: jplace.parse_jplace({"tree": "", "fields": d["fields"], "placements": d["placements"]})[2]
    rep = jplace.compare(mk(g["TEST"]), mk(g["GOLDEN"]))
    assert rep.ok and rep.n_queries == 3, rep.summary()


def test_ar53_tree_string_matches_golden_jplace():
    """Newick parse -> numbering -> Newick_gtree.to_string ~with_node_numbers:true on the ar53
    refpkg tree (1416 quoted labels with ':' and ';', bootstrap values as node labels) reproduces
    the GOLDEN jplace "tree" byte for byte."""
    path = os.path.join(configs.GOLDEN, "_ar53", "tree.nwk")
    if not os.path.exists(path):
        pytest.skip("tests/golden/_ar53 fixture not present (make_fixtures.py /root/reference --ar53)")
    gold = configs.ar53_golden()["GOLDEN"]
    s = newick.parse_file(path).to_newick(with_node_numbers=True)
    assert len(s) == gold["tree_len"]
    assert hashlib.sha256(s.encode()).hexdigest() == gold["tree_sha256"]
