#!/usr/bin/env python
"""Run the CPU oracle on the reference's only end-to-end regression (donovan-test ar53, WAG +
Price-CAT, 6968 taxa x 8062 columns, 3 genomes) and compare with ar53_test_genomes.GOLDEN.json.
Needs /root/reference (this container only), ~56 GB of file-backed GLV arrays under --mmap-dir and
several minutes of CPU.  Writes tests/golden/ar53_oracle.json (small, committed) so the result
travels; see tests/test_oracle_ar53.py."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from pplacer_b200 import newick, seqio, models, reference
from oracle import binding as ob

ap = argparse.ArgumentParser()
ap.add_argument("--base", default="/root/reference/donovan-test/ar53_pplacer_testset")
ap.add_argument("--mmap-dir", default="/tmp")
ap.add_argument("--threads", type=int, default=3)
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ar53_oracle.json"))
a = ap.parse_args()
rp = os.path.join(a.base, "gtdb_r226_ar53.refpkg")
contents = json.load(open(os.path.join(rp, "CONTENTS.json")))["files"]
t0 = time.time()
tree = newick.parse_file(os.path.join(rp, contents["tree"]))
aln = seqio.read_fasta(os.path.join(rp, contents["aln_fasta"]))
model = models.model_from_json_file(os.path.join(rp, contents["phylo_model"]))
queries = seqio.read_fasta(os.path.join(a.base, "ar53-test_genomes.fasta.gz"))
prob = reference.build_problem(tree, model, aln, queries)
print("parsed: nodes", tree.n_nodes, "L_raw", prob.L_raw, "L", prob.L, "Q", len(prob.query_chars), "%.1fs" % (time.time() - t0), flush=True)
o = ob.Oracle(prob, mmap_dir=a.mmap_dir)
t0 = time.time(); o.precompute(); t_pre = time.time() - t0
print("precompute %.1fs" % t_pre, flush=True)
t0 = time.time()
out, n_out, status, info = o.place(prob.query_chars, ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(tree)), n_threads=a.threads, debug=True)
t_place = time.time() - t0
print("place %.1fs" % t_place, status, n_out, flush=True)
gold = json.load(open(os.path.join(a.base, "ar53_test_genomes.GOLDEN.json")))
test = json.load(open(os.path.join(a.base, "ar53_test_genomes.TEST.json")))
fields = gold["fields"]
def tab(j):
    return {pq["nm"][0][0] if "nm" in pq else pq["n"][0]: [dict(zip(j["fields"], p)) for p in pq["p"]] for pq in j["placements"]}
G, T = tab(gold), tab(test)
res = {"L": prob.L, "n_nodes": tree.n_nodes, "t_precompute_s": t_pre, "t_place_s": t_place, "queries": {}}
for qi, names in enumerate(prob.query_names):
    ps = [{k: (int(out[qi][j][k]) if k == "edge" else float(out[qi][j][k])) for k in ("edge", "log_like", "pendant", "distal", "ml_ratio")} for j in range(n_out[qi])]
    res["queries"][names[0]] = {"status": int(status[qi]), "n_q": int(seqio.informative(prob.query_chars[qi]).sum()),
                                "placements": ps, "n_pitches": int(info["n_pitches"][qi]), "n_like_calls": int(info["n_like_calls"][qi])}
    g = G[names[0]][0]
    p = ps[0]
    print(names[0], "oracle", p, "\n   golden", g, "\n   test  ", T[names[0]][0])
json.dump(res, open(a.out, "w"), indent=1)
print("wrote", a.out)
