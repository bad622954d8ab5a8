#!/usr/bin/env python
"""Copy the reference's own known-answer DATA for this path into tests/golden/ so the tests can run
where /root/reference does not exist (the GPU box).  Data files only - no reference source code.

  like/{gtr,wag,jtt}/   alignment, PhyML tree, phylo_model JSON and PhyML statistics file of tests/pplacer/test_like.ml
                        (expected PhyML log-likelihoods: test_like.ml:71-90)
  ar53_golden.json      the three GOLDEN / TEST placements of donovan-test/ar53_pplacer_testset
  _ar53/ (git-ignored)  the ar53 refpkg (tree, model JSON, MSA as .npz) + the 3 query genomes, for
                        the optional GPU end-to-end regression; too large for history, still
                        travels to the GPU box with the repo snapshot.
"""
import gzip, json, os, shutil, sys
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

for d, files in (("gtr", ["hiv.fasta", "hiv.phy_phyml_tree.txt", "phylo_model.jplace", "hiv.phy_phyml_stats.txt"]),
                 ("wag", ["actin.fasta", "actin.phy_phyml_tree.txt", "phylo_model.jplace", "actin.phy_phyml_stats.txt"]),
                 ("jtt", ["actin.fasta", "actin.phy_phyml_tree.txt", "phylo_model.jplace", "actin.phy_phyml_stats.txt"])):
    os.makedirs(os.path.join(HERE, "like", d), exist_ok=True)
    for f in files:
        dst = os.path.join(HERE, "like", d, f)
        shutil.copyfile(os.path.join(REF, "tests", "data", "like", d, f), dst)
        os.chmod(dst, 0o644)

base = os.path.join(REF, "donovan-test", "ar53_pplacer_testset")
out = {}
for tag in ("GOLDEN", "TEST"):
    j = json.load(open(os.path.join(base, "ar53_test_genomes.%s.json" % tag)))
    import hashlib
    out[tag] = {"fields": j["fields"], "version": j.get("version"),
                "tree_sha256": hashlib.sha256(j["tree"].encode()).hexdigest(), "tree_len": len(j["tree"]),
                "invocation": j["metadata"]["invocation"],
                "placements": [{"nm": pq.get("nm", pq.get("n")), "p": pq["p"]} for pq in j["placements"]]}
json.dump(out, open(os.path.join(HERE, "ar53_golden.json"), "w"), indent=1)

if "--ar53" in sys.argv:
    from pplacer_b200 import seqio
    rp = os.path.join(base, "gtdb_r226_ar53.refpkg")
    dst = os.path.join(HERE, "_ar53")
    os.makedirs(dst, exist_ok=True)
    contents = json.load(open(os.path.join(rp, "CONTENTS.json")))["files"]
    shutil.copyfile(os.path.join(rp, contents["tree"]), os.path.join(dst, "tree.nwk"))
    shutil.copyfile(os.path.join(rp, contents["phylo_model"]), os.path.join(dst, "phylo_model.json"))
    shutil.copyfile(os.path.join(base, "ar53-test_genomes.fasta.gz"), os.path.join(dst, "queries.fasta.gz"))
    aln = seqio.read_fasta(os.path.join(rp, contents["aln_fasta"]))
    names = np.array([n for n, _ in aln])
    chars = seqio.to_char_matrix([s for _, s in aln])
    np.savez_compressed(os.path.join(dst, "msa.npz"), names=names, chars=chars)
    for f in os.listdir(dst):
        os.chmod(os.path.join(dst, f), 0o644)
    print("ar53 fixture:", {f: os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst)})
print("fixtures written under", HERE)
