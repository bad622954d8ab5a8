"""Host-side logic either side of the kernels: Newick numbering, pre-masking, de-duplication,
alphabet checks, config builders, query sharding.  CPU only."""
import numpy as np
import pytest

from pplacer_b200 import configs, models, newick, reference, seqio, shard, synth


def test_newick_postorder_numbering():
    # newick_parser.mly:49-62: ids assigned in post-order = jplace edge numbers
    t = newick.parse("((A:2,B:9):7,C:5,D:1):0;")
    assert t.n_nodes == 6 and t.root == 5
    assert t.children[5] == [2, 3, 4] and t.children[2] == [0, 1]
    assert [t.node_label[i] for i in (0, 1, 3, 4)] == ["A", "B", "C", "D"]
    assert np.allclose(t.bl, [2, 9, 7, 5, 1, 0])
    assert t.to_newick() == "((A:2{0},B:9{1}):7{2},C:5{3},D:1{4}){5};"
    # quoted labels may hold ':' ';' ',' (ar53 has 1416 of them); {n} overrides numbering
    t2 = newick.parse("('x:1;y',(b,'c d')'100.0:p__X;c__Y':0.5);")
    assert t2.node_label[0] == "x:1;y" and t2.node_label[3] == "100.0:p__X;c__Y"
    t3 = newick.parse("((A:1{1},B:2{0}):1{2},C:3{3}){4};")
    assert t3.node_label[1] == "A" and t3.node_label[0] == "B"
    rt = newick.parse(t3.to_newick())
    assert rt.children == t3.children
    with pytest.raises(newick.NewickError):
        newick.parse("((A,B);")


def test_premask_and_dedup():
    tree = newick.parse("((A:0.1,B:0.2):0.1,C:0.3,D:0.1):0;")
    m = models.ModelSpec(4, 0, "GTR", models.gtr_exch(1, 2, 1, 1, 2, 1), None, True, 4, 0.5, None, None)
    ref = [("A", "AC-GT-A"), ("B", "AC-GA-A"), ("C", "AT-GT-C"), ("D", "AC-GT-A")]
    qs = [("q1", "ac-g---"), ("q2", "AC-G---"), ("q3", "-CAG---")]
    p = reference.build_problem(tree, m, ref, qs)
    # columns 2 and 5 are all-gap in the reference, 4 and 6 all-gap in the queries
    assert list(np.nonzero(p.mask)[0]) == [0, 1, 3]
    assert p.L == 3 and p.L_raw == 7
    assert p.query_names == [["q1", "q2"], ["q3"]]            # upper-cased then de-duplicated
    assert p.query_chars.tobytes() == b"ACG-CG"
    assert list(p.leaf_row[:2]) == [0, 1]
    with pytest.raises(ValueError, match="not a known base"):
        reference.build_problem(tree, m, ref, [("bad", "ACJGT-A")])
    with pytest.raises(ValueError, match="not the same length"):
        reference.build_problem(tree, m, ref, [("short", "ACG")])
    with pytest.raises(ValueError, match="doesn't overlap"):
        reference.build_problem(tree, m, ref, [("nov", "--A----")])


def test_gcat_site_categories_follow_premask():
    prob = configs.gcat_variant("gtr")
    assert prob.model.model_class == 1
    assert len(prob.site_cat) == prob.L
    assert np.array_equal(prob.site_cat, prob.model.site_cat[prob.mask])


def test_synthetic_problem_shapes():
    prob = configs.synthetic(24, 120, 16, seed=3, min_len=30)
    assert prob.tree.n_nodes == 2 * 24 - 2 and len(prob.tree.children[prob.tree.root]) == 3
    assert prob.ref_chars.shape[0] == 24 and prob.query_chars.shape[1] == prob.L
    inf = seqio.informative(prob.query_chars).sum(axis=1)
    assert inf.min() >= 1
    again = configs.synthetic(24, 120, 16, seed=3, min_len=30)
    assert np.array_equal(prob.query_chars, again.query_chars)  # seeded


def test_shard_ranges_cover_and_balance():
    for Q, W in ((10, 1), (10, 3), (7, 8), (1_000_003, 8), (0, 4)):
        spans = [shard.shard_range(Q, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == Q
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_refpkg_loader_and_error_messages(tmp_path):
    """refpkg_parse.ml:86-104 / refpkg.ml:26-63: CONTENTS.json -> tree, aln_fasta, phylo_model."""
    import json, os, shutil
    from pplacer_b200 import refpkg
    src = os.path.join(configs.GOLDEN, "like", "gtr")
    rp = tmp_path / "hiv.refpkg"
    rp.mkdir()
    for f in ("hiv.fasta", "hiv.phy_phyml_tree.txt", "phylo_model.jplace"):
        shutil.copyfile(os.path.join(src, f), rp / f)
    files = {"aln_fasta": "hiv.fasta", "tree": "hiv.phy_phyml_tree.txt", "phylo_model": "phylo_model.jplace"}
    (rp / "CONTENTS.json").write_text(json.dumps({"files": files}))
    tree, aln, model = refpkg.load(str(rp))
    assert tree.n_nodes == 38 and len(aln) == 20 and model.n_states == 4 and model.n_rates == 4
    assert all(s == s.upper() for _, s in aln)                      # refpkg.ml:37 Alignment.uppercase
    with pytest.raises(ValueError, match="not a reference package"):
        refpkg.load(str(tmp_path))
    (rp / "CONTENTS.json").write_text(json.dumps({"files": {k: v for k, v in files.items() if k != "phylo_model"}}))
    with pytest.raises(ValueError, match="phylo_model"):
        refpkg.load(str(rp))


def test_phyml_stats_file_gives_the_json_model(tmp_path):
    """parse_stats.ml:223-276 on the reference's own PhyML statistics files (tests/data/like/*): the model they describe
    is the model of the phylo_model JSON next to them (same rates, alpha, categories)."""
    import os
    from pplacer_b200 import stats
    for name, stem in (("gtr", "hiv"), ("wag", "actin"), ("jtt", "actin")):
        d = os.path.join(configs.GOLDEN, "like", name)
        got = stats.model_from_stats_file(os.path.join(d, stem + ".phy_phyml_stats.txt"))
        want = models.model_from_json_file(os.path.join(d, "phylo_model.jplace"))
        assert (got.n_states, got.model_class, got.subs_model, got.n_rates) == (want.n_states, 0, want.subs_model, want.n_rates)
        assert got.gamma_alpha == want.gamma_alpha and np.array_equal(got.exch, want.exch)
    assert stats.parse_stats(os.path.join(configs.GOLDEN, "like", "gtr", "hiv.phy_phyml_stats.txt")) == \
        ("GTR", 4, 0.771, [1.17728, 3.08642, 0.60698, 0.92672, 4.74366, 1.0])
    # RAxML 7.2.x info file (parse_stats.ml:84-126) and the re-estimated form (:129-190)
    info = tmp_path / "RAxML_info.x"
    info.write_text("This is RAxML version 7.2.6 released by Alexandros Stamatakis\n\nPartition: 0\nAlignment Patterns: 3\nName: No Name Provided\n"
                    "DataType: DNA\nSubstitution Matrix: GTR\n\nInference[0]: Time 1.0 GAMMA-based likelihood -1.0, best rearrangement setting 5\n"
                    "alpha[0]: 0.5421 rates[0] ac ag at cg ct gt: 1.1 2.2 3.3 4.4 5.5 1.000000\n")
    assert stats.parse_stats(str(info)) == ("GTR", 4, 0.5421, [1.1, 2.2, 3.3, 4.4, 5.5, 1.0])
    info.write_text("This is RAxML version 7.2.6\n\nraxmlHPC -f e -t tree -m GTRGAMMA -s aln -n x\n\nPartition: 0\nAlignment Patterns: 3\n"
                    "DataType: AA\nSubstitution Matrix: WAG\nalpha: 0.9\n")
    assert stats.parse_stats(str(info)) == ("WAG", 4, 0.9, None)
    info.write_text("nothing useful\n")
    with pytest.raises(stats.StatsParsingError):
        stats.parse_stats(str(info))


def test_refpkg_with_stockholm_alignment_and_stats_file(tmp_path):
    """refpkg.ml:33-36,51-59: aln_sto is read when aln_fasta is absent ('.' gaps -> '-', interleaved blocks joined, names
    sorted), and a tree_stats file stands in for the phylo_model JSON."""
    import json, os, shutil
    from pplacer_b200 import refpkg
    src = os.path.join(configs.GOLDEN, "like", "gtr")
    aln = seqio.read_fasta(os.path.join(src, "hiv.fasta"))
    rp = tmp_path / "sto.refpkg"
    rp.mkdir()
    half = len(aln[0][1]) // 2
    with open(rp / "hiv.sto", "w") as fh:
        fh.write("# STOCKHOLM 1.0\n#=GF ID hiv\n\n")
        for n, s in aln:
            fh.write("%s  %s\n" % (n, s[:half].replace("-", ".")))
        fh.write("#=GC RF xxxx\n\n")
        for n, s in aln:
            fh.write("%s\t%s   \n" % (n, s[half:].lower()))
        fh.write("//\n")
    shutil.copyfile(os.path.join(src, "hiv.phy_phyml_tree.txt"), rp / "t.nwk")
    shutil.copyfile(os.path.join(src, "hiv.phy_phyml_stats.txt"), rp / "stats.txt")
    (rp / "CONTENTS.json").write_text(json.dumps({"files": {"aln_sto": "hiv.sto", "tree": "t.nwk", "tree_stats": "stats.txt"}}))
    tree, got_aln, model = refpkg.load(str(rp))
    assert got_aln == sorted((n, s.upper()) for n, s in aln)
    want = models.model_from_json_file(os.path.join(src, "phylo_model.jplace"))
    assert np.array_equal(model.exch, want.exch) and model.gamma_alpha == want.gamma_alpha and model.n_rates == 4
    (rp / "bad.sto").write_text("A  ACGT\n//\n")
    with pytest.raises(ValueError, match="before header"):
        seqio.read_stockholm(str(rp / "bad.sto"))


def test_figs_of_tree_by_hand():
    """Fig._figs_of_gtree (fig.ml:70-99) on a tree small enough to work through by hand (see the comments)."""
    from pplacer_b200 import fig
    t = newick.parse("((A:1,B:1):1,(C:1,D:1):1,E:5):0;")      # ids: A0 B1 (AB)2 C3 D4 (CD)5 E6 root7
    assert fig.figs_of_tree(0.0, t) is None                    # Dummy
    # cutoff 2: A-B and C-D are 2 apart -> two figs of two leaves, E alone; closed at the root: only leaf edges inside
    assert fig.figs_of_tree(2.0, t) == [(6, [6]), (3, [3, 4]), (0, [0, 1])]
    # cutoff 0.5: every cherry splits at its own node, which is added to both halves (add_edge_to_figs); the root then
    # appends its children's (now empty) leftovers after the accumulated figs, in reverse child order
    assert fig.figs_of_tree(0.5, t) == [(4, [4, 5]), (3, [3, 5]), (1, [1, 2]), (0, [0, 2]), (6, [6]), (3, []), (0, [])]
    # the enumeration: representatives sorted, groups merged within strike_box, members sorted, nothing yielded twice
    score = {0: -10.0, 1: -11.0, 2: -10.5, 3: -20.0, 4: -19.0, 5: -30.0, 6: -12.5}.get
    figs = fig.figs_of_tree(0.5, t)
    order = [e for _, e in fig.enum_by_score(figs, score, 3.0, t.nonroot_ids())]
    # reps sorted: 0(-10) 0(-10, the empty fig) 1(-11) 6(-12.5) 4(-19) 3(-20) 3(-20); group 1 = {0,2} u {} u {1,2} u {6} (all within 3 of -10)
    assert order == [0, 2, 1, 6, 4, 3, 5]
    assert [e for _, e in fig.enum_by_score(None, score, 3.0, t.nonroot_ids())] == [0, 2, 1, 6, 4, 3, 5]
    order = [e for _, e in fig.enum_by_score(figs, score, 0.6, t.nonroot_ids())]
    assert order == [0, 2, 1, 6, 4, 5, 3]                      # strike_box 0.6: {0,2} | {1,(2)} | {6} | {4,5} | {3,(5)}
