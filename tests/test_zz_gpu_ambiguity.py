"""Two- and three-base IUPAC codes through the log-table prefilter forms of the DNA Gamma-4 model (6: lanes on edges,
7: query lanes, 8: query lanes on the pair-of-sites table): these sites select the zero column in the main kernels and
get their exact terms from prefilter_qlane_fix_kernel / the inline slow term - same scores as ten_masked_logdot_c."""
import numpy as np
import pytest

from pplacer_b200 import configs, reference
from pplacer_b200.placer import Placer

pytestmark = pytest.mark.gpu


def test_ambiguity_codes_through_the_log_table_forms():
    from oracle import binding as ob
    prob = configs.c1()
    q = prob.query_chars.copy()
    codes = "RYMKSWHBVDN"
    rng = np.random.default_rng(5)
    for i in range(0, q.shape[0], 2):          # every other query gets ambiguity codes at random informative sites
        inf = np.flatnonzero((q[i] != ord("-")) & (q[i] != ord("?")))
        pick = rng.choice(inf, size=min(12, len(inf)), replace=False)
        q[i, pick] = np.frombuffer(rng.choice(list(codes), size=len(pick)).astype("S1").tobytes(), dtype=np.uint8)
    o = ob.Oracle(prob)
    o.precompute()
    _, _, st, info = o.place(q, ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree)), n_threads=4, debug=True)
    want = info["scores"]
    o.close()
    root = prob.tree.root
    keep = [i for i in range(prob.tree.n_nodes) if i != root]
    with Placer(prob) as p:
        p.precompute()
        for kernel in (1, 6, 7, 8):
            got = p.prefilter_scores(q, kernel=kernel)
            assert p.prefilter_kernel_used() == kernel
            g, w = got[:, keep], want[:, keep]
            assert np.all(np.abs(g - w) <= 1e-10 * np.abs(w)), (kernel, float(np.abs((g - w) / w).max()))
