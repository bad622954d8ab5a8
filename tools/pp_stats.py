#!/usr/bin/env python
"""How much work is a --pp placement?  Runs the CPU oracle on the first N queries of C5 and prints, per placement, the
number of outer (distal) and inner (pendant) integrand evaluations of the nested QNG and the ETOL retries."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as ob
from pplacer_b200 import configs, reference
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prob = configs.c2(Q=n)
o = ob.Oracle(prob); o.precompute()
pf = ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree), calc_pp=1, keep_at_most=40, keep_factor=0.001, max_strikes=20)
L = ob.lib(); st = (C.c_longlong * 4)()
L.orc_pp_stats(st, 1)
res, n_out, status, _ = o.place(prob.query_chars, pf, n_threads=os.cpu_count())
L.orc_pp_stats(st, 0)
pl, outer, inner, retry = list(st)
print("placements %d (kept per query %.2f)  outer evals/placement %.1f  inner evals/outer %.1f  inner evals/placement %.0f  ETOL retries/placement %.3f"
      % (pl, n_out.mean(), outer / pl, inner / max(outer, 1), inner / pl, retry / pl))
