#!/usr/bin/env python
"""How far do the second-stage engines sit from the oracle, and why?  Places the first N queries of a workload with the
oracle and with each engine, and prints the distribution of the differences (log-likelihood, branch lengths, ml_ratio,
objective calls).  Differences beyond rounding come from Brent taking a different path at a comparison of two objective
values that agree to rounding (the reference's own conditioning): this script measures how often and how far.
  python tools/parity_stats.py c2 512 3 4
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob                      # noqa: E402  (diagnostic tool: the oracle is the checker)
from pplacer_b200 import configs, reference           # noqa: E402
from pplacer_b200.placer import Placer                # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    engines = [int(a) for a in sys.argv[3:]] or [3, 4]
    prob = configs.c2(Q=n) if wl == "c2" else configs.c4(Q=n)
    o = ob.Oracle(prob)
    o.precompute()
    op = ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree))
    exp, n_exp, st_exp, oi = o.place(prob.query_chars, op, n_threads=os.cpu_count(), debug=True)
    rep = {"workload": wl, "queries": n}
    with Placer(prob) as p:
        p.precompute()
        for eng in engines:
            p.set_three_tax_engine(eng)
            got, n_got, st_got, gi = p.place(prob.query_chars, p.prefs(), debug=True)
            dll, dpend, ddist, dratio, worst = [], [], [], [], []
            sets_differ = best_differ = 0
            for q in range(n):
                g, e = got[q][:n_got[q]], exp[q][:n_exp[q]]
                ge = {int(r["edge"]): r for r in g}
                ee = {int(r["edge"]): r for r in e}
                if set(ge) != set(ee):
                    sets_differ += 1
                if int(g[0]["edge"]) != int(e[0]["edge"]):
                    best_differ += 1
                for edge in set(ge) & set(ee):
                    a, b = ge[edge], ee[edge]
                    dll.append(abs(a["log_like"] - b["log_like"]))
                    dpend.append(abs(a["pendant"] - b["pendant"]))
                    ddist.append(abs(a["distal"] - b["distal"]))
                    dratio.append(abs(a["ml_ratio"] - b["ml_ratio"]))
                    if dratio[-1] > 1e-6:
                        worst.append((float(dratio[-1]), q, edge, float(dll[-1]), float(dpend[-1]), float(ddist[-1]),
                                      int(gi["n_like_calls"][q]), int(oi["n_like_calls"][q])))
            dll, dpend, ddist, dratio = map(np.array, (dll, dpend, ddist, dratio))
            pct = lambda v: {"p50": float(np.percentile(v, 50)), "p99": float(np.percentile(v, 99)), "max": float(v.max())}
            rep["engine_%d" % eng] = {
                "kept_sets_differ": sets_differ, "best_edge_differs": best_differ, "records": int(len(dll)),
                "abs_dll": pct(dll), "abs_dpend": pct(dpend), "abs_ddist": pct(ddist), "abs_dratio": pct(dratio),
                "records_dratio_gt_1e-5": int((dratio > 1e-5).sum()), "records_dll_gt_1e-6": int((dll > 1e-6).sum()),
                "queries_same_call_count": float((gi["n_like_calls"] == oi["n_like_calls"]).mean()),
                "pitch_counts_equal": float((gi["n_pitches"] == oi["n_pitches"]).mean()),
                "worst": sorted(worst, reverse=True)[:12],
            }
        p.set_three_tax_engine(0)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
