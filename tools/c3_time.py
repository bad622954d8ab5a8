#!/usr/bin/env python
"""C3 (donovan-test ar53: WAG + Price-CAT 20, 6968 taxa x 8035 columns, 3 bundled genomes) on one GPU:
precompute and placement times, and the comparison with the reference's GOLDEN jplace.  Needs the
git-ignored fixture tests/golden/_ar53 (tests/golden/make_fixtures.py /root/reference --ar53)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pplacer_b200 import configs, jplace
from pplacer_b200.placer import Placer

t0 = time.perf_counter(); prob = configs.ar53(); t_load = time.perf_counter() - t0
gold = configs.ar53_golden()["GOLDEN"]
want = jplace.parse_jplace({"tree": "", "fields": gold["fields"], "placements": gold["placements"]})[2]
with Placer(prob) as p:
    t0 = time.perf_counter(); p.precompute(); torch.cuda.synchronize(); t_pre = time.perf_counter() - t0
    times = []
    for _ in range(3):
        p.phase_times_ms(reset=True)
        t0 = time.perf_counter(); out, n_out, st, _ = p.place(prob.query_chars, p.prefs()); torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    phases = p.phase_times_ms()
got = jplace.pqueries_from_results(prob.query_names, out, n_out, st)
rep = jplace.compare(got, want)
inf = int(((prob.query_chars != ord("-")) & (prob.query_chars != ord("?"))).sum())
print(json.dumps({"workload": "C3 ar53 (WAG + Price-CAT 20): N=6968, E=13933, L=%d, 3 bundled genomes" % prob.L,
                  "load_s": t_load, "precompute_s": t_pre, "place_s": times, "phase_ms_last": [float(x) for x in phases],
                  "informative_sites": inf, "golden_compare": rep.summary(), "golden_ok": rep.ok,
                  "glv_bytes": 3 * prob.tree.n_nodes * 20 * (prob.L + 31) // 32 * 32 * 8}))
