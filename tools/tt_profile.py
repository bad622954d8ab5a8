#!/usr/bin/env python3
"""Per-phase cycle profile of the second-stage block engine (library built with -DPPB_TT_PROFILE).
usage: PPB_NVCC_FLAGS=-DPPB_TT_PROFILE python -m pplacer_b200.build --force && python tools/tt_profile.py [workload] [Q]
Prints where thread 0 of each block spent its cycles (scalar optimiser, matrix builds, passes by command)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from pplacer_b200.placer import Placer

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
prob = bench.build_workload(wl, Q, 0)
p = Placer(prob, device=0)
p.precompute()
w = bench.WORKLOADS[wl]
prefs = p.prefs(pp=w[4], **w[5])
for _ in range(2):
    p.place(prob.query_chars, prefs)
p.three_tax_profile(reset=True)
p.phase_times_ms(reset=True)
t0 = time.perf_counter()
p.place(prob.query_chars, prefs)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
pr = p.three_tax_profile()
ph = p.phase_times_ms()
if p.three_tax_engine_used() == 4:
    names = {1: "PEND (1 or 5 points)", 2: "BUILD_Y + PEND", 3: "DIST1", 4: "DIST4"}
else:
    names = {1: "DIST1", 2: "PEND1_FUSED", 3: "PEND1_CACHED", 4: "DIST_BATCH", 5: "PEND_BATCH_FUSED", 6: "PEND_BATCH_CACHED"}
tot = pr[25] or 1.0
out = {"workload": wl, "Q": Q, "wall_ms": wall * 1e3, "phase_ms": ph.tolist(), "candidates": pr[24],
       "pendant_calls": pr[0], "distal_calls": pr[1], "cycles_per_candidate": pr[25] / max(pr[24], 1),
       "scalar_frac": pr[4] / tot, "build_frac": pr[5] / tot, "passes_with_build": pr[6], "build_jobs": pr[7], "passes": {}}
for k, nm in names.items():
    out["passes"][nm] = {"n": pr[16 + k], "frac": pr[8 + k] / tot, "cycles_each": pr[8 + k] / max(pr[16 + k], 1)}
out["scalar_cycles_per_pass"] = pr[4] / max(sum(pr[17:23]), 1)
out["build_cycles_each"] = pr[5] / max(pr[6], 1)
print(json.dumps(out, indent=1))
