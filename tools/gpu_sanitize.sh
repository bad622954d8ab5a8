#!/bin/bash
# compute-sanitizer over the second-stage engines on a small problem (memcheck, then racecheck on shared memory)
set -u
TAG=${1:-san}
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from pplacer_b200 import configs
from pplacer_b200.placer import Placer
prob = configs.synthetic(40, 200, 12, seed=7, min_len=40)
with Placer(prob) as p:
    p.precompute()
    for eng in (3, 2):
        p.set_three_tax_engine(eng)
        out, n, st, _ = p.place(prob.query_chars, p.prefs())
        out2, n2, st2, _ = p.place(prob.query_chars[:4], p.prefs(pp=True))
        print("engine", eng, "placed", int((st == 0).sum()), "kept", n.tolist(), "pp kept", n2.tolist())
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_case.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|engine|Invalid|out of bounds" gpurun_out/${TAG}_memcheck.log | head -12
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python /tmp/san_case.py > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|engine|ERROR|WARN" gpurun_out/${TAG}_racecheck.log | head -30
