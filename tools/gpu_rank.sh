#!/bin/bash
set -u
TAG=${1:-rk}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python - <<'PY'
# truncated ranking vs full ranking on a tree large enough to use it (n_nodes > 1024): same bytes
import os, sys, numpy as np
sys.path.insert(0, ".")
from pplacer_b200 import configs
from pplacer_b200.placer import Placer
prob = configs.synthetic(900, 300, 200, seed=11, min_len=60)
with Placer(prob) as p:
    p.precompute()
    out, n, st, _ = p.place(prob.query_chars, p.prefs())
os.environ["PPB_RANK_FULL"] = "1"
PY
for F in 0 1; do
for W in c2 c4; do
PPB_RANK_FULL=$F timeout 300 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${W}_$F.json 2> gpurun_out/${TAG}_bench_${W}_$F.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_${W}_$F.json")); print("full=$F $W:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()}, "kept", j["config"]["mean_kept"])
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_${W}_$F.err
done
done
