#!/bin/bash
set -u
TAG=${1:-pf}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "prefilter or zero_cat or c2_properties or phyml or place_matches" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${TAG}_pytest.log)"; tail -12 gpurun_out/${TAG}_pytest.log | head -11
for W in c2 c4 c4_long; do
timeout 300 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_$W.json")); print("$W:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()}, "kept", j["config"]["mean_kept"], "ok", j["config"]["placed_ok_frac"])
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_$W.err
done
