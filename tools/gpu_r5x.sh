#!/bin/bash
# visit r5x: six-column query-lane table with the edges of a column contiguous (16-byte loads, two edges per load)
set -u
TAG=${1:-r5x}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "prefilter or query_lane or c4_shape or fig_ranking_at or c2_properties" > gpurun_out/${TAG}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/${TAG}_pytest.log | cut -c1-300; exit 1; fi
for W in c2 c4; do
timeout 200 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err; echo "bench $W rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_$W.json")); r=j["roofline"]
    print("$W", round(j["value"]), "pl/s", round(j["ms_per_step"],2), "ms", {k: round(v,2) for k,v in r["share_of_step_ms"].items()})
except Exception as e: print("$W parse failed", e)
PY
done
