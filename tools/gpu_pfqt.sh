#!/bin/bash
set -u
TAG=${1:-qt}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "prefilter or zero_cat or c2_properties or phyml or edge_cases" 2>&1 | tail -2
for G in ${GEOMS:-0 3 1}; do
for W in c2 c4; do
PPB_PF_GEOM=$G timeout 300 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${W}_$G.json 2> gpurun_out/${TAG}_bench_${W}_$G.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_${W}_$G.json")); print("geom=$G $W:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()})
except Exception as e: print("parse failed", e)
PY
done
done
