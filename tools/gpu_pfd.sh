#!/bin/bash
set -u
TAG=${1:-pfd}
mkdir -p gpurun_out
for G in ${GEOMS:-0 1 2 3 4}; do
PPB_PFD_GEOM=$G timeout 100 python -m pytest tests -m gpu -x -q -k "prefilter_kernel_variants" 2>&1 | tail -1
PPB_PFD_GEOM=$G timeout 300 python bench.py --prefilter direct --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$G.json 2> gpurun_out/${TAG}_bench_$G.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_$G.json")); r=j["roofline"]["other_kernel"]; print("dgeom=$G:", round(j["value"]), "prefilter ms %.1f"%j["roofline"]["share_of_step_ms"]["prefilter"], r["kernel"], "frac %.3f"%r["frac"])
except Exception as e: print("parse failed", e)
PY
done
