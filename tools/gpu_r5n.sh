#!/bin/bash
# visit r5n: engine-4 geometries with the eigenbasis gather passes
set -u
TAG=${1:-r5n}
mkdir -p gpurun_out
for V in ${VARIANTS:-1 2 3}; do
PPB_TTE_VARIANT=$V timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench_v$V.json 2> gpurun_out/${TAG}_bench_v$V.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_v$V.json")); r=j["roofline"]
    print("variant $V:", round(j["value"]), "pl/s", {k: round(v,2) for k,v in r["share_of_step_ms"].items()}, "frac", round(r["frac"],4))
except Exception as e: print("variant $V parse failed", e)
PY
done
