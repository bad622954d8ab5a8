#!/bin/bash
# visit r5v: pendant passes with the global overflow of y read one site ahead; SiteProd slow path marked unlikely
set -u
TAG=${1:-r5v}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for W in c2 c4; do
timeout 200 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err; echo "bench $W rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_$W.json")); r=j["roofline"]
    print("$W", round(j["value"]), "pl/s", round(j["ms_per_step"],2), "ms", {k: round(v,2) for k,v in r["share_of_step_ms"].items()}, "frac", round(r["frac"],4))
except Exception as e: print("$W parse failed", e)
PY
done
