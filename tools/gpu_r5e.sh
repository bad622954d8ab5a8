#!/bin/bash
# visit r5e: speculative golden-section points in the pendant passes (C2 bench), ncu capture of the general block engine on C3
set -u
TAG=${1:-r5e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench.json")); r=j["roofline"]
print(round(j["value"]), "pl/s", round(j["ms_per_step"],2), "ms", {k: round(v,2) for k,v in r["share_of_step_ms"].items()}, "frac", round(r["frac"],4), r.get("objective_evaluations_per_step"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:three_tax_kernel -s 0 -c 1 -o gpurun_out/${TAG}_c3_full -f python bench.py --workload c3 --queries 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_c3_ncu.log 2>&1; echo "ncu c3 rc=$?"
tail -2 gpurun_out/${TAG}_c3_ncu.log | cut -c1-200
