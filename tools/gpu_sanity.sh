#!/bin/bash
# last check of the tree as it stands: smoke + the GPU parity suite + one short default bench line
set -u
TAG=${1:-san}
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench.json")); r=j["roofline"]
print(round(j["value"]), "pl/s", round(j["ms_per_step"],2), "ms e2e", round(j["e2e"]["value"]), "frac", round(r["frac"],4), "traffic", r.get("traffic"), "direct", round(r["prefilter_direct"]["frac"],3))
PY
