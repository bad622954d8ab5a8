#!/bin/bash
# last check of the tree with the pair-of-sites table as the default: GPU parity suite + the default bench line
set -u
TAG=${1:-san}
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -2 gpurun_out/${TAG}_pytest.log
timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench.json")); r=j["roofline"]
print(round(j["value"]), "pl/s", round(j["ms_per_step"],2), "ms e2e", round(j["e2e"]["value"]), {k: round(v,2) for k,v in r["share_of_step_ms"].items()}, "frac", round(r["frac"],4), "other", r["other_kernel"]["kernel"], round(r["other_kernel"]["frac"],3))
PY
