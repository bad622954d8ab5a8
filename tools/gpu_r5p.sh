#!/bin/bash
# visit r5p: cheaper running product (SiteProd: one range test + one DMUL per factor)
set -u
TAG=${1:-r5p}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "three_tax or place_matches or engines_agree or c2_properties" > gpurun_out/${TAG}_pytest_tt.log 2>&1; rc=$?; echo "pytest tt rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_tt.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/${TAG}_pytest_tt.log | cut -c1-300; exit 1; fi
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench.json")); r=j["roofline"]
print(round(j["value"]), "pl/s", round(j["ms_per_step"],2), "ms", {k: round(v,2) for k,v in r["share_of_step_ms"].items()}, "frac", round(r["frac"],4))
PY
