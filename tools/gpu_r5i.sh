#!/bin/bash
# visit r5i: protein query-lane prefilter + pair-load pendant pass (fixed trip count); short timeouts around the new kernels
set -u
TAG=${1:-r5i}
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -k "gcat or wag or protein or ar53" > gpurun_out/${TAG}_pytest_aa.log 2>&1; rc=$?; echo "pytest aa rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_aa.log
if [ $rc -ne 0 ]; then tail -30 gpurun_out/${TAG}_pytest_aa.log; exit 1; fi
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
TO=240
timeout $TO python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_c3.json")); r=j["roofline"]
    print(round(j["value"],1), "genomes/s", round(j["ms_per_step"],1), "ms", {k: round(v,1) for k,v in r["share_of_step_ms"].items()}, "e2e", round(j["e2e"]["value"],1))
except Exception as e: print("parse failed", e)
PY
tail -3 gpurun_out/${TAG}_bench_c3.err
