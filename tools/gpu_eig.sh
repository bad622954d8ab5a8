#!/bin/bash
# engine 4 bring-up: second-stage parity tests, then bench lines for the geometry variants
set -u
TAG=${1:-e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "three_tax or place_matches or edge_cases or c2_properties or work_counters" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest.log
for V in ${VARIANTS:-0 1 2 3}; do
PPB_TTE_VARIANT=$V timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_v$V.json 2> gpurun_out/${TAG}_bench_v$V.err; echo "bench v$V rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_v$V.json")); print("variant $V:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()}, "frac", round(j["roofline"]["frac"],3))
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_v$V.err
done
