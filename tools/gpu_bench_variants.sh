#!/bin/bash
# bench lines for the engine-4 geometry variants (no tests)
set -u
TAG=${1:-e}
mkdir -p gpurun_out
for V in ${VARIANTS:-0 1 2 3}; do
PPB_TTE_VARIANT=$V timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench_v$V.json 2> gpurun_out/${TAG}_bench_v$V.err; echo "bench v$V rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_v$V.json")); print("variant $V:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()}, "frac", round(j["roofline"]["frac"],3))
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_v$V.err
done
