#!/bin/bash
set -u
TAG=${1:-v}
mkdir -p gpurun_out
for V in ${VARIANTS:-3 4 1}; do
PPB_TTG_VARIANT=$V timeout 150 python -m pytest tests -m gpu -x -q -k "engines_agree or c2_properties or edge_cases" > gpurun_out/${TAG}_pytest_v$V.log 2>&1; echo "variant $V pytest rc=$? $(tail -1 gpurun_out/${TAG}_pytest_v$V.log)"
PPB_TTG_VARIANT=$V timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_v$V.json 2> gpurun_out/${TAG}_bench_v$V.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_v$V.json")); print("variant $V:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()})
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_v$V.err
done
