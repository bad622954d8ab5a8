#!/bin/bash
# sweep of the warp engine's runtime knobs (PPB_TT_FLAGS: bit0 recompute w, bit1 L2 prefetch; blocks per SM)
set -u
TAG=${1:-sw}
mkdir -p gpurun_out
for F in ${FLAGS_LIST:-0 2}; do
PPB_TT_FLAGS=$F timeout 600 python -m pytest tests -m gpu -x -q -k "engines_agree or c2_properties" > gpurun_out/${TAG}_pytest_f$F.log 2>&1; echo "flags $F pytest rc=$? $(tail -1 gpurun_out/${TAG}_pytest_f$F.log)"
for B in 3 2; do
PPB_TT_FLAGS=$F timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tt-blocks-per-sm $B > gpurun_out/${TAG}_bench_f${F}_b$B.json 2> gpurun_out/${TAG}_bench_f${F}_b$B.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_f${F}_b$B.json")); print("flags $F blocks/SM $B:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()})
except Exception as e: print("parse failed", e)
PY
done
done
