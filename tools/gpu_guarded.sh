#!/bin/bash
# guarded GPU visit for a new kernel: every step under a short timeout so that a hang costs seconds, not minutes
set -u
TAG=${1:-g}
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q -k "engines_agree" > gpurun_out/${TAG}_pytest1.log 2>&1; rc=$?; echo "pytest engines rc=$rc"; tail -8 gpurun_out/${TAG}_pytest1.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 python -m pytest tests -m gpu -x -q -k "three_tax or place or c2 or edge_cases or device_resident or refpkg" > gpurun_out/${TAG}_pytest2.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -8 gpurun_out/${TAG}_pytest2.log
if [ $rc -ne 0 ]; then exit 1; fi
for E in group warp; do
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tt-engine $E > gpurun_out/${TAG}_bench_$E.json 2> gpurun_out/${TAG}_bench_$E.err; echo "bench $E rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_$E.json")); print("engine $E:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()})
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_$E.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tt_group_kernel' -c 1 -o gpurun_out/${TAG}_full_ttgroup -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
