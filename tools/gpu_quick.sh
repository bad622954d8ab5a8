#!/bin/bash
# quick GPU visit: second-stage parity tests + bench lines for the engine variants + one ncu capture
set -u
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "three_tax or place or c2 or edge_cases or device_resident" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${TAG}_pytest.log
for E in group warp; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tt-engine $E > gpurun_out/${TAG}_bench_$E.json 2> gpurun_out/${TAG}_bench_$E.err; echo "bench $E rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_$E.json")); print("engine $E:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()})
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_$E.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tt_group_kernel' -c 1 -o gpurun_out/${TAG}_full_ttgroup -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
