#!/bin/bash
# quick GPU visit: second-stage parity tests + bench lines for the engine variants
set -u
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "three_tax or place or c2 or edge_cases or device_resident" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
for B in 3 2; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tt-blocks-per-sm $B > gpurun_out/${TAG}_bench_b$B.json 2> gpurun_out/${TAG}_bench_b$B.err; echo "bench b$B rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_b$B.json")); print("blocks/SM $B:", j["value"], j["roofline"]["share_of_step_ms"])
except Exception as e: print("parse failed", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tt_warp_kernel' -c 1 -o gpurun_out/${TAG}_full_ttwarp -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tt-blocks-per-sm 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
