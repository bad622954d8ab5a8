#!/bin/bash
set -u
TAG=${1:-c4}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "zero_cat or refpkg" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench_c4.json")); print("C4:", round(j["value"]), j["roofline"]["share_of_step_ms"], "kept", j["config"]["mean_kept"], "precompute", j["precompute_ms"])
PY
tail -3 gpurun_out/${TAG}_bench_c4.err
