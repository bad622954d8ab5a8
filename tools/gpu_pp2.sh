#!/bin/bash
set -u
TAG=${1:-pp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "posterior or priors or engines_agree or c5_shape or c1_refpkg or stub" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; echo "bench c5 rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench_c5.json")); print(round(j["value"]), round(j["ms_per_step"],1), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()})
PY
tail -3 gpurun_out/${TAG}_bench_c5.err
