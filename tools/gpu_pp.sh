#!/bin/bash
set -u
TAG=${1:-pp}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "posterior or priors or engines_agree or refpkg" > gpurun_out/${TAG}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc $(tail -1 gpurun_out/${TAG}_pytest.log)"; tail -14 gpurun_out/${TAG}_pytest.log | head -13
if [ $rc -ne 0 ]; then exit 1; fi
for P in 1 0; do
PPB_TTG_PP=$P timeout 200 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5_p$P.json 2> gpurun_out/${TAG}_bench_c5_p$P.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_c5_p$P.json")); print("pp-on-group $P:", round(j["value"]), {k: round(v,1) for k,v in j["roofline"]["share_of_step_ms"].items()}, "kept", j["config"]["mean_kept"])
except Exception as e: print("parse failed", e)
PY
tail -2 gpurun_out/${TAG}_bench_c5_p$P.err
done
