#!/bin/bash
# prefilter bring-up: parity of every form, then launch times of the forms (and staged geometry variants) at the C2 shape
set -u
TAG=${1:-pf}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "prefilter or c4_shape or c2_properties" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest.log
for V in ${VARIANTS:-0 1 2 3}; do
PPB_PFS_VARIANT=$V timeout 300 python - > gpurun_out/${TAG}_times_v$V.json 2> gpurun_out/${TAG}_times_v$V.err <<PY
import json, sys
sys.path.insert(0, ".")
import bench
from pplacer_b200.placer import Placer
out = {}
for wl, Q in (("c2", 10000), ("c4", 4000)):
    if wl == "c4" and "$V" not in ("0",) and "${C4ALL:-0}" == "0":
        continue
    prob = bench.build_workload(wl, Q, 0)
    with Placer(prob) as p:
        p.precompute()
        out[wl] = {str(k): p.prefilter_time_ms(prob.query_chars, kernel=k, reps=3) for k in (6, 2, 1, 4, 5)}
print(json.dumps(out))
PY
echo "variant $V rc=$?"; cat gpurun_out/${TAG}_times_v$V.json; tail -3 gpurun_out/${TAG}_times_v$V.err
done
