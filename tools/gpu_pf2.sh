#!/bin/bash
# log-table prefilter: parity, then launch times for its geometry variants (PPB_PF_GEOM) at C2 and C4 shapes
set -u
TAG=${1:-pf}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "prefilter or c4_shape or c2_properties or place_matches or edge_cases" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest.log
for V in ${VARIANTS:-0 1 2 3}; do
PPB_PF_GEOM=$V timeout 300 python - > gpurun_out/${TAG}_times_g$V.json 2> gpurun_out/${TAG}_times_g$V.err <<PY
import json, sys
sys.path.insert(0, ".")
import bench
from pplacer_b200.placer import Placer
out = {}
for wl, Q in (("c2", 10000), ("c4", 10000)):
    prob = bench.build_workload(wl, Q, 0)
    with Placer(prob) as p:
        p.precompute()
        out[wl] = {str(k): p.prefilter_time_ms(prob.query_chars, kernel=k, reps=3) for k in ((6, 2) if "$V" != "0" else (6, 2, 5))}
print(json.dumps(out))
PY
echo "geom $V rc=$?"; cat gpurun_out/${TAG}_times_g$V.json; tail -3 gpurun_out/${TAG}_times_g$V.err
done
