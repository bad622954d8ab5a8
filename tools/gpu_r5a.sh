#!/bin/bash
# visit r5a: full GPU parity suite, qlane prefilter form timed against the default on C2 and C4, C3 bench line
set -u
TAG=${1:-r5a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
for W in c2 c4; do
for PF in default qlane; do
for V in ${VARIANTS:-0 1 2 3}; do
[ $PF = default ] && [ $V != 0 ] && continue
PPB_PFQ_VARIANT=$V timeout 300 python bench.py --workload $W --prefilter $PF --steps 3 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench_${W}_${PF}$V.json 2> gpurun_out/${TAG}_bench_${W}_${PF}$V.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_${W}_${PF}$V.json")); print("$W $PF v$V:", round(j["value"]), {k: round(v,2) for k,v in j["roofline"]["share_of_step_ms"].items()})
except Exception as e: print("$W $PF v$V parse failed", e)
PY
done; done; done
bash tools/gpu_c3.sh ${TAG}
