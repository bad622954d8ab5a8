#!/bin/bash
# one GPU visit of round 2: full parity suite (no -x: every failure is wanted), engine parity statistics, default bench line
set -u
TAG=${1:-r4a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt; free -g >> gpurun_out/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
if [ "${STATS:-1}" = "1" ]; then
timeout 600 python tools/parity_stats.py c2 768 3 4 > gpurun_out/${TAG}_parity_stats.json 2> gpurun_out/${TAG}_parity_stats.err; echo "stats rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_parity_stats.json"))
    for k,v in j.items():
        if isinstance(v,dict): print(k, {a:b for a,b in v.items() if a!="worst"}); print("  worst", v["worst"][:4])
except Exception as e: print("stats parse failed", e)
PY
tail -3 gpurun_out/${TAG}_parity_stats.err
fi
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
