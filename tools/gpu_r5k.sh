#!/bin/bash
# visit r5k: larger batches (memory-scaled), fig ranking through the query-lane table, C4 at 10k / 40k queries, ncu of the block engine on C3
set -u
TAG=${1:-r5k}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for Q in 10000 40000; do
timeout 300 python bench.py --workload c4 --queries $Q --steps 2 --warmup 2 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_bench_c4_$Q.json 2> gpurun_out/${TAG}_bench_c4_$Q.err; echo "bench c4 $Q rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_c4_$Q.json")); r=j["roofline"]
    print("c4 Q=$Q", round(j["value"]), "pl/s", round(j["ms_per_step"],1), "ms", {k: round(v,1) for k,v in r["share_of_step_ms"].items()}, "e2e", round(j["e2e"]["value"]))
except Exception as e: print("parse failed", e)
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:three_tax_kernel -s 0 -c 1 -o gpurun_out/${TAG}_c3_full -f python bench.py --workload c3 --queries 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_c3_ncu.log 2>&1; echo "ncu c3 rc=$?"
