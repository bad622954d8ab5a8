#!/bin/bash
# the final visit without the two large ncu reports (gpurun copies back at most 64 MiB): smoke, parity suite, every bench line,
# launch list, DRAM traffic of one step; tools/gpu_ncu_kernel.sh captures the full reports in a separate visit
set -u
TAG=${1:-fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference rc=$?"
cut -c1-200 gpurun_out/${TAG}_bench_reference.json
for W in c4 c5 c3; do
timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err; echo "bench $W rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_$W.json")); r=j["roofline"]
    print("$W", round(j["value"],1), "pl/s", round(j["ms_per_step"],1), "ms", {k: round(v,1) for k,v in r["share_of_step_ms"].items()}, "e2e", round(j["e2e"]["value"],1), "frac", r.get("frac"))
except Exception as e: print("$W parse failed", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
bash tools/gpu_traffic2.sh ${TAG} c2 | cut -c1-300
du -sh gpurun_out
