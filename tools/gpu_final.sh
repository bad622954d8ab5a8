#!/bin/bash
# final visit of a round: parity suite, every bench line (C2 default with CPU baseline, reference arm, C4, C5, C3), launch list,
# DRAM traffic of one step, full ncu captures of the two dominant kernels (outputs under gpurun_out/, summaries copied to profiles/)
set -u
TAG=${1:-fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference rc=$?"
cut -c1-400 gpurun_out/${TAG}_bench_reference.json
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
bash tools/gpu_traffic2.sh ${TAG} c2 | cut -c1-600
bash tools/gpu_ncu_kernel.sh ${TAG}_tt tt_eigen_kernel
SKIP=0 bash tools/gpu_ncu_kernel.sh ${TAG}_pf prefilter_qlane_kernel --workload c4
