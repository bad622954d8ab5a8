#!/bin/bash
# one GPU visit: parity tests, bench lines, ncu launch list and full captures (outputs under gpurun_out/)
set -u
TAG=${1:-r1c}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
for B in 1 3; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tt-blocks-per-sm $B > gpurun_out/${TAG}_bench_b$B.json 2> gpurun_out/${TAG}_bench_b$B.err; echo "bench b$B rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_b$B.json")); print("blocks/SM $B:", j["value"], j["roofline"]["share_of_step_ms"])
except Exception as e: print("parse failed", e)
PY
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tt-engine block > gpurun_out/${TAG}_bench_block.json 2> gpurun_out/${TAG}_bench_block.err; echo "bench block rc=$?"
cat gpurun_out/${TAG}_bench_block.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tt_warp_kernel' -c 2 -o gpurun_out/${TAG}_full_ttwarp -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --queries 4096 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
