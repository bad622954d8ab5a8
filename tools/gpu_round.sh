#!/bin/bash
# one GPU visit: parity tests, bench lines, ncu launch list and full captures (outputs under gpurun_out/)
set -u
TAG=${1:-r1b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --steps 5 --warmup 3 --prefilter direct --no-cpu-baseline > gpurun_out/${TAG}_bench_direct.json 2> gpurun_out/${TAG}_bench_direct.err; echo "bench direct rc=$?"
cat gpurun_out/${TAG}_bench_direct.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'prefilter_table_kernel|three_tax_kernel|mtable_kernel' -c 4 -o gpurun_out/${TAG}_full_default -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --queries 4096 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'prefilter_direct_kernel' -c 1 -o gpurun_out/${TAG}_full_direct -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --queries 4096 --prefilter direct > gpurun_out/${TAG}_ncu_full_direct.log 2>&1; echo "ncu full direct rc=$?"
ls -la gpurun_out
