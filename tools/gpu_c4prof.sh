#!/bin/bash
set -u
TAG=${1:-c4p}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_c4_launches.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c4_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:'prefilter_table_kernel' --launch-skip 3 -c 1 -o gpurun_out/${TAG}_c4_prefilter -f python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c4_full.log 2>&1; echo "ncu full rc=$?"
