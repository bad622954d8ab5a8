#!/bin/bash
# one ncu --set full capture (with source) of the second-stage kernel of a C2 step
set -u
TAG=${1:-n}
PAT=${2:-tt_eigen_kernel}
mkdir -p gpurun_out
PPB_TTE_VARIANT=${VARIANT:-1} timeout 900 ncu --set full --clock-control none --import-source on -k regex:$PAT -s ${SKIP:-0} -c 1 -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --queries ${QUERIES:-10000} > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_full.ncu-rep
