#!/bin/bash
# one ncu --set full capture (with source) of a named kernel of a C2 step:  gpu_ncu_kernel.sh TAG KERNEL_REGEX [bench args...]
set -u
TAG=${1:-n}; PAT=${2:-prefilter_direct_kernel}; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$PAT -s ${SKIP:-0} -c ${COUNT:-1} -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-direct-prefilter "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
ls -la gpurun_out/${TAG}_full.ncu-rep
