#!/bin/bash
# one full ncu capture of the pair-of-sites prefilter kernel on a C2 step (short: fits a one-minute visit)
set -u
mkdir -p gpurun_out
timeout 70 ncu --set full --clock-control none --import-source on -k regex:prefilter_qlane_kernel -s 1 -c 1 -o gpurun_out/r6a_pf_pairs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-direct-prefilter > gpurun_out/r6a_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r6a_ncu.log | cut -c1-200
