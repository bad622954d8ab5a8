#!/bin/bash
# second-stage phase profile (library must have been built with -DPPB_TT_PROFILE)
set -u
TAG=${1:-p}
mkdir -p gpurun_out
for V in ${VARIANTS:-0}; do
PPB_TTE_VARIANT=$V timeout 600 python tools/tt_profile.py c2 10000 > gpurun_out/${TAG}_ttprof_c2_v$V.json 2> gpurun_out/${TAG}_ttprof_c2_v$V.err; echo "rc=$?"
cat gpurun_out/${TAG}_ttprof_c2_v$V.json; tail -3 gpurun_out/${TAG}_ttprof_c2_v$V.err
done
