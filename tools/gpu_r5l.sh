#!/bin/bash
# visit r5l: generic query-lane table with the any-state column, DNA Price-CAT through it; full suite
set -u
TAG=${1:-r5l}
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -k "generic_query_lane or prefilter" > gpurun_out/${TAG}_pytest_pf.log 2>&1; rc=$?; echo "pytest pf rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_pf.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/${TAG}_pytest_pf.log; exit 1; fi
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
