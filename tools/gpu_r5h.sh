#!/bin/bash
# visit r5h: Gcat pendant pass with 16-byte pair loads, w kept in registers - parity suite + C3 bench
set -u
TAG=${1:-r5h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
bash tools/gpu_c3.sh ${TAG}
