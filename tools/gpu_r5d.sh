#!/bin/bash
# visit r5d: Gcat second stage on compact (category, site)-ordered copies - parity suite + C3 bench
set -u
TAG=${1:-r5d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
bash tools/gpu_c3.sh ${TAG}
