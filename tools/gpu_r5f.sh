#!/bin/bash
# visit r5f: general block engine with shared-memory (not generic) matrix reads, Gcat pendant pass with its loads hoisted - parity suite + C3 bench
set -u
TAG=${1:-r5f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
bash tools/gpu_c3.sh ${TAG}
