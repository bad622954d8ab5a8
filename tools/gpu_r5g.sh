#!/bin/bash
# visit r5g: general block engine - transposed xit in build_P, three sites per thread in the Gcat pendant pass; parity, C3 bench, ncu
set -u
TAG=${1:-r5g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
bash tools/gpu_c3.sh ${TAG}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:three_tax_kernel -s 0 -c 1 -o gpurun_out/${TAG}_c3_full -f python bench.py --workload c3 --queries 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_c3_ncu.log 2>&1; echo "ncu c3 rc=$?"
