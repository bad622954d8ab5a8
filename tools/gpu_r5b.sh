#!/bin/bash
# visit r5b: ncu full captures (with source) of the second-stage kernel on C2 and of the query-lane prefilter on C4; bench lines
set -u
TAG=${1:-r5b}
mkdir -p gpurun_out
bash tools/gpu_ncu_kernel.sh ${TAG}_tt tt_eigen_kernel
SKIP=3 bash tools/gpu_ncu_kernel.sh ${TAG}_pf prefilter_qlane_kernel --workload c4
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 400 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
cat gpurun_out/${TAG}_bench_c4.json
