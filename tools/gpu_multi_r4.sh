#!/bin/bash
# N-GPU visit: the multi-device handle test, the default bench line at N (weak, gather inside e2e) and the strong-scaled C4 line
set -u
TAG=${1:-m}; N=${2:-2}; C4Q=${3:-80000}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or stub" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench weak rc=$?"
cat gpurun_out/${TAG}_bench_n$N.json | cut -c1-900; tail -3 gpurun_out/${TAG}_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c4 --scaling strong --total-queries $C4Q --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_c4_strong_n$N.json 2> gpurun_out/${TAG}_bench_c4_strong_n$N.err; echo "bench c4 strong rc=$?"
cat gpurun_out/${TAG}_bench_c4_strong_n$N.json | cut -c1-1200; tail -3 gpurun_out/${TAG}_bench_c4_strong_n$N.err
