#!/bin/bash
set -u
N=${1:-8}; TAG=${2:-mg8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
cat gpurun_out/${TAG}_bench_n$N.json | cut -c1-400; tail -2 gpurun_out/${TAG}_bench_n$N.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c4 --queries 125000 --steps 1 --warmup 3 > gpurun_out/${TAG}_bench_c4_1M_n$N.json 2> gpurun_out/${TAG}_bench_c4_1M_n$N.err; echo "bench c4 1M N=$N rc=$?"
cat gpurun_out/${TAG}_bench_c4_1M_n$N.json | cut -c1-400; tail -2 gpurun_out/${TAG}_bench_c4_1M_n$N.err
