#!/bin/bash
set -u
N=${1:-8}; TAG=${2:-mg}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_n$N.json; tail -2 gpurun_out/${TAG}_bench_n$N.err
