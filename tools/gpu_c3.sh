#!/bin/bash
# C3 (ar53) bench line: 256 genomes on one GPU
set -u
TAG=${1:-c3}
mkdir -p gpurun_out
timeout 1200 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_c3.json")); r=j["roofline"]
    print(round(j["value"],1), "genomes/s", round(j["ms_per_step"],1), "ms", {k: round(v,1) for k,v in r["share_of_step_ms"].items()}, "e2e", round(j["e2e"]["value"],1))
    o = r if r["kernel"].startswith("prefilter") else r["other_kernel"]
    print("prefilter:", o["bound"], round(o["achieved"],1), o["unit"], "frac", o["frac"])
except Exception as e: print("parse failed", e)
PY
tail -3 gpurun_out/${TAG}_bench_c3.err
