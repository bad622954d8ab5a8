#!/bin/bash
set -u
TAG=${1:-ll}
mkdir -p gpurun_out
PPB_TT_FLAGS=${2:-2} timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tt-blocks-per-sm ${3:-2} > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/${TAG}_launches.csv") if not l.startswith("=="))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); 
sel=[(r[ki][:40], float(r[vi].replace(",",""))) for r in rows[1:] if "tt_warp" in r[ki] or "prefilter" in r[ki] or "three_tax" in r[ki]]
# last step's launches: print the final 12
for k,v in sel[-12:]: print(k, round(v/1e6,3), "ms")
PY
