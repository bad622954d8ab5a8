#!/bin/bash
# visit r5c: full GPU suite with the category-ordered site lists (Gcat second stage), C3 bench, C2 launch list
set -u
TAG=${1:-r5c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
bash tools/gpu_c3.sh ${TAG}
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/${TAG}_launches.csv") if not l.startswith("=="))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); mi=h.index("Metric Name"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{"k":r[ki][:34]})[r[mi]]=float(r[vi].replace(",",""))
sel=[v for v in d.values() if "tt_eigen" in v["k"] or "prefilter_q" in v["k"]]
for v in sel[-8:]: print(v["k"], round(v.get("gpu__time_duration.sum",0)/1e6,3), "ms grid", v.get("launch__grid_size"))
PY
