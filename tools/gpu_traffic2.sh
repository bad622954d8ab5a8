#!/bin/bash
# dram bytes of the prefilter and of every second-stage launch of ONE bench step, for roofline.traffic (profiles/traffic.json)
#   gpu_traffic2.sh TAG [workload]
set -u
TAG=${1:-tr}; WL=${2:-c2}
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'tt_eigen_kernel|prefilter_qlane_kernel' -c 400 --csv --log-file gpurun_out/${TAG}_traffic_${WL}.csv python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-direct-prefilter > gpurun_out/${TAG}_traffic_${WL}.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, json
rows=[r for r in csv.reader(l for l in open("gpurun_out/${TAG}_traffic_${WL}.csv") if not l.startswith("=="))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); mi=h.index("Metric Name"); ii=h.index("ID"); ui=h.index("Metric Unit")
d={}
def val(r):
    v=float(r[vi].replace(",","")); u=r[ui]
    return v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9,"ns":1,"us":1e3,"ms":1e6,"s":1e9}.get(u,1)
for r in rows[1:]:
    d.setdefault(int(r[ii]),{"k":r[ki]})[r[mi]]=val(r)
ids=sorted(d)
pf=[i for i in ids if "prefilter_qlane_kernel" in d[i]["k"]]
last=pf[-1]
step=[d[i] for i in ids if i>=last]
out={"prefilter_qlane_kernel":{"per":"launch","capture":"${TAG}","dram_bytes":step[0]["dram__bytes_read.sum"]+step[0]["dram__bytes_write.sum"],"launches":1}}
tt=[x for x in step if "tt_eigen" in x["k"]]
out["tt_eigen_kernel"]={"per":"step (%d launches: baseball waves + refinement)"%len(tt),"capture":"${TAG}","launches":len(tt),
  "dram_bytes":sum(x["dram__bytes_read.sum"]+x["dram__bytes_write.sum"] for x in tt),
  "per_launch_read_gb":[round(x["dram__bytes_read.sum"]/1e9,3) for x in tt],"per_launch_write_gb":[round(x["dram__bytes_write.sum"]/1e9,3) for x in tt],
  "per_launch_ms":[round(x["gpu__time_duration.sum"]/1e6,3) for x in tt]}
json.dump(out,open("gpurun_out/${TAG}_traffic_${WL}.json","w"),indent=1); print(json.dumps(out))
PY
