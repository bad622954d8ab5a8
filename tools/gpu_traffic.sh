#!/bin/bash
# dram__bytes of every second-stage launch of ONE bench step (ncu --set full), for roofline.traffic
set -u
TAG=${1:-tr}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:'tt_group_kernel|prefilter_table_kernel' --launch-skip 28 -c 7 -o gpurun_out/${TAG}_traffic -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_traffic.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_traffic.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:40], r[h.index('gpu__time_duration.sum')], r[h.index('dram__bytes_read.sum')], r[h.index('dram__bytes_write.sum')], r[h.index('launch__grid_size')])
"
