#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
  python tools/ncu_summary.py raw gpurun_out/prof_prefilter_r1.ncu-rep > profiles/r1_prefilter_ncu.txt

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (count, total
time, share of the captured command).  `raw` prints the headline metrics of every captured launch
of an `--set full` report (needs the ncu CLI, no GPU).
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
]


def launches(path):
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.3f ms of kernel time (ncu per-launch times: cold cache, serialised)" % (path, n, tot / 1e6))
    print("%-100s %6s %12s %7s" % ("kernel", "n", "total_ms", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-100s %6d %12.3f %6.1f%%" % (k[:100], a[0], a[1] / 1e6, 100 * a[1] / tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("## %s  (id %s)" % (r[name_i][:120], r[0]))
        for i, h in enumerate(hdr):
            if any(h == k or h.startswith(k + ".") and h.count(".") == k.count(".") for k in RAW_KEYS) and r[i] != "":
                print("  %-80s %14s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
