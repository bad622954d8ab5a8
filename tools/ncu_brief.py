#!/usr/bin/env python
"""Brief of an `ncu --set full` report: headline metrics + top warp-stall reasons per captured launch.
  python tools/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg.per_second"]
for r in rows[2:]:
    print("## %s  (id %s)" % (r[hdr.index("Kernel Name")][:90], r[hdr.index("ID")]))
    for w in want:
        if w in hdr:
            print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    for v, h in sorted(st, reverse=True)[:7]:
        print("   stall %-30s %.2f warps per issue" % (h, v))
