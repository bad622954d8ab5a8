#!/usr/bin/env python3
"""Per-source-line stall samples and executed instructions from an ncu report.
usage: ncu_source_lines.py report.ncu-rep kernel_regex [launch_skip] [top_n]
Reads `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo + --import-source on)."""
import collections
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + pat, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8:
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] != "" and hdr:
            try:
                agg.append((cur, int(r[0]), r[1], int(r[4] or 0), int(r[7] or 0), dict(zip(hdr[30:47], r[30:47]))))
            except ValueError:
                pass
    ts = sum(a[3] for a in agg) or 1
    ti = sum(a[4] for a in agg) or 1
    print("total samples", ts, "warp instructions", ti)
    bf, bi = collections.Counter(), collections.Counter()
    for f, l, s, sm, ins, st in agg:
        bf[f] += sm
        bi[f] += ins
    for f in bf:
        print("  %-24s %5.1f%% samples %5.1f%% instructions" % (f, 100 * bf[f] / ts, 100 * bi[f] / ti))
    agg.sort(key=lambda a: -a[3])
    for a in agg[:top]:
        st = sorted(((int(v or 0), k) for k, v in a[5].items()), reverse=True)[:2]
        print("%-20s %4d %5.2f%% smp %5.2f%% ins  [%s] %s" % (a[0], a[1], 100 * a[3] / ts, 100 * a[4] / ti,
              ",".join("%s %d" % (k.replace("stall_", ""), v) for v, k in st), a[2].strip()[:90]))


if __name__ == "__main__":
    main()
