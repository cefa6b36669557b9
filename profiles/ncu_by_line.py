#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line.

usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
       python profiles/ncu_by_line.py src.csv [top]
Prints, per CUDA source line: stall samples, executed warp instructions, SASS instruction count.
"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, cur_line, cur_src = None, None, None, ""
    agg = defaultdict(lambda: [0, 0, 0, ""])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_samp = hdr.index("# Samples")
            i_inst = hdr.index("Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur_line, cur_src = r[0], r[1]
            continue
        if len(r) <= i_inst or r[2] in ("...", "-"):
            continue
        try:
            s = int(r[i_samp])
            n = int(r[i_inst])
        except ValueError:
            continue
        a = agg[(cur_file, cur_line)]
        a[0] += s
        a[1] += n
        a[2] += 1
        a[3] = cur_src
    tot_s = sum(a[0] for a in agg.values()) or 1
    tot_n = sum(a[1] for a in agg.values()) or 1
    print("total samples %d, total warp instructions %d, lines %d" % (tot_s, tot_n, len(agg)))
    print("%-22s %8s %6s %12s %6s %5s  %s" % ("file:line", "samples", "%", "warp-inst", "%", "sass", "source"))
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %8d %6.1f %12d %6.1f %5d  %s" % ("%s:%s" % (f, l), a[0], 100.0 * a[0] / tot_s, a[1],
                                                     100.0 * a[1] / tot_n, a[2], a[3].strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
