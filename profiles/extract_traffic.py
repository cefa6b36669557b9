#!/usr/bin/env python
"""Write profiles/ncu_traffic.json from `ncu --set full` captures of k_render / k_prepare:
dram__bytes_read.sum + dram__bytes_write.sum per launch (bench.py reports it as roofline.traffic).
usage: python profiles/extract_traffic.py <render.ncu-rep> <prepare.ncu-rep> <tag>"""
import csv
import json
import os
import subprocess
import sys


def metrics(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))
    units = dict(zip(hdr, rows[1]))

    def mb(key):
        v, u = float(d[key]), units[key]
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
    return {"kernel": d.get("Kernel Name", "?"), "grid": d.get("launch__grid_size"),
            "issue_slots_busy_pct": float(d["sm__inst_issued.avg.pct_of_peak_sustained_active"]),
            "fma_pipe_active_pct": float(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
            "warp_instructions": float(d["smsp__inst_executed.sum"]),
            "dram_pct_of_peak": float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]),
            "dram_read_mb": mb("dram__bytes_read.sum"), "dram_write_mb": mb("dram__bytes_write.sum"),
            "duration_us_under_ncu": float(d["gpu__time_duration.sum"])}


if __name__ == "__main__":
    r, p, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    out = {"tag": tag, "sources_per_launch": 7,
           "note": "ncu --set full --clock-control none, bench.py --utterances 7 (14 cfg2 sources per step, 96 MB chunk budget -> "
                   "7 sources per launch, the same launch size as the default bench)",
           "k_render": metrics(r), "k_prepare": metrics(p)}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
