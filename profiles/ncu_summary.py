#!/usr/bin/env python
"""Key metrics of an ncu report: `python profiles/ncu_summary.py X.ncu-rep`."""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit", "sm__throughput.avg.pct",
        "sm__inst_issued.avg.pct", "sm__warps_active.avg.pct", "smsp__pcsamp_warps_issue_stalled", "smsp__pcsamp_sample_count",
        "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts", "sass__inst_executed_local", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg", "l1tex__lsu_writeback", "lts__t_sectors_srcunit_tex_op_read.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) for k in KEEP) and "not_issued" not in h:
                print("%-88s %-10s %s" % (h, u, v))


if __name__ == "__main__":
    main(sys.argv[1])
