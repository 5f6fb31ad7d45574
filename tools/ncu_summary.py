#!/usr/bin/env python
# coding=utf-8
"""Summarise an .ncu-rep (ncu --set full) into the JSON kept under profiles/: one record per captured launch with the
metrics the roofline discussion uses.  Usage: python tools/ncu_summary.py report.ncu-rep > profiles/rN_ncu_full_<what>.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                rec["{} [{}]".format(k, units[i]) if units[i] else k] = r[i]
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
