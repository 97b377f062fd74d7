#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into the text kept under profiles/.
usage: ncu_summary.py gpurun_out/prof_X.ncu-rep > profiles/X_summary.txt"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sector_hit_rate\.pct$", r"^l1tex__t_sector_hit_rate\.pct$",
    r"^l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|st)\.sum$", r"^l1tex__t_requests_pipe_lsu_mem_global_op_(ld|st)\.sum$",
    r"^lts__t_sectors_srcunit_tex_op_(read|write)\.sum$",
    r"^launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|waves_per_multiprocessor)$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_(alu|fma|xu|lsu)\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__average_warps?_issue_stalled_\w+_per_issue_active\.ratio$", r"^smsp__average_warp_latency_issue_stalled_\w+\.ratio$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^smsp__inst_executed\.sum$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    print(f"# ncu --set full summary of {path}")
    for r in data:
        print(f"\n## {r[name_col]}")
        for i, h in enumerate(hdr):
            if any(re.search(k, h) for k in KEYS):
                print(f"{h:85s} {r[i]:>18s} {units[i]}")
        try:
            # this ncu prints each column in a unit of its own choosing (read in Gbyte beside write in Mbyte): sum in bytes
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            i_rd, i_wr = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
            total = float(r[i_rd]) * scale[units[i_rd]] + float(r[i_wr]) * scale[units[i_wr]]
            print(f"{'TRAFFIC dram read+write':85s} {total / 1e9:18.6f} Gbyte")
        except Exception:  # noqa: BLE001
            pass


if __name__ == "__main__":
    main(sys.argv[1])
