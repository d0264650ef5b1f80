#!/usr/bin/env python
"""Summarise an ncu report (ncu -i rep --page raw --csv) as a markdown table of the counters
this repository's roofline discussion uses.  Usage: ncu_summary.py rep.ncu-rep|rep.raw.csv [...]
(a .csv argument is the already exported raw page: gpurun brings back at most 64 MiB)."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-slot %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    for rep in sys.argv[1:]:
        if rep.endswith(".csv"):
            out = open(rep).read()
        else:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"### {rep.split('/')[-1]}\n")
        print("| kernel | " + " | ".join(label for _, label in WANT) + " |")
        print("|---|" + "---|" * len(WANT))
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            cells = []
            for key, _ in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    v = r[i]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {units[i]}".strip())
                else:
                    cells.append("-")
            print(f"| {name} | " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main()
