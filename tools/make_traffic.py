#!/usr/bin/env python
"""Regenerates profiles/traffic.json from a round's `ncu --set full` captures (raw pages exported with
`ncu -i rep --page raw --csv`): per kernel the DRAM bytes and executed warp instructions per sample of ONE launch, its
launch shape, duration under ncu, and the hash of the kernel's SASS in the library that was profiled
(rlshaders_b200/kernel_hashes.json, written by __graft_entry__.build).  bench.py accepts an entry only while that hash
equals the running library's, so a capture never describes a kernel that has since changed.

Usage: make_traffic.py TAG raw.csv [raw.csv ...]      (run where the library of the capture is the one in the tree)
Samples per launch: grid x block for the one-thread-per-sample kernels (x 2 / x 4 for kernels named *_x2 / *_x4); the
sweep kernels take theirs from --sweep-samples (cells x spp, default 65536 x 4096)."""
import argparse
import csv
import json
import os
import re
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}


def num(row, hdr, units, key):
    if key not in hdr:
        return None
    i = hdr.index(key)
    try:
        v = float(row[i].replace(",", ""))
    except ValueError:
        return None
    return v * SCALE.get(units[i], 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("csvs", nargs="+")
    ap.add_argument("--sweep-samples", type=float, default=65536.0 * 4096.0)
    ap.add_argument("--samples-per-thread", type=float, default=1.0)
    ap.add_argument("-o", default=os.path.join(ROOT, "profiles", "traffic.json"))
    a = ap.parse_args()
    with open(os.path.join(ROOT, "rlshaders_b200", "kernel_hashes.json")) as fh:
        hashes = json.load(fh)
    kernels = {}
    for path in a.csvs:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            print("skip (empty):", path)
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            full = r[hdr.index("Kernel Name")]
            m = re.search(r"(k_[A-Za-z0-9_]+)", full)
            if not m:
                continue
            name = m.group(1)
            grid = num(r, hdr, units, "launch__grid_size")
            block = num(r, hdr, units, "launch__block_size")
            per_thread = 2.0 if name.endswith("_x2") else (4.0 if name.endswith("_x4") else a.samples_per_thread)
            n = a.sweep_samples if "sweep" in name else grid * block * per_thread
            rd, wr = num(r, hdr, units, "dram__bytes_read.sum"), num(r, hdr, units, "dram__bytes_write.sum")
            inst = num(r, hdr, units, "smsp__inst_executed.sum")
            e = {"sass_sha16": hashes.get(name), "samples_per_launch": n, "grid": grid, "block": block,
                 "registers": num(r, hdr, units, "launch__registers_per_thread"),
                 "dram_bytes_per_sample": (rd + wr) / n, "dram_read_bytes": rd, "dram_write_bytes": wr,
                 "warp_instr_per_32_samples": inst / (n / 32.0) if inst else None,
                 "ms_under_ncu": num(r, hdr, units, "gpu__time_duration.sum") * 1e3,
                 "issue_slot_pct": num(r, hdr, units, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "dram_pct_of_peak": num(r, hdr, units, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "achieved_occupancy_pct": num(r, hdr, units, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                 "threads_per_inst": num(r, hdr, units, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                 "fma_pipe_pct": num(r, hdr, units, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                 "alu_pipe_pct": num(r, hdr, units, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                 "fp64_pipe_pct": num(r, hdr, units, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                 "xu_pipe_pct": num(r, hdr, units, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                 "l2_hit_pct": num(r, hdr, units, "lts__t_sector_hit_rate.pct"),
                 "capture": os.path.basename(path), "kernel_signature": full.split("(")[0]}
            kernels[name] = e
    out = {"captured": f"{a.tag}, {time.strftime('%Y-%m-%d')}",
           "_source": "tools/make_traffic.py over the round's ncu --set full raw pages (tools/final_profile.sh): "
                      "(dram__bytes_read.sum + dram__bytes_write.sum) and smsp__inst_executed.sum of one launch / "
                      "samples of that launch",
           "kernels": kernels}
    with open(a.o, "w") as fh:
        json.dump(out, fh, indent=1)
    for k, e in sorted(kernels.items()):
        print(f"{k:34s} {e['dram_bytes_per_sample']:8.2f} B/sample  {e['warp_instr_per_32_samples'] or 0:8.1f} instr/32  "
              f"issue {e['issue_slot_pct'] or 0:5.1f}%  dram {e['dram_pct_of_peak'] or 0:5.1f}%  sass {e['sass_sha16']}")


if __name__ == "__main__":
    main()
