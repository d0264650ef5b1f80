#!/usr/bin/env python
"""Writes the TABLES of profiles/<tag>_ncu_summary.md from the round's evidence: the ncu counter table of every raw page
(tools/ncu_summary.py), the per-sample table from profiles/traffic.json, and the launch list grouped by kernel.  The
hand-written "Reading" section lives in profiles/<tag>_ncu_reading.md and is appended verbatim.
Usage: make_ncu_summary.py TAG   (after tools/final_profile.sh TAG and tools/make_traffic.py TAG ...)"""
import collections
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = os.path.join(ROOT, "profiles")
B_ALG = {"k_ggx_dielectric": 113, "k_ggx_dielectric_tol": 113, "k_disney_sample_eval_pdf": 176, "k_disney_sample_eval_pdf_tol": 176,
         "k_skin_profile": 40, "k_skin_profile_tol_x2": 40, "k_ggx_sample_eval_pdf": 88, "k_ggx_sample_eval_pdf_tol": 88,
         "k_albedo_sweep": 0, "k_albedo_sweep_tol": 0}

raws = sorted(glob.glob(os.path.join(P, f"{tag}_prof_*.raw.csv")))
t = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")] + raws, capture_output=True, text=True).stdout
hdr = sep = None
rows = []
for block in t.split("### ")[1:]:
    lines = [ln for ln in block.splitlines() if ln.startswith("|")]
    hdr, sep = lines[0], lines[1]
    rows += lines[2:]
out = [f"# ncu evidence, {tag} (B200, `ncu --set full --clock-control none`, one launch per kernel inside `bench.py`; raw pages: `profiles/{tag}_prof_*.raw.csv`)\n",
       f"Captured by `tools/final_profile.sh {tag}` on the final kernels of the round (SASS hashes in `profiles/traffic.json`, which `bench.py` "
       f"checks against the running library).  Times under ncu are cold-cache and serialised; the shares of the step agree with the CUDA-event "
       f"timings of `profiles/{tag}_bench_n1.json`.\n", hdr, sep] + rows + [""]
tr = json.load(open(os.path.join(P, "traffic.json")))["kernels"]
out += ["## Per sample (profiles/traffic.json)\n",
        "| kernel | DRAM B / sample (ncu) | algorithmic B / sample | warp instr / 32 samples | issue slots | DRAM % of ncu peak | registers | occupancy | ms under ncu |",
        "|---|---|---|---|---|---|---|---|---|"]
for k in B_ALG:
    if k in tr:
        e = tr[k]
        out.append(f"| `{k}` | {e['dram_bytes_per_sample']:.1f} | {B_ALG[k]} | {e['warp_instr_per_32_samples']:.0f} | {e['issue_slot_pct']:.1f} % | "
                   f"{e['dram_pct_of_peak']:.1f} % | {e['registers']:.0f} | {e['achieved_occupancy_pct']:.0f} % | {e['ms_under_ncu']:.3f} |")
out.append("")
reading = os.path.join(P, f"{tag}_ncu_reading.md")
if os.path.exists(reading):
    out.append(open(reading).read())
launches = os.path.join(P, f"{tag}_launches.csv")
if os.path.exists(launches):
    agg, h = collections.OrderedDict(), None
    for r in csv.reader(io.StringIO(open(launches).read())):
        if r and r[0] == "ID":
            h = r
            continue
        if h and len(r) == len(h):
            m = re.search(r"(k_[A-Za-z0-9_]+|elementwise_kernel|reduce_kernel|fill|memset)", r[h.index("Kernel Name")])
            a = agg.setdefault(m.group(1) if m else r[h.index("Kernel Name")][:40], [0, 0.0])
            a[0] += 1
            a[1] += float(r[h.index("Metric Value")])
    tot = sum(a[1] for a in agg.values())
    out += [f"## Launch list (`profiles/{tag}_launches.csv`: `ncu --metrics gpu__time_duration.sum` over one default `bench.py` run)\n",
            "| kernel | launches | total ms | share | us per launch |", "|---|---|---|---|---|"]
    for k, (c, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {c} | {tt / 1e6:.3f} | {100 * tt / tot:.1f} % | {tt / c / 1e3:.1f} |")
    out.append("\n(`elementwise_kernel` = torch copies that build the pinned e2e prefix and slice inputs: set-up, outside every timed region.)")
open(os.path.join(P, f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("wrote", os.path.join(P, f"{tag}_ncu_summary.md"))
