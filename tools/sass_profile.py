#!/usr/bin/env python
"""Join an ncu SASS source page (per-instruction executed counts) with nvdisasm line info,
and aggregate executed warp-instructions per CUDA source line / file.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<k> > sass.csv
    cuobjdump -xelf all librls_b200.so ; nvdisasm -g -c *.cubin > all.sass
    python tools/sass_profile.py sass.csv all.sass <mangled-kernel-substring> <n_warps>
"""
import csv
import re
import sys
from collections import defaultdict


def load_counts(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    hdr = rows[hi]
    ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    isrc = hdr.index("Source")
    out = []
    base = None
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":
            break                      # only the first captured launch
        if len(r) <= it or r[ia] == "Address":
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        out.append((a - base, int(r[ie]), int(r[it]), r[isrc].strip()))
    return out


def load_lines(path, kernel):
    lines = {}
    cur = None
    inside = False
    for ln in open(path):
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            lines[int(m.group(1), 16)] = cur
    return lines


def main():
    counts = load_counts(sys.argv[1])
    lines = load_lines(sys.argv[2], sys.argv[3])
    nw = float(sys.argv[4])
    per_line, per_file, per_op = defaultdict(int), defaultdict(int), defaultdict(int)
    tot = thr = 0
    for off, ie, it, src in counts:
        key = lines.get(off, ("?", 0))
        per_line[key] += ie
        per_file[key[0]] += ie
        per_op[src.split()[0] if not src.startswith("@") else src.split()[1]] += ie
        tot += ie
        thr += it
    print(f"total warp-instructions per warp: {tot / nw:.1f}; thread efficiency {thr / max(tot, 1) / 32:.3f}; static {len(counts)}")
    print("--- per file")
    for f, v in sorted(per_file.items(), key=lambda kv: -kv[1]):
        print(f"{v / nw:9.1f}  {f}")
    print("--- top lines")
    for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1])[:60]:
        print(f"{v / nw:9.1f}  {f}:{l}")
    print("--- top opcodes")
    for op, v in sorted(per_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{v / nw:9.1f}  {op}")


if __name__ == "__main__":
    main()
