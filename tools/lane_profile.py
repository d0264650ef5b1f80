#!/usr/bin/env python
"""Where a kernel's warps run with inactive lanes: per-instruction executed counts and active-thread counts of an
ncu source page, grouped into contiguous regions and (joined with nvdisasm -g line info) into source lines.

    ncu -i prof.ncu-rep --page source --csv > src.csv            # capture taken with --import-source on, -lineinfo build
    cuobjdump -xelf all librls_b200.so ; nvdisasm -g -c *.cubin > all.sass
    python tools/lane_profile.py src.csv all.sass <mangled-kernel-substring> <n_warps> [max_lanes=29.5]

Prints (1) contiguous SASS regions that most warps execute with fewer than `max_lanes` active lanes -- a branch whose
sides run the SAME operations on different operands shows as two regions of ~complementary lane counts and is a
candidate for the select form; sides with different operations can only be reclaimed by re-grouping
samples -- and (2) the source lines behind them.  Used for profiles/r01_ncu_summary.md (rlDisney, rough dielectric)."""
import sys
from collections import defaultdict

import sass_profile as sp


def main():
    counts = sp.load_counts(sys.argv[1])
    lines = sp.load_lines(sys.argv[2], sys.argv[3])
    nw = float(sys.argv[4])
    max_lanes = float(sys.argv[5]) if len(sys.argv) > 5 else 29.5
    total = sum(ie for _, ie, _, _ in counts)
    threads = sum(it for _, _, it, _ in counts)
    print(f"warp instructions per warp {total / nw:.1f}; threads per instruction {threads / max(total, 1):.2f}")
    regions, cur = [], None
    for off, ie, it, src in counts:
        if ie == 0:
            continue
        lanes = it / ie
        if cur and abs(cur["ie"] - ie) < 0.02 * nw and abs(cur["lanes"] - lanes) < 1.0:
            cur["n"] += 1
            cur["slots"] += ie
        else:
            cur = dict(ie=ie, lanes=lanes, n=1, slots=ie, first=src, off=off)
            regions.append(cur)
    print("--- regions executed by >= 30 % of the warps with inactive lanes (>= 3 slots per warp)")
    for r in regions:
        if r["slots"] / nw >= 3 and r["ie"] > 0.3 * nw and r["lanes"] < max_lanes:
            key = lines.get(r["off"], ("?", 0))
            print(f"{r['slots'] / nw:7.1f} slots/warp  {r['lanes']:5.1f} lanes  in {r['ie'] / nw:5.3f} of the warps  "
                  f"{key[0]}:{key[1]}  {r['first'][:44]}")
    per_line = defaultdict(lambda: [0, 0])
    for off, ie, it, _ in counts:
        if ie > 0.3 * nw and it / ie < max_lanes:
            k = lines.get(off, ("?", 0))
            per_line[k][0] += ie
            per_line[k][1] += it
    print("--- source lines behind them")
    for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f"{v[0] / nw:7.1f} slots/warp  {v[1] / v[0]:5.1f} lanes  {k[0]}:{k[1]}")


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    main()
