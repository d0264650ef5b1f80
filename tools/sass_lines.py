#!/usr/bin/env python
"""List the SASS instructions (with executed warp-instructions per warp) that nvdisasm
attributes to a source-line range.  Companion of sass_profile.py.

    python tools/sass_lines.py sass.csv all.sass <kernel-substring> <n_warps> <file> <first> <last>
"""
import sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
import sass_profile as sp


def main():
    counts = sp.load_counts(sys.argv[1])
    lines = sp.load_lines(sys.argv[2], sys.argv[3])
    nw = float(sys.argv[4])
    fname, lo, hi = sys.argv[5], int(sys.argv[6]), int(sys.argv[7])
    tot = 0.0
    for off, ie, it, src in counts:
        f, l = lines.get(off, ("?", 0))
        if f == fname and lo <= l <= hi:
            tot += ie / nw
            print(f"{off:6x} {ie / nw:6.2f} {f}:{l:<4d} {src}")
    print(f"total {tot:.1f}")


if __name__ == "__main__":
    main()
