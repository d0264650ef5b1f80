#!/usr/bin/env python
"""How far is "within tolerance of the reference" reachable at all?  For every output of the four fused units this prints,
sample by sample against a binary64 evaluation of the SAME algorithm (oracle/rls_oracle_f64.c):

    ref   |reference (binary32) - f64|   the reference's own rounding noise
    tol   |RLS_ARITH_TOLERANT  - f64|   the tolerance policy's error (CPU build of csrc/rls_tol.cuh, tests/native)
    t-r   |RLS_ARITH_TOLERANT  - reference|

as the fraction of samples within 1e-6 / 1e-5 / ... (directions: absolute, max component; values: relative).  Samples the
band tracker sends to the bit-exact re-run, and samples on which the binary64 evaluation takes a different discrete branch
(flags differ), are left out of all three columns.  CPU only.   Usage: tol_vs_f64.py [n] [--json out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

import oracle_lib as ol
import parity
import tol_host as th
from rlshaders_b200 import _abi as abi


def table(kinds, ref, tolr, f64, keep):
    rows = {}
    for name, kind in kinds.items():
        if kind == "flags":
            continue
        cols = {}
        for label, a, b in (("ref", ref[name], f64[name]), ("tol", tolr[name], f64[name]), ("t-r", tolr[name], ref[name])):
            e = th.errors(kind, a, b)[keep]
            cols[label] = {"within": {f"{t:g}": float((e <= t).mean()) for t in th.TOLS[kind]},
                           "p50": float(np.median(e)), "p99": float(np.quantile(e, 0.99)), "p9999": float(np.quantile(e, 0.9999))}
        rows[name] = cols
    return rows


def run(n):
    lib, ref, f64 = th.load(), ol.load_ref() or ol.load_port(), ol.load_f64()
    ref.set_threads(0)
    out = {}
    sg, kw, rx, ry = parity.ggx_dielectric_inputs(n, 0x5EED0002, False)
    p = abi.ggx_params(**kw)
    t, rerun = th.ggx_dielectric(lib, sg, p, rx, ry)
    r, d = ref.ggx_dielectric(sg, p, rx, ry), f64.ggx_dielectric(sg, p, rx, ry)
    keep = (rerun == 0) & (r["flags"] == d["flags"])
    out["ggx_dielectric"] = {"n": n, "kept": float(keep.mean()), "outputs": table(th.KINDS_DIELECTRIC, r, t, d, keep)}
    sg, p, rx, ry = ol.workload_ggx_conductor(n)
    t, rerun = th.ggx_conductor(lib, sg, p, rx, ry)
    r, d = ref.ggx_sample_eval_pdf(sg, p, rx, ry), f64.ggx_sample_eval_pdf(sg, p, rx, ry)
    keep = (rerun == 0) & (r["flags"] == d["flags"])
    out["ggx_conductor"] = {"n": n, "kept": float(keep.mean()), "outputs": table(th.KINDS_GGX, r, t, d, keep)}
    sg, kw, u = parity.disney_inputs(n, 0x5EED0003)
    p = abi.disney_params(**kw)
    t, rerun = th.disney(lib, sg, p, *u)
    r, d = ref.disney_sample_eval_pdf(sg, p, *u), f64.disney_sample_eval_pdf(sg, p, *u)
    keep = (rerun == 0) & (r["flags"] == d["flags"])
    out["disney"] = {"n": n, "kept": float(keep.mean()), "outputs": table(th.KINDS_DISNEY, r, t, d, keep)}
    kw, rx = parity.skin_inputs(n, 0x5EED0004)
    p = abi.skin_params(**kw)
    t, rerun = th.skin_profile(lib, p, rx)
    r, d = ref.skin_profile(p, rx), f64.skin_profile(p, rx)
    keep = (rerun == 0) & (r["flags"] == d["flags"])
    out["skin_profile"] = {"n": n, "kept": float(keep.mean()), "outputs": table(th.KINDS_SKIN, r, t, d, keep)}
    return out


def show(out):
    for unit, blk in out.items():
        print(f"{unit}: n = {blk['n']}, compared {blk['kept'] * 100:.3f} % (rest: band re-run or a different branch in binary64)")
        for name, cols in blk["outputs"].items():
            for label in ("ref", "tol", "t-r"):
                c = cols[label]
                w = "  ".join(f"<={t}: {v * 100:8.4f}%" for t, v in c["within"].items())
                print(f"  {name:9s} {label:4s} {w}  p50 {c['p50']:.1e} p99 {c['p99']:.1e} p99.99 {c['p9999']:.1e}")


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 1 << 20
    res = run(n)
    show(res)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
            json.dump(res, fh, indent=1)
