#!/usr/bin/env python
"""CPU hunt for flag mismatches of the tolerance policy on the BENCH distributions (tests/native host build of
csrc/rls_tol.cuh against the reference library; tools/tol_flag_hunt.py is the device version, tests/hunts/tol_stress_hunt.py the
one on stress distributions).  Each repetition draws 2^22 samples per unit from the recipes of BASELINE configs 2
(isotropic and anisotropic alternating) and 3 with a new seed; a failure is a sample whose flags differ from the
reference's without the band tracker having listed it for the bit-exact re-run.

    python tests/hunts/tol_flag_hunt_host.py [repetitions] [--ulp] [--seed K]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle_lib as ol
import parity
import tol_host as th
from rlshaders_b200 import _abi as abi


def run(reps, ulp, seed0=0):
    lib = th.load(ulp=ulp)
    orc = ol.load_ref() or ol.load_port()
    orc.set_threads(0)
    n = 1 << 22
    tot = dict(dielectric=0, disney=0)
    listed = dict(dielectric=0, disney=0)
    t0 = time.time()
    for rep in range(reps):
        seed = 0x5EED0002 + 7919 * (rep + 100 + 100000 * seed0)
        sg, kw, rx, ry = parity.ggx_dielectric_inputs(n, seed, rep % 2 == 1)
        p = abi.ggx_params(**kw)
        o = orc.ggx_dielectric(sg, p, rx, ry)
        t, rr = th.ggx_dielectric(lib, sg, p, rx, ry)
        bad = (t["flags"] != o["flags"]) & (rr == 0)
        tot["dielectric"] += int(bad.sum()); listed["dielectric"] += int(rr.sum())
        if bad.any():
            print("dielectric rep", rep, "first bad sample", int(bad.nonzero()[0][0]), flush=True)
        sg, kw, u = parity.disney_inputs(n, seed + 1)
        p = abi.disney_params(**kw)
        o = orc.disney_sample_eval_pdf(sg, p, *u)
        t, rr = th.disney(lib, sg, p, *u)
        bad = (t["flags"] != o["flags"]) & (rr == 0)
        tot["disney"] += int(bad.sum()); listed["disney"] += int(rr.sum())
        if bad.any():
            print("disney rep", rep, "first bad sample", int(bad.nonzero()[0][0]), flush=True)
        if rep % 8 == 7 or rep == reps - 1:
            print(f"{rep + 1} x {n} samples per unit ({'ulp-perturbed' if ulp else 'plain'}): flag failures {tot}, listed fraction "
                  f"{ {k: round(v / ((rep + 1) * n), 6) for k, v in listed.items()} }  [{time.time() - t0:.0f} s]", flush=True)
    return tot


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 8
    seed0 = int(sys.argv[sys.argv.index("--seed") + 1]) if "--seed" in sys.argv else 0
    bad = run(reps, "--ulp" in sys.argv, seed0)
    sys.exit(1 if any(bad.values()) else 0)
