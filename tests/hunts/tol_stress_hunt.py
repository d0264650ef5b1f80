#!/usr/bin/env python
"""CPU hunt for flag mismatches of the tolerance policy on STRESS distributions (tests/native host build of
csrc/rls_tol.cuh against the reference library): parameters drawn from mixtures that put mass on the end points and
thresholds of every range (roughness 0 / 1e-3 / 1, ior 1 / 1 +- 1e-4 / 0.47 / 1e-5, anisotropic 0 / 1, rlDisney
parameters exactly 0 and 1, clearcoat_gloss at both ends), views from grazing to normal incidence and below the horizon,
uniforms at 2^-24, 1 - 2^-24, the lobe boundaries and the thirds.  A sample counts as a failure when its flags differ
from the reference's AND the band tracker did not list it for the bit-exact re-run.

    python tests/hunts/tol_stress_hunt.py [repetitions of 2^20 samples] [--ulp] [--quat] [--seed K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np

import oracle_lib as ol
import tol_host as th
from rlshaders_b200 import _abi as abi

f32 = np.float32


def mix(rng, n, special, lo, hi, p_special=0.5):
    """Half the mass on the listed special values, half uniform in [lo, hi]."""
    u = rng.uniform(lo, hi, n).astype(f32)
    s = rng.choice(np.asarray(special, f32), n).astype(f32)
    return np.where(rng.random(n) < p_special, s, u).astype(f32)


def uniforms(rng, n, extra=()):
    special = [2.0 ** -24, 1 - 2.0 ** -24, 0.5, 0.5 + 2.0 ** -24, 0.3333, 0.6666, 0.25, 0.75] + list(extra)
    return np.clip(mix(rng, n, special, 0.0, 1.0, 0.3), 2.0 ** -24, 1 - 2.0 ** -24).astype(f32)


SEED = int(sys.argv[sys.argv.index("--seed") + 1]) if "--seed" in sys.argv else 0      # another family of random streams
QUAT = "--quat" in sys.argv      # frames decoded from unit quaternions (compact host forms): orthonormal to ~3e-7 only


def shading(rng, n, seed):
    sg = ol.make_shading(n, seed, cos_lo=-0.05, cos_hi=1.0, backfacing_fraction=0.3)
    if QUAT:
        sg.update(ol.frame_from_quaternion(ol.quaternion_from_frame(sg)))
    k = n // 16                                   # views exactly along the normal / in the tangent plane
    for c in "xyz":
        sg["wo" + c][:k] = sg["N" + c][:k]
        sg["wo" + c][k:2 * k] = sg["U" + c][k:2 * k]
    return sg


def run(reps, ulp):
    lib = th.load(ulp=ulp)
    orc = ol.load_ref() or ol.load_port()
    orc.set_threads(0)
    n = 1 << 20
    bad_total = {"dielectric": 0, "conductor": 0, "disney": 0, "skin": 0}
    listed = {k: 0 for k in bad_total}
    for rep in range(reps):
        rng = np.random.default_rng(1000 + rep + 100003 * SEED)
        sg = shading(rng, n, 0x57E55 + rep + 7919 * SEED)
        rough = mix(rng, n, [0.0, 1e-3, 0.01, 0.05, 0.3, 0.999, 1.0], 0.0, 1.0)
        ior = mix(rng, n, [1.0, 1.0001, 0.9999, 0.47, 1e-5, 1.5, 2.5, 1.33], 0.2, 3.0)
        aniso = mix(rng, n, [0.0, 1.0, 0.5, 0.999], 0.0, 1.0)
        rx, ry = uniforms(rng, n), uniforms(rng, n)
        p = abi.ggx_params(specularRoughness=rough, ior=ior, anisotropic=aniso)
        t, rr = th.ggx_dielectric(lib, sg, p, rx, ry)
        o = orc.ggx_dielectric(sg, p, rx, ry)
        bad = (t["flags"] != o["flags"]) & (rr == 0)
        bad_total["dielectric"] += int(bad.sum()); listed["dielectric"] += int(rr.sum())
        if bad.any():
            i = int(np.nonzero(bad)[0][0])
            print("dielectric", rep, i, "flags %08x %08x" % (t["flags"][i], o["flags"][i]), rough[i], ior[i], aniso[i], rx[i], ry[i], flush=True)
        sgc = dict(sg); sgc["backfacing"] = None
        pc = abi.ggx_params(KsColor=(1.0, 0.5, 0.25), specularRoughness=rough, ior=ior, anisotropic=aniso)
        t, rr = th.ggx_conductor(lib, sgc, pc, rx, ry)
        o = orc.ggx_sample_eval_pdf(sgc, pc, rx, ry)
        bad = (t["flags"] != o["flags"]) & (rr == 0)
        bad_total["conductor"] += int(bad.sum()); listed["conductor"] += int(rr.sum())
        if bad.any():
            i = int(np.nonzero(bad)[0][0])
            print("conductor", rep, i, "flags %08x %08x" % (t["flags"][i], o["flags"][i]), rough[i], ior[i], aniso[i], rx[i], ry[i], flush=True)
        names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
                 "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
        kw = {nm: mix(rng, n, [0.0, 1.0, 0.5, 1e-3, 0.999], 0.0, 1.0) for nm in names}
        kw["base_color"] = tuple(mix(rng, n, [0.0, 1.0, 0.18], 0.0, 1.0) for _ in range(3))
        u = [uniforms(rng, n) for _ in range(4)]
        # rx_s next to the lobe boundary 1 / (1 + clearcoat / 4), from both sides
        k = n // 8
        gw = (f32(1.0) / (f32(1.0) + kw["clearcoat"][:k] * f32(0.25))).astype(f32)
        u[0][:k] = np.clip(gw * (f32(1.0) + rng.choice(np.asarray([-3e-7, -1e-7, 0.0, 1e-7, 3e-7, -1e-4, -3e-5], f32), k)), 2.0 ** -24, 1 - 2.0 ** -24).astype(f32)
        pd = abi.disney_params(**kw)
        t, rr = th.disney(lib, sgc, pd, *u)
        o = orc.disney_sample_eval_pdf(sgc, pd, *u)
        bad = (t["flags"] != o["flags"]) & (rr == 0)
        bad_total["disney"] += int(bad.sum()); listed["disney"] += int(rr.sum())
        if bad.any():
            i = int(np.nonzero(bad)[0][0])
            print("disney", rep, i, "flags %08x %08x" % (t["flags"][i], o["flags"][i]), {nm: float(kw[nm][i]) for nm in names}, [float(x[i]) for x in u], flush=True)
        dist = tuple(mix(rng, n, [0.0, 1e-4, 9e-5, 1.1e-4, 1.0, 100.0], 0.0, 3.0) for _ in range(3))
        mult = mix(rng, n, [1.0, 0.0, 1e-3], 0.0, 2.0)
        ps = abi.skin_params(sss_scatter_dist=dist, sss_dist_multiplier=mult)
        t, rr = th.skin_profile(lib, ps, rx)
        o = orc.skin_profile(ps, rx)
        bad = (t["flags"] != o["flags"]) & (rr == 0)
        bad_total["skin"] += int(bad.sum()); listed["skin"] += int(rr.sum())
        if bad.any():
            i = int(np.nonzero(bad)[0][0])
            print("skin", rep, i, "flags %08x %08x" % (t["flags"][i], o["flags"][i]), [float(c[i]) for c in dist], float(mult[i]), float(rx[i]), flush=True)
        if rep % 8 == 7 or rep == reps - 1:
            print(f"{rep + 1} x {n} samples per unit ({'ulp-perturbed' if ulp else 'plain'}): flag failures {bad_total}, "
                  f"listed fraction { {k: round(v / ((rep + 1) * n), 4) for k, v in listed.items()} }", flush=True)
    return bad_total


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 16
    tot = run(reps, "--ulp" in sys.argv)
    sys.exit(1 if any(tot.values()) else 0)
