#!/usr/bin/env python
"""GPU: the stress distributions of tests/hunts/tol_stress_hunt.py through the CUDA library against the reference library.
  * default (bit-exact) policy: every output of every sample bit-identical (NaN = NaN), every flag equal;
  * tolerance policy: every flag equal (the device's real MUFU errors instead of the host build's emulation).
    python tests/hunts/gpu_stress_parity.py [repetitions of 2^20 samples]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hunts")):
    sys.path.insert(0, p)

import numpy as np
import torch

import oracle_lib as ol
import parity
import tol_stress_hunt as H
from rlshaders_b200 import _abi as abi
from rlshaders_b200 import api


def same_bits(g, c):
    g, c = np.ascontiguousarray(g), np.ascontiguousarray(c)
    eq = (g.view(np.uint32) == c.view(np.uint32)) | (np.isnan(g) & np.isnan(c)) if g.dtype == np.float32 else (g == c)
    return eq.all(axis=0) if eq.ndim == 2 else eq


def main(reps):
    orc = ol.load_ref() or ol.load_port()
    orc.set_threads(0)
    e, t = api.Context(0), api.Context(0)
    t.set_arith_policy("tolerant")
    n = 1 << 20
    dev = lambda a: parity.to_dev(a, e.device)      # noqa: E731
    fails = {}

    def check(unit, out_e, out_t, cpu):
        for k, v in cpu.items():
            g = out_e[k].cpu().numpy()
            g = g.view(np.uint32) if k == "flags" else g
            bad = ~same_bits(g, v)
            if bad.any():
                fails[f"{unit}.{k} (bit-exact policy)"] = fails.get(f"{unit}.{k} (bit-exact policy)", 0) + int(bad.sum())
                print(unit, k, "first bad sample", int(np.nonzero(bad)[0][0]), flush=True)
        bad = out_t["flags"].cpu().numpy().view(np.uint32) != cpu["flags"]
        if bad.any():
            fails[f"{unit}.flags (tolerance policy)"] = fails.get(f"{unit}.flags (tolerance policy)", 0) + int(bad.sum())
            print(unit, "tolerance flags, first bad sample", int(np.nonzero(bad)[0][0]), flush=True)

    for rep in range(reps):
        rng = np.random.default_rng(1000 + rep)
        sg = H.shading(rng, n, 0x57E55 + rep)
        rough = H.mix(rng, n, [0.0, 1e-3, 0.01, 0.05, 0.3, 0.999, 1.0], 0.0, 1.0)
        ior = H.mix(rng, n, [1.0, 1.0001, 0.9999, 0.47, 1e-5, 1.5, 2.5, 1.33], 0.2, 3.0)
        aniso = H.mix(rng, n, [0.0, 1.0, 0.5, 0.999], 0.0, 1.0)
        rx, ry = H.uniforms(rng, n), H.uniforms(rng, n)
        kw = dict(specularRoughness=rough, ior=ior, anisotropic=aniso)
        cpu = orc.ggx_dielectric(sg, abi.ggx_params(**kw), rx, ry)
        outs = []
        for c in (e, t):
            s = api.GgxSampler(c, api.ShadingBatch.from_numpy(sg, c.device), **parity.params_to_dev(kw, c.device))
            outs.append(s.dielectricSampleEvalPdf(dev(rx), dev(ry)))
        torch.cuda.synchronize()
        check("dielectric", outs[0], outs[1], cpu)
        sgc = dict(sg); sgc["backfacing"] = None
        kwc = dict(KsColor=(1.0, 0.5, 0.25), **kw)
        cpu = orc.ggx_sample_eval_pdf(sgc, abi.ggx_params(**kwc), rx, ry)
        outs = []
        for c in (e, t):
            s = api.GgxSampler(c, api.ShadingBatch.from_numpy(sgc, c.device), **parity.params_to_dev(kwc, c.device))
            outs.append(s.sampleEvalPdf(dev(rx), dev(ry)))
        torch.cuda.synchronize()
        check("conductor", outs[0], outs[1], cpu)
        names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
                 "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
        kd = {nm: H.mix(rng, n, [0.0, 1.0, 0.5, 1e-3, 0.999], 0.0, 1.0) for nm in names}
        kd["base_color"] = tuple(H.mix(rng, n, [0.0, 1.0, 0.18], 0.0, 1.0) for _ in range(3))
        u = [H.uniforms(rng, n) for _ in range(4)]
        k = n // 8
        gw = (np.float32(1.0) / (np.float32(1.0) + kd["clearcoat"][:k] * np.float32(0.25))).astype(np.float32)
        u[0][:k] = np.clip(gw * (np.float32(1.0) + rng.choice(np.asarray([-3e-7, -1e-7, 0.0, 1e-7, 3e-7, -1e-4, -3e-5], np.float32), k)),
                           2.0 ** -24, 1 - 2.0 ** -24).astype(np.float32)
        cpu = orc.disney_sample_eval_pdf(sgc, abi.disney_params(**kd), *u)
        outs = []
        for c in (e, t):
            s = api.DisneySampler(c, api.ShadingBatch.from_numpy(sgc, c.device), **parity.params_to_dev(kd, c.device))
            outs.append(s.sampleEvalPdf(*[dev(x) for x in u]))
        torch.cuda.synchronize()
        check("disney", outs[0], outs[1], cpu)
        dist = tuple(H.mix(rng, n, [0.0, 1e-4, 9e-5, 1.1e-4, 1.0, 100.0], 0.0, 3.0) for _ in range(3))
        mult = H.mix(rng, n, [1.0, 0.0, 1e-3], 0.0, 2.0)
        ks = dict(sss_scatter_dist=dist, sss_dist_multiplier=mult)
        cpu = orc.skin_profile(abi.skin_params(**ks), rx)
        outs = []
        for c in (e, t):
            s = api.SkinProfile(c, n, **parity.params_to_dev(ks, c.device))
            outs.append(s.sampleEvalPdf(dev(rx)))
        torch.cuda.synchronize()
        check("skin", outs[0], outs[1], cpu)
        print(f"{rep + 1} x {n} stress samples per unit: failures {fails}", flush=True)
    e.close(); t.close()
    return fails


if __name__ == "__main__":
    f = main(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
    sys.exit(1 if f else 0)
