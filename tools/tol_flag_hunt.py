"""GPU diagnostic: run the fused units under the bit-exact and the tolerance policy on many device-generated samples and
dump every sample whose flags differ (inputs + both outputs) to gpurun_out/tol_flag_hunt.npz for host-side replay."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlshaders_b200 import api  # noqa: E402

n = 1 << 24
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
e, t = api.Context(0), api.Context(0)
t.set_arith_policy("tolerant")
found = {}
for rep in range(reps):
    seed = 0x5EED0002 + 7919 * rep
    sg = e.synth_shading(n, seed, 0, 0.02, 1.0, 0.25)
    u = [e.synth_uniform(n, seed, s) for s in range(4)]
    rough, ior = e.synth_uniform(n, seed, 2, 0, 0.05, 1.0), e.synth_uniform(n, seed, 3, 0, 1.05, 2.5)
    a = api.GgxSampler(e, sg, specularRoughness=rough, ior=ior).dielectricSampleEvalPdf(u[0], u[1])
    b = api.GgxSampler(t, sg, specularRoughness=rough, ior=ior).dielectricSampleEvalPdf(u[0], u[1])
    torch.cuda.synchronize()
    idx = torch.nonzero(a["flags"] != b["flags"]).flatten()
    print("dielectric rep", rep, "mismatches", idx.numel(), flush=True)
    if idx.numel():
        d = found.setdefault("dielectric", [])
        d.append(dict(U=sg.U[:, idx].cpu().numpy(), V=sg.V[:, idx].cpu().numpy(), N=sg.N[:, idx].cpu().numpy(),
                      wo=sg.wo[:, idx].cpu().numpy(), back=sg.backfacing[idx].cpu().numpy(), rough=rough[idx].cpu().numpy(),
                      ior=ior[idx].cpu().numpy(), rx=u[0][idx].cpu().numpy(), ry=u[1][idx].cpu().numpy(),
                      **{"exact_" + k: v[..., idx].cpu().numpy() for k, v in a.items()},
                      **{"tol_" + k: v[..., idx].cpu().numpy() for k, v in b.items()}))
    del a, b
    names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
             "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
    kw = {nm: e.synth_uniform(n, seed, 20 + j) for j, nm in enumerate(names)}
    kw["base_color"] = tuple(e.synth_uniform(n, seed, 30 + j) for j in range(3))
    sg3 = api.ShadingBatch(sg.U, sg.V, sg.N, sg.wo)
    a = api.DisneySampler(e, sg3, **kw).sampleEvalPdf(*u)
    b = api.DisneySampler(t, sg3, **kw).sampleEvalPdf(*u)
    torch.cuda.synchronize()
    idx = torch.nonzero(a["flags"] != b["flags"]).flatten()
    print("disney rep", rep, "mismatches", idx.numel(), flush=True)
    if idx.numel():
        d = found.setdefault("disney", [])
        d.append(dict(U=sg.U[:, idx].cpu().numpy(), V=sg.V[:, idx].cpu().numpy(), N=sg.N[:, idx].cpu().numpy(),
                      wo=sg.wo[:, idx].cpu().numpy(), u=np.stack([x[idx].cpu().numpy() for x in u]),
                      **{"p_" + k: (np.stack([c[idx].cpu().numpy() for c in v]) if isinstance(v, tuple) else v[idx].cpu().numpy())
                         for k, v in kw.items()},
                      **{"exact_" + k: v[..., idx].cpu().numpy() for k, v in a.items()},
                      **{"tol_" + k: v[..., idx].cpu().numpy() for k, v in b.items()}))
    del a, b, kw, sg3
    dist = tuple(e.synth_uniform(n, seed, 50 + j, 0, 0.05, 2.0) for j in range(3))
    a = api.SkinProfile(e, n, sss_scatter_dist=dist).sampleEvalPdf(u[0])
    b = api.SkinProfile(t, n, sss_scatter_dist=dist).sampleEvalPdf(u[0])
    torch.cuda.synchronize()
    idx = torch.nonzero(a["flags"] != b["flags"]).flatten()
    print("skin rep", rep, "mismatches", idx.numel(), flush=True)
    if idx.numel():
        d = found.setdefault("skin", [])
        d.append(dict(dist=np.stack([c[idx].cpu().numpy() for c in dist]), rx=u[0][idx].cpu().numpy(),
                      **{"exact_" + k: v[..., idx].cpu().numpy() for k, v in a.items()},
                      **{"tol_" + k: v[..., idx].cpu().numpy() for k, v in b.items()}))
    del a, b, dist, sg, u, rough, ior
    torch.cuda.empty_cache()
flat = {}
for w, lst in found.items():
    for key in lst[0]:
        flat[w + "." + key] = np.concatenate([x[key] for x in lst], axis=-1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "tol_flag_hunt.npz"), **flat)
print("saved", {k: v.shape for k, v in flat.items() if k.endswith("flags")})
