"""Device-timed rlDisney fused kernel (BASELINE configs[2] inputs, 2^26 samples) -- A/B of a library switch:
   RLS_DISNEY_LOBE_SORT=0 python tools/disney_ab.py ; RLS_DISNEY_LOBE_SORT=1 python tools/disney_ab.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlshaders_b200 import api  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
SEED = 0x5EED0003
ctx = api.Context(0)
sg = ctx.synth_shading(n, SEED, 0, 0.02, 1.0, 0.0)
names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic", "sheen", "sheen_tint",
         "clearcoat", "clearcoat_gloss"]
kw = {nm: ctx.synth_uniform(n, SEED, 20 + j, 0) for j, nm in enumerate(names)}
base = tuple(ctx.synth_uniform(n, SEED, 30 + j, 0) for j in range(3))
kw["base_color"] = base
u = [ctx.synth_uniform(n, SEED, s, 0) for s in range(4)]
smp = api.DisneySampler(ctx, sg, **kw)
out = smp.alloc_out(u[0])
step = lambda: smp.sampleEvalPdf(*u, out=out)  # noqa: E731
for _ in range(3):
    step()
torch.cuda.synchronize()
ctx.fallback_count(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
chk = int(sum(int(v.view(torch.int32).sum(dtype=torch.int64)) for v in out.values()) & 0xFFFFFFFFFFFF)
print(json.dumps({"kernel": "k_disney_sample_eval_pdf", "lobe_sort": os.environ.get("RLS_DISNEY_LOBE_SORT", "default"),
                  "samples": n, "ms": ms, "samples_per_s": n / (ms * 1e-3), "output_checksum": chk,
                  "exact_rerun_fraction": ctx.fallback_count(reset=True) / float(n * steps)}))
