"""Device-timed throughput of the fused GaussianProfile unit (k_gauss_profile): 2 input + 3 output floats
per sample = 20 algorithmic bytes, against the measured HBM peak.  Usage: python tools/gauss_bench.py [log2_n]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlshaders_b200 import api  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 27)
ctx = api.Context(0)
dist_x = ctx.synth_uniform(n, 0x5EED0006, 1, 0, 0.05, 2.0)
rx = ctx.synth_uniform(n, 0x5EED0006, 0, 0)
out = [ctx.empty(n) for _ in range(3)]


def step():
    rc = ctx.lib.rls_gaussprofile_sample_eval_pdf(ctx.handle, n, dist_x.data_ptr(), rx.data_ptr(),
                                                  out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
    assert rc == 0


peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6541.1)
res = {"kernel": "k_gauss_profile", "samples": n, "algorithmic_bytes_per_sample": 20, "peak_gbs": peak,
       "l2": "working set %.1f GB >> 126 MB L2" % (n * 20 / 1e9)}
for policy in ("exact", "fast"):
    ctx.set_arith_policy(policy)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ctx.fallback_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gbs = n * 20 / (ms * 1e-3) / 1e9
    res[policy] = {"ms": ms, "samples_per_s": n / (ms * 1e-3), "achieved_gbs": gbs, "hbm_frac": gbs / peak,
                   "exact_rerun_fraction": ctx.fallback_count(reset=True) / float(n * steps)}
print(json.dumps(res))
