#!/usr/bin/env python
"""PCIe ceiling vs the *_host pipeline: pinned H2D / D2H / both-way copy bandwidth, then the
rough-dielectric host entry point at several chunk sizes.  Prints one line each."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rlshaders_b200 import api

ctx = api.Context(0)
dev = ctx.device
nbytes = 1 << 30
h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


print("H2D  %.1f GB/s" % (nbytes / timed(lambda: d_a.copy_(h_in, non_blocking=True)) / 1e9))
print("D2H  %.1f GB/s" % (nbytes / timed(lambda: h_out.copy_(d_b, non_blocking=True)) / 1e9))
print("both %.1f GB/s each way" % (nbytes / timed(both) / 1e9))
del h_in, h_out, d_a, d_b

n = 1 << 25
sg = ctx.synth_shading(n, 1, 0, 0.02, 1.0, 0.25)
rough = ctx.synth_uniform(n, 1, 2, 0, 0.05, 1.0)
ior = ctx.synth_uniform(n, 1, 3, 0, 1.05, 2.5)
rx, ry = ctx.synth_uniform(n, 1, 0, 0), ctx.synth_uniform(n, 1, 1, 0)


def pin(t):
    o = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    o.copy_(t)
    return o


hsg = api.ShadingBatch(pin(sg.U), pin(sg.V), pin(sg.N), pin(sg.wo), pin(sg.backfacing))
hs = api.GgxSampler(ctx, hsg, specularRoughness=pin(rough), ior=pin(ior))
hrx, hry = pin(rx), pin(ry)
out = hs.alloc_dielectric_out(hrx)
for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    f = lambda: hs.dielectricSampleEvalPdf(hrx, hry, out=out, chunk=chunk)   # noqa: E731
    t = timed(f, 3)
    print("chunk 2^%d: %.3f G samples/s  (%.1f GB/s up, %.1f GB/s down)" % (chunk.bit_length() - 1, n / t / 1e9, n * 65 / t / 1e9, n * 48 / t / 1e9))
