#!/usr/bin/env python
"""Copy-only probe of the *_host staging pattern (no kernel): what the host <-> device link of THIS box sustains for the
rough-dielectric host entry point's traffic (16 input arrays + 1 byte array, 12 output arrays per chunk), under

  A  the round-1 schedule: K stage streams, each running upload -> (kernel) -> download of its chunk in order;
  B  dedicated streams: every upload on one stream, every download on another, chunk-wise event dependencies
     (download of chunk c waits for its upload, upload of chunk c waits for the download of chunk c - depth);

for the full frame (65 B up, 48 B down per sample) and the compact frame (45 B up).  Run alone or under torchrun: every
rank drives its own GPU at the same time and rank 0 prints the per-GPU and the summed rates -- the host-side ceiling of
the e2e number at N GPUs.

    python tools/pcie_pattern_probe.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_pattern_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("gloo")          # CPU rendezvous only: the GPUs carry nothing but the copies


def run(pattern, n_in, n_out, chunk_samples=1 << 21, chunks=24, depth=3):
    nb = chunk_samples * 4
    hin = [torch.empty(nb * chunks, dtype=torch.uint8, pin_memory=True) for _ in range(n_in)]
    hout = [torch.empty(nb * chunks, dtype=torch.uint8, pin_memory=True) for _ in range(n_out)]
    din = [[torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(n_in)] for _ in range(depth)]
    dout = [[torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(n_out)] for _ in range(depth)]
    if pattern == "A":
        ss = [torch.cuda.Stream() for _ in range(depth)]

        def go():
            for c in range(chunks):
                s = c % depth
                with torch.cuda.stream(ss[s]):
                    for k in range(n_in):
                        din[s][k].copy_(hin[k][c * nb:(c + 1) * nb], non_blocking=True)
                    for k in range(n_out):
                        hout[k][c * nb:(c + 1) * nb].copy_(dout[s][k], non_blocking=True)
    else:
        up, down = torch.cuda.Stream(), torch.cuda.Stream()

        def go():
            e_up = [None] * chunks
            e_down = [None] * chunks
            for c in range(chunks):
                s = c % depth
                with torch.cuda.stream(up):
                    if c >= depth:
                        up.wait_event(e_down[c - depth])
                    for k in range(n_in):
                        din[s][k].copy_(hin[k][c * nb:(c + 1) * nb], non_blocking=True)
                    e_up[c] = up.record_event()
                with torch.cuda.stream(down):
                    down.wait_event(e_up[c])
                    for k in range(n_out):
                        hout[k][c * nb:(c + 1) * nb].copy_(dout[s][k], non_blocking=True)
                    e_down[c] = down.record_event()
    go()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    go()
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    up_gbs, dn_gbs = nb * n_in * chunks / t / 1e9, nb * n_out * chunks / t / 1e9
    rates = [None] * world
    if world > 1:
        dist.all_gather_object(rates, (up_gbs, dn_gbs, chunk_samples * chunks / t / 1e9))
    else:
        rates = [(up_gbs, dn_gbs, chunk_samples * chunks / t / 1e9)]
    if rank == 0:
        print(json.dumps({"pattern": pattern, "arrays_in": n_in, "arrays_out": n_out, "gpus": world,
                          "up_GBs_per_gpu": [round(r[0], 1) for r in rates], "down_GBs_per_gpu": [round(r[1], 1) for r in rates],
                          "up_GBs_sum": round(sum(r[0] for r in rates), 1), "down_GBs_sum": round(sum(r[1] for r in rates), 1),
                          "G_samples_s_sum": round(sum(r[2] for r in rates), 3)}), flush=True)
    del hin, hout, din, dout


for pat in ("A", "B"):
    run(pat, 16, 12)          # full frame: 16 float arrays (+ the byte array, neglected) up, 12 down
    run(pat, 11, 12)          # compact frame (unit quaternion): 11 float arrays up
if world > 1:
    dist.destroy_process_group()
