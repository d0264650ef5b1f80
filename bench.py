#!/usr/bin/env python
"""bench.py -- BSDF samples/s (eval+sample+pdf) of the B200-native rlShaders hot path.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

One "step" = one pass of the fused rough-dielectric unit (BASELINE.json configs[1]:
GGX reflection + refraction, Walter 2007, 2^26 samples, per-sample roughness and IOR)
over one synthetic batch resident in HBM.  Under torchrun (N > 1) every rank owns one GPU
and an independent 2^26-sample slice of the index-addressed synthetic stream (weak
scaling, no data-path collective); rank 0 prints ONE JSON line.

Keys beyond the base contract:
  roofline      dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own C++ (oracle/_ref) or its C port timed on this box's
                host cores on a bounded prefix of the same input bits (rank 0, N = 1)
  e2e           same metric through the host-buffer C-ABI entry point (pinned host
                buffers, H2D + kernel + D2H inside the timed region)
  other_workloads  device-timed throughput of BASELINE.json configs 1, 3, 4, 5
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "BSDF samples/sec (eval+sample+pdf)"
UNIT = "samples/s"
N_DIELECTRIC = 1 << 26          # configs[1]
SEED = 0x5EED0002
# Algorithmic bytes per sample (SURVEY.md 8(a) size table; DESIGN.md "Data layout"):
B_ALG = {"ggx_conductor": 88, "ggx_dielectric": 113, "disney": 176, "skin_profile": 52}
# Algorithmic FP32 work per sample (every add/mul/compare/select = 1, an FMA = 2, each
# division, root and transcendental = 1, worst-case branch), and the nominal FP32 peak of one B200.
F_ALG = {"ggx_conductor": 430, "ggx_dielectric": 680, "disney": 595, "skin_profile": 130}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.4; FFMA-only microbenchmark on this pool: 67.5
WORKLOAD = "ggx_dielectric_64M (BASELINE configs[1]: GGX rough dielectric reflection+refraction, " \
           "2^26 samples/GPU, per-sample roughness~U[0.05,1] ior~U[1.05,2.5], 25% back-facing)"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel):
    """Per-sample DRAM bytes of `kernel` from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """Polls SM clock and clock-event reasons while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ============================================================ reference arm (CPU)
def host_dielectric_inputs(n, seed, first_index=0):
    """numpy inputs of configs[1] (same recipe as the device generators)."""
    import oracle_lib as ol
    sg = ol.make_shading(n, seed, backfacing_fraction=0.25)
    kw = dict(specularRoughness=ol.hash_uniform(n, seed, 2, first_index, 0.05, 1.0),
              ior=ol.hash_uniform(n, seed, 3, first_index, 1.05, 2.5))
    return sg, kw, ol.hash_uniform(n, seed, 0, first_index), ol.hash_uniform(n, seed, 1, first_index)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as ol
    from rlshaders_b200 import _abi as abi
    orc = ol.load_ref() or ol.load_port()
    kind = "reference" if orc.kind == "reference" else "port"
    orc.set_threads(0)
    cores = orc.max_threads()
    n = 1 << 23     # bounded sample of the 2^26-sample workload per step
    sg, kw, rx, ry = host_dielectric_inputs(n, SEED)
    params = abi.ggx_params(**kw)
    for _ in range(args.warmup):
        orc.ggx_dielectric(sg, params, rx, ry)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.ggx_dielectric(sg, params, rx, ry)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} samples/step (first 2^23 of the 2^26-sample workload), {args.steps} steps, all host threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host
    buffers of the e2e leg are allocated (first touch) on the GPU's own NUMA node and the
    H2D / D2H copies do not cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ============================================================== product arm (GPU)
def run_product(args):
    import torch
    import torch.distributed as dist
    from rlshaders_b200 import _abi as abi
    from rlshaders_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product arm has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    ctx.set_arith_policy(args.arith)
    dev = ctx.device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    peak, peak_src = measured_peak()

    # ---- main workload: configs[1], this rank's slice of the synthetic stream
    n = args.samples
    first = rank * n
    sg = ctx.synth_shading(n, SEED, first, 0.02, 1.0, 0.25)
    rough = ctx.synth_uniform(n, SEED, 2, first, 0.05, 1.0)
    ior = ctx.synth_uniform(n, SEED, 3, first, 1.05, 2.5)
    rx = ctx.synth_uniform(n, SEED, 0, first)
    ry = ctx.synth_uniform(n, SEED, 1, first)
    sampler = api.GgxSampler(ctx, sg, specularRoughness=rough, ior=ior)
    out = sampler.alloc_dielectric_out(rx)
    step = lambda: sampler.dielectricSampleEvalPdf(rx, ry, out=out)   # noqa: E731

    clocks = ClockSampler(local)
    for _ in range(args.warmup):
        step()
    barrier()
    ctx.fallback_count(reset=True)
    clocks.start()
    launches1 = ctx.kernel_launches
    ms = timed(step, args.steps, 0)
    gpu_launches = ctx.kernel_launches - launches1
    # samples the fast arithmetic policy re-ran with the guarded IEEE operators (rls_fp.cuh)
    arith = {"policy": args.arith, "exact_rerun_fraction": ctx.fallback_count(reset=True) / float(n * args.steps)}
    value = world * n / (ms * 1e-3)
    achieved = n * B_ALG["ggx_dielectric"] / (ms * 1e-3) / 1e9
    traffic = measured_traffic("k_ggx_dielectric")
    roofline = {"bound": "hbm", "kernel": "k_ggx_dielectric", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_sample": B_ALG["ggx_dielectric"],
                "traffic": traffic * n if traffic else None}

    # ---- e2e: host (pinned) buffers through the *_host C-ABI entry point
    # The staged pipeline is PCIe-bound and its rate does not depend on n beyond a few chunks,
    # so the e2e leg runs on a prefix of the step's batch: all of it at N = 1 (bounded by
    # --e2e-samples), 2^24 samples per rank at N > 1 to keep pinned host memory per box small.
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ne = min(n, args.e2e_samples if world == 1 else min(args.e2e_samples, 1 << 24))

    def pin(t):
        src = t[..., :ne]
        out_t = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        out_t.copy_(src)
        return out_t
    hsg = api.ShadingBatch(pin(sg.U), pin(sg.V), pin(sg.N), pin(sg.wo), pin(sg.backfacing))
    hrough, hior, hrx, hry = pin(rough), pin(ior), pin(rx), pin(ry)
    hsampler = api.GgxSampler(ctx, hsg, specularRoughness=hrough, ior=hior)
    hout = hsampler.alloc_dielectric_out(hrx)
    h2d = sum(t.numel() * t.element_size() for t in (hsg.U, hsg.V, hsg.N, hsg.wo, hsg.backfacing, hrough, hior, hrx, hry))
    d2h = sum(t.numel() * t.element_size() for t in hout.values())
    hstep = lambda: hsampler.dielectricSampleEvalPdf(hrx, hry, out=hout)   # noqa: E731
    hstep()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hstep()
        _ = float(hout["pdf_r"][0])      # device->host result is read on the host
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_s = max_over_ranks(e2e_s * 1e3) * 1e-3
    clock_info = clocks.stop()
    # FP32 roofline beside the HBM one: algorithmic flops, and the flops the kernel really issues
    # (no FMA contraction, IEEE division / root sequences, the host libm's algorithms).
    issued_flops = measured_traffic("fp32_flops_per_sample_issued")
    per_gpu = n / (ms * 1e-3)
    roofline["fp32"] = {"flops_per_sample_algorithmic": F_ALG["ggx_dielectric"],
                        "achieved_tflops": per_gpu * F_ALG["ggx_dielectric"] / 1e12,
                        "peak_tflops": FP32_PEAK_TFLOPS, "frac": per_gpu * F_ALG["ggx_dielectric"] / 1e12 / FP32_PEAK_TFLOPS,
                        "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (FFMA-only microbenchmark: 67.5, "
                                       "profiles/r01_ffma2_microbench.txt)"}
    if isinstance(issued_flops, dict) and issued_flops.get("k_ggx_dielectric"):
        roofline["fp32"]["flops_per_sample_issued"] = issued_flops["k_ggx_dielectric"]
        roofline["fp32"]["issued_tflops"] = per_gpu * issued_flops["k_ggx_dielectric"] / 1e12
    # The kernel is issue bound, not HBM bound: also report the fraction of the SM issue
    # rate it sustains = executed warp instructions per launch (ncu count committed under profiles/)
    # / launch time / (SMs x 4 schedulers x SM clock under load).
    wi = measured_traffic("warp_instr_per_32_samples")
    if isinstance(wi, dict) and wi.get("k_ggx_dielectric") and clock_info.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issued = wi["k_ggx_dielectric"] * (n / 32.0) / (ms * 1e-3)
        roofline["issue"] = {"warp_instr_per_32_samples": wi["k_ggx_dielectric"],
                             "frac_of_issue_peak": issued / (sms * 4 * clock_info["sm_mhz"] * 1e6)}
    e2e = {"value": world * ne / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "ms_per_step": e2e_s * 1e3, "samples_per_gpu_per_step": ne,
           "host_cpus_bound_to_gpu_numa_node": numa}
    # e2e outputs must equal the device-resident outputs bit for bit
    same = all(torch.equal(hout[k], out[k][..., :ne].cpu()) for k in hout)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "samples_per_gpu": n,
                       "l2": "inputs+outputs per step = %.1f GB >> 126 MB L2, no flush needed" % (n * 113 / 1e9),
                       "sharding": "independent index ranges per rank, no collective"},
            "roofline": roofline, "e2e": e2e, "e2e_matches_device": bool(same), "arith": arith,
            "gpu_launches": int(gpu_launches), "clocks": clock_info}

    # ---- CPU baseline beside it (rank 0, N = 1): same input bits, bounded prefix
    if world == 1 and not args.no_cpu:
        import numpy as np
        import oracle_lib as ol
        import parity
        orc = ol.load_ref() or ol.load_port()
        orc.set_threads(0)
        nc = min(ne, 1 << 24)
        hs = {}
        for name, t in (("U", hsg.U), ("V", hsg.V), ("N", hsg.N), ("wo", hsg.wo)):
            for j, c in enumerate("xyz"):
                hs[name + c] = np.ascontiguousarray(t[j, :nc].numpy())
        hs["backfacing"] = np.ascontiguousarray(hsg.backfacing[:nc].numpy())
        cp = abi.ggx_params(specularRoughness=np.ascontiguousarray(hrough[:nc].numpy()),
                            ior=np.ascontiguousarray(hior[:nc].numpy()))
        crx, cry = np.ascontiguousarray(hrx[:nc].numpy()), np.ascontiguousarray(hry[:nc].numpy())
        orc.ggx_dielectric(hs, cp, crx[:1 << 16], cry[:1 << 16])      # warm-up
        best, cpu_out = None, None
        for _ in range(2):
            t0 = time.perf_counter()
            cpu_out = orc.ggx_dielectric(hs, cp, crx, cry)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        line["cpu_baseline"] = {"value": nc / best, "unit": UNIT, "cores": orc.max_threads(),
                                "kind": "reference" if orc.kind == "reference" else "port",
                                "sample": f"first {nc} samples of the GPU workload's own input bits, best of 2"}
        kinds = dict(fresnel="rel", wi_r="dir", f_r="rel", pdf_r="rel", wi_t="dir", f_t="rel", weight_t="rel",
                     flags="flags")
        gp = {k: v[..., :nc].cpu().numpy() for k, v in out.items()}
        st = parity.summarize(gp, cpu_out, kinds)
        line["parity_vs_oracle"] = {k: ({"flag_mismatches": v["mismatches"]} if "mismatches" in v else
                                        {"bit_exact": v["bit_exact"], "within_tol": v["within"]})
                                    for k, v in st.items()}

    # ---- the other BASELINE configs, device-timed (informational)
    if not args.main_only:
        del hsg, hsampler, hout, hrough, hior, hrx, hry
        others = {}
        steps_o, warm_o = max(3, args.steps // 2), 3

        # config 1: GGX conductor, 2^20 samples, uniform parameters (gold fixture)
        n1 = 1 << 20
        s1 = api.GgxSampler(ctx, api.ShadingBatch(sg.U[:, :n1].contiguous(), sg.V[:, :n1].contiguous(),
                                                  sg.N[:, :n1].contiguous(), sg.wo[:, :n1].contiguous()),
                            KsColor=(1.0, 1.0, 1.0), specularRoughness=0.3, ior=0.47)
        rx1, ry1 = rx[:n1].contiguous(), ry[:n1].contiguous()
        o1 = s1.alloc_out(rx1, want_fresnel=False)
        # 88 MB working set fits in L2: flush by writing a 512 MB buffer between iterations
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

        def step1():
            flush.zero_()
            s1.sampleEvalPdf(rx1, ry1, out=o1)
        t_flush = timed(lambda: flush.zero_(), steps_o, warm_o)
        ctx.fallback_count(reset=True)
        t1 = timed(step1, steps_o, warm_o) - t_flush
        fb1 = ctx.fallback_count(reset=True) / float(n1 * (steps_o + warm_o))
        others["ggx_conductor_1M"] = {"samples_per_s": world * n1 / (t1 * 1e-3), "ms": t1,
                                      "hbm_frac": n1 * B_ALG["ggx_conductor"] / (t1 * 1e-3) / 1e9 / peak,
                                      "exact_rerun_fraction": fb1,
                                      "l2": "flushed between iterations (512 MB memset, its time subtracted)"}
        # The same 2^20-sample launch replayed from a CUDA graph: R x (L2 flush + kernel) captured once
        # on a side stream through a context bound to that stream, minus a graph of the R flushes alone.
        try:
            R = 20
            gs = torch.cuda.Stream()
            gs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(gs):
                cg = api.Context(rank % torch.cuda.device_count())
                s1g = api.GgxSampler(cg, s1.sg, KsColor=(1.0, 1.0, 1.0), specularRoughness=0.3, ior=0.47)
                for _ in range(3):
                    flush.zero_()
                    s1g.sampleEvalPdf(rx1, ry1, out=o1)
            gs.synchronize()
            g_full, g_flush = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_flush, stream=gs):
                for _ in range(R):
                    flush.zero_()
            with torch.cuda.graph(g_full, stream=gs):
                for _ in range(R):
                    flush.zero_()
                    s1g.sampleEvalPdf(rx1, ry1, out=o1)
            tg = (timed(g_full.replay, 3, 2) - timed(g_flush.replay, 3, 2)) / R
            others["ggx_conductor_1M"]["cuda_graph"] = {"samples_per_s": world * n1 / (tg * 1e-3), "ms": tg,
                                                        "launches_per_replay": R}
            del g_full, g_flush, s1g
            cg.close()
        except Exception as exc:            # a capture problem must not cost the bench line
            others["ggx_conductor_1M"]["cuda_graph"] = {"unavailable": repr(exc)[:200]}
        del flush, s1, o1
        del sampler, out, rough, ior
        torch.cuda.empty_cache()

        # config 4: skin profile, 2^28 samples
        n4 = args.samples * 4
        color = tuple(ctx.synth_uniform(n4, 0x5EED0004, 40 + j, rank * n4, 0.05, 1.0) for j in range(3))
        distv = tuple(ctx.synth_uniform(n4, 0x5EED0004, 50 + j, rank * n4, 0.05, 2.0) for j in range(3))
        rx4 = ctx.synth_uniform(n4, 0x5EED0004, 0, rank * n4)
        s4 = api.SkinProfile(ctx, n4, sss_color=color, sss_scatter_dist=distv, sss_dist_multiplier=1.0)
        o4 = s4.alloc_out(rx4)
        t4 = timed(lambda: s4.sampleEvalPdf(rx4, out=o4), steps_o, warm_o)
        others["skin_profile_256M"] = {"samples_per_s": world * n4 / (t4 * 1e-3), "ms": t4,
                                       "exact_rerun_fraction": ctx.fallback_count(reset=True) / float(n4 * (steps_o + warm_o)),
                                       "hbm_frac": n4 * B_ALG["skin_profile"] / (t4 * 1e-3) / 1e9 / peak}
        del color, distv, rx4, s4, o4
        torch.cuda.empty_cache()

        # config 3: Disney, 2^28 samples, every parameter spatially varying
        n3 = args.samples * 4
        sg3 = ctx.synth_shading(n3, 0x5EED0003, rank * n3)
        names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
                 "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
        kw = {nm: ctx.synth_uniform(n3, 0x5EED0003, 20 + j, rank * n3) for j, nm in enumerate(names)}
        kw["base_color"] = tuple(ctx.synth_uniform(n3, 0x5EED0003, 30 + j, rank * n3) for j in range(3))
        u = [ctx.synth_uniform(n3, 0x5EED0003, s, rank * n3) for s in range(4)]
        s3 = api.DisneySampler(ctx, sg3, **kw)
        o3 = s3.alloc_out(u[0])
        t3 = timed(lambda: s3.sampleEvalPdf(*u, out=o3), steps_o, warm_o)
        others["disney_256M"] = {"samples_per_s": world * n3 / (t3 * 1e-3), "ms": t3,
                                 "exact_rerun_fraction": ctx.fallback_count(reset=True) / float(n3 * (steps_o + warm_o)),
                                 "hbm_frac": n3 * B_ALG["disney"] / (t3 * 1e-3) / 1e9 / peak}
        del sg3, kw, u, s3, o3
        torch.cuda.empty_cache()

        # config 5: albedo sweep 64 x 64 x 16 cells x 4096 spp, spp sharded over ranks, NCCL sum
        grid = abi.SweepGrid(64, 64, 16, 0.02, 1.0, 1.0, 2.5)
        spp = 4096
        k0, k1 = rank * spp // world, (rank + 1) * spp // world
        table = torch.empty(64 * 64 * 16, abi.SWEEP_VALUES_PER_CELL, dtype=torch.float64, device=dev)

        def step5():
            ctx.albedo_sweep(grid, 0x5EED0005, k0, k1, out=table)
            if world > 1:
                dist.all_reduce(table, op=dist.ReduceOp.SUM)
        t5 = timed(step5, max(2, steps_o // 2), 1)
        others["albedo_sweep_65536x4096"] = {"samples_per_s": 64 * 64 * 16 * spp / (t5 * 1e-3), "ms": t5,
                                             "exact_rerun_fraction": ctx.fallback_count(reset=True) /
                                             float(64 * 64 * 16 * (k1 - k0) * (max(2, steps_o // 2) + 1)),
                                             "collective": "nccl all_reduce(sum) of the 2.6 MB FP64 table" if world > 1 else "none (1 GPU)"}
        # FP32 roofline of each (algorithmic flops per sample; the sweep has no per-sample memory
        # traffic, so this is its governing roofline) -- per GPU, against the nominal FP32 peak
        for key, unit in (("ggx_conductor_1M", "ggx_conductor"), ("skin_profile_256M", "skin_profile"),
                          ("disney_256M", "disney"), ("albedo_sweep_65536x4096", "ggx_conductor")):
            others[key]["fp32_frac"] = others[key]["samples_per_s"] / world * F_ALG[unit] / 1e12 / FP32_PEAK_TFLOPS
        issued = measured_traffic("fp32_flops_per_sample_issued") or {}
        for key, kern in (("skin_profile_256M", "k_skin_profile"), ("disney_256M", "k_disney_sample_eval_pdf")):
            if isinstance(issued, dict) and issued.get(kern):
                others[key]["fp32_issued_tflops"] = others[key]["samples_per_s"] / world * issued[kern] / 1e12
        line["other_workloads"] = others

    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """Write the ONE JSON line to the process's real stdout (see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Libraries (NCCL's version banner, torchrun notices) print to stdout; the contract is ONE
    # JSON line there.  Route fd 1 to stderr for the duration and keep the real stdout for emit().
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--samples", type=int, default=N_DIELECTRIC, help="samples per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-samples", type=int, default=N_DIELECTRIC, help="samples per e2e step (per GPU)")
    ap.add_argument("--main-only", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--arith", choices=["fast", "exact", "tolerant"], default="fast",
                    help="arithmetic policy of the fused kernels: fast / exact = same bits (csrc/rls_fp.cuh); "
                         "tolerant = stated tolerance, bit-exact flags (csrc/rls_tol.cuh)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
