"""Parity harness: run one workload through the CUDA library (via the C ABI) and through
a CPU oracle on identical input bits, and summarise the agreement.

TEST INFRASTRUCTURE.  Tolerances (BASELINE.json north_star): directions 1e-6 absolute,
f / pdf / radii 1e-5 relative, flags and lobe choices bit-exact.
"""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
for _p in (_ROOT, _os.path.join(_ROOT, "tests")):
    if _p not in _sys.path:
        _sys.path.insert(0, _p)

import numpy as np
import torch

import oracle_lib as ol
from rlshaders_b200 import _abi as abi
from rlshaders_b200 import api

DIR_TOL = 1e-6
REL_TOL = 1e-5


def to_dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def params_to_dev(kw, device):
    """numpy arrays in a node-parameter dict -> device tensors (scalars stay uniform)."""
    out = {}
    for k, v in kw.items():
        if isinstance(v, np.ndarray):
            out[k] = to_dev(v, device)
        elif isinstance(v, tuple) and isinstance(v[0], np.ndarray):
            out[k] = tuple(to_dev(t, device) for t in v)
        else:
            out[k] = v
    return out


def stat_dir(g, c):
    """Vector outputs [3, n]: absolute error per sample (max over components)."""
    g, c = np.asarray(g), np.asarray(c)
    err = np.max(np.abs(g.astype(np.float64) - c.astype(np.float64)), axis=0)
    both_nan = np.all(np.isnan(g) == np.isnan(c), axis=0)
    err = np.where(np.isnan(err), np.where(both_nan, 0.0, np.inf), err)
    exact = np.all((g.view(np.uint32) == c.view(np.uint32)) | (np.isnan(g) & np.isnan(c)), axis=0)
    return dict(n=err.size, bit_exact=float(exact.mean()), within=float((err <= DIR_TOL).mean()),
                within_1e4=float((err <= 1e-4).mean()), max=float(np.max(err)))


def stat_rel(g, c):
    """Scalar/colour outputs: relative error per element (absolute where the oracle is 0)."""
    g, c = np.asarray(g, dtype=np.float32), np.asarray(c, dtype=np.float32)
    g64, c64 = g.astype(np.float64), c.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        err = np.abs(g64 - c64) / np.maximum(np.abs(c64), 1e-30)
    same = (g.view(np.uint32) == c.view(np.uint32)) | (np.isnan(g) & np.isnan(c)) | (g == c)
    err = np.where(same, 0.0, err)
    err = np.where(np.isnan(err), np.inf, err)
    if err.ndim == 2:
        err, same = err.max(axis=0), same.all(axis=0)
    return dict(n=err.size, bit_exact=float(same.mean()), within=float((err <= REL_TOL).mean()),
                within_1e3=float((err <= 1e-3).mean()), max=float(np.max(err)))


def stat_flags(g, c):
    g, c = np.asarray(g).astype(np.uint32), np.asarray(c).astype(np.uint32)
    return dict(n=g.size, mismatches=int((g != c).sum()))


def summarize(gpu, cpu, kinds):
    """kinds: name -> 'dir' | 'rel' | 'flags'."""
    out = {}
    for name, kind in kinds.items():
        gv = gpu[name].cpu().numpy() if hasattr(gpu[name], "cpu") else gpu[name]
        fn = dict(dir=stat_dir, rel=stat_rel, flags=stat_flags)[kind]
        out[name] = fn(gv, cpu[name])
    return out


# ----------------------------------------------------------------- workloads
def run_ggx_conductor(ctx, oracle, n, seed=0x5EED0001):
    sg, p, rx, ry = ol.workload_ggx_conductor(n, seed)
    cpu = oracle.ggx_sample_eval_pdf(sg, p, rx, ry)
    dsg = api.ShadingBatch.from_numpy(sg, ctx.device)
    s = api.GgxSampler(ctx, dsg, KsColor=(1.0, 1.0, 1.0), specularRoughness=0.3, ior=0.47, anisotropic=0.0)
    gpu = s.sampleEvalPdf(to_dev(rx, ctx.device), to_dev(ry, ctx.device))
    ctx.synchronize()
    kinds = dict(wi="dir", f="rel", pdf="rel", fresnel="rel", flags="flags")
    return summarize(gpu, cpu, kinds), gpu, cpu, (sg, p, s)


def ggx_dielectric_inputs(n, seed=0x5EED0002, aniso=False):
    sg = ol.make_shading(n, seed, backfacing_fraction=0.25)
    kw = dict(specularRoughness=ol.hash_uniform(n, seed, 2, lo=0.05, hi=1.0),
              ior=ol.hash_uniform(n, seed, 3, lo=1.05, hi=2.5))
    if aniso:
        kw["anisotropic"] = ol.hash_uniform(n, seed, 4)
    return sg, kw, ol.hash_uniform(n, seed, 0), ol.hash_uniform(n, seed, 1)


def run_ggx_dielectric(ctx, oracle, n, seed=0x5EED0002, aniso=False):
    sg, kw, rx, ry = ggx_dielectric_inputs(n, seed, aniso)
    cpu = oracle.ggx_dielectric(sg, abi.ggx_params(**kw), rx, ry)
    dsg = api.ShadingBatch.from_numpy(sg, ctx.device)
    s = api.GgxSampler(ctx, dsg, **params_to_dev(kw, ctx.device))
    gpu = s.dielectricSampleEvalPdf(to_dev(rx, ctx.device), to_dev(ry, ctx.device))
    ctx.synchronize()
    kinds = dict(fresnel="rel", wi_r="dir", f_r="rel", pdf_r="rel", wi_t="dir", f_t="rel", weight_t="rel",
                 flags="flags")
    return summarize(gpu, cpu, kinds), gpu, cpu, (sg, kw, s)


def disney_inputs(n, seed=0x5EED0003):
    sg = ol.make_shading(n, seed)
    u = [ol.hash_uniform(n, seed, s) for s in range(4)]
    names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
             "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
    kw = {nm: ol.hash_uniform(n, seed, 20 + j) for j, nm in enumerate(names)}
    kw["base_color"] = tuple(ol.hash_uniform(n, seed, 30 + j) for j in range(3))
    return sg, kw, u


def run_disney(ctx, oracle, n, seed=0x5EED0003):
    sg, kw, u = disney_inputs(n, seed)
    cpu = oracle.disney_sample_eval_pdf(sg, abi.disney_params(**kw), *u)
    dsg = api.ShadingBatch.from_numpy(sg, ctx.device)
    s = api.DisneySampler(ctx, dsg, **params_to_dev(kw, ctx.device))
    gpu = s.sampleEvalPdf(*[to_dev(t, ctx.device) for t in u])
    ctx.synchronize()
    kinds = dict(wi_s="dir", f_s="rel", pdf_s="rel", wi_d="dir", f_d="rel", pdf_d="rel", flags="flags")
    return summarize(gpu, cpu, kinds), gpu, cpu, (sg, kw, s)


def skin_inputs(n, seed=0x5EED0004):
    kw = dict(sss_color=tuple(ol.hash_uniform(n, seed, 40 + j, lo=0.05, hi=1.0) for j in range(3)),
              sss_scatter_dist=tuple(ol.hash_uniform(n, seed, 50 + j, lo=0.05, hi=2.0) for j in range(3)),
              sss_dist_multiplier=1.0)
    return kw, ol.hash_uniform(n, seed, 0)


def run_skin(ctx, oracle, n, seed=0x5EED0004):
    kw, rx = skin_inputs(n, seed)
    cpu = oracle.skin_profile(abi.skin_params(**kw), rx)
    s = api.SkinProfile(ctx, n, **params_to_dev(kw, ctx.device))
    gpu = s.sampleEvalPdf(to_dev(rx, ctx.device))
    ctx.synchronize()
    kinds = dict(r="rel", pdf="rel", Rd="rel", flags="flags")
    return summarize(gpu, cpu, kinds), gpu, cpu, (kw, s)


_HOST_LIBM = None


def host_libm_status():
    """RLS_HOST_LIBM_CHECK: does THIS host's libm return the bits the device port reproduces (glibc 2.39, x86-64 FMA
    builds; rlshaders_b200/csrc/rls_libm.cuh)?  A few hundred thousand known-answer arguments per transcendental, the
    device port's own source compiled for the host against the host libm (tests/native/libm_check).  Returns
    ("ok", "") or ("degraded", "<function>: k mismatches ..."): on a degraded host the GPU parity tests fall back from
    "bit-identical" to their fractional tolerances and say so, instead of failing on a libm they were not written for.
    RLS_HOST_LIBM_CHECK=0 skips the probe (assumes ok)."""
    global _HOST_LIBM
    if _HOST_LIBM is not None:
        return _HOST_LIBM
    import subprocess
    if _os.environ.get("RLS_HOST_LIBM_CHECK", "1") == "0":
        _HOST_LIBM = ("ok", "probe skipped (RLS_HOST_LIBM_CHECK=0)")
        return _HOST_LIBM
    src = _os.path.join(_ROOT, "tests", "native", "libm_check.cpp")
    exe = _os.path.join(_ROOT, "tests", "native", "libm_check")
    deps = [src, _os.path.join(_ROOT, "rlshaders_b200", "csrc", "rls_libm.cuh")]
    try:
        if not _os.path.exists(exe) or any(_os.path.getmtime(d) > _os.path.getmtime(exe) for d in deps):
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-o", exe, src, "-lm"],
                           check=True)
        bad = []
        for fn in ("sincos", "tan", "atan", "acos", "exp", "log", "atan2", "pow"):
            out = subprocess.run([exe, fn, "16411"], check=True, capture_output=True, text=True).stdout.split()
            if int(out[4]) != 0:
                bad.append(f"{fn}: {out[4]} of {out[2]} arguments differ")
        _HOST_LIBM = ("ok", "") if not bad else ("degraded", "; ".join(bad))
    except Exception as e:        # noqa: BLE001
        _HOST_LIBM = ("degraded", f"probe could not run: {e}")
    return _HOST_LIBM


def format_report(title, stats):
    lines = [title]
    for name, s in stats.items():
        if "mismatches" in s:
            lines.append(f"  {name:10s} flag mismatches {s['mismatches']} / {s['n']}")
        else:
            extra = "within_1e4" if "within_1e4" in s else "within_1e3"
            lines.append(f"  {name:10s} bit-exact {s['bit_exact'] * 100:9.5f}%  within tol {s['within'] * 100:9.5f}%  "
                         f"{extra} {s[extra] * 100:9.5f}%  max {s['max']:.3e}")
    return "\n".join(lines)


if __name__ == "__main__":
    import sys
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    ctx = api.Context(0)
    for orc in (ol.load_ref(), ol.load_port()):
        if orc is None:
            continue
        print(f"=== oracle kind: {orc.kind}, n = {n}")
        print(format_report("GGX conductor (config 1)", run_ggx_conductor(ctx, orc, n)[0]))
        print(format_report("GGX dielectric (config 2)", run_ggx_dielectric(ctx, orc, n)[0]))
        print(format_report("GGX dielectric, anisotropic", run_ggx_dielectric(ctx, orc, n, aniso=True)[0]))
        print(format_report("Disney (config 3)", run_disney(ctx, orc, n)[0]))
        print(format_report("Skin profile (config 4)", run_skin(ctx, orc, n)[0]))


# ------------------------------------------- callers of the triple
def skin_layers_inputs(n, k, seed=0x5EED00F2, with_radiance=True):
    """P shading points x K samples (sample-major), every rlSkin layer parameter varying; a few
    points have a zero / tiny layer weight or a black layer colour (the node's skip branches)."""
    sg = ol.make_shading(n, seed)
    u = lambda s, lo=0.0, hi=1.0: ol.hash_uniform(n, seed, s, lo=lo, hi=hi)     # noqa: E731
    sel = ol.hash_uniform(n, seed, 99)
    kw = dict(sheen_color=tuple(u(20 + j) for j in range(3)), sheen_weight=np.where(sel < 0.1, 0.0, u(23)).astype(np.float32),
              sheen_roughness=u(24, 0.05, 1.0), sheen_ior=u(25, 1.05, 2.0),
              specular_color=tuple(np.where((sel > 0.9) & (sel < 0.95), 0.0, u(26 + j)).astype(np.float32) for j in range(3)),
              specular_weight=np.where(sel > 0.97, 5e-5, u(29)).astype(np.float32),
              specular_roughness=u(30, 0.05, 1.0), specular_ior=u(31, 1.05, 2.0), sss_weight=u(32))
    uu = [ol.hash_uniform(n * k, seed, 40 + j) for j in range(4)]
    li = [np.stack([ol.hash_uniform(n * k, seed, 50 + 3 * j + c, lo=0.0, hi=4.0) for c in range(3)]) for j in range(2)] \
        if with_radiance else [None, None]
    return sg, kw, uu, li


def run_skin_layers(ctx, oracle, n, k, seed=0x5EED00F2, with_radiance=True):
    sg, kw, uu, li = skin_layers_inputs(n, k, seed, with_radiance)
    cpu = oracle.skin_glossy_layers(sg, abi.skin_params(**kw), k, *uu, li[0], li[1])
    dsg = api.ShadingBatch.from_numpy(sg, ctx.device)
    s = api.SkinProfile(ctx, n, **params_to_dev(kw, ctx.device))
    dli = [to_dev(t, ctx.device) if t is not None else None for t in li]
    gpu = s.glossyLayers(dsg, k, *[to_dev(t, ctx.device) for t in uu], dli[0], dli[1])
    ctx.synchronize()
    kinds = dict(sheen="rel", specular="rel", sheen_fresnel="rel", specular_fresnel="rel", sss_weight="rel", flags="flags")
    return summarize(gpu, cpu, kinds), gpu, cpu


def light_inputs(n, sg, seed):
    """A light sample per shading point: direction in the upper hemisphere of the frame (a few
    below it and a few zero vectors), radiance, pdf (some zero); and the same light evaluated
    along the BRDF-sampled direction."""
    N = np.stack([sg["N" + c] for c in "xyz"]); U = np.stack([sg["U" + c] for c in "xyz"]); V = np.stack([sg["V" + c] for c in "xyz"])
    cz = ol.hash_uniform(n, seed, 70, lo=-0.2, hi=1.0)
    ph = ol.hash_uniform(n, seed, 71, lo=0.0, hi=2.0 * np.pi)
    sr = np.sqrt(np.maximum(0.0, 1.0 - cz.astype(np.float64) ** 2))
    Ld = U * (sr * np.cos(ph)) + V * (sr * np.sin(ph)) + N * cz
    Ld = (Ld / np.linalg.norm(Ld, axis=0)).astype(np.float32)
    sel = ol.hash_uniform(n, seed, 72)
    Ld[:, sel < 0.02] = 0.0
    Li = np.stack([ol.hash_uniform(n, seed, 73 + c, lo=0.0, hi=10.0) for c in range(3)])
    pl = np.where(sel > 0.97, 0.0, ol.hash_uniform(n, seed, 76, lo=0.01, hi=5.0)).astype(np.float32)
    Lib = np.stack([ol.hash_uniform(n, seed, 77 + c, lo=0.0, hi=10.0) for c in range(3)])
    plb = np.where((sel > 0.5) & (sel < 0.55), 0.0, ol.hash_uniform(n, seed, 80, lo=0.01, hi=5.0)).astype(np.float32)
    return np.ascontiguousarray(Ld), Li, pl, Lib, plb


MIS_KINDS = dict(rgb="rel", w_light="rel", w_brdf="rel")
