"""TEST INFRASTRUCTURE: drives tests/native/libtol_host*.so (the tolerance-policy units of csrc/rls_tol.cuh compiled
for the CPU) on the same inputs as the oracle and reports what RLS_ARITH_TOLERANT promises: flag mismatches among the
samples the band tracker keeps (must be 0), the re-run fraction, and error percentiles of every output.

    python tests/tol_host.py [n] [--ulp]      # prints the report for every workload
"""
import ctypes as C
import os
import subprocess
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (_ROOT, os.path.join(_ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

import oracle_lib as ol
from rlshaders_b200 import _abi as abi

NATIVE = os.path.join(_ROOT, "tests", "native")
f32 = np.float32


def load(ulp=False):
    path = os.path.join(NATIVE, "libtol_host_ulp.so" if ulp else "libtol_host.so")
    src = [os.path.join(NATIVE, "tol_host.cpp"), os.path.join(_ROOT, "rlshaders_b200", "csrc", "rls_tol.cuh")]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
        subprocess.run([os.path.join(NATIVE, "build_tol_host.sh")], check=True)
    return C.CDLL(path)


def _z(n, dtype=f32):
    return np.zeros(n, dtype=dtype)


def _v3(n):
    return (_z(n), _z(n), _z(n))


def ggx_dielectric(lib, sg, params, rx, ry):
    n = len(rx)
    F, wir, fr, pr, wit, ft, wt, fl = _z(n), _v3(n), _z(n), _z(n), _v3(n), _z(n), _z(n), _z(n, np.uint32)
    rerun = _z(n, np.uint8)
    out = abi.GgxDielectricOut(F.ctypes.data, abi.vec3(wir), fr.ctypes.data, pr.ctypes.data,
                               abi.vec3(wit), ft.ctypes.data, wt.ctypes.data, fl.ctypes.data)
    lib.tol_ggx_dielectric(C.c_size_t(n), C.byref(ol.shading_struct(sg)), C.byref(params),
                           C.c_void_p(rx.ctypes.data), C.c_void_p(ry.ctypes.data), C.byref(out),
                           C.c_void_p(rerun.ctypes.data))
    return dict(fresnel=F, wi_r=np.stack(wir), f_r=fr, pdf_r=pr, wi_t=np.stack(wit), f_t=ft, weight_t=wt, flags=fl), rerun


def ggx_conductor(lib, sg, params, rx, ry):
    n = len(rx)
    wi, f, pdf, F, fl = _v3(n), _v3(n), _z(n), _z(n), _z(n, np.uint32)
    rerun = _z(n, np.uint8)
    out = abi.BsdfOut(abi.vec3(wi), abi.vec3(f), pdf.ctypes.data, F.ctypes.data, fl.ctypes.data)
    lib.tol_ggx_sample_eval_pdf(C.c_size_t(n), C.byref(ol.shading_struct(sg)), C.byref(params),
                                C.c_void_p(rx.ctypes.data), C.c_void_p(ry.ctypes.data), C.byref(out),
                                C.c_void_p(rerun.ctypes.data))
    return dict(wi=np.stack(wi), f=np.stack(f), pdf=pdf, fresnel=F, flags=fl), rerun


def disney(lib, sg, params, rx_s, ry_s, rx_d, ry_d):
    n = len(rx_s)
    wis, fs, ps, wid, fd, pd, fl = _v3(n), _v3(n), _z(n), _v3(n), _v3(n), _z(n), _z(n, np.uint32)
    rerun = _z(n, np.uint8)
    out = abi.DisneyOut(abi.vec3(wis), abi.vec3(fs), ps.ctypes.data, abi.vec3(wid), abi.vec3(fd), pd.ctypes.data,
                        fl.ctypes.data)
    lib.tol_disney(C.c_size_t(n), C.byref(ol.shading_struct(sg)), C.byref(params), C.c_void_p(rx_s.ctypes.data),
                   C.c_void_p(ry_s.ctypes.data), C.c_void_p(rx_d.ctypes.data), C.c_void_p(ry_d.ctypes.data),
                   C.byref(out), C.c_void_p(rerun.ctypes.data))
    return dict(wi_s=np.stack(wis), f_s=np.stack(fs), pdf_s=ps, wi_d=np.stack(wid), f_d=np.stack(fd), pdf_d=pd,
                flags=fl), rerun


def skin_profile(lib, params, rx):
    n = len(rx)
    r, pdf, rd, fl = _z(n), _z(n), _v3(n), _z(n, np.uint32)
    rerun = _z(n, np.uint8)
    out = abi.ProfileOut(r.ctypes.data, pdf.ctypes.data, abi.vec3(rd), fl.ctypes.data)
    lib.tol_skin_profile(C.c_size_t(n), C.byref(params), C.c_void_p(rx.ctypes.data), C.byref(out),
                         C.c_void_p(rerun.ctypes.data))
    return dict(r=r, pdf=pdf, Rd=np.stack(rd), flags=fl), rerun


# ----------------------------------------------------------------- error statistics
TOLS = dict(dir=(1e-6, 1e-5, 1e-4, 1e-3), rel=(1e-5, 1e-4, 1e-3, 1e-2))


def errors(kind, g, c):
    """Per-sample error: directions = max abs component difference; values = relative (absolute where the oracle is 0)."""
    g64, c64 = np.asarray(g, np.float64), np.asarray(c, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "dir":
            e = np.abs(g64 - c64)
        else:
            e = np.abs(g64 - c64) / np.maximum(np.abs(c64), 1e-30)
    e = np.where(np.asarray(g) == np.asarray(c), 0.0, e)
    e = np.where(np.isnan(g64) & np.isnan(c64), 0.0, e)
    e = np.where(np.isnan(e), np.inf, e)
    return e.max(axis=0) if e.ndim == 2 else e


def compare(tolr, rerun, orc, kinds):
    """The samples the band tracker flags get the bit-exact result on the device: substitute the oracle's there."""
    keep = rerun == 0
    out = {"rerun_fraction": float(1.0 - keep.mean()), "n": int(keep.size)}
    for name, kind in kinds.items():
        if kind == "flags":
            bad = (tolr[name] != orc[name]) & keep
            out[name] = {"mismatches": int(bad.sum()), "mismatches_without_rerun": int((tolr[name] != orc[name]).sum())}
            out["_bad_index"] = np.nonzero(bad)[0][:8].tolist()
        else:
            e = np.where(keep, errors(kind, tolr[name], orc[name]), 0.0)
            out[name] = {"within": {t: float((e <= t).mean()) for t in TOLS[kind]}, "max": float(e.max()),
                         "p50": float(np.median(e))}
    return out


def report(title, st):
    lines = [f"{title}: n = {st['n']}, re-run fraction {st['rerun_fraction']:.3e}"]
    for name, s in st.items():
        if not isinstance(s, dict):
            continue
        if "mismatches" in s:
            lines.append(f"  {name:9s} flag mismatches {s['mismatches']}  (without the re-run: {s['mismatches_without_rerun']})")
        else:
            w = "  ".join(f"<={t:g}: {v * 100:8.4f}%" for t, v in s["within"].items())
            lines.append(f"  {name:9s} {w}  median {s['p50']:.1e}  max {s['max']:.2e}")
    return "\n".join(lines)


KINDS_DIELECTRIC = dict(fresnel="rel", wi_r="dir", f_r="rel", pdf_r="rel", wi_t="dir", f_t="rel", weight_t="rel", flags="flags")
KINDS_GGX = dict(wi="dir", f="rel", pdf="rel", fresnel="rel", flags="flags")
KINDS_DISNEY = dict(wi_s="dir", f_s="rel", pdf_s="rel", wi_d="dir", f_d="rel", pdf_d="rel", flags="flags")
KINDS_SKIN = dict(r="rel", pdf="rel", Rd="rel", flags="flags")


def run_dielectric(lib, orc, n, seed=0x5EED0002, aniso=False):
    import parity
    sg, kw, rx, ry = parity.ggx_dielectric_inputs(n, seed, aniso)
    p = abi.ggx_params(**kw)
    t, rerun = ggx_dielectric(lib, sg, p, rx, ry)
    return compare(t, rerun, orc.ggx_dielectric(sg, p, rx, ry), KINDS_DIELECTRIC)


def run_conductor(lib, orc, n, seed=0x5EED0001):
    sg, p, rx, ry = ol.workload_ggx_conductor(n, seed)
    t, rerun = ggx_conductor(lib, sg, p, rx, ry)
    return compare(t, rerun, orc.ggx_sample_eval_pdf(sg, p, rx, ry), KINDS_GGX)


def run_disney(lib, orc, n, seed=0x5EED0003):
    import parity
    sg, kw, u = parity.disney_inputs(n, seed)
    p = abi.disney_params(**kw)
    t, rerun = disney(lib, sg, p, *u)
    return compare(t, rerun, orc.disney_sample_eval_pdf(sg, p, *u), KINDS_DISNEY)


def run_skin(lib, orc, n, seed=0x5EED0004):
    import parity
    kw, rx = parity.skin_inputs(n, seed)
    p = abi.skin_params(**kw)
    t, rerun = skin_profile(lib, p, rx)
    return compare(t, rerun, orc.skin_profile(p, rx), KINDS_SKIN)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 1 << 20
    lib = load(ulp="--ulp" in sys.argv)
    orc = ol.load_ref() or ol.load_port()
    orc.set_threads(0)
    which = [a for a in sys.argv[1:] if a in ("dielectric", "aniso", "conductor", "disney", "skin")] or \
            ["dielectric", "aniso", "conductor", "disney", "skin"]
    for w in which:
        fn = dict(dielectric=lambda: run_dielectric(lib, orc, n), aniso=lambda: run_dielectric(lib, orc, n, aniso=True),
                  conductor=lambda: run_conductor(lib, orc, n), disney=lambda: run_disney(lib, orc, n),
                  skin=lambda: run_skin(lib, orc, n))[w]
        if not hasattr(lib, dict(dielectric="tol_ggx_dielectric", aniso="tol_ggx_dielectric", conductor="tol_ggx_sample_eval_pdf",
                                 disney="tol_disney", skin="tol_skin_profile")[w]):
            continue
        st = fn()
        print(report(w, st))
        if st.get("_bad_index"):
            print("   first flag mismatches at", st["_bad_index"])
