"""Test-side loader for the CPU oracles (oracle/oracle_api.h) and numpy workload builders.

TEST INFRASTRUCTURE: the product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
for _p in (_ROOT, _os.path.join(_ROOT, "tests")):
    if _p not in _sys.path:
        _sys.path.insert(0, _p)

import numpy as np

from rlshaders_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(ROOT, "oracle", "librls_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "librls_ref.so")
if os.environ.get("RLS_ORACLE_DIR"):        # tools/sanitize.sh host: ASan / UBSan builds of the same two libraries
    PORT_SO = os.path.join(os.environ["RLS_ORACLE_DIR"], "librls_oracle.so")
    REF_SO = os.path.join(os.environ["RLS_ORACLE_DIR"], "librls_ref.so")
F64_SO = os.path.join(ROOT, "oracle", "librls_oracle_f64.so")
REFERENCE_SRC = "/root/reference/src"

f32 = np.float32


def build_oracles():
    """(Re)build what can be built here: always the port; the reference library only
    where /root/reference exists (the build container)."""
    target = ["port"]
    if os.path.exists(os.path.join(REFERENCE_SRC, "rlGgx.h")):
        target.append("ref")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")] + target, check=True)


# ----------------------------------------------------------------- numpy hash
_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xD1B54A32D192ED03)
_M3 = np.uint64(0xBF58476D1CE4E5B9)
_M4 = np.uint64(0x94D049BB133111EB)


def hash_uniform(n, seed, stream, first_index=0, lo=0.0, hi=1.0):
    """numpy restatement of the synthetic generators' integer hash (oracle_common.h,
    rlshaders_b200/csrc/rls_synth.cuh): bit-identical uniforms on host and device."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(first_index)
        z = np.uint64(seed) + _M1 * (idx + np.uint64(1)) + _M2 * np.uint64(stream)
        z = (z ^ (z >> np.uint64(30))) * _M3
        z = (z ^ (z >> np.uint64(27))) * _M4
        z = z ^ (z >> np.uint64(31))
    k = (z >> np.uint64(40)).astype(np.uint32)
    k = np.maximum(k, np.uint32(1))
    u = k.astype(f32) * f32(2.0 ** -24)
    return (f32(lo) + (f32(hi) - f32(lo)) * u).astype(f32)


def make_shading(n, seed, cos_lo=0.02, cos_hi=1.0, backfacing_fraction=0.0):
    """Random orthonormal frames + view vectors (host construction; the device
    generator rls_synth_shading follows the same recipe but its sin/cos bits differ, so
    parity tests always move ONE side's arrays to the other)."""
    u1 = hash_uniform(n, seed, 10)
    u2 = hash_uniform(n, seed, 11)
    u3 = hash_uniform(n, seed, 12)
    u4 = hash_uniform(n, seed, 13)
    u5 = hash_uniform(n, seed, 14)
    u6 = hash_uniform(n, seed, 15)
    nz = (f32(1.0) - f32(2.0) * u1).astype(f32)
    rn = np.sqrt(np.maximum(f32(0.0), f32(1.0) - nz * nz)).astype(f32)
    phn = (f32(2.0 * np.pi) * u2).astype(f32)
    N = np.stack([rn * np.cos(phn), rn * np.sin(phn), nz]).astype(f32)
    # tangent: any vector not parallel to N, Gram-Schmidt, then rotate by a random angle
    a = np.where(np.abs(N[0]) < 0.9, 1.0, 0.0).astype(f32)
    A = np.stack([a, f32(1.0) - a, np.zeros(n, f32)]).astype(f32)
    T = A - N * np.sum(A * N, axis=0, dtype=f32)
    T = (T / np.sqrt(np.sum(T * T, axis=0, dtype=f32))).astype(f32)
    B = np.cross(N.T, T.T).T.astype(f32)
    pht = (f32(2.0 * np.pi) * u3).astype(f32)
    U = (T * np.cos(pht) + B * np.sin(pht)).astype(f32)
    U = (U / np.sqrt(np.sum(U * U, axis=0, dtype=f32))).astype(f32)
    V = np.cross(N.T, U.T).T.astype(f32)
    cz = (f32(cos_lo) + f32(cos_hi - cos_lo) * u4).astype(f32)
    sr = np.sqrt(np.maximum(f32(0.0), f32(1.0) - cz * cz)).astype(f32)
    phv = (f32(2.0 * np.pi) * u5).astype(f32)
    wo = (U * (sr * np.cos(phv)) + V * (sr * np.sin(phv)) + N * cz).astype(f32)
    wo = (wo / np.sqrt(np.sum(wo * wo, axis=0, dtype=f32))).astype(f32)
    sg = {}
    for name, M in (("U", U), ("V", V), ("N", N), ("wo", wo)):
        for j, c in enumerate("xyz"):
            sg[name + c] = np.ascontiguousarray(M[j], dtype=f32)
    sg["backfacing"] = (u6 < f32(backfacing_fraction)).astype(np.uint8) if backfacing_fraction > 0 else None
    return sg


def shading_struct(sg):
    return abi.shading((sg["Ux"], sg["Uy"], sg["Uz"]), (sg["Vx"], sg["Vy"], sg["Vz"]),
                       (sg["Nx"], sg["Ny"], sg["Nz"]), (sg["wox"], sg["woy"], sg["woz"]),
                       sg.get("backfacing"))


def _z(n, dtype=f32):
    return np.zeros(n, dtype=dtype)


def _v3(n):
    return (_z(n), _z(n), _z(n))


def _cv3(t):
    return abi.vec3(tuple(np.ascontiguousarray(a, dtype=f32) for a in t))


class Oracle:
    """ctypes front-end over one oracle library; every method returns numpy arrays."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = C.CDLL(path)
        self.lib.oracle_kind.restype = C.c_char_p
        self.lib.oracle_max_threads.restype = C.c_int
        self.kind = self.lib.oracle_kind().decode()

    def set_threads(self, n):
        self.lib.oracle_set_threads(C.c_int(n))

    def max_threads(self):
        return self.lib.oracle_max_threads()

    def set_flag_probe(self, on):
        """Timed runs switch the RLS_FLAG_SLOPE_EARLY_OUT probe of the reference library off (oracle_api.h)."""
        self.lib.oracle_set_flag_probe(C.c_int(1 if on else 0))

    # ---- rlGgx
    def ggx_eval_sample(self, sg, params, rx, ry):
        n = len(rx)
        wi, F = _v3(n), _z(n)
        self.lib.oracle_ggx_eval_sample(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                        C.c_void_p(rx.ctypes.data), C.c_void_p(ry.ctypes.data),
                                        abi.vec3(wi), C.c_void_p(F.ctypes.data))
        return dict(wi=np.stack(wi), fresnel=F)

    def ggx_eval_brdf(self, sg, params, wi):
        n = wi.shape[1]
        f = _v3(n)
        keep = [np.ascontiguousarray(wi[j], dtype=f32) for j in range(3)]
        self.lib.oracle_ggx_eval_brdf(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                      abi.vec3(keep), abi.vec3(f))
        return np.stack(f)

    def ggx_eval_pdf(self, sg, params, wi):
        n = wi.shape[1]
        pdf = _z(n)
        keep = [np.ascontiguousarray(wi[j], dtype=f32) for j in range(3)]
        self.lib.oracle_ggx_eval_pdf(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                     abi.vec3(keep), C.c_void_p(pdf.ctypes.data))
        return pdf

    def ggx_sample_eval_pdf(self, sg, params, rx, ry):
        n = len(rx)
        wi, f, pdf, F, fl = _v3(n), _v3(n), _z(n), _z(n), _z(n, np.uint32)
        out = abi.BsdfOut(abi.vec3(wi), abi.vec3(f), pdf.ctypes.data, F.ctypes.data, fl.ctypes.data)
        self.lib.oracle_ggx_sample_eval_pdf(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                            C.c_void_p(rx.ctypes.data), C.c_void_p(ry.ctypes.data),
                                            C.byref(out))
        return dict(wi=np.stack(wi), f=np.stack(f), pdf=pdf, fresnel=F, flags=fl)

    def ggx_dielectric(self, sg, params, rx, ry):
        n = len(rx)
        F, wir, fr, pr, wit, ft, wt, fl = _z(n), _v3(n), _z(n), _z(n), _v3(n), _z(n), _z(n), _z(n, np.uint32)
        out = abi.GgxDielectricOut(F.ctypes.data, abi.vec3(wir), fr.ctypes.data, pr.ctypes.data,
                                   abi.vec3(wit), ft.ctypes.data, wt.ctypes.data, fl.ctypes.data)
        self.lib.oracle_ggx_dielectric_sample_eval_pdf(
            C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
            C.c_void_p(rx.ctypes.data), C.c_void_p(ry.ctypes.data), C.byref(out))
        return dict(fresnel=F, wi_r=np.stack(wir), f_r=fr, pdf_r=pr, wi_t=np.stack(wit), f_t=ft,
                    weight_t=wt, flags=fl)

    def ggx_refract_direction(self, sg, params, m):
        n = m.shape[1]
        wi, fl = _v3(n), _z(n, np.uint32)
        keep = [np.ascontiguousarray(m[j], dtype=f32) for j in range(3)]
        self.lib.oracle_ggx_refract_direction(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                              abi.vec3(keep), abi.vec3(wi), C.c_void_p(fl.ctypes.data))
        return np.stack(wi), fl

    def ggx_eval_btdf(self, sg, params, wi):
        n = wi.shape[1]
        ft = _z(n)
        keep = [np.ascontiguousarray(wi[j], dtype=f32) for j in range(3)]
        self.lib.oracle_ggx_eval_btdf(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params), abi.vec3(keep),
                                      C.c_void_p(ft.ctypes.data))
        return ft

    def ggx_sample_weight(self, sg, params, wi, m):
        n = wi.shape[1]
        w = _z(n)
        k1 = [np.ascontiguousarray(wi[j], dtype=f32) for j in range(3)]
        k2 = [np.ascontiguousarray(m[j], dtype=f32) for j in range(3)]
        self.lib.oracle_ggx_sample_weight(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params), abi.vec3(k1),
                                          abi.vec3(k2), C.c_void_p(w.ctypes.data))
        return w

    def prepare_ggx_dielectric(self, sg, params, rx, ry):
        """For timing: outputs allocated AND touched once, ctypes arguments built once.  Returns (call, outputs) where
        call() is nothing but the library's oracle_ggx_dielectric_sample_eval_pdf on those buffers."""
        n = len(rx)
        F, wir, fr, pr, wit, ft, wt, fl = _z(n), _v3(n), _z(n), _z(n), _v3(n), _z(n), _z(n), _z(n, np.uint32)
        out = abi.GgxDielectricOut(F.ctypes.data, abi.vec3(wir), fr.ctypes.data, pr.ctypes.data,
                                   abi.vec3(wit), ft.ctypes.data, wt.ctypes.data, fl.ctypes.data)
        for a in (F, fr, pr, ft, wt, fl) + wir + wit:
            a.fill(0)
        args = (C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params), C.c_void_p(rx.ctypes.data),
                C.c_void_p(ry.ctypes.data), C.byref(out))
        keep = (sg, params, rx, ry, out)
        fn = self.lib.oracle_ggx_dielectric_sample_eval_pdf

        def call():
            fn(*args)
            return keep is None
        outs = dict(fresnel=F, wi_r=wir, f_r=fr, pdf_r=pr, wi_t=wit, f_t=ft, weight_t=wt, flags=fl)
        return call, outs

    def prepare_disney(self, sg, params, rx_s, ry_s, rx_d, ry_d):
        """As prepare_ggx_dielectric, for oracle_disney_sample_eval_pdf."""
        n = len(rx_s)
        wis, fs, ps, wid, fd, pd, fl = _v3(n), _v3(n), _z(n), _v3(n), _v3(n), _z(n), _z(n, np.uint32)
        out = abi.DisneyOut(abi.vec3(wis), abi.vec3(fs), ps.ctypes.data, abi.vec3(wid), abi.vec3(fd),
                            pd.ctypes.data, fl.ctypes.data)
        for a in (ps, pd, fl) + wis + fs + wid + fd:
            a.fill(0)
        args = (C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params), C.c_void_p(rx_s.ctypes.data),
                C.c_void_p(ry_s.ctypes.data), C.c_void_p(rx_d.ctypes.data), C.c_void_p(ry_d.ctypes.data), C.byref(out))
        keep = (sg, params, rx_s, ry_s, rx_d, ry_d, out)
        fn = self.lib.oracle_disney_sample_eval_pdf

        def call():
            fn(*args)
            return keep is None
        return call, dict(wi_s=wis, f_s=fs, pdf_s=ps, wi_d=wid, f_d=fd, pdf_d=pd, flags=fl)

    # ---- rlDisney
    def disney_eval_sample(self, sg, params, sample_type, rx, ry):
        n = len(rx)
        wi, fl = _v3(n), _z(n, np.uint32)
        self.lib.oracle_disney_eval_sample(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                           C.c_int(sample_type), C.c_void_p(rx.ctypes.data),
                                           C.c_void_p(ry.ctypes.data), abi.vec3(wi),
                                           C.c_void_p(fl.ctypes.data))
        return dict(wi=np.stack(wi), flags=fl)

    def disney_eval_brdf(self, sg, params, sample_type, wi):
        n = wi.shape[1]
        f = _v3(n)
        keep = [np.ascontiguousarray(wi[j], dtype=f32) for j in range(3)]
        self.lib.oracle_disney_eval_brdf(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                         C.c_int(sample_type), abi.vec3(keep), abi.vec3(f))
        return np.stack(f)

    def disney_eval_pdf(self, sg, params, sample_type, wi):
        n = wi.shape[1]
        pdf = _z(n)
        keep = [np.ascontiguousarray(wi[j], dtype=f32) for j in range(3)]
        self.lib.oracle_disney_eval_pdf(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                        C.c_int(sample_type), abi.vec3(keep), C.c_void_p(pdf.ctypes.data))
        return pdf

    def disney_sample_eval_pdf(self, sg, params, rx_s, ry_s, rx_d, ry_d):
        n = len(rx_s)
        wis, fs, ps, wid, fd, pd, fl = _v3(n), _v3(n), _z(n), _v3(n), _v3(n), _z(n), _z(n, np.uint32)
        out = abi.DisneyOut(abi.vec3(wis), abi.vec3(fs), ps.ctypes.data, abi.vec3(wid), abi.vec3(fd),
                            pd.ctypes.data, fl.ctypes.data)
        self.lib.oracle_disney_sample_eval_pdf(
            C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
            C.c_void_p(rx_s.ctypes.data), C.c_void_p(ry_s.ctypes.data),
            C.c_void_p(rx_d.ctypes.data), C.c_void_p(ry_d.ctypes.data), C.byref(out))
        return dict(wi_s=np.stack(wis), f_s=np.stack(fs), pdf_s=ps, wi_d=np.stack(wid),
                    f_d=np.stack(fd), pdf_d=pd, flags=fl)

    # ---- NDProfile / rlSkin
    @staticmethod
    def _profile_struct(prof):
        return abi.NdProfileSoA(_cv3(prof["distance"]), _cv3(prof["C1"]), _cv3(prof["C2"]),
                                prof["max_radius"].ctypes.data)

    def ndprofile_set_distance(self, dist, albedo):
        n = dist.shape[1]
        d, c1, c2, R = _v3(n), _v3(n), _v3(n), _z(n)
        out = abi.NdProfileSoA(abi.vec3(d), abi.vec3(c1), abi.vec3(c2), R.ctypes.data)
        kd = [np.ascontiguousarray(dist[j], dtype=f32) for j in range(3)]
        ka = [np.ascontiguousarray(albedo[j], dtype=f32) for j in range(3)]
        self.lib.oracle_ndprofile_set_distance(C.c_size_t(n), abi.vec3(kd), abi.vec3(ka), C.byref(out))
        return dict(distance=np.stack(d), C1=np.stack(c1), C2=np.stack(c2), max_radius=R)

    def ndprofile_get_radius(self, prof, rx):
        n = len(rx)
        r, fl = _z(n), _z(n, np.uint32)
        keep = {k: (np.ascontiguousarray(v) if v.ndim == 1 else [np.ascontiguousarray(v[j]) for j in range(3)])
                for k, v in prof.items()}
        s = abi.NdProfileSoA(abi.vec3(keep["distance"]), abi.vec3(keep["C1"]), abi.vec3(keep["C2"]),
                             keep["max_radius"].ctypes.data)
        self.lib.oracle_ndprofile_get_radius(C.c_size_t(n), C.byref(s), C.c_void_p(rx.ctypes.data),
                                             C.c_void_p(r.ctypes.data), C.c_void_p(fl.ctypes.data))
        return dict(r=r, flags=fl)

    def ndprofile_get_pdf(self, prof, r):
        n = len(r)
        pdf = _z(n)
        keep = {k: (np.ascontiguousarray(v) if v.ndim == 1 else [np.ascontiguousarray(v[j]) for j in range(3)])
                for k, v in prof.items()}
        s = abi.NdProfileSoA(abi.vec3(keep["distance"]), abi.vec3(keep["C1"]), abi.vec3(keep["C2"]),
                             keep["max_radius"].ctypes.data)
        self.lib.oracle_ndprofile_get_pdf(C.c_size_t(n), C.byref(s), C.c_void_p(r.ctypes.data),
                                          C.c_void_p(pdf.ctypes.data))
        return pdf

    def ndprofile_eval_profile(self, prof, r):
        n = len(r)
        rd = _v3(n)
        keep = {k: (np.ascontiguousarray(v) if v.ndim == 1 else [np.ascontiguousarray(v[j]) for j in range(3)])
                for k, v in prof.items()}
        s = abi.NdProfileSoA(abi.vec3(keep["distance"]), abi.vec3(keep["C1"]), abi.vec3(keep["C2"]),
                             keep["max_radius"].ctypes.data)
        self.lib.oracle_ndprofile_eval_profile(C.c_size_t(n), C.byref(s), C.c_void_p(r.ctypes.data),
                                               abi.vec3(rd))
        return np.stack(rd)

    # ---- GaussianProfile (src/rlSss.h:63-97)
    @staticmethod
    def _gauss_struct(prof):
        keep = {k: np.ascontiguousarray(prof[k], dtype=f32) for k in ("variance", "max_radius", "norm")}
        return keep, abi.GaussProfileSoA(keep["variance"].ctypes.data, keep["max_radius"].ctypes.data,
                                         keep["norm"].ctypes.data)

    def gaussprofile_set_distance(self, dist, albedo):
        n = dist.shape[1]
        out = dict(variance=_z(n), max_radius=_z(n), norm=_z(n))
        _, s = self._gauss_struct(out)
        s = abi.GaussProfileSoA(out["variance"].ctypes.data, out["max_radius"].ctypes.data, out["norm"].ctypes.data)
        kd = [np.ascontiguousarray(dist[j], dtype=f32) for j in range(3)]
        ka = [np.ascontiguousarray(albedo[j], dtype=f32) for j in range(3)]
        self.lib.oracle_gaussprofile_set_distance(C.c_size_t(n), abi.vec3(kd), abi.vec3(ka), C.byref(s))
        return out

    def _gauss_unary(self, fn, prof, x):
        n = len(x)
        x = np.ascontiguousarray(x, dtype=f32)
        out = _z(n)
        keep, s = self._gauss_struct(prof)
        fn(C.c_size_t(n), C.byref(s), C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data))
        return out

    def gaussprofile_get_radius(self, prof, rx):
        return self._gauss_unary(self.lib.oracle_gaussprofile_get_radius, prof, rx)

    def gaussprofile_get_pdf(self, prof, r):
        return self._gauss_unary(self.lib.oracle_gaussprofile_get_pdf, prof, r)

    def gaussprofile_eval_profile(self, prof, r):
        return self._gauss_unary(self.lib.oracle_gaussprofile_eval_profile, prof, r)

    def gaussprofile(self, dist_x, rx):
        n = len(rx)
        dist_x, rx = np.ascontiguousarray(dist_x, dtype=f32), np.ascontiguousarray(rx, dtype=f32)
        out = dict(r=_z(n), pdf=_z(n), Rd=_z(n))
        self.lib.oracle_gaussprofile_sample_eval_pdf(C.c_size_t(n), C.c_void_p(dist_x.ctypes.data),
                                                     C.c_void_p(rx.ctypes.data), C.c_void_p(out["r"].ctypes.data),
                                                     C.c_void_p(out["pdf"].ctypes.data), C.c_void_p(out["Rd"].ctypes.data))
        return out

    def skin_profile(self, params, rx):
        n = len(rx)
        r, pdf, rd, fl = _z(n), _z(n), _v3(n), _z(n, np.uint32)
        out = abi.ProfileOut(r.ctypes.data, pdf.ctypes.data, abi.vec3(rd), fl.ctypes.data)
        self.lib.oracle_skin_profile_sample_eval_pdf(C.c_size_t(n), C.byref(params),
                                                     C.c_void_p(rx.ctypes.data), C.byref(out))
        return dict(r=r, pdf=pdf, Rd=np.stack(rd), flags=fl)

    def skin_layer_weights(self, params, avg_f_sheen, avg_f_spec):
        n = len(avg_f_sheen)
        a, b = _z(n), _z(n)
        self.lib.oracle_skin_layer_weights(C.c_size_t(n), C.byref(params),
                                           C.c_void_p(avg_f_sheen.ctypes.data),
                                           C.c_void_p(avg_f_spec.ctypes.data),
                                           C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data))
        return dict(specular_scale=a, sss_weight=b)

    def skin_probe_ray(self, sg, params, rx, ry):
        n = len(rx)
        r, o, d, md, fl = _z(n), _v3(n), _v3(n), _z(n), _z(n, np.uint32)
        out = abi.ProbeOut(r.ctypes.data, abi.vec3(o), abi.vec3(d), md.ctypes.data, fl.ctypes.data)
        self.lib.oracle_skin_probe_ray(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                       C.c_void_p(rx.ctypes.data), C.c_void_p(ry.ctypes.data),
                                       C.byref(out))
        return dict(r=r, origin=np.stack(o), dir=np.stack(d), maxdist=md, flags=fl)

    def skin_probe_mis_pdf(self, sg, params, disp, hit_normal):
        n = disp.shape[1]
        pdf = _z(n)
        kd = [np.ascontiguousarray(disp[j], dtype=f32) for j in range(3)]
        kn = [np.ascontiguousarray(hit_normal[j], dtype=f32) for j in range(3)]
        self.lib.oracle_skin_probe_mis_pdf(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params),
                                           abi.vec3(kd), abi.vec3(kn), C.c_void_p(pdf.ctypes.data))
        return pdf

    # ---- SURVEY.md 8(f) f2-f4
    def skin_glossy_layers(self, sg, params, k, rx_a, ry_a, rx_b, ry_b, li_a=None, li_b=None):
        n = len(rx_a) // k
        sheen, spec, sf, pf, w, fl = _v3(n), _v3(n), _z(n), _z(n), _z(n), _z(n, np.uint32)
        out = abi.SkinLayersOut(abi.vec3(sheen), abi.vec3(spec), sf.ctypes.data, pf.ctypes.data, w.ctypes.data,
                                fl.ctypes.data)
        ka = [np.ascontiguousarray(li_a[j], dtype=f32) for j in range(3)] if li_a is not None else None
        kb = [np.ascontiguousarray(li_b[j], dtype=f32) for j in range(3)] if li_b is not None else None
        self.lib.oracle_skin_glossy_layers(C.c_size_t(n), C.c_uint32(k), C.byref(shading_struct(sg)), C.byref(params),
                                           C.c_void_p(rx_a.ctypes.data), C.c_void_p(ry_a.ctypes.data),
                                           C.c_void_p(rx_b.ctypes.data), C.c_void_p(ry_b.ctypes.data),
                                           abi.vec3(ka), abi.vec3(kb), C.byref(out))
        return dict(sheen=np.stack(sheen), specular=np.stack(spec), sheen_fresnel=sf, specular_fresnel=pf,
                    sss_weight=w, flags=fl)

    def _light_sample(self, fn, head, sg, params, Ld, Li, pl, rx, ry, Li_b, pl_b):
        n = len(pl)
        keep = [[np.ascontiguousarray(a[j], dtype=f32) for j in range(3)] for a in (Ld, Li)]
        light = abi.light_sample(keep[0], keep[1], pl)
        at_l = None
        if Li_b is not None:
            kb = [np.ascontiguousarray(Li_b[j], dtype=f32) for j in range(3)]
            at_l = abi.light_sample(None, kb, pl_b)
        rgb, wl, wb = _v3(n), _z(n), _z(n)
        fn(C.c_size_t(n), C.byref(shading_struct(sg)), C.byref(params), *head, C.byref(light),
           C.c_void_p(rx.ctypes.data if rx is not None else None), C.c_void_p(ry.ctypes.data if ry is not None else None),
           C.byref(at_l) if at_l is not None else None, abi.vec3(rgb), C.c_void_p(wl.ctypes.data), C.c_void_p(wb.ctypes.data))
        return dict(rgb=np.stack(rgb), w_light=wl, w_brdf=wb)

    def ggx_light_sample(self, sg, params, Ld, Li, pl, rx=None, ry=None, Li_b=None, pl_b=None):
        return self._light_sample(self.lib.oracle_ggx_evaluate_light_sample, (), sg, params, Ld, Li, pl, rx, ry, Li_b, pl_b)

    def disney_light_sample(self, sg, params, sample_type, Ld, Li, pl, rx=None, ry=None, Li_b=None, pl_b=None):
        return self._light_sample(self.lib.oracle_disney_evaluate_light_sample, (C.c_int(sample_type),), sg, params,
                                  Ld, Li, pl, rx, ry, Li_b, pl_b)

    def sample_writer(self, node, sg, params, point, sample_type, width, height, rx=None, ry=None):
        """writeRadiance, then (if samples are given) writeSample painted over it."""
        image = np.zeros((3, height, width), dtype=f32)
        self.lib.oracle_sample_writer_radiance(C.c_int(node), C.byref(shading_struct(sg)), C.byref(params),
                                               C.c_size_t(point), C.c_int(sample_type), C.c_int(width), C.c_int(height),
                                               C.c_void_p(image.ctypes.data))
        missing = np.zeros(1, dtype=np.uint32)
        if rx is not None:
            self.lib.oracle_sample_writer_scatter(C.c_int(node), C.byref(shading_struct(sg)), C.byref(params),
                                                  C.c_size_t(point), C.c_int(sample_type), C.c_size_t(len(rx)),
                                                  C.c_void_p(rx.ctypes.data), C.c_void_p(ry.ctypes.data),
                                                  C.c_int(width), C.c_int(height), C.c_void_p(image.ctypes.data),
                                                  C.c_void_p(missing.ctypes.data))
        return image, int(missing[0])

    def albedo_sweep(self, grid, seed, spp_begin, spp_end):
        cells = grid.n_rough * grid.n_cos * grid.n_ior
        table = np.zeros((cells, abi.SWEEP_VALUES_PER_CELL), dtype=np.float64)
        self.lib.oracle_albedo_sweep(C.byref(grid), C.c_uint64(seed), C.c_uint32(spp_begin),
                                     C.c_uint32(spp_end), C.c_void_p(table.ctypes.data))
        return table

    def synth_uniform(self, n, seed, stream, first_index=0, lo=0.0, hi=1.0):
        out = _z(n)
        self.lib.oracle_synth_uniform(C.c_size_t(n), C.c_uint64(seed), C.c_uint32(stream),
                                      C.c_uint64(first_index), C.c_float(lo), C.c_float(hi),
                                      C.c_void_p(out.ctypes.data))
        return out


def load_port():
    if not os.path.exists(PORT_SO):
        build_oracles()
    return Oracle(PORT_SO)


class _PrefixedLib:
    """oracle_X -> f64_oracle_X (oracle/rls_oracle_f64.c renames its exports)."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        return getattr(self._lib, "f64_" + name)


class F64Oracle(Oracle):
    """oracle/librls_oracle_f64.so: the port re-typed to binary64 -- a yardstick for the reference's own rounding noise,
    NOT an oracle of its bits.  The ABI structs keep binary32 arrays; bare `float *` arguments are `double *` there, so
    the four fused calls below hand the uniforms over as float64."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = _PrefixedLib(C.CDLL(path))
        self.lib.oracle_max_threads.restype = C.c_int
        self.kind = "port-f64"

    @staticmethod
    def _d(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    def ggx_sample_eval_pdf(self, sg, params, rx, ry):
        return Oracle.ggx_sample_eval_pdf(self, sg, params, self._d(rx), self._d(ry))

    def ggx_dielectric(self, sg, params, rx, ry):
        return Oracle.ggx_dielectric(self, sg, params, self._d(rx), self._d(ry))

    def disney_sample_eval_pdf(self, sg, params, rx_s, ry_s, rx_d, ry_d):
        return Oracle.disney_sample_eval_pdf(self, sg, params, self._d(rx_s), self._d(ry_s), self._d(rx_d), self._d(ry_d))

    def skin_profile(self, params, rx):
        return Oracle.skin_profile(self, params, self._d(rx))


def load_f64():
    if not os.path.exists(F64_SO):
        build_oracles()
    return F64Oracle(F64_SO)


def load_ref():
    """The reference library, or None when it has not been built (no /root/reference
    and no prebuilt .so shipped)."""
    if not os.path.exists(REF_SO) and os.path.exists(os.path.join(REFERENCE_SRC, "rlGgx.h")):
        build_oracles()
    return Oracle(REF_SO) if os.path.exists(REF_SO) else None


# ------------------------------------------------ compact frames (include/rls_b200.h rls_shading_quat_soa)
def frame_from_quaternion(q):
    """numpy restatement of the header's DEFINITION of the frame of a unit quaternion q = (x, y, z, w) [4, n]: every
    operation below is one binary32 operation (numpy float32 arithmetic rounds each), in the header's order.  Returns
    the dict-of-arrays shading form's U*, V*, N* entries."""
    x, y, z, w = (np.ascontiguousarray(c, dtype=f32) for c in q)
    one = f32(1.0)
    x2, y2, z2 = x + x, y + y, z + z
    xx, yy, zz, xy, xz, yz = x * x2, y * y2, z * z2, x * y2, x * z2, y * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    out = {"Ux": one - (yy + zz), "Uy": xy + wz, "Uz": xz - wy,
           "Vx": xy - wz, "Vy": one - (xx + zz), "Vz": yz + wx,
           "Nx": xz + wy, "Ny": yz - wx, "Nz": one - (xx + yy)}
    return {k: np.ascontiguousarray(v, dtype=f32) for k, v in out.items()}


def quaternion_from_frame(sg):
    """A unit quaternion [4, n] (float32) whose rotation matrix has columns ~ U, V, N of the dict-of-arrays shading form
    (computed in float64, largest-component branch; how a renderer would encode its frames)."""
    m = np.array([[sg["Ux"], sg["Vx"], sg["Nx"]], [sg["Uy"], sg["Vy"], sg["Ny"]], [sg["Uz"], sg["Vz"], sg["Nz"]]], dtype=np.float64)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = (m[i, j] for i in range(3) for j in range(3))
    cand = np.stack([
        np.stack([1 + m00 - m11 - m22, m01 + m10, m02 + m20, m21 - m12]),      # x largest
        np.stack([m01 + m10, 1 - m00 + m11 - m22, m12 + m21, m02 - m20]),      # y largest
        np.stack([m02 + m20, m12 + m21, 1 - m00 - m11 + m22, m10 - m01]),      # z largest
        np.stack([m21 - m12, m02 - m20, m10 - m01, 1 + m00 + m11 + m22]),      # w largest
    ])                                                                           # [4 candidates, 4 comps, n]
    pick = np.argmax(np.stack([cand[0, 0], cand[1, 1], cand[2, 2], cand[3, 3]]), axis=0)
    q = np.take_along_axis(cand, pick[None, None, :], axis=0)[0]
    q = q / np.sqrt(np.sum(q * q, axis=0))
    return np.ascontiguousarray(q, dtype=f32)


def shading_from_quaternion(q, sg):
    """The shading dict the oracle sees for a compact batch: decoded frame + the batch's own wo / backfacing."""
    out = frame_from_quaternion(q)
    for k in ("wox", "woy", "woz"):
        out[k] = sg[k]
    out["backfacing"] = sg.get("backfacing")
    return out


# ------------------------------------------------ BASELINE.json workload recipes
def workload_ggx_conductor(n, seed=0x5EED0001):
    """Config 1: gold fixture (testsuite/mtoa/0002/data/ggx_gold.ass:17-21)."""
    sg = make_shading(n, seed)
    rx, ry = hash_uniform(n, seed, 0), hash_uniform(n, seed, 1)
    return sg, abi.ggx_params(KsColor=(1.0, 1.0, 1.0), specularRoughness=0.3, ior=0.47, anisotropic=0.0), rx, ry


def workload_ggx_dielectric(n, seed=0x5EED0002, aniso=False):
    """Config 2: per-sample roughness ~U[0.05,1], ior ~U[1.05,2.5], 25% back-facing."""
    sg = make_shading(n, seed, backfacing_fraction=0.25)
    rx, ry = hash_uniform(n, seed, 0), hash_uniform(n, seed, 1)
    rough = hash_uniform(n, seed, 2, lo=0.05, hi=1.0)
    ior = hash_uniform(n, seed, 3, lo=1.05, hi=2.5)
    kw = dict(specularRoughness=rough, ior=ior)
    if aniso:
        kw["anisotropic"] = hash_uniform(n, seed, 4)
    return sg, abi.ggx_params(**kw), rx, ry


def workload_disney(n, seed=0x5EED0003):
    """Config 3: every parameter spatially varying in [0,1]."""
    sg = make_shading(n, seed)
    u = [hash_uniform(n, seed, s) for s in range(4)]
    names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
             "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
    kw = {nm: hash_uniform(n, seed, 20 + j) for j, nm in enumerate(names)}
    kw["base_color"] = tuple(hash_uniform(n, seed, 30 + j) for j in range(3))
    return sg, abi.disney_params(**kw), u


def workload_skin(n, seed=0x5EED0004):
    """Config 4: sss_color ~U[0.05,1]^3, sss_scatter_dist ~U[0.05,2]^3, multiplier 1."""
    rx = hash_uniform(n, seed, 0)
    color = tuple(hash_uniform(n, seed, 40 + j, lo=0.05, hi=1.0) for j in range(3))
    dist = tuple(hash_uniform(n, seed, 50 + j, lo=0.05, hi=2.0) for j in range(3))
    return abi.skin_params(sss_color=color, sss_scatter_dist=dist, sss_dist_multiplier=1.0), rx
