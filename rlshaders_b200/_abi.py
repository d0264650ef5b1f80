"""ctypes mirror of include/rls_b200.h (structs, constants, flag bits).

Pure data layout: nothing here touches a device.  The product binding
(`rlshaders_b200._lib`) builds its argument blocks from these classes; the parity tests
reuse them so that one descriptor feeds both sides of a comparison.
"""
import ctypes as C

ABI_VERSION = 1

RLS_OK = 0
ARITH_FAST, ARITH_EXACT, ARITH_TOLERANT = 0, 1, 2
RLS_ERR_INVALID_ARGUMENT = -1
RLS_ERR_CUDA = -2
RLS_ERR_NO_DEVICE = -3
RLS_ERR_OUT_OF_MEMORY = -4

# DisneySampler::setSampleType values (AI_RAY_DIFFUSE / AI_RAY_GLOSSY), rlDisney.cpp:194
RLS_RAY_DIFFUSE = 0x20
RLS_RAY_GLOSSY = 0x40
GGX_SAMPLER_VNDF = 0
GGX_SAMPLER_NDF = 1

FLAG_ZERO_L = 0x0001
FLAG_BELOW_HORIZON = 0x0002
FLAG_PDF_ZERO = 0x0004
FLAG_F_BLACK = 0x0008
FLAG_ENTERING = 0x0010
FLAG_TIR = 0x0020
FLAG_PDF_FLOORED = 0x0040
FLAG_SLOPE_EARLY_OUT = 0x0080
FLAG_LOBE_SHIFT = 8
FLAG_LOBE_MASK = 0x0300
FLAG_EXP_LOBE = 0x0400
FLAG_DEGENERATE = 0x0800
FLAG_PROBE_AXIS_SHIFT = 12
FLAG_PROBE_AXIS_MASK = 0x3000
FLAG_DIFFUSE_SHIFT = 16

SWEEP_VALUES_PER_CELL = 5

_fp = C.POINTER(C.c_float)


class CVec3(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p)]


Vec3 = CVec3  # same layout; constness is a C-side notion


class Param1(C.Structure):
    _fields_ = [("value", C.c_float), ("array", C.c_void_p)]


class Param3(C.Structure):
    _fields_ = [("value", C.c_float * 3), ("array", CVec3)]


class ShadingSoA(C.Structure):
    _fields_ = [("U", CVec3), ("V", CVec3), ("N", CVec3), ("wo", CVec3),
                ("backfacing", C.c_void_p)]


class ShadingQuatSoA(C.Structure):
    # rls_shading_quat_soa: unit quaternion (x, y, z, w) in place of U, V, N (include/rls_b200.h)
    _fields_ = [("qx", C.c_void_p), ("qy", C.c_void_p), ("qz", C.c_void_p), ("qw", C.c_void_p), ("wo", CVec3),
                ("backfacing", C.c_void_p)]


class GgxParams(C.Structure):
    # names = rlGgx node parameters, reference src/rlGgx.cpp:172-186
    _fields_ = [("KsColor", Param3), ("Ks", Param1), ("specularRoughness", Param1),
                ("ior", Param1), ("anisotropic", Param1),
                ("KdColor", Param3), ("Kd", Param1), ("diffuseRoughness", Param1),
                ("KtColor", Param3), ("Kt", Param1), ("opacity", Param1),
                ("opacity_color", Param3), ("normal_sampler", C.c_int32)]


class DisneyParams(C.Structure):
    # names = rlDisney node parameters, reference src/rlDisney.cpp:606-625
    _fields_ = [("base_color", Param3), ("subsurface", Param1), ("metallic", Param1),
                ("specular", Param1), ("specular_tint", Param1), ("roughness", Param1),
                ("anisotropic", Param1), ("sheen", Param1), ("sheen_tint", Param1),
                ("clearcoat", Param1), ("clearcoat_gloss", Param1),
                ("opacity", Param3), ("indirectDiffuseScale", Param1),
                ("indirectSpecularScale", Param1), ("sample_from_visible_normal", C.c_int32)]


class SkinParams(C.Structure):
    # names = rlSkin node parameters, reference src/rlSkin.cpp:109-131
    _fields_ = [("sss_color", Param3), ("sss_weight", Param1), ("sss_dist_multiplier", Param1),
                ("sss_scatter_dist", Param3), ("sss_cavity_fadeout", C.c_int32),
                ("specular_color", Param3), ("specular_weight", Param1),
                ("specular_roughness", Param1), ("specular_ior", Param1),
                ("sheen_color", Param3), ("sheen_weight", Param1),
                ("sheen_roughness", Param1), ("sheen_ior", Param1),
                ("opacity", Param1), ("opacity_color", Param3)]


class BsdfOut(C.Structure):
    _fields_ = [("wi", Vec3), ("f", Vec3), ("pdf", C.c_void_p), ("fresnel", C.c_void_p),
                ("flags", C.c_void_p)]


class GgxDielectricOut(C.Structure):
    _fields_ = [("fresnel", C.c_void_p), ("wi_r", Vec3), ("f_r", C.c_void_p),
                ("pdf_r", C.c_void_p), ("wi_t", Vec3), ("f_t", C.c_void_p),
                ("weight_t", C.c_void_p), ("flags", C.c_void_p)]


class DisneyOut(C.Structure):
    _fields_ = [("wi_s", Vec3), ("f_s", Vec3), ("pdf_s", C.c_void_p),
                ("wi_d", Vec3), ("f_d", Vec3), ("pdf_d", C.c_void_p),
                ("flags", C.c_void_p)]


class NdProfileSoA(C.Structure):
    _fields_ = [("distance", Vec3), ("C1", Vec3), ("C2", Vec3), ("max_radius", C.c_void_p)]


class GaussProfileSoA(C.Structure):
    _fields_ = [("variance", C.c_void_p), ("max_radius", C.c_void_p), ("norm", C.c_void_p)]


class ProfileOut(C.Structure):
    _fields_ = [("r", C.c_void_p), ("pdf", C.c_void_p), ("Rd", Vec3), ("flags", C.c_void_p)]


class ProbeOut(C.Structure):
    _fields_ = [("r", C.c_void_p), ("origin", Vec3), ("dir", Vec3), ("maxdist", C.c_void_p),
                ("flags", C.c_void_p)]


class SkinLayersOut(C.Structure):
    _fields_ = [("sheen", Vec3), ("specular", Vec3), ("sheen_fresnel", C.c_void_p),
                ("specular_fresnel", C.c_void_p), ("sss_weight", C.c_void_p), ("flags", C.c_void_p)]


class LightSample(C.Structure):
    _fields_ = [("dir", CVec3), ("radiance", CVec3), ("pdf", C.c_void_p)]


def light_sample(direction, radiance, pdf):
    """rls_light_sample from ([3] arrays or None, [3] arrays, array)."""
    s = LightSample(vec3(direction), vec3(radiance), _addr(pdf))
    s._keepalive = (direction, radiance, pdf)
    return s


NODE_GGX, NODE_DISNEY = 0, 1
SKIN_SHEEN_EVALUATED, SKIN_SPECULAR_EVALUATED, SKIN_SSS_SKIPPED = 1, 2, 4


class SweepGrid(C.Structure):
    _fields_ = [("n_rough", C.c_int32), ("n_cos", C.c_int32), ("n_ior", C.c_int32),
                ("roughness_lo", C.c_float), ("roughness_hi", C.c_float),
                ("ior_lo", C.c_float), ("ior_hi", C.c_float)]


# Node-parameter defaults, from the reference's node_parameters blocks.
GGX_DEFAULTS = dict(KsColor=(1.0, 1.0, 1.0), Ks=0.5, specularRoughness=0.0, ior=1.0,
                    anisotropic=0.0, KdColor=(1.0, 1.0, 1.0), Kd=0.5, diffuseRoughness=0.0,
                    KtColor=(1.0, 1.0, 1.0), Kt=0.0, opacity=1.0,
                    opacity_color=(1.0, 1.0, 1.0), normal_sampler=0)    # rlGgx.cpp:172-186
DISNEY_DEFAULTS = dict(base_color=(1.0, 1.0, 1.0), subsurface=0.0, metallic=0.0, specular=0.0,
                       specular_tint=0.0, roughness=0.0, anisotropic=0.0, sheen=0.0,
                       sheen_tint=0.0, clearcoat=0.0, clearcoat_gloss=0.0,
                       opacity=(1.0, 1.0, 1.0), indirectDiffuseScale=1.0,
                       indirectSpecularScale=1.0, sample_from_visible_normal=1)  # rlDisney.cpp:606-628
SKIN_DEFAULTS = dict(sss_color=(1.0, 1.0, 1.0), sss_weight=1.0, sss_dist_multiplier=1.0,
                     sss_scatter_dist=(1.0, 1.0, 1.0), sss_cavity_fadeout=1,
                     specular_color=(1.0, 1.0, 1.0), specular_weight=0.6,
                     specular_roughness=0.5, specular_ior=1.44,
                     sheen_color=(1.0, 1.0, 1.0), sheen_weight=0.0, sheen_roughness=0.35,
                     sheen_ior=1.44, opacity=1.0, opacity_color=(1.0, 1.0, 1.0))  # rlSkin.cpp:109-131


def _addr(a):
    """Address of an array-like: numpy array, torch tensor, int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):          # torch.Tensor
        return a.data_ptr()
    if hasattr(a, "ctypes"):            # numpy.ndarray
        return a.ctypes.data
    raise TypeError(f"cannot take the address of {type(a)!r}")


def vec3(xyz):
    """Build an rls_cvec3 / rls_vec3 from a 3-tuple of arrays (or None)."""
    if xyz is None:
        return CVec3(None, None, None)
    x, y, z = xyz
    return CVec3(_addr(x), _addr(y), _addr(z))


def param1(v):
    """Scalar -> uniform; array -> per-sample."""
    if isinstance(v, (int, float)):
        return Param1(float(v), None)
    return Param1(0.0, _addr(v))


def param3(v):
    """3 scalars -> uniform; 3 arrays -> per-sample."""
    a, b, c = v
    if all(isinstance(t, (int, float)) for t in (a, b, c)):
        return Param3((C.c_float * 3)(float(a), float(b), float(c)), CVec3(None, None, None))
    return Param3((C.c_float * 3)(0.0, 0.0, 0.0), vec3((a, b, c)))


def _fill_params(cls, defaults, overrides):
    unknown = set(overrides) - set(defaults)
    if unknown:
        raise TypeError(f"unknown node parameter(s) {sorted(unknown)} for {cls.__name__}")
    merged = dict(defaults)
    merged.update(overrides)
    p = cls()
    keep = []
    for name, ctype in cls._fields_:
        v = merged[name]
        if ctype is Param1:
            setattr(p, name, param1(v))
        elif ctype is Param3:
            setattr(p, name, param3(v))
        else:
            setattr(p, name, int(v))
        keep.append(v)
    p._keepalive = keep      # hold references to any arrays
    return p


def ggx_params(**kw):
    return _fill_params(GgxParams, GGX_DEFAULTS, kw)


def disney_params(**kw):
    return _fill_params(DisneyParams, DISNEY_DEFAULTS, kw)


def skin_params(**kw):
    return _fill_params(SkinParams, SKIN_DEFAULTS, kw)


def shading_quat(q, wo, backfacing=None):
    """q: four arrays (x, y, z, w)."""
    s = ShadingQuatSoA(_addr(q[0]), _addr(q[1]), _addr(q[2]), _addr(q[3]), vec3(wo), _addr(backfacing))
    s._keepalive = (q, wo, backfacing)
    return s


def shading(U, V, N, wo, backfacing=None):
    s = ShadingSoA(vec3(U), vec3(V), vec3(N), vec3(wo), _addr(backfacing))
    s._keepalive = (U, V, N, wo, backfacing)
    return s
