"""Loads rlshaders_b200/librls_b200.so (the hand-written CUDA library behind
include/rls_b200.h) and declares its C ABI for ctypes.

There is no fallback: a missing library raises ImportError from `load()`, and
`rls_init` itself fails when no sm_100 device is present.
"""
import ctypes as C
import os

from . import _abi as abi

# RLS_B200_LIB lets a developer point at an experimental build of the SAME library (tuning
# sweeps); there is still exactly one implementation and no fallback.
LIB_PATH = os.environ.get("RLS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                          "librls_b200.so")

# Every extern "C" symbol include/rls_b200.h declares (tests check the export list).
SYMBOLS = [
    "rls_init", "rls_shutdown", "rls_synchronize", "rls_last_error_string", "rls_abi_version",
    "rls_kernel_launch_count", "rls_node_name", "rls_set_arith_policy", "rls_fallback_count",
    "rls_ggx_eval_sample", "rls_ggx_eval_brdf", "rls_ggx_eval_pdf", "rls_ggx_sample_eval_pdf",
    "rls_ggx_dielectric_sample_eval_pdf", "rls_ggx_refract_direction", "rls_ggx_eval_btdf", "rls_ggx_sample_weight",
    "rls_disney_eval_sample", "rls_disney_eval_brdf", "rls_disney_eval_pdf",
    "rls_disney_sample_eval_pdf",
    "rls_ndprofile_set_distance", "rls_ndprofile_get_radius", "rls_ndprofile_get_pdf",
    "rls_ndprofile_eval_profile", "rls_skin_profile_sample_eval_pdf", "rls_skin_layer_weights",
    "rls_skin_probe_ray", "rls_skin_probe_mis_pdf",
    "rls_gaussprofile_set_distance", "rls_gaussprofile_get_radius", "rls_gaussprofile_get_pdf",
    "rls_gaussprofile_eval_profile", "rls_gaussprofile_sample_eval_pdf",
    "rls_ggx_sample_eval_pdf_host", "rls_ggx_dielectric_sample_eval_pdf_host",
    "rls_disney_sample_eval_pdf_host", "rls_skin_profile_sample_eval_pdf_host",
    "rls_frame_from_quaternion", "rls_ggx_sample_eval_pdf_hostq", "rls_ggx_dielectric_sample_eval_pdf_hostq",
    "rls_disney_sample_eval_pdf_hostq",
    "rls_host_alloc", "rls_host_free",
    "rls_albedo_sweep", "rls_synth_uniform", "rls_synth_shading", "rls_debug_libm", "rls_debug_policy_check",
    "rls_skin_glossy_layers", "rls_ggx_evaluate_light_sample", "rls_disney_evaluate_light_sample",
    "rls_sample_writer_radiance", "rls_sample_writer_scatter",
    "rls_get_arith_policy", "rls_stream", "rls_device_alloc", "rls_device_free", "rls_memcpy_to_host",
    "rls_multi_init", "rls_multi_shutdown", "rls_multi_device_count", "rls_multi_context", "rls_multi_synchronize",
    "rls_multi_last_error_string", "rls_multi_timer_begin", "rls_multi_timer_end", "rls_multi_albedo_sweep",
    "rls_multi_graph_replays", "rls_multi_set_nccl_library", "rls_multi_partition",
]

# librls_b200_experiments.so: the same sources built with -DRLS_EXPERIMENTS (csrc/experiments/, kernels that
# measured slower + their environment switches); only tests/test_experiments.py and the A/B tools load it.
EXPERIMENTS_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librls_b200_experiments.so")

_libs = {}


def load(path=None):
    """Return the ctypes handle of the product library (or of `path`, an alternative build of the SAME library),
    loading it on first use.  Raises ImportError (never falls back) when the CUDA library has not been built --
    run __graft_entry__.build()."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension is not built "
            "(python -c 'import __graft_entry__ as g; g.build()'). rlshaders_b200 has no CPU path.")
    lib = C.CDLL(path)
    vp, sz, u64, u32, i32, f = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, C.c_float
    P = C.POINTER
    sig = {
        "rls_init": [i32, vp, P(vp)],
        "rls_shutdown": [vp],
        "rls_synchronize": [vp],
        "rls_ggx_eval_sample": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), vp, vp, abi.Vec3, vp],
        "rls_ggx_eval_brdf": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), abi.CVec3, abi.Vec3],
        "rls_ggx_eval_pdf": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), abi.CVec3, vp],
        "rls_ggx_refract_direction": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), abi.CVec3, abi.Vec3, vp],
        "rls_ggx_eval_btdf": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), abi.CVec3, vp],
        "rls_ggx_sample_weight": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), abi.CVec3, abi.CVec3, vp],
        "rls_ggx_sample_eval_pdf": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), vp, vp, P(abi.BsdfOut)],
        "rls_ggx_dielectric_sample_eval_pdf": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), vp, vp,
                                               P(abi.GgxDielectricOut)],
        "rls_disney_eval_sample": [vp, sz, P(abi.ShadingSoA), P(abi.DisneyParams), i32, vp, vp, abi.Vec3, vp],
        "rls_disney_eval_brdf": [vp, sz, P(abi.ShadingSoA), P(abi.DisneyParams), i32, abi.CVec3, abi.Vec3],
        "rls_disney_eval_pdf": [vp, sz, P(abi.ShadingSoA), P(abi.DisneyParams), i32, abi.CVec3, vp],
        "rls_disney_sample_eval_pdf": [vp, sz, P(abi.ShadingSoA), P(abi.DisneyParams), vp, vp, vp, vp,
                                       P(abi.DisneyOut)],
        "rls_ndprofile_set_distance": [vp, sz, abi.CVec3, abi.CVec3, P(abi.NdProfileSoA)],
        "rls_ndprofile_get_radius": [vp, sz, P(abi.NdProfileSoA), vp, vp, vp],
        "rls_ndprofile_get_pdf": [vp, sz, P(abi.NdProfileSoA), vp, vp],
        "rls_ndprofile_eval_profile": [vp, sz, P(abi.NdProfileSoA), vp, abi.Vec3],
        "rls_skin_profile_sample_eval_pdf": [vp, sz, P(abi.SkinParams), vp, P(abi.ProfileOut)],
        "rls_skin_layer_weights": [vp, sz, P(abi.SkinParams), vp, vp, vp, vp],
        "rls_skin_probe_ray": [vp, sz, P(abi.ShadingSoA), P(abi.SkinParams), vp, vp, P(abi.ProbeOut)],
        "rls_skin_probe_mis_pdf": [vp, sz, P(abi.ShadingSoA), P(abi.SkinParams), abi.CVec3, abi.CVec3, vp],
        "rls_gaussprofile_set_distance": [vp, sz, abi.CVec3, abi.CVec3, P(abi.GaussProfileSoA)],
        "rls_gaussprofile_get_radius": [vp, sz, P(abi.GaussProfileSoA), vp, vp],
        "rls_gaussprofile_get_pdf": [vp, sz, P(abi.GaussProfileSoA), vp, vp],
        "rls_gaussprofile_eval_profile": [vp, sz, P(abi.GaussProfileSoA), vp, vp],
        "rls_gaussprofile_sample_eval_pdf": [vp, sz, vp, vp, vp, vp, vp],
        "rls_ggx_sample_eval_pdf_host": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), vp, vp,
                                         P(abi.BsdfOut), sz],
        "rls_ggx_dielectric_sample_eval_pdf_host": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), vp, vp,
                                                    P(abi.GgxDielectricOut), sz],
        "rls_disney_sample_eval_pdf_host": [vp, sz, P(abi.ShadingSoA), P(abi.DisneyParams), vp, vp, vp, vp,
                                            P(abi.DisneyOut), sz],
        "rls_skin_profile_sample_eval_pdf_host": [vp, sz, P(abi.SkinParams), vp, P(abi.ProfileOut), sz],
        "rls_frame_from_quaternion": [vp, sz, vp, vp, vp, vp, abi.Vec3, abi.Vec3, abi.Vec3],
        "rls_ggx_sample_eval_pdf_hostq": [vp, sz, P(abi.ShadingQuatSoA), P(abi.GgxParams), vp, vp, P(abi.BsdfOut), sz],
        "rls_ggx_dielectric_sample_eval_pdf_hostq": [vp, sz, P(abi.ShadingQuatSoA), P(abi.GgxParams), vp, vp,
                                                     P(abi.GgxDielectricOut), sz],
        "rls_disney_sample_eval_pdf_hostq": [vp, sz, P(abi.ShadingQuatSoA), P(abi.DisneyParams), vp, vp, vp, vp,
                                             P(abi.DisneyOut), sz],
        "rls_host_alloc": [vp, sz, P(vp)],
        "rls_host_free": [vp, vp],
        "rls_albedo_sweep": [vp, P(abi.SweepGrid), u64, u32, u32, vp],
        "rls_synth_uniform": [vp, sz, u64, u32, u64, f, f, vp],
        "rls_synth_shading": [vp, sz, u64, u64, f, f, f, P(abi.ShadingSoA)],
        "rls_debug_libm": [vp, i32, sz, vp, vp, vp, vp],
        "rls_debug_policy_check": [vp, i32, u32, u64, u32, f, vp],
        "rls_skin_glossy_layers": [vp, sz, u32, P(abi.ShadingSoA), P(abi.SkinParams), vp, vp, vp, vp,
                                   abi.CVec3, abi.CVec3, P(abi.SkinLayersOut)],
        "rls_ggx_evaluate_light_sample": [vp, sz, P(abi.ShadingSoA), P(abi.GgxParams), P(abi.LightSample), vp, vp,
                                          P(abi.LightSample), abi.Vec3, vp, vp],
        "rls_disney_evaluate_light_sample": [vp, sz, P(abi.ShadingSoA), P(abi.DisneyParams), i32, P(abi.LightSample),
                                             vp, vp, P(abi.LightSample), abi.Vec3, vp, vp],
        "rls_sample_writer_radiance": [vp, i32, P(abi.ShadingSoA), vp, sz, i32, i32, i32, vp],
        "rls_sample_writer_scatter": [vp, i32, P(abi.ShadingSoA), vp, sz, i32, sz, vp, vp, i32, i32, vp, vp, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.rls_last_error_string.argtypes = [vp]
    lib.rls_last_error_string.restype = C.c_char_p
    lib.rls_abi_version.argtypes = []
    lib.rls_abi_version.restype = C.c_int
    lib.rls_kernel_launch_count.argtypes = [vp]
    lib.rls_kernel_launch_count.restype = C.c_uint64
    lib.rls_set_arith_policy.argtypes = [vp, i32]
    lib.rls_set_arith_policy.restype = C.c_int
    lib.rls_fallback_count.argtypes = [vp, C.POINTER(C.c_uint64), i32]
    lib.rls_fallback_count.restype = C.c_int
    lib.rls_node_name.argtypes = [i32]
    lib.rls_node_name.restype = C.c_char_p
    lib.rls_get_arith_policy.argtypes = [vp]
    lib.rls_get_arith_policy.restype = C.c_int
    lib.rls_stream.argtypes = [vp]
    lib.rls_stream.restype = vp
    for name, argtypes in {"rls_device_alloc": [vp, sz, P(vp)], "rls_device_free": [vp, vp], "rls_memcpy_to_host": [vp, vp, vp, sz],
                           "rls_multi_init": [i32, P(i32), P(vp)], "rls_multi_shutdown": [vp], "rls_multi_synchronize": [vp],
                           "rls_multi_timer_begin": [vp], "rls_multi_timer_end": [vp, P(f), P(f)],
                           "rls_multi_albedo_sweep": [vp, P(abi.SweepGrid), u64, u32, P(vp), i32],
                           "rls_multi_set_nccl_library": [C.c_char_p],
                           "rls_multi_partition": [u64, i32, i32, P(u64), P(u64)]}.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.rls_multi_device_count.argtypes = [vp]
    lib.rls_multi_device_count.restype = C.c_int
    lib.rls_multi_context.argtypes = [vp, i32]
    lib.rls_multi_context.restype = vp
    lib.rls_multi_last_error_string.argtypes = [vp]
    lib.rls_multi_last_error_string.restype = C.c_char_p
    lib.rls_multi_graph_replays.argtypes = [vp]
    lib.rls_multi_graph_replays.restype = C.c_uint64
    if lib.rls_abi_version() != abi.ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.rls_abi_version()} != {abi.ABI_VERSION}")
    _libs[path] = lib
    return lib
