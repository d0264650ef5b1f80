"""CPU: the C-ABI library loads without a GPU, exports every symbol include/rls_b200.h
declares, its struct layouts match the ctypes mirror, and it fails loudly (no CPU path)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from rlshaders_b200 import _abi as abi
from rlshaders_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rls_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rls_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/rls_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names


def test_struct_layouts_match_header():
    """Compile a C probe against the real header and compare sizeof/offsetof."""
    structs = {"rls_cvec3": abi.CVec3, "rls_param1": abi.Param1, "rls_param3": abi.Param3,
               "rls_shading_soa": abi.ShadingSoA, "rls_shading_quat_soa": abi.ShadingQuatSoA, "rls_ggx_params": abi.GgxParams,
               "rls_disney_params": abi.DisneyParams, "rls_skin_params": abi.SkinParams,
               "rls_bsdf_out": abi.BsdfOut, "rls_ggx_dielectric_out": abi.GgxDielectricOut,
               "rls_disney_out": abi.DisneyOut, "rls_ndprofile_soa": abi.NdProfileSoA,
               "rls_gaussprofile_soa": abi.GaussProfileSoA,
               "rls_profile_out": abi.ProfileOut, "rls_probe_out": abi.ProbeOut,
               "rls_sweep_grid": abi.SweepGrid, "rls_skin_layers_out": abi.SkinLayersOut,
               "rls_light_sample": abi.LightSample}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "probe.c"), os.path.join(td, "probe")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c11", "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_flag_constants_match_header():
    text = open(HEADER).read()
    for cname, val in (("RLS_FLAG_ZERO_L", abi.FLAG_ZERO_L), ("RLS_FLAG_BELOW_HORIZON", abi.FLAG_BELOW_HORIZON),
                       ("RLS_FLAG_PDF_ZERO", abi.FLAG_PDF_ZERO), ("RLS_FLAG_F_BLACK", abi.FLAG_F_BLACK),
                       ("RLS_FLAG_ENTERING", abi.FLAG_ENTERING), ("RLS_FLAG_TIR", abi.FLAG_TIR),
                       ("RLS_FLAG_PDF_FLOORED", abi.FLAG_PDF_FLOORED), ("RLS_FLAG_SLOPE_EARLY_OUT", abi.FLAG_SLOPE_EARLY_OUT),
                       ("RLS_ARITH_FAST", abi.ARITH_FAST), ("RLS_ARITH_EXACT", abi.ARITH_EXACT), ("RLS_ARITH_TOLERANT", abi.ARITH_TOLERANT),
                       ("RLS_FLAG_EXP_LOBE", abi.FLAG_EXP_LOBE),
                       ("RLS_FLAG_DEGENERATE", abi.FLAG_DEGENERATE), ("RLS_FLAG_LOBE_MASK", abi.FLAG_LOBE_MASK),
                       ("RLS_RAY_DIFFUSE", abi.RLS_RAY_DIFFUSE), ("RLS_RAY_GLOSSY", abi.RLS_RAY_GLOSSY)):
        m = re.search(rf"#define\s+{cname}\s+(0x[0-9a-fA-F]+|\d+)", text)
        assert m and int(m.group(1), 0) == val, cname


def test_node_names_and_version():
    lib = _lib.load()
    assert lib.rls_abi_version() == abi.ABI_VERSION
    # same enumeration contract as NodeLoader (reference src/_PluginMain.cpp:16-47)
    assert [lib.rls_node_name(i) for i in range(4)] == [b"rlGgx", b"rlDisney", b"rlSkin", None]


def test_parameter_names_are_the_node_parameter_names():
    """Drop-in surface: field names = node parameter names of the reference
    (src/rlGgx.cpp:172-186, src/rlDisney.cpp:606-625, src/rlSkin.cpp:109-131)."""
    ggx = [f for f, _ in abi.GgxParams._fields_]
    assert ggx[:5] == ["KsColor", "Ks", "specularRoughness", "ior", "anisotropic"]
    assert {"KdColor", "Kd", "diffuseRoughness", "KtColor", "Kt", "opacity", "opacity_color"} <= set(ggx)
    disney = [f for f, _ in abi.DisneyParams._fields_]
    assert disney[:11] == ["base_color", "subsurface", "metallic", "specular", "specular_tint", "roughness",
                           "anisotropic", "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
    skin = {f for f, _ in abi.SkinParams._fields_}
    assert {"sss_color", "sss_weight", "sss_dist_multiplier", "sss_scatter_dist", "sss_cavity_fadeout",
            "specular_color", "specular_weight", "specular_roughness", "specular_ior", "sheen_color",
            "sheen_weight", "sheen_roughness", "sheen_ior"} <= skin
    with pytest.raises(TypeError):
        abi.ggx_params(roughness=0.3)      # not a rlGgx parameter name


def test_no_cpu_fallback():
    """Without a usable sm_100 device rls_init must fail with an explanatory error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.rls_init(0, None, C.byref(h))
    assert rc == abi.RLS_ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in lib.rls_last_error_string(None)
    from rlshaders_b200 import api
    with pytest.raises(api.RlsError):
        api.Context(0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under rlshaders_b200/ may name it, and no tool under tools/ may load it
    (the oracle-using hunts live in tests/hunts/; tools/sanitize.sh only hands its `host` leg over to them).  The header's
    comments cite where the oracle restates a definition; it includes nothing from there."""
    assert "#include \"../oracle" not in open(HEADER).read() and "#include \"oracle" not in open(HEADER).read()
    for sub in ("rlshaders_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    for needle in ("librls_oracle", "librls_ref", "oracle_lib", "oracle/", "oracle_api", "_ref/"):
                        assert needle not in text, f"{sub}/{f} mentions {needle}"
