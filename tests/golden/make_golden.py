"""Generates tests/golden/*.npz from the REFERENCE library (oracle/_ref/librls_ref.so =
the reference's own sources compiled here against oracle/shim/ai.h).

Run in the build container only (needs /root/reference to build the library):
    python tests/golden/make_golden.py
Each file holds the inputs and the reference outputs of one op on a few thousand seeded
samples, plus named fixtures mined from the reference's .ass scenes:
teflon / gold / anisotropic rlGgx, rlDisney 0004-0008, rlSkin 0009.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from rlshaders_b200 import _abi as abi  # noqa: E402

N = 2048


def flat(prefix, d):
    return {f"{prefix}{k}": v for k, v in d.items() if v is not None}


def load_reference():
    ref = ol.load_ref()
    assert ref is not None and ref.kind == "reference", "reference library not built"
    ref.set_threads(1)
    return ref


def gaussian(ref):
    """GaussianProfile (src/rlSss.h:63-97): varying radii incl. 0 (variance 0 -> NaN, as written) + the 0009 distance 1."""
    rx = ol.hash_uniform(N, 106, 0)
    dist_x = ol.hash_uniform(N, 106, 1, lo=0.05, hi=2.0)
    dist_x[:8] = 0.0
    dist_x[8:16] = 1.0
    with np.errstate(all="ignore"):
        out = dict(rx=rx, dist_x=dist_x)
        out.update(flat("out_", ref.gaussprofile(dist_x, rx)))
        dist = np.stack([dist_x, np.zeros_like(dist_x), np.zeros_like(dist_x)])
        prof = ref.gaussprofile_set_distance(dist, np.ones_like(dist))
        out.update(flat("state_", prof))
        r = ol.hash_uniform(N, 106, 2, lo=0.0, hi=2.5)
        out.update(r=r, pdf_at_r=ref.gaussprofile_get_pdf(prof, r), rd_at_r=ref.gaussprofile_eval_profile(prof, r),
                   radius_at_rx=ref.gaussprofile_get_radius(prof, rx))
    np.savez_compressed(os.path.join(HERE, "gaussian_profile.npz"), **out)


def main():
    ref = load_reference()
    if sys.argv[1:] == ["gaussian"]:        # regenerate this one file only
        gaussian(ref)
        return

    # ---- rlGgx named fixtures: testsuite/mtoa/0001-0003 (roughness, ior, anisotropic)
    sg = ol.make_shading(N, 101)
    rx, ry = ol.hash_uniform(N, 101, 0), ol.hash_uniform(N, 101, 1)
    out = dict(flat("sg_", sg), rx=rx, ry=ry)
    for name, (rough, ior, aniso) in dict(teflon=(0.35, 1.35, 0.0), gold=(0.35, 0.47, 0.0),
                                          anisotropic=(0.3, 0.47, 1.0), gold_bench=(0.3, 0.47, 0.0)).items():
        p = abi.ggx_params(specularRoughness=rough, ior=ior, anisotropic=aniso)
        out.update(flat(f"{name}_", ref.ggx_sample_eval_pdf(sg, p, rx, ry)))
        out[f"{name}_params"] = np.array([rough, ior, aniso], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "ggx_fixtures.npz"), **out)

    # ---- rlGgx dielectric, per-sample parameters, anisotropy on, 25% back-facing
    sg = ol.make_shading(N, 102, backfacing_fraction=0.25)
    rx, ry = ol.hash_uniform(N, 102, 0), ol.hash_uniform(N, 102, 1)
    kw = dict(specularRoughness=ol.hash_uniform(N, 102, 2, lo=0.05, hi=1.0),
              ior=ol.hash_uniform(N, 102, 3, lo=1.05, hi=2.5), anisotropic=ol.hash_uniform(N, 102, 4))
    out = dict(flat("sg_", sg), rx=rx, ry=ry, **{f"p_{k}": v for k, v in kw.items()})
    out.update(flat("out_", ref.ggx_dielectric(sg, abi.ggx_params(**kw), rx, ry)))
    np.savez_compressed(os.path.join(HERE, "ggx_dielectric.npz"), **out)

    # ---- rlDisney: varying parameters + the five scene fixtures 0004-0008
    sg = ol.make_shading(N, 103)
    u = [ol.hash_uniform(N, 103, s) for s in range(4)]
    names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
             "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
    kw = {nm: ol.hash_uniform(N, 103, 20 + j) for j, nm in enumerate(names)}
    base = tuple(ol.hash_uniform(N, 103, 30 + j) for j in range(3))
    out = dict(flat("sg_", sg), u0=u[0], u1=u[1], u2=u[2], u3=u[3], base_r=base[0], base_g=base[1], base_b=base[2],
               **{f"p_{k}": v for k, v in kw.items()})
    out.update(flat("out_", ref.disney_sample_eval_pdf(sg, abi.disney_params(base_color=base, **kw), *u)))
    scenes = dict(default=dict(roughness=0.5, specular=0.5), subsurface=dict(roughness=0.5, specular=0.5, subsurface=1.0),
                  metallic=dict(metallic=1.0, roughness=0.3), specular=dict(specular=1.0, roughness=0.5),
                  aniso=dict(metallic=1.0, roughness=0.2, anisotropic=1.0),
                  clearcoat=dict(roughness=0.6, clearcoat=1.0, clearcoat_gloss=0.8, sheen=0.5, sheen_tint=0.5))
    for name, params in scenes.items():
        p = abi.disney_params(base_color=(0.8, 0.4, 0.2), **params)
        out.update(flat(f"{name}_", ref.disney_sample_eval_pdf(sg, p, *u)))
    np.savez_compressed(os.path.join(HERE, "disney.npz"), **out)

    # ---- rlSkin / NDProfile: varying + the 0009 fixture (colour 1,0.84235,0.5; dist 1,1,1)
    rx = ol.hash_uniform(N, 104, 0)
    color = tuple(ol.hash_uniform(N, 104, 40 + j, lo=0.05, hi=1.0) for j in range(3))
    dist = tuple(ol.hash_uniform(N, 104, 50 + j, lo=0.05, hi=2.0) for j in range(3))
    out = dict(rx=rx, color_r=color[0], color_g=color[1], color_b=color[2], dist_x=dist[0], dist_y=dist[1], dist_z=dist[2])
    out.update(flat("out_", ref.skin_profile(abi.skin_params(sss_color=color, sss_scatter_dist=dist), rx)))
    out.update(flat("scene0009_", ref.skin_profile(abi.skin_params(sss_color=(1.0, 0.84235, 0.5),
                                                                    sss_scatter_dist=(1.0, 1.0, 1.0)), rx)))
    sgp = ol.make_shading(N, 105)
    ryp = ol.hash_uniform(N, 105, 1)
    out.update(flat("probe_sg_", sgp))
    out["probe_ry"] = ryp
    out.update(flat("probe_", ref.skin_probe_ray(sgp, abi.skin_params(sss_color=color, sss_scatter_dist=dist), rx, ryp)))
    np.savez_compressed(os.path.join(HERE, "skin.npz"), **out)

    gaussian(ref)

    # ---- albedo sweep on a reduced grid
    g = abi.SweepGrid(4, 4, 2, 0.05, 1.0, 1.0, 2.0)
    np.savez_compressed(os.path.join(HERE, "sweep.npz"), table=ref.albedo_sweep(g, 0x5EED0005, 0, 512),
                        grid=np.array([4, 4, 2], dtype=np.int32), ranges=np.array([0.05, 1.0, 1.0, 2.0], dtype=np.float32),
                        seed=np.uint64(0x5EED0005), spp=np.int32(512))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
