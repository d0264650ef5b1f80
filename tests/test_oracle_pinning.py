"""CPU: pins the C port of the oracle (oracle/rls_oracle.c) against
  (1) the golden vectors generated from the reference's own sources (tests/golden/), and
  (2) the reference library itself, bit for bit, where it is built (oracle/_ref/).
"""
import numpy as np
import pytest

import golden_io as gio
import oracle_lib as ol
from rlshaders_b200 import _abi as abi


@pytest.fixture(scope="module")
def port():
    return ol.load_port()


@pytest.fixture(scope="module")
def ref():
    r = ol.load_ref()
    if r is None:
        pytest.skip("reference library not built (no /root/reference here)")
    return r


def assert_same(a, b, what):
    for k in a:
        assert gio.bits_equal(a[k], b[k]), f"{what}: {k} differs"


def test_port_vs_golden_ggx_fixtures(port):
    g = gio.load("ggx_fixtures")
    sg = gio.shading(g)
    for name in ("teflon", "gold", "anisotropic", "gold_bench"):
        rough, ior, aniso = [float(x) for x in g[name + "_params"]]
        got = port.ggx_sample_eval_pdf(sg, abi.ggx_params(specularRoughness=rough, ior=ior, anisotropic=aniso),
                                       g["rx"], g["ry"])
        want = {k: g[f"{name}_{k}"] for k in got}
        assert_same(want, got, name)


def test_port_vs_golden_dielectric(port):
    g = gio.load("ggx_dielectric")
    kw = gio.group(g, "p_")
    got = port.ggx_dielectric(gio.shading(g), abi.ggx_params(**kw), g["rx"], g["ry"])
    assert_same(gio.group(g, "out_"), got, "dielectric")


def test_port_vs_golden_disney(port):
    g = gio.load("disney")
    sg = gio.shading(g)
    u = [g[f"u{j}"] for j in range(4)]
    kw = gio.group(g, "p_")
    kw["base_color"] = (g["base_r"], g["base_g"], g["base_b"])
    assert_same(gio.group(g, "out_"), port.disney_sample_eval_pdf(sg, abi.disney_params(**kw), *u), "disney")
    scenes = dict(default=dict(roughness=0.5, specular=0.5), subsurface=dict(roughness=0.5, specular=0.5, subsurface=1.0),
                  metallic=dict(metallic=1.0, roughness=0.3), specular=dict(specular=1.0, roughness=0.5),
                  aniso=dict(metallic=1.0, roughness=0.2, anisotropic=1.0),
                  clearcoat=dict(roughness=0.6, clearcoat=1.0, clearcoat_gloss=0.8, sheen=0.5, sheen_tint=0.5))
    for name, params in scenes.items():
        got = port.disney_sample_eval_pdf(sg, abi.disney_params(base_color=(0.8, 0.4, 0.2), **params), *u)
        assert_same(gio.group(g, name + "_"), got, name)


def test_port_vs_golden_skin(port):
    g = gio.load("skin")
    color = (g["color_r"], g["color_g"], g["color_b"])
    dist = (g["dist_x"], g["dist_y"], g["dist_z"])
    assert_same(gio.group(g, "out_"), port.skin_profile(abi.skin_params(sss_color=color, sss_scatter_dist=dist), g["rx"]), "skin")
    got = port.skin_profile(abi.skin_params(sss_color=(1.0, 0.84235, 0.5), sss_scatter_dist=(1.0, 1.0, 1.0)), g["rx"])
    assert_same(gio.group(g, "scene0009_"), got, "scene 0009")
    sgp = gio.shading(g, "probe_sg_")
    got = port.skin_probe_ray(sgp, abi.skin_params(sss_color=color, sss_scatter_dist=dist), g["rx"], g["probe_ry"])
    want = {k: g["probe_" + k] for k in ("r", "origin", "dir", "maxdist", "flags")}
    assert_same(want, got, "probe ray")


def test_port_vs_golden_gaussian_profile(port):
    g = gio.load("gaussian_profile")
    with np.errstate(all="ignore"):
        assert_same(gio.group(g, "out_"), port.gaussprofile(g["dist_x"], g["rx"]), "gaussian fused")
        dist = np.stack([g["dist_x"], np.zeros_like(g["dist_x"]), np.zeros_like(g["dist_x"])])
        prof = port.gaussprofile_set_distance(dist, np.ones_like(dist))
        assert_same(gio.group(g, "state_"), prof, "gaussian setDistance")
        assert gio.bits_equal(g["pdf_at_r"], port.gaussprofile_get_pdf(prof, g["r"]))
        assert gio.bits_equal(g["rd_at_r"], port.gaussprofile_eval_profile(prof, g["r"]))
        assert gio.bits_equal(g["radius_at_rx"], port.gaussprofile_get_radius(prof, g["rx"]))


def test_port_vs_golden_sweep(port):
    g = gio.load("sweep")
    n = [int(x) for x in g["grid"]]
    r = [float(x) for x in g["ranges"]]
    grid = abi.SweepGrid(n[0], n[1], n[2], r[0], r[1], r[2], r[3])
    got = port.albedo_sweep(grid, int(g["seed"]), 0, int(g["spp"]))
    assert np.array_equal(got, g["table"])


N = 1 << 16


def test_port_vs_reference_ggx(port, ref):
    sg, p, rx, ry = ol.workload_ggx_conductor(N)
    a = ref.ggx_sample_eval_pdf(sg, p, rx, ry)
    assert_same(a, port.ggx_sample_eval_pdf(sg, p, rx, ry), "ggx fused")
    sg, p, rx, ry = ol.workload_ggx_dielectric(N, aniso=True)
    d = ref.ggx_dielectric(sg, p, rx, ry)
    assert_same(d, port.ggx_dielectric(sg, p, rx, ry), "ggx dielectric")
    assert_same(ref.ggx_eval_sample(sg, p, rx, ry), port.ggx_eval_sample(sg, p, rx, ry), "ggx evalSample")
    assert gio.bits_equal(ref.ggx_eval_brdf(sg, p, d["wi_r"]), port.ggx_eval_brdf(sg, p, d["wi_r"]))
    assert gio.bits_equal(ref.ggx_eval_pdf(sg, p, d["wi_t"]), port.ggx_eval_pdf(sg, p, d["wi_t"]))


def test_port_vs_reference_refraction_half(port, ref):
    """getRefractDirection / refraction / getSampleWeight at given directions (src/rlGgx.h:277-328): the port equals
    the reference's own members bit for bit, including TIR, back-facing samples and a microfacet normal that is not
    the one the sampler drew."""
    n = 1 << 16
    sg = ol.make_shading(n, 0x5EED0002, backfacing_fraction=0.25)
    kw = dict(specularRoughness=ol.hash_uniform(n, 0x5EED0002, 2, lo=0.05, hi=1.0), ior=ol.hash_uniform(n, 0x5EED0002, 3, lo=1.05, hi=2.5),
              anisotropic=ol.hash_uniform(n, 0x5EED0002, 4))
    rx, ry = ol.hash_uniform(n, 0x5EED0002, 0), ol.hash_uniform(n, 0x5EED0002, 1)
    p = abi.ggx_params(**kw)
    d = ref.ggx_dielectric(sg, p, rx, ry)
    wo = np.stack([sg["wo" + c] for c in "xyz"])
    m = d["wi_r"] + wo                                 # reflection about m: m ~ (wi_r + wo) / |.|
    m = (m / np.linalg.norm(m, axis=0)).astype(np.float32)
    m[:, :64] = np.stack([sg["N" + c] for c in "xyz"])[:, :64]
    for o in (port, ref):
        o.set_flag_probe(True)
    wr, fr = ref.ggx_refract_direction(sg, p, m)
    wp, fp = port.ggx_refract_direction(sg, p, m)
    assert gio.bits_equal(wr, wp) and np.array_equal(fr, fp)
    assert 0 < int((fr & abi.FLAG_TIR != 0).sum()) < n
    wt = d["wi_t"]
    assert gio.bits_equal(ref.ggx_eval_btdf(sg, p, wt), port.ggx_eval_btdf(sg, p, wt))
    assert gio.bits_equal(ref.ggx_sample_weight(sg, p, wt, m), port.ggx_sample_weight(sg, p, wt, m))


def test_port_vs_reference_disney(port, ref):
    sg, p, u = ol.workload_disney(N)
    a = ref.disney_sample_eval_pdf(sg, p, *u)
    assert_same(a, port.disney_sample_eval_pdf(sg, p, *u), "disney fused")
    for t in (abi.RLS_RAY_DIFFUSE, abi.RLS_RAY_GLOSSY):
        assert_same(ref.disney_eval_sample(sg, p, t, u[0], u[1]), port.disney_eval_sample(sg, p, t, u[0], u[1]), "evalSample")
        assert gio.bits_equal(ref.disney_eval_brdf(sg, p, t, a["wi_s"]), port.disney_eval_brdf(sg, p, t, a["wi_s"]))
        assert gio.bits_equal(ref.disney_eval_pdf(sg, p, t, a["wi_d"]), port.disney_eval_pdf(sg, p, t, a["wi_d"]))


def test_port_vs_reference_profile(port, ref):
    p, rx = ol.workload_skin(N)
    assert_same(ref.skin_profile(p, rx), port.skin_profile(p, rx), "skin fused")
    dist = np.stack([ol.hash_uniform(N, 9, j, lo=0.0, hi=2.0) for j in range(3)])
    dist[:, :64] = 0.0            # degenerate profiles: maxRadius < eps
    dist[0, 64:128] = 5e-5        # one channel below eps
    albedo = np.stack([ol.hash_uniform(N, 9, 3 + j) for j in range(3)])
    pa, pb = ref.ndprofile_set_distance(dist, albedo), port.ndprofile_set_distance(dist, albedo)
    assert_same(pa, pb, "setDistance")
    ra, rb = ref.ndprofile_get_radius(pa, rx), port.ndprofile_get_radius(pb, rx)
    assert_same(ra, rb, "getRadius")
    with np.errstate(all="ignore"):
        assert gio.bits_equal(ref.ndprofile_get_pdf(pa, ra["r"]), port.ndprofile_get_pdf(pb, rb["r"]))
        assert gio.bits_equal(ref.ndprofile_eval_profile(pa, ra["r"]), port.ndprofile_eval_profile(pb, rb["r"]))
    sg = ol.make_shading(N, 77)
    ry = ol.hash_uniform(N, 5, 9)
    assert_same(ref.skin_probe_ray(sg, p, rx, ry), port.skin_probe_ray(sg, p, rx, ry), "probe ray")
    f1, f2 = ol.hash_uniform(N, 6, 1), ol.hash_uniform(N, 6, 2)
    sp = abi.skin_params(sheen_weight=ol.hash_uniform(N, 6, 3), specular_weight=ol.hash_uniform(N, 6, 4),
                         sss_weight=ol.hash_uniform(N, 6, 5))
    assert_same(ref.skin_layer_weights(sp, f1, f2), port.skin_layer_weights(sp, f1, f2), "layer weights")


def test_port_vs_reference_gaussian_profile(port, ref):
    """GaussianProfile (src/rlSss.h:63-97): the reference class itself vs the C restatement, every method."""
    rx = ol.hash_uniform(N, 11, 0)
    dist = np.stack([ol.hash_uniform(N, 11, 1 + j, lo=0.0, hi=3.0) for j in range(3)])
    dist[0, :64] = 0.0            # variance 0: NaN / Inf exactly as the reference writes them
    dist[0, 64:128] = 1e-30       # variance underflows to 0
    dist[0, 128:192] = 1e20       # variance overflows
    albedo = np.ones_like(dist)
    with np.errstate(all="ignore"):
        pa, pb = ref.gaussprofile_set_distance(dist, albedo), port.gaussprofile_set_distance(dist, albedo)
        assert_same(pa, pb, "gaussian setDistance")
        ra, rb = ref.gaussprofile_get_radius(pa, rx), port.gaussprofile_get_radius(pb, rx)
        assert gio.bits_equal(ra, rb)
        r = ol.hash_uniform(N, 11, 5, lo=0.0, hi=4.0)
        assert gio.bits_equal(ref.gaussprofile_get_pdf(pa, r), port.gaussprofile_get_pdf(pb, r))
        assert gio.bits_equal(ref.gaussprofile_eval_profile(pa, r), port.gaussprofile_eval_profile(pb, r))
        assert_same(ref.gaussprofile(dist[0], rx), port.gaussprofile(dist[0], rx), "gaussian fused")


def test_port_vs_reference_sweep(port, ref):
    g = abi.SweepGrid(3, 5, 2, 0.02, 1.0, 1.0, 2.5)
    assert np.array_equal(ref.albedo_sweep(g, 123, 0, 128), port.albedo_sweep(g, 123, 0, 128))


def test_port_vs_reference_sampler_variants_and_mis(port, ref):
    """NDFKernel as the GgxSamplerT argument (src/rlGgx.h:24-56), DisneySampler with
    mSampleFromVisibleNormal = false (src/rlDisney.cpp:377-379,541-542), probe-hit MIS pdf
    (src/rlSss.h:252-263)."""
    n = N
    sg = ol.make_shading(n, 3, backfacing_fraction=0.25)
    rx, ry = ol.hash_uniform(n, 3, 0), ol.hash_uniform(n, 3, 1)
    kw = dict(specularRoughness=ol.hash_uniform(n, 3, 2, lo=0.05, hi=1.0), ior=ol.hash_uniform(n, 3, 3, lo=1.05, hi=2.5),
              anisotropic=ol.hash_uniform(n, 3, 4), KsColor=tuple(ol.hash_uniform(n, 3, 60 + j) for j in range(3)),
              normal_sampler=abi.GGX_SAMPLER_NDF)
    p = abi.ggx_params(**kw)
    a = ref.ggx_sample_eval_pdf(sg, p, rx, ry)
    assert_same(a, port.ggx_sample_eval_pdf(sg, p, rx, ry), "ndf fused")
    assert_same(ref.ggx_dielectric(sg, p, rx, ry), port.ggx_dielectric(sg, p, rx, ry), "ndf dielectric")
    assert gio.bits_equal(ref.ggx_eval_pdf(sg, p, a["wi"]), port.ggx_eval_pdf(sg, p, a["wi"]))
    # the NDF pdf is not floored: it differs from the VNDF pdf on the same directions
    vndf = port.ggx_eval_pdf(sg, abi.ggx_params(**dict(kw, normal_sampler=abi.GGX_SAMPLER_VNDF)), a["wi"])
    assert not gio.bits_equal(vndf, port.ggx_eval_pdf(sg, p, a["wi"]))

    sgd, _, u = ol.workload_disney(n)
    names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic",
             "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss"]
    kwd = {nm: ol.hash_uniform(n, 0x5EED0003, 20 + j) for j, nm in enumerate(names)}
    kwd["base_color"] = tuple(ol.hash_uniform(n, 0x5EED0003, 30 + j) for j in range(3))
    pd = abi.disney_params(sample_from_visible_normal=0, **kwd)
    b = ref.disney_sample_eval_pdf(sgd, pd, *u)
    assert_same(b, port.disney_sample_eval_pdf(sgd, pd, *u), "disney non-visible-normal")
    assert gio.bits_equal(ref.disney_eval_pdf(sgd, pd, abi.RLS_RAY_GLOSSY, b["wi_s"]),
                          port.disney_eval_pdf(sgd, pd, abi.RLS_RAY_GLOSSY, b["wi_s"]))

    sp, _ = ol.workload_skin(n)
    disp = np.stack([ol.hash_uniform(n, 8, j, lo=-1.5, hi=1.5) for j in range(3)])
    hn = ol.make_shading(n, 9)
    hnv = np.stack([hn["Nx"], hn["Ny"], hn["Nz"]])
    assert gio.bits_equal(ref.skin_probe_mis_pdf(sg, sp, disp, hnv), port.skin_probe_mis_pdf(sg, sp, disp, hnv))


# ------------------------------------------- callers of the triple
def test_port_vs_reference_skin_glossy_layers(port, ref):
    import parity
    n, k = 1 << 12, 9
    for with_li in (True, False):
        sg, kw, uu, li = parity.skin_layers_inputs(n, k, with_radiance=with_li)
        p = abi.skin_params(**kw)
        a = port.skin_glossy_layers(sg, p, k, *uu, li[0], li[1])
        b = ref.skin_glossy_layers(sg, p, k, *uu, li[0], li[1])
        assert_same(b, a, f"skin glossy layers (radiance={with_li})")
        assert (a["flags"] & abi.SKIN_SHEEN_EVALUATED).any() and not (a["flags"] & abi.SKIN_SHEEN_EVALUATED).all()
    # a skipped layer leaves its Fresnel term at 0 and the estimate black (src/rlSkin.cpp:191,207)
    off = (a["flags"] & abi.SKIN_SHEEN_EVALUATED) == 0
    assert not a["sheen_fresnel"][off].any() and not a["sheen"][:, off].any()


def test_port_vs_reference_light_sample(port, ref):
    import parity
    n = 1 << 14
    sg, kw, rx, ry = parity.ggx_dielectric_inputs(n, 0x5EED00F3, aniso=True)
    Ld, Li, pl, Lib, plb = parity.light_inputs(n, sg, 0x5EED00F3)
    p = abi.ggx_params(KsColor=(0.9, 0.6, 0.3), **kw)
    for half in (False, True):
        extra = (rx, ry, Lib, plb) if half else ()
        a, b = port.ggx_light_sample(sg, p, Ld, Li, pl, *extra), ref.ggx_light_sample(sg, p, Ld, Li, pl, *extra)
        assert_same(b, a, f"ggx light sample (brdf half={half})")
    assert (a["w_light"] == 0).any() and (a["w_brdf"] == 1).any() and (a["w_light"] > 0).any()
    dsg, dkw, du = parity.disney_inputs(n, 0x5EED00F3)
    dp = abi.disney_params(**dkw)
    Ld, Li, pl, Lib, plb = parity.light_inputs(n, dsg, 0x5EED00F4)
    for st, (ux, uy) in ((abi.RLS_RAY_GLOSSY, du[:2]), (abi.RLS_RAY_DIFFUSE, du[2:])):
        a = port.disney_light_sample(dsg, dp, st, Ld, Li, pl, ux, uy, Lib, plb)
        b = ref.disney_light_sample(dsg, dp, st, Ld, Li, pl, ux, uy, Lib, plb)
        assert_same(b, a, f"disney light sample type {st:#x}")


def test_port_vs_reference_sample_writer(port, ref):
    import parity
    n = 64
    sg, kw, rx, ry = parity.ggx_dielectric_inputs(n, 0x5EED00F5)
    # the writer's own use (src/rlGgx.cpp:202-224): identity frame, view at 45 degrees
    for name, val in (("U", (1, 0, 0)), ("V", (0, 1, 0)), ("N", (0, 0, 1)), ("wo", (np.sqrt(0.5), 0, np.sqrt(0.5)))):
        for c, v in zip("xyz", val):
            sg[name + c][0] = np.float32(v)
    sx, sy = ol.hash_uniform(4096, 7, 0), ol.hash_uniform(4096, 7, 1)
    p = abi.ggx_params(specularRoughness=0.35, ior=1.5)
    for point in (0, 5):
        a, ma = port.sample_writer(abi.NODE_GGX, sg, p, point, 0, 96, 48, sx, sy)
        b, mb = ref.sample_writer(abi.NODE_GGX, sg, p, point, 0, 96, 48, sx, sy)
        assert gio.bits_equal(a, b) and ma == mb
    assert (a[1] == 1.0).any()                          # green scatter pixels (plane order B, G, R)
    dsg, dkw, _ = parity.disney_inputs(n, 0x5EED00F6)
    dp = abi.disney_params(**{k: (float(v[3]) if not isinstance(v, tuple) else tuple(float(t[3]) for t in v)) for k, v in dkw.items()})
    for st in (abi.RLS_RAY_GLOSSY, abi.RLS_RAY_DIFFUSE):
        a, ma = port.sample_writer(abi.NODE_DISNEY, dsg, dp, 3, st, 64, 32, sx, sy)
        b, mb = ref.sample_writer(abi.NODE_DISNEY, dsg, dp, 3, st, 64, 32, sx, sy)
        assert gio.bits_equal(a, b) and ma == mb


def test_exr_writer_round_trip(tmp_path):
    from rlshaders_b200 import exr
    rng = np.random.default_rng(5)
    img = rng.random((3, 17, 33), dtype=np.float32) * 4.0
    for half in (True, False):
        path = exr.write_scanline_exr(str(tmp_path / f"w{int(half)}.exr"), img, ("B", "G", "R"), half=half)
        planes, attrs = exr.read_scanline_exr(path)
        assert sorted(planes) == ["B", "G", "R"] and attrs["compression"][1] == b"\0"
        for j, nm in enumerate(("B", "G", "R")):
            want = img[j].astype(np.float16) if half else img[j]
            assert planes[nm].dtype == want.dtype and np.array_equal(planes[nm], want)
    assert open(path, "rb").read(4) == bytes([0x76, 0x2f, 0x31, 0x01])       # OpenEXR magic
