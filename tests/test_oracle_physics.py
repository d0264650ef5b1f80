"""CPU: known-answer / physics checks the reference implies but never tests
(SURVEY.md 4 "Consequence for the build"), run against the oracle."""
import numpy as np
import pytest

import oracle_lib as ol
from rlshaders_b200 import _abi as abi

f32 = np.float32


@pytest.fixture(scope="module")
def orc():
    return ol.load_ref() or ol.load_port()


def canonical_shading(n, cosv):
    z, o = np.zeros(n, f32), np.ones(n, f32)
    cosv = np.full(n, cosv, f32) if np.isscalar(cosv) else cosv.astype(f32)
    sinv = np.sqrt(f32(1) - cosv * cosv).astype(f32)
    return dict(Ux=o, Uy=z, Uz=z, Vx=z, Vy=o, Vz=z, Nx=z, Ny=z, Nz=o, wox=sinv, woy=z.copy(), woz=cosv, backfacing=None)


def test_fresnel_normal_incidence_is_f0(orc):
    # fresnel(normal incidence, ior 1.5) = ((1.5-1)/(1.5+1))^2 = 0.04 (src/rlGgx.h:249-270)
    n = 16
    sg = canonical_shading(n, 1.0)
    wi = np.stack([np.zeros(n, f32), np.zeros(n, f32), np.ones(n, f32)])
    f = orc.ggx_eval_brdf(sg, abi.ggx_params(ior=1.5, specularRoughness=1.0), wi)
    # f = F * G * D / 4 with G = 1, D(n) = 1/(pi a^2) = 1/pi at alpha = 1
    assert np.allclose(f[0], 0.04 / (4 * np.pi), rtol=1e-6)


def test_pdf_floor_and_no_zero_guard(orc):
    # rlGgx evalPdf is floored at AI_EPSILON and has no zero-L guard (src/rlGgx.h:79,121-127)
    n = 8
    sg = canonical_shading(n, 0.5)
    wi = np.zeros((3, n), f32)
    p = abi.ggx_params(ior=1.5, specularRoughness=0.3)
    assert np.all(orc.ggx_eval_pdf(sg, p, wi) >= f32(1e-4))
    assert np.all(orc.ggx_eval_brdf(sg, p, wi) == 0)             # black on zero indir (:112)
    # rlDisney: zero indir -> black and pdf 0 (src/rlDisney.cpp:124-127,141-144)
    dp = abi.disney_params(roughness=0.5)
    for t in (abi.RLS_RAY_DIFFUSE, abi.RLS_RAY_GLOSSY):
        assert np.all(orc.disney_eval_pdf(sg, dp, t, wi) == 0)
        assert np.all(orc.disney_eval_brdf(sg, dp, t, wi) == 0)


def test_vndf_samples_integrate_to_fresnel_weighted_albedo(orc):
    # E[f/pdf] over visible-normal samples stays in (0, 1] for a white conductor-like GGX lobe
    n = 1 << 16
    sg = canonical_shading(n, 0.7)
    rx, ry = ol.hash_uniform(n, 1, 0), ol.hash_uniform(n, 1, 1)
    o = orc.ggx_sample_eval_pdf(sg, abi.ggx_params(ior=0.47, specularRoughness=0.3), rx, ry)
    ok = (o["flags"] & (abi.FLAG_BELOW_HORIZON | abi.FLAG_ZERO_L)) == 0
    w = o["f"][0][ok] / o["pdf"][ok]
    assert 0.5 < w.mean() <= 1.0 + 1e-3
    # unit length except where the sampled normal faces away from wo: reflectDirection's ABS
    # (src/rlUtil.h:33) then no longer reflects -- a reference quirk the kernels reproduce
    assert np.mean(np.abs(np.linalg.norm(o["wi"], axis=0) - 1.0) < 1e-5) > 0.995


def test_disney_lobe_selection_boundary(orc):
    # GTR2 vs GTR1 by rx < 1/(clearcoat*0.25 + 1) (src/rlDisney.cpp:373-375)
    n = 1 << 14
    sg = canonical_shading(n, 0.8)
    rx, ry = ol.hash_uniform(n, 2, 0), ol.hash_uniform(n, 2, 1)
    o = orc.disney_eval_sample(sg, abi.disney_params(roughness=0.4, clearcoat=1.0), abi.RLS_RAY_GLOSSY, rx, ry)
    lobe = (o["flags"] & abi.FLAG_LOBE_MASK) >> abi.FLAG_LOBE_SHIFT
    assert np.array_equal(lobe, (rx >= f32(1.0) / (f32(1.0) * f32(0.25) + f32(1.0))).astype(np.uint32))
    assert 0.15 < lobe.mean() < 0.25
    o0 = orc.disney_eval_sample(sg, abi.disney_params(roughness=0.4, clearcoat=0.0), abi.RLS_RAY_GLOSSY, rx, ry)
    assert not np.any(o0["flags"] & abi.FLAG_LOBE_MASK)


def test_disney_diffuse_is_cosine_weighted(orc):
    n = 1 << 16
    sg = canonical_shading(n, 0.6)
    rx, ry = ol.hash_uniform(n, 3, 0), ol.hash_uniform(n, 3, 1)
    o = orc.disney_eval_sample(sg, abi.disney_params(roughness=0.5), abi.RLS_RAY_DIFFUSE, rx, ry)
    assert abs(o["wi"][2].mean() - 2.0 / 3.0) < 5e-3          # E[cos] = 2/3
    pdf = orc.disney_eval_pdf(sg, abi.disney_params(roughness=0.5), abi.RLS_RAY_DIFFUSE, o["wi"])
    assert np.allclose(pdf, np.maximum(1e-4, o["wi"][2] / np.pi), rtol=1e-6)


def test_ndprofile_normalisation(orc):
    # int evalProfile 2 pi r dr = 1 per channel over [0, inf) and ~0.96.. over r < maxRadius;
    # int getPdf dA = 1 over the disc r < maxRadius (src/rlSss.cpp:42-43,68-84)
    n = 1 << 16
    d = np.array([0.3, 0.7, 1.3], f32)
    dist = np.repeat(d[:, None], n, axis=1)
    prof = orc.ndprofile_set_distance(dist, np.ones((3, n), f32))
    R = float(prof["max_radius"][0])
    assert R == pytest.approx(3.0 * 1.3, rel=1e-6)
    r = ((np.arange(n, dtype=np.float64) + 0.5) / n * R).astype(f32)
    dr = R / n
    pdf = orc.ndprofile_get_pdf(prof, r).astype(np.float64)
    assert np.sum(pdf * 2 * np.pi * r * dr) == pytest.approx(1.0, abs=2e-4)
    rd = orc.ndprofile_eval_profile(prof, r).astype(np.float64)
    for ch in range(3):
        full = 1.0 - 0.25 * np.exp(-R / d[ch]) - 0.75 * np.exp(-R / (3 * d[ch]))
        assert np.sum(rd[ch] * 2 * np.pi * r * dr) == pytest.approx(full, abs=3e-3)


def test_gaussian_profile_normalisation_and_inverse_cdf(orc):
    # GaussianProfile (src/rlSss.h:63-97): getPdf integrates to 1 over the disc r < maxRadius,
    # getRadius inverts its radial CDF, getRadius(1) = maxRadius, variance = R^2 / 12.46.
    n = 1 << 16
    R = 1.7
    dist = np.stack([np.full(n, R, f32), np.zeros(n, f32), np.zeros(n, f32)])
    prof = orc.gaussprofile_set_distance(dist, np.ones((3, n), f32))
    assert float(prof["max_radius"][0]) == np.float32(R)
    v = float(prof["variance"][0])
    assert v == pytest.approx(R * R / 12.46, rel=1e-6)
    norm = float(prof["norm"][0])
    assert norm == pytest.approx(1.0 - np.exp(-12.46 / 2.0), rel=1e-5)
    r = ((np.arange(n, dtype=np.float64) + 0.5) / n * R).astype(f32)
    pdf = orc.gaussprofile_get_pdf(prof, r).astype(np.float64)
    assert np.sum(pdf * 2 * np.pi * r * (R / n)) == pytest.approx(1.0, abs=2e-4)
    rd = orc.gaussprofile_eval_profile(prof, r).astype(np.float64)
    assert np.allclose(rd, pdf * norm, rtol=1e-6)
    rx = ((np.arange(n, dtype=np.float64) + 0.5) / n).astype(f32)
    rad = orc.gaussprofile_get_radius(prof, rx).astype(np.float64)
    cdf = (1.0 - np.exp(-rad * rad / (2 * v))) / norm
    assert np.allclose(cdf, rx, atol=2e-5)
    assert np.all(np.diff(rad) >= 0) and rad.max() <= R * (1 + 1e-5)
    assert float(orc.gaussprofile_get_radius(prof, np.ones(n, f32))[0]) == pytest.approx(R, rel=1e-4)


def test_ndprofile_radius_limits_and_thirds(orc):
    n = 6
    dist = np.ones((3, n), f32)
    prof = orc.ndprofile_set_distance(dist, np.ones((3, n), f32))
    rx = np.array([2.0 ** -24, 0.3332, 0.3334, 0.6665, 0.6667, 1 - 2.0 ** -24], f32)
    o = orc.ndprofile_get_radius(prof, rx)
    ch = (o["flags"] & abi.FLAG_LOBE_MASK) >> abi.FLAG_LOBE_SHIFT
    assert list(ch) == [0, 0, 1, 1, 2, 2]                      # thirds at 0.3333f / 0.6666f (src/rlSss.h:32-40)
    assert o["r"][0] < 1e-6 and o["r"][-1] == pytest.approx(3.0, rel=1e-4)   # rx -> 1 gives maxRadius
    assert np.all(np.isfinite(o["r"])) and np.all(o["r"] >= 0)
    # r < eps returns white, degenerate profile returns black / pdf 1 (src/rlSss.cpp:70-72,88-93)
    assert np.all(orc.ndprofile_eval_profile(prof, np.full(n, 5e-5, f32)) == 1.0)
    zero = orc.ndprofile_set_distance(np.zeros((3, n), f32), np.ones((3, n), f32))
    with np.errstate(all="ignore"):
        assert np.all(orc.ndprofile_eval_profile(zero, np.ones(n, f32)) == 0.0)
        assert np.all(orc.ndprofile_get_pdf(zero, np.ones(n, f32)) == 1.0)
        assert np.all(orc.ndprofile_get_radius(zero, rx)["r"] == 0.0)


def test_dielectric_tir_and_entering_flags(orc):
    n = 1 << 14
    sg = ol.make_shading(n, 11, backfacing_fraction=0.5)
    rx, ry = ol.hash_uniform(n, 11, 0), ol.hash_uniform(n, 11, 1)
    o = orc.ggx_dielectric(sg, abi.ggx_params(ior=1.5, specularRoughness=0.2), rx, ry)
    entering = (o["flags"] & abi.FLAG_ENTERING) != 0
    assert np.array_equal(entering, sg["backfacing"] == 0)      # wo is in the Nf hemisphere
    tir = (o["flags"] & abi.FLAG_TIR) != 0
    assert tir.any() and np.all(o["f_t"][tir] == 0)
    assert np.all((o["fresnel"] >= 0) & (o["fresnel"] <= 1))


def test_skin_layer_weights(orc):
    n = 4
    sp = abi.skin_params(sheen_weight=0.5, specular_weight=0.6, sss_weight=1.0)
    o = orc.skin_layer_weights(sp, np.full(n, 0.2, f32), np.full(n, 0.1, f32))
    sheenF, specF = f32(0.2) * f32(0.5), f32(0.1) * f32(0.6)
    assert np.all(o["specular_scale"] == f32(0.6) * (f32(1) - sheenF))
    assert np.all(o["sss_weight"] == f32(1) * (f32(1) - specF * (f32(1) - sheenF)))
    o0 = orc.skin_layer_weights(abi.skin_params(sheen_weight=0.0, specular_weight=5e-5), np.ones(n, f32), np.ones(n, f32))
    assert np.all(o0["sss_weight"] == 1.0)                      # weights <= eps skip the layer (:191,214)
