// rls_callers.cuh -- the callers immediately above the BRDF triple:
//   f2  rlSkin's glossy layers with per-shading-point average Fresnel (src/rlSkin.cpp:184-238)
//   f3  one MIS light sample, the shape of AiEvaluateLightSample (src/rlGgx.h:167-170)
//   f4  the SampleWriter lat-long dumps (src/rlUtil.h:98-156)
// Arnold's integrators are proprietary; what they compute is DEFINED in include/rls_b200.h and
// restated identically by both oracles.  The BRDF triples themselves are the reference's.
#pragma once
#include "rls_fused.cuh"

namespace rls {

// ----------------------------------------------------------------------------- f2
struct SkinLayerDev { P3 color; P1 weight, roughness, ior; };
struct SkinLayersDev { SkinLayerDev sheen, spec; P1 sss_weight; };
struct SkinLayers1 { f3 sheen, spec; float sheenF, specF, sssW; uint32_t flags; };

// One layer: GgxSampler(sg, color, ior, roughness) + integrateGlossy's K triples
// (src/rlGgx.h:172-179) + getAvgReflectWeight (:181-184).  Samples are read sample-major.
template <class Fp>
RLS_DEV f3 skin_layer(Fp &fp, const Shading &sh, f3 color, float ior, float rough, uint32_t K, size_t P, uint32_t p,
                      const float *rx, const float *ry, const CV3 &li, float &avgF)
{
    Ggx g;
    ggx_init(fp, g, sh, color, ior, rough, 0.0f);
    float reflectWeight = 0.0f, count = 0.0f;          // mReflectWeight, mMisSampleCount (both float)
    f3 acc = mk3(0.0f, 0.0f, 0.0f);
    // integrateGlossy returns black without sampling when the colour is small (:174-176)
    const bool small = abs_m(color.x) < kEps && abs_m(color.y) < kEps && abs_m(color.z) < kEps;
    if (!small) {
        for (uint32_t k = 0; k < K; k++) {
            const size_t idx = (size_t)k * P + p;
            GgxBsdf o = ggx_unit(fp, g, __ldg(rx + idx), __ldg(ry + idx));
            reflectWeight += o.fresnel;                // src/rlGgx.h:103
            count += 1.0f;                             // :104
            f3 w = mk3(fp.div_pz(o.f.x, o.pdf), fp.div_pz(o.f.y, o.pdf), fp.div_pz(o.f.z, o.pdf));   // pdf >= 1e-4
            if (li.x) w = mk3(w.x * __ldg(li.x + idx), w.y * __ldg(li.y + idx), w.z * __ldg(li.z + idx));
            acc = acc + w;
        }
    }
    avgF = count > 0.0f ? fp.div_pz(reflectWeight, count) : 1.0f;    // :181-184 (the sum is 0 for ior 1)
    const float invK = 1.0f / (float)K;                // K >= 1: exact policy not needed (uniform value)
    return acc * invK;
}

template <class Fp>
RLS_DEV SkinLayers1 skin_layers_unit(Fp &fp, const Shading &sh, const SkinLayersDev &sp, uint32_t K, size_t P, uint32_t p,
                                     const float *rx_a, const float *ry_a, const float *rx_b, const float *ry_b,
                                     const CV3 &li_a, const CV3 &li_b)
{
    SkinLayers1 o;
    o.flags = 0;
    float sheenFresnel = 0.0f, specularFresnel = 0.0f;
    f3 sheen = mk3(0.0f, 0.0f, 0.0f), specular = mk3(0.0f, 0.0f, 0.0f);
    const float sheenWeight = fetch(sp.sheen.weight, p);
    if (sheenWeight > kEps) {                          // src/rlSkin.cpp:191
        float avg;
        sheen = skin_layer(fp, sh, fetch(sp.sheen.color, p), fetch(sp.sheen.ior, p), fetch(sp.sheen.roughness, p),
                           K, P, p, rx_a, ry_a, li_a, avg);
        sheenFresnel = avg * sheenWeight;              // :204
        o.flags |= 0x1u;
    }
    sheen = sheen * sheenWeight;                       // :207
    const float specularWeight = fetch(sp.spec.weight, p);
    if (specularWeight > kEps) {                       // :214
        float avg;
        specular = skin_layer(fp, sh, fetch(sp.spec.color, p), fetch(sp.spec.ior, p), fetch(sp.spec.roughness, p),
                              K, P, p, rx_b, ry_b, li_b, avg);
        specularFresnel = avg * specularWeight;        // :228
        o.flags |= 0x2u;
    }
    specular = specular * (specularWeight * (1.0f - sheenFresnel));   // :231
    float sssWeight = fetch(sp.sss_weight, p);
    sssWeight *= 1.0f - specularFresnel * (1.0f - sheenFresnel);      // :238
    if (sssWeight < kEps) o.flags |= 0x4u;             // :244
    o.sheen = sheen; o.spec = specular;
    o.sheenF = sheenFresnel; o.specF = specularFresnel; o.sssW = sssWeight;
    return o;
}

// ----------------------------------------------------------------------------- f3
struct LightDev { CV3 dir, radiance; const float *pdf; };
struct Mis1 { f3 rgb; float w_light, w_brdf; };

RLS_DEV float power_heuristic(float a, float b) { float a2 = a * a; return a2 / (a2 + b * b); }

// The two halves combined.  f_l / p_bl: evalBrdf / evalPdf at the light direction; L, f_b, p_b:
// the BRDF sample and its evaluation.  A half with a zero direction or a zero pdf contributes 0.
RLS_DEV Mis1 mis_combine(f3 Ld, f3 Li, float p_l, f3 f_l, float p_bl, bool have_brdf_half, f3 L, f3 f_b, float p_b,
                         f3 Li_b, float p_lb)
{
    Mis1 o;
    o.rgb = mk3(0.0f, 0.0f, 0.0f);
    o.w_light = 0.0f; o.w_brdf = 0.0f;
    if (!is_zero(Ld) && p_l > 0.0f) {
        o.w_light = power_heuristic(p_l, p_bl);
        const float s = o.w_light / p_l;
        o.rgb = mk3(f_l.x * Li.x * s, f_l.y * Li.y * s, f_l.z * Li.z * s);
    }
    if (have_brdf_half && !is_zero(L) && p_b > 0.0f) {
        o.w_brdf = power_heuristic(p_b, p_lb);
        const float s = o.w_brdf / p_b;
        o.rgb = o.rgb + mk3(f_b.x * Li_b.x * s, f_b.y * Li_b.y * s, f_b.z * Li_b.z * s);
    }
    return o;
}

// ----------------------------------------------------------------------------- f4
// Pixel of a sampled direction (writeSample, src/rlUtil.h:127-141); returns false for a zero L.
RLS_DEV bool scatter_pixel(f3 dir, int W, int H, int &i, int &j, bool &red)
{
    if (is_zero(dir)) return false;
    FpExact fp;
    float theta = rlm::acosf_(fp, dir.z);
    float phi = rlm::atan2f_(fp, dir.y, dir.x);
    if (phi < 0.0f) phi += kTwoPi;
    const float kInv2Pi = 0.15915494309189533577f;      // AI_ONEOVER2PI
    int ii = (int)(phi * kInv2Pi * (float)W);
    int jj = (int)(theta / kHalfPi * (float)H);
    i = ii < 0 ? 0 : (ii > W - 1 ? W - 1 : ii);
    j = jj < 0 ? 0 : (jj > H - 1 ? H - 1 : jj);
    red = theta > kHalfPi;
    return true;
}

} // namespace rls
