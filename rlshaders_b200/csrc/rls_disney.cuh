// rls_disney.cuh -- device restatement of DisneySampler (reference src/rlDisney.cpp:105-602).
#pragma once
#include "rls_ggx.cuh"

namespace rls {

constexpr int kRayDiffuse = 0x20;   // AI_RAY_DIFFUSE
constexpr int kRayGlossy  = 0x40;   // AI_RAY_GLOSSY

struct DisneyParamsDev {
    P3 base_color;
    P1 subsurface, metallic, specular, specular_tint, roughness, anisotropic, sheen, sheen_tint,
       clearcoat, clearcoat_gloss;
    int sample_from_visible_normal;
};

struct Disney {
    f3 U, V, N, wo;
    f3 base, F0, sheenColor;
    float roughness, subsurface, metallic, clearcoat, clearcoatGloss;
    float specRough, ax, ay;
    bool visibleNormal;   // mSampleFromVisibleNormal (src/rlDisney.cpp:191)
};

// src/rlDisney.cpp:155-192
template <bool kArrays = false, bool kReload = false, class Fp>
RLS_DEV void disney_init(Fp &fp, Disney &d, const Shading &sh, const DisneyParamsDev &p, uint32_t i)
{
    d.U = sh.U; d.V = sh.V; d.N = sh.N; d.wo = sh.wo;
    d.base = fetch_t<kArrays, kReload>(p.base_color, i);
    d.roughness = fetch_t<kArrays, kReload>(p.roughness, i);
    d.subsurface = fetch_t<kArrays, kReload>(p.subsurface, i);
    float specular = fetch_t<kArrays, kReload>(p.specular, i) * 0.08f;            // :163
    float specularTint = fetch_t<kArrays, kReload>(p.specular_tint, i);
    d.metallic = fetch_t<kArrays, kReload>(p.metallic, i);
    float sheen = fetch_t<kArrays, kReload>(p.sheen, i);
    float sheenTint = fetch_t<kArrays, kReload>(p.sheen_tint, i);
    float anisotropic = fetch_t<kArrays, kReload>(p.anisotropic, i);
    d.clearcoat = fetch_t<kArrays, kReload>(p.clearcoat, i) * 0.25f;              // :169
    d.clearcoatGloss = fetch_t<kArrays, kReload>(p.clearcoat_gloss, i);

    float aspect = fp.sqrt(1.0f - anisotropic * 0.9f);        // :177
    d.ax = max_m(1e-2f, fp.div_pz(sqr(d.roughness), aspect)); // :178 (floor 1e-2, not 1e-4)
    d.ay = max_m(1e-2f, sqr(d.roughness) * aspect);
    d.specRough = sqr(d.roughness);                           // :181

    float luminance = color_to_luminance(d.base);             // :185
    f3 white = mk3(1.0f, 1.0f, 1.0f);
    f3 tint = white;
    if (luminance > 0.0f)                                     // a channel may be exactly 0
        tint = mk3(fp.div_pz(d.base.x, luminance), fp.div_pz(d.base.y, luminance), fp.div_pz(d.base.z, luminance));
    f3 metallicColor = lerp_m(specularTint, white, tint) * specular;   // :187
    d.F0 = lerp_m(d.metallic, metallicColor, d.base);                  // :188
    d.sheenColor = lerp_m(sheenTint, white, tint) * sheen;             // :190
    d.visibleNormal = p.sample_from_visible_normal != 0;
}
// src/rlDisney.cpp:570-577
template <class Fp>
RLS_DEV float smithG_GGX(Fp &fp, float NdotV, float alphaG)
{
    float a = alphaG * alphaG;
    float b = NdotV * NdotV;
    return fp.rcp(NdotV + fp.sqrt(a + b - a * b));
}
// src/rlDisney.cpp:545-551
template <class Fp>
RLS_DEV float D_GTR1(Fp &fp, const Disney &d, float MdotN2)
{
    float alpha = lerp_m(d.clearcoatGloss, 0.1f, 0.001f);
    float a2 = sqr(alpha);
    float denominator = rlm::logf_(fp, a2) * (1.0f + (a2 - 1.0f) * MdotN2);
    return fp.div((a2 - 1.0f) * kInvPi, denominator);
}
// src/rlDisney.cpp:561-568
template <class Fp>
RLS_DEV float D_GTR2Aniso(Fp &fp, const Disney &d, f3 m, float MdotN2)
{
    float HdotU = dot(m, d.U);
    float HdotV = dot(m, d.V);
    float denominator = d.ax * d.ay * sqr(sqr(fp.div(HdotU, d.ax)) + sqr(fp.div(HdotV, d.ay)) + MdotN2);
    return fp.div(kInvPi, denominator);
}
// src/rlDisney.cpp:199-236
template <class Fp>
RLS_DEV f3 disney_eval_diffuse(Fp &fp, const Disney &d, f3 L)
{
    float LdotN = dot(L, d.N);
    float VdotN = dot(d.wo, d.N);
    if (LdotN < kEps || VdotN < kEps) return mk3(0.0f, 0.0f, 0.0f);
    f3 H = normalize(fp, L + d.wo);
    float LdotH = dot(L, H);
    float NdotH = dot(d.wo, H);   // sic: V.H (:210)
    if (NdotH < kEps || LdotH < kEps) return mk3(0.0f, 0.0f, 0.0f);
    float LdotH2 = sqr(LdotH);
    float FL = rlm::pow5_unit_<Fp::kSmemTables>(clamp_m(1.0f - LdotN, 0.0f, 1.0f));
    float FV = rlm::pow5_unit_<Fp::kSmemTables>(clamp_m(1.0f - VdotN, 0.0f, 1.0f));
    float F90 = 0.5f + 2.0f * d.roughness * LdotH2;
    float diffuseFactor = lerp_m(FL, 1.0f, F90) * lerp_m(FV, 1.0f, F90);
    float Fss90 = d.roughness * LdotH2;
    float Fss = lerp_m(FL, 1.0f, Fss90) * lerp_m(FV, 1.0f, Fss90);
    float ssFactor = 1.25f * (Fss * (fp.rcp(LdotN + VdotN) - 0.5f) + 0.5f);
    f3 diffuse = d.base * kInvPi * lerp_m(d.subsurface, diffuseFactor, ssFactor);
    return diffuse * (1.0f - d.metallic);
}
// src/rlDisney.cpp:318-356
template <class Fp>
RLS_DEV f3 disney_eval_specular(Fp &fp, const Disney &d, f3 L)
{
    float LdotN = dot(L, d.N);
    float VdotN = dot(d.wo, d.N);
    if (LdotN < kEps || VdotN < kEps) return mk3(0.0f, 0.0f, 0.0f);
    f3 M = normalize(fp, L + d.wo);
    float LdotM = dot(L, M);
    float NdotM = dot(d.N, M);
    if (NdotM < kEps || LdotM < kEps) return mk3(0.0f, 0.0f, 0.0f);
    float NdotM2 = sqr(NdotM);
    float Ds = D_GTR2Aniso(fp, d, M, NdotM2);
    float FH = rlm::pow5_unit_<Fp::kSmemTables>(clamp_m(1.0f - LdotM, 0.0f, 1.0f));
    f3 Fs = lerp_m(FH, d.F0, mk3(1.0f, 1.0f, 1.0f));
    float Gs = smithG_GGX(fp, LdotN, d.specRough) * smithG_GGX(fp, VdotN, d.specRough);
    float Dr = D_GTR1(fp, d, NdotM2);
    float Fr = lerp_m(FH, 0.04f, 1.0f);
    float Gr = smithG_GGX(fp, LdotN, 0.25f) * smithG_GGX(fp, VdotN, 0.25f);
    f3 Fsheen = d.sheenColor * FH * (1.0f - d.metallic);
    f3 spec = Fs * Ds * Gs;
    float coat = d.clearcoat * Dr * Fr * Gr;
    return mk3(spec.x + coat, spec.y + coat, spec.z + coat) + Fsheen;
}
// src/rlDisney.cpp:120-137
template <class Fp>
RLS_DEV f3 disney_eval_brdf(Fp &fp, const Disney &d, int type, f3 L)
{
    if (is_zero(L)) return mk3(0.0f, 0.0f, 0.0f);
    float NdotL = dot(d.N, L);
    f3 f = (type == kRayDiffuse) ? disney_eval_diffuse(fp, d, L) : disney_eval_specular(fp, d, L);
    return f * NdotL;
}
// src/rlDisney.cpp:359-365
template <class Fp>
RLS_DEV f3 disney_sample_diffuse(Fp &fp, const Disney &d, float rx, float ry)
{
    f2 k = concentric_disk_sample(fp, rx, ry);
    f3 omega = mk3(k.x, k.y, fp.sqrt(max_m(0.0f, 1.0f - sqr(k.x) - sqr(k.y))));
    return rotate_to_frame(omega, d.U, d.V, d.N);
}
// src/rlDisney.cpp:393-404 (a2 = roughness^2, NOT the clearcoat-gloss alpha of D_GTR1)
template <class Fp>
RLS_DEV f3 disney_sample_gtr1(Fp &fp, const Disney &d, float rx, float ry)
{
    float phiH = kTwoPi * rx;
    float a2 = sqr(d.roughness);
    float cosThetaH = (a2 == 1.0f) ? fp.sqrt(1.0f - ry)
                                   : fp.sqrt(fp.div(1.0f - rlm::powf_<Fp::kSmemTables>(a2, 1.0f - ry), 1.0f - a2));
    f3 omega = spherical_direction(fp, cosThetaH, phiH);
    return normalize(fp, rotate_to_frame(omega, d.U, d.V, d.N));
}
// src/rlDisney.cpp:367-390; lobe: 0 = GTR2 (visible normals), 1 = GTR1 (clearcoat)
template <class Fp>
RLS_DEV f3 disney_sample_specular(Fp &fp, const Disney &d, float rx, float ry, uint32_t &lobe, bool *early = nullptr)
{
    f3 M;
    if (early) *early = false;
    float gtr2Weight = fp.rcp(d.clearcoat + 1.0f);
    if (rx < gtr2Weight) {
        rx = fp.div(rx, gtr2Weight);
        // :377-379; sampleGTR2AnisoDirection (:406-414) is NDF sampling with (ry, rx)
        M = d.visibleNormal ? sample_visible_normal(fp, d.wo, d.U, d.V, d.N, d.ax, d.ay, rx, ry, early)
                            : sample_ndf_normal(fp, d.U, d.V, d.N, d.ax, d.ay, ry, rx);
        lobe = 0;
    } else {
        rx = fp.div_pz(rx - gtr2Weight, 1.0f - gtr2Weight);
        M = disney_sample_gtr1(fp, d, rx, ry);
        lobe = 1;
    }
    if (dot(d.N, M) < 0.0f) return mk3(0.0f, 0.0f, 0.0f);
    return reflect_direction(d.wo, M);
}
// src/rlDisney.cpp:515-518
RLS_DEV float disney_diffuse_pdf(const Disney &d, f3 i) { return max_m(1e-4f, dot(i, d.N) * kInvPi); }
// src/rlDisney.cpp:520-543 (mSampleFromVisibleNormal == true)
template <class Fp>
RLS_DEV float disney_specular_pdf(Fp &fp, const Disney &d, f3 i)
{
    f3 m = normalize(fp, i + d.wo);
    float IdotM = abs_m(dot(i, m));
    float MdotN = dot(m, d.N);
    if (MdotN < 0.0f) return 0.0f;
    float MdotN2 = sqr(MdotN);
    float clearcoatWeight = fp.div_pz(d.clearcoat, d.clearcoat + 1.0f);   // clearcoat 0 is the default
    if (!d.visibleNormal) {                                   // :541-542
        float D0 = lerp_m(clearcoatWeight, D_GTR2Aniso(fp, d, m, MdotN2), D_GTR1(fp, d, MdotN2));
        return fp.div(D0 * abs_m(MdotN) * 0.25f, IdotM);
    }
    float VdotN = max_m(1e-4f, dot(d.wo, d.N));
    float Dw = fp.div(smithG_GGX(fp, IdotM, d.specRough) * D_GTR2Aniso(fp, d, m, MdotN2) * 2.0f * IdotM, VdotN);
    float D = lerp_m(clearcoatWeight, Dw, fp.div(D_GTR1(fp, d, MdotN2) * abs_m(MdotN), IdotM));
    return D * 0.25f;
}
// src/rlDisney.cpp:139-152
template <class Fp>
RLS_DEV float disney_eval_pdf(Fp &fp, const Disney &d, int type, f3 L)
{
    if (is_zero(L)) return 0.0f;
    return (type == kRayDiffuse) ? disney_diffuse_pdf(d, L) : disney_specular_pdf(fp, d, L);
}

} // namespace rls
