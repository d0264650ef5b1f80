// rls_fused.cuh -- the fused units of work (construct + sample + eval + pdf) with every
// sub-expression that the reference evaluates more than once computed ONCE.
//
// The reference's callbacks re-derive the half vector, D, the view-side masking term and the
// IOR ratio on every call (src/rlGgx.h:304-357 is entered through evalBrdf, evalPdf,
// getSampleWeight and refraction separately).  In a fused kernel all of them see the same
// operand bits, so sharing is bit-exact as long as the shared value is produced by the same
// IEEE operations in the same order -- the identities used are noted inline:
//   * a + b == b + a, a * b == b * a bitwise (IEEE commutativity);
//   * dot(v, s*h) == s*dot(v, h) and (s*x)^2 == x^2 bitwise for s = +-1 (negation is exact and
//     round-to-nearest is sign-symmetric);
//   * 2/d == 2*(1/d) bitwise while neither side over/underflows (power-of-two scaling is exact);
//   * x/1 == x bitwise: one of mIorIn, mIorOut is the unit index (src/rlGgx.h:138-142), so the
//     two quotients mIorOut/mIorIn (:258) and mIorIn/mIorOut (:282) are b and 1/b in some order.
// tests/test_gpu_parity.py asserts fused == separate entry points == oracle, bit for bit.
#pragma once
#include "rls_ggx.cuh"
#include "rls_disney.cuh"

namespace rls {

struct Dielectric { float F, f_r, pdf_r, f_t, w_t; f3 wi_r, wi_t; uint32_t flags; };

// G1's value part with the final 2/d as 2*(1/d): d = 1 + sqrt(..) lies in [2, 2^64] or is inf/NaN.
template <class Fp>
RLS_DEV float ggx_G1_value2(Fp &fp, const Ggx &g, float VdotN)
{
    float cosSqr = sqr(VdotN);
    float tanSqr = fp.rcp(cosSqr) - 1.0f;
    float denominator = 1.0f + fp.sqrt(1.0f + sqr(g.rough) * tanSqr);
    return 2.0f * fp.rcp_in_window(denominator);      // 1 + sqrt(t), t tracked: in [1, 2^60 + 1]
}

// One GGX reflection evaluation + pdf at direction L, sharing the half vector, D and the
// view-side G1 between evalBrdf (src/rlGgx.h:304-313) and evalPdf (:121-127, :72-80).
struct GgxShared {
    float VdotN, absVdotN, sgnV, G1v, ratio2;
    float eta;       // mIorIn / mIorOut (getRefractDirection, src/rlGgx.h:282)
};
template <class Fp>
RLS_DEV GgxShared ggx_shared(Fp &fp, const Ggx &g)
{
    GgxShared s;
    s.VdotN = dot(g.wo, g.N);
    s.absVdotN = fp.abs_nz(s.VdotN);                  // a factor of the divisors of refl and pdf
    s.sgnV = sgn_m(s.VdotN);
    s.G1v = ggx_G1_value2(fp, g, s.VdotN);
    // entering: iorIn = 1, iorOut = b  ->  ratio = b/1 = b, eta = 1/b; leaving: the other way round
    const float invB = fp.rcp(g.iorB);
    s.ratio2 = sqr(g.entering ? g.iorB : invB);
    s.eta = g.entering ? invB : g.iorB;
    return s;
}
// Returns reflection(V, L, N) * dot(L, N) (KsColor applied by the caller) and the pdf.
template <class Fp>
RLS_DEV void ggx_reflect_eval_pdf(Fp &fp, const Ggx &g, const GgxShared &s, f3 L, float LdotN, float G1l,
                                  float &refl_cos, float &pdf)
{
    f3 H = normalize(fp, L + g.wo);                   // o + i (brdf)  ==  V + L (pdf), bitwise
    float VH = dot(g.wo, H);
    float LH = dot(L, H);
    // D(H): also D(hr) for hr = +-H
    float D_H = ggx_D(fp, g, H);
    if (g.ndf) {                                      // NDFKernel::evalPdf (src/rlGgx.h:45-50)
        pdf = fp.div_pz(D_H * abs_m(dot(H, g.N)) * 0.25f, abs_m(VH));
    } else {                                          // VNDFKernel::evalPdf: G1(V, H, N), floored
        float G1_pdf = (VH * s.VdotN < 0.0f) ? 0.0f : s.G1v;
        pdf = max_m(fp.div_pz(D_H * G1_pdf, s.absVdotN) * 0.25f, kEps);
    }
    // brdf: hr = sgn(V.N) * H
    float VHr = VH * s.sgnV, LHr = LH * s.sgnV;
    float F = ggx_fresnel_c(fp, s.ratio2, fabsf(VHr));
    float G1i = (VHr * s.VdotN < 0.0f) ? 0.0f : s.G1v;
    float G1o = (LHr * LdotN < 0.0f) ? 0.0f : G1l;
    float D_hr = (s.sgnV != 0.0f) ? D_H : __int_as_float(0x7f800000);   // D(0 vector) = 1/0
    float refl = fp.div_pz(F * (G1i * G1o) * D_hr * 0.25f, fp.abs_nz(LdotN) * s.absVdotN);   // F or G may be 0
    refl_cos = refl;
}

// The rough-dielectric unit: src/rlGgx.h:228-243 loop body with the in-tree
// getRefractDirection standing in for Arnold's AiRefractRay (same composition as the oracle).
// kFlat: the refraction branch (src/rlGgx.h:230-236) as selects.  Under total internal reflection a
// lane runs the refraction evaluation on the reflected direction and drops the result; the reflection
// and refraction evaluations then sit in ONE basic block and the compiler interleaves the two
// independent instruction streams (with the branch they are consecutive blocks).  Measured on B200:
// neutral (+-0.3 %) at 56 and 64 registers, and 10 % more exact re-runs -- the kernels use kFlat = false.
template <bool kFlat = false, class Fp>
RLS_DEV Dielectric dielectric_unit(Fp &fp, const Shading &sh, float ior, float rough, float aniso, float rx, float ry,
                                   bool ndf = false)
{
    Dielectric r;
    Ggx g;
    ggx_init(fp, g, sh, mk3(1.0f, 1.0f, 1.0f), ior, rough, aniso);
    g.ndf = ndf;
    const GgxShared s = ggx_shared(fp, g);
    bool early;
    f3 m = ggx_sample_normal(fp, g, rx, ry, &early);
    float Vm = dot(g.wo, m);
    // reflectDirection(V, m) = 2|V.m| m - V
    r.wi_r = m * (2.0f * fp.abs_nz(Vm)) - g.wo;       // Vm is the tracked numerator of w_t below
    r.F = ggx_fresnel_c(fp, s.ratio2, fabsf(dot(r.wi_r, m)));

    // evalBrdf(wi_r) with white KsColor, evalPdf(wi_r)
    const f3 L = r.wi_r;
    const float LdotN = dot(L, g.N);
    const float G1l = ggx_G1_value2(fp, g, LdotN);
    float refl;
    ggx_reflect_eval_pdf(fp, g, s, L, LdotN, G1l, refl, r.pdf_r);
    const bool zeroL = is_zero(L);
    r.f_r = zeroL ? 0.0f : refl * LdotN;              // (1 * refl) * dot(L, N)
    uint32_t fl = 0;
    if (zeroL) fl |= 0x0001u;
    if (LdotN <= 0.0f) fl |= 0x0002u;
    if (r.pdf_r == 0.0f) fl |= 0x0004u;
    if (r.f_r == 0.0f) fl |= 0x0008u;
    if (r.pdf_r == kEps) fl |= 0x0040u;
    if (g.entering) fl |= 0x0010u;
    if (early) fl |= 0x0080u;                         // RLS_FLAG_SLOPE_EARLY_OUT

    // getRefractDirection(m, V): src/rlGgx.h:277-291
    const float eta = s.eta;
    const float cosThetaTSqr = 1.0f + eta * (sqr(Vm) - 1.0f);
    const float mN = dot(m, g.N);
    float TdotN, G1t;
    if (kFlat) {
        const bool tir = cosThetaTSqr < 0.0f;
        if (tir) fl |= 0x0020u;
        const float sc = eta * Vm - s.sgnV * fp.sqrt(tir ? 1.0f : cosThetaTSqr);
        const f3 Tr = m * sc - g.wo * eta;
        const f3 T = tir ? L : Tr;
        r.wi_t = T;
        TdotN = dot(T, g.N);                          // == LdotN bitwise under TIR (same operands)
        G1t = ggx_G1_value2(fp, g, TdotN);            // == G1l bitwise under TIR
        f3 ht = -normalize(fp, g.wo * g.iorIn + T * g.iorOut);
        float IdotH = dot(g.wo, ht);
        float OdotH = dot(T, ht);
        float refractWeight = 1.0f - ggx_fresnel_c(fp, s.ratio2, fabsf(IdotH));
        float denominator = fp.abs_nz(TdotN) * s.absVdotN * sqr(g.iorIn * IdotH + g.iorOut * OdotH);
        float G1i = (IdotH * s.VdotN < 0.0f) ? 0.0f : s.G1v;
        float G1o = (OdotH * TdotN < 0.0f) ? 0.0f : G1t;
        const float f_t = fp.div_pz(abs_m(OdotH * IdotH) * sqr(g.iorOut) * refractWeight * (G1i * G1o) * ggx_D(fp, g, ht), denominator);
        r.f_t = tir ? 0.0f : f_t;
    } else if (cosThetaTSqr < 0.0f) {                 // total internal reflection: reflect about m
        fl |= 0x0020u;
        r.wi_t = r.wi_r;
        r.f_t = 0.0f;
        TdotN = LdotN;
        G1t = G1l;
    } else {
        float sc = eta * Vm - s.sgnV * fp.sqrt(cosThetaTSqr);
        f3 T = m * sc - g.wo * eta;
        r.wi_t = T;
        TdotN = dot(T, g.N);
        G1t = ggx_G1_value2(fp, g, TdotN);
        // refraction(V, T, N): src/rlGgx.h:316-328
        f3 ht = -normalize(fp, g.wo * g.iorIn + T * g.iorOut);
        float IdotH = dot(g.wo, ht);
        float OdotH = dot(T, ht);
        float refractWeight = 1.0f - ggx_fresnel_c(fp, s.ratio2, fabsf(IdotH));
        float denominator = fp.abs_nz(TdotN) * s.absVdotN * sqr(g.iorIn * IdotH + g.iorOut * OdotH);
        float G1i = (IdotH * s.VdotN < 0.0f) ? 0.0f : s.G1v;
        float G1o = (OdotH * TdotN < 0.0f) ? 0.0f : G1t;
        r.f_t = fp.div_pz(abs_m(OdotH * IdotH) * sqr(g.iorOut) * refractWeight * (G1i * G1o) * ggx_D(fp, g, ht), denominator);
    }
    // getSampleWeight(V, wi_t, m): src/rlGgx.h:294-301
    {
        float G1i = (Vm * s.VdotN < 0.0f) ? 0.0f : s.G1v;
        float G1o = (dot(r.wi_t, m) * TdotN < 0.0f) ? 0.0f : G1t;
        r.w_t = (G1i * G1o) * fp.abs_nz(fp.div(Vm, s.absVdotN * fp.abs_nz(mN)));   // a tracked quotient is not 0
    }
    r.flags = fl;
    return r;
}

// Fused rlGgx unit with a KsColor: ctor + evalSample + evalBrdf + evalPdf (+ the Fresnel term).
struct GgxBsdf { f3 L, f; float pdf, fresnel; uint32_t flags; };
template <class Fp>
RLS_DEV GgxBsdf ggx_unit(Fp &fp, const Ggx &g, float rx, float ry)
{
    GgxBsdf o;
    const GgxShared s = ggx_shared(fp, g);
    bool early;
    f3 m = ggx_sample_normal(fp, g, rx, ry, &early);
    o.L = m * (2.0f * abs_m(dot(g.wo, m))) - g.wo;
    o.fresnel = ggx_fresnel_c(fp, s.ratio2, fabsf(dot(o.L, m)));
    const float LdotN = dot(o.L, g.N);
    float refl;
    ggx_reflect_eval_pdf(fp, g, s, o.L, LdotN, ggx_G1_value2(fp, g, LdotN), refl, o.pdf);
    const bool black = is_zero(o.L) || (abs_m(g.ks.x) < kEps && abs_m(g.ks.y) < kEps && abs_m(g.ks.z) < kEps);
    o.f = black ? mk3(0.0f, 0.0f, 0.0f) : g.ks * refl * LdotN;
    o.flags = bsdf_flags(o.L, g.N, o.f, o.pdf);
    if (g.entering) o.flags |= 0x0010u;
    if (early) o.flags |= 0x0080u;                    // RLS_FLAG_SLOPE_EARLY_OUT
    return o;
}

} // namespace rls

namespace rls {

// ------------------------------------------------------------------ rlDisney fused unit
// ctor + glossy triple + diffuse triple with the terms the reference recomputes shared:
//   * the half vector M = normalize(L + V), L.M, N.M of evalSpecular (src/rlDisney.cpp:328-331)
//     and evalSpecularPdf (:522-524) are the same IEEE operations on the same bits;
//   * D_GTR2Aniso(M) and D_GTR1 (incl. its logf of a per-sample constant) are evaluated by both
//     (:339,347 and :536-537);
//   * V.N and the two view-side smithG_GGX terms depend on the shading point only.
// The local microfacet normal of both specular lobes goes through ONE rotate/normalize/reflect tail.
struct DisneyOut1 { f3 Ls, fs, Ld, fd; float ps, pd; uint32_t flags; };

template <class Fp>
RLS_DEV DisneyOut1 disney_unit(Fp &fp, const Disney &d, float rx_s, float ry_s, float rx_d, float ry_d)
{
    DisneyOut1 o;
    const float VdotN = dot(d.wo, d.N);
    // per-sample constants of D_GTR1 (src/rlDisney.cpp:547-549)
    const float gtr1_alpha = lerp_m(d.clearcoatGloss, 0.1f, 0.001f);
    const float gtr1_a2 = sqr(gtr1_alpha);
    const float gtr1_log = rlm::logf_(fp, gtr1_a2);
    auto D_GTR1_shared = [&](float MdotN2) {
        float denominator = gtr1_log * (1.0f + (gtr1_a2 - 1.0f) * MdotN2);
        return fp.div((gtr1_a2 - 1.0f) * kInvPi, denominator);
    };

    // ---- specular sample (src/rlDisney.cpp:367-390).  Both lobes go through ONE sincosf and
    // one rotate/normalize tail: GTR2 (visible normals) needs sincosf(phi or 2 pi ry), GTR1
    // sincosf(2 pi rx'), plain GTR2 sincosf(2 pi rx) -- the angle is selected per lane.
    uint32_t lobe;
    bool early = false;                                           // GTR2 visible-normal lobe only
    f3 M;
    {
        float gtr2Weight = fp.rcp(d.clearcoat + 1.0f);
        VndfState st;
        float rx, angle, cosThetaH = 0.0f, g = 0.0f;
        if (rx_s < gtr2Weight) {
            lobe = 0;
            rx = fp.div(rx_s, gtr2Weight);
            if (d.visibleNormal) {
                st = vndf_prepare(fp, d.wo, d.U, d.V, d.N, d.ax, d.ay);
                angle = vndf_angle(st, ry_s);
            } else {                                              // sampleGTR2AnisoDirection (:406-414)
                g = fp.sqrt(fp.div(ry_s, 1.0f - ry_s));           // sample_ndf_normal(.., ry_s, rx)
                angle = kTwoPi * rx;
            }
        } else {
            lobe = 1;
            rx = fp.div_pz(rx_s - gtr2Weight, 1.0f - gtr2Weight);
            angle = kTwoPi * rx;                                  // sampleGTR1Direction (:393-404)
            float a2 = sqr(d.roughness);
            cosThetaH = (a2 == 1.0f) ? fp.sqrt(1.0f - ry_s)
                                     : fp.sqrt(fp.div(1.0f - rlm::powf_<Fp::kSmemTables>(a2, 1.0f - ry_s), 1.0f - a2));
        }
        float s, c;
        rlm::sincosf_(fp, angle, &s, &c);
        f3 omega;
        if (lobe == 0) {
            omega = d.visibleNormal ? vndf_omega(fp, st, s, c, d.ax, d.ay, rx, ry_s, &early)
                                    : mk3(g * d.ax * c, g * d.ay * s, 1.0f);
        } else {                                                  // sphericalDirection(cosThetaH, phiH)
            float r = fp.sqrt(1.0f - sqr(cosThetaH));
            omega = mk3(r * c, r * s, cosThetaH);
        }
        M = normalize(fp, rotate_to_frame(omega, d.U, d.V, d.N));
    }
    const bool zeroS = dot(d.N, M) < 0.0f;
    o.Ls = zeroS ? mk3(0.0f, 0.0f, 0.0f) : reflect_direction(d.wo, M);

    // ---- specular eval + pdf at Ls, sharing the half vector and both D terms
    if (zeroS) {
        o.fs = mk3(0.0f, 0.0f, 0.0f);
        o.ps = 0.0f;
    } else {
        const f3 L = o.Ls;
        const float LdotN = dot(L, d.N);                     // == dot(N, L) bitwise
        const f3 H = normalize(fp, L + d.wo);
        const float LdotM = dot(L, H);
        const float NdotM = dot(d.N, H);                     // == dot(H, N) bitwise
        const float NdotM2 = sqr(NdotM);
        const float Ds = D_GTR2Aniso(fp, d, H, NdotM2);
        const float Dr = D_GTR1_shared(NdotM2);
        // pdf (:520-543)
        if (NdotM < 0.0f) {
            o.ps = 0.0f;
        } else {
            const float IdotM = fp.abs_nz(LdotM);            // the divisor of both pdf forms
            const float clearcoatWeight = fp.div_pz(d.clearcoat, d.clearcoat + 1.0f);
            if (d.visibleNormal) {
                const float Vn = max_m(1e-4f, VdotN);
                const float Dw = fp.div(smithG_GGX(fp, IdotM, d.specRough) * Ds * 2.0f * IdotM, Vn);
                o.ps = lerp_m(clearcoatWeight, Dw, fp.div(Dr * abs_m(NdotM), IdotM)) * 0.25f;
            } else {
                o.ps = fp.div(lerp_m(clearcoatWeight, Ds, Dr) * abs_m(NdotM) * 0.25f, IdotM);
            }
        }
        // eval (:318-356) x N.L (:136)
        if (LdotN < kEps || VdotN < kEps || NdotM < kEps || LdotM < kEps) {
            o.fs = mk3(0.0f, 0.0f, 0.0f) * LdotN;            // black * NdotL keeps the sign of zero
        } else {
            const float FH = rlm::pow5_unit_<Fp::kSmemTables>(clamp_m(1.0f - LdotM, 0.0f, 1.0f));
            const f3 Fs = lerp_m(FH, d.F0, mk3(1.0f, 1.0f, 1.0f));
            const float Gs = smithG_GGX(fp, LdotN, d.specRough) * smithG_GGX(fp, VdotN, d.specRough);
            const float Fr = lerp_m(FH, 0.04f, 1.0f);
            const float Gr = smithG_GGX(fp, LdotN, 0.25f) * smithG_GGX(fp, VdotN, 0.25f);
            const f3 Fsheen = d.sheenColor * FH * (1.0f - d.metallic);
            const f3 spec = Fs * Ds * Gs;
            const float coat = d.clearcoat * Dr * Fr * Gr;
            o.fs = (mk3(spec.x + coat, spec.y + coat, spec.z + coat) + Fsheen) * LdotN;
        }
    }

    // ---- diffuse triple (src/rlDisney.cpp:359-365, 199-236, 515-518)
    o.Ld = disney_sample_diffuse(fp, d, rx_d, ry_d);
    o.fd = disney_eval_brdf(fp, d, kRayDiffuse, o.Ld);
    o.pd = disney_eval_pdf(fp, d, kRayDiffuse, o.Ld);

    const uint32_t fls = (bsdf_flags(o.Ls, d.N, o.fs, o.ps) & ~0x0040u) | (lobe << 8) | (early ? 0x0080u : 0u);
    const uint32_t fld = bsdf_flags(o.Ld, d.N, o.fd, o.pd);
    o.flags = fls | (fld << 16);
    return o;
}

} // namespace rls
