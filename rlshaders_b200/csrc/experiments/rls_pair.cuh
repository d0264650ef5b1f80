// rls_pair.cuh -- fused units under the FAST arithmetic policy with same-shaped sub-expressions of
// ONE sample evaluated as the two lanes of packed f32x2 instructions.
//
// Why: the fused kernels are instruction-issue bound with the FMA pipe half idle
// (profiles/r01_ncu_summary.md).  A packed FADD2 / FFMA2 is the same IEEE-754 binary32 operation
// on two lanes for one issue slot.  Pairing two SAMPLES per thread needs 128 registers and loses
// (profiles/r01_packed_experiment.md); pairing two evaluations of the SAME sample does not: the
// rough-dielectric unit evaluates the reflection BRDF at L and the refraction BTDF at T through
// the same sequence -- half vector, normalize, D, Fresnel, two masking terms, one final quotient
// (src/rlGgx.h:304-313 and :316-328) -- so lane 0 carries the reflection and lane 1 the refraction,
// and the operands both share (view vector, frame, roughness, IOR ratio) are broadcast operands.
//
// Bit-exactness: every lane operation is the scalar code's operation on the same operand bits, in
// the same order (rls_fused.cuh dielectric_unit is the statement-by-statement reference; the
// sign identities used are noted inline).  Lanes that the scalar code does not evaluate (the
// refraction lane under total internal reflection) run on the reflection lane's operands, their
// results are discarded, and anything they do to the operand tracker can only cause an exact
// re-run.  tests/test_gpu_parity.py: paired == scalar exact policy == both oracles, bit for bit.
#pragma once
#include "../rls_fused.cuh"
#include "rls_f2.cuh"

namespace rls {
namespace pk {

RLS_DEV V2 bc3(f3 a) { return mk3(bc(a.x), bc(a.y), bc(a.z)); }
RLS_DEV V2 zip3(f3 a, f3 b) { return mk3(mk(a.x, b.x), mk(a.y, b.y), mk(a.z, b.z)); }

// rls::ggx_fresnel_c; lanes with gSqr < 0 return 1 and feed the root an in-window dummy
RLS_DEV F2 ggx_fresnel_c(Fp2 &fp, F2 ratio2, F2 c)
{
    F2 gSqr = ratio2 - 1.0f + c * c;
    B2 tir = lt(gSqr, 0.0f);
    F2 gg = fp.sqrt(sel(tir, 1.0f, gSqr));
    F2 gmc = gg - c;
    F2 gpc = gg + c;
    F2 v = sqr(fp.div_pz(gmc, gpc)) * 0.5f * (bc(1.0f) + sqr(fp.div(c * gpc - 1.0f, c * gmc + 1.0f)));
    return sel(tir, 1.0f, v);
}
// rls::ggx_G1_value2 for two cosines; rough2 = sqr(mRoughness)
RLS_DEV F2 ggx_G1_value2(Fp2 &fp, float rough2, F2 VdotN)
{
    F2 cosSqr = sqr(VdotN);
    F2 tanSqr = fp.rcp(cosSqr) - 1.0f;
    F2 denominator = bc(1.0f) + fp.sqrt(bc(1.0f) + bc(rough2) * tanSqr);
    return fp.rcp_in_window(denominator) * 2.0f;
}
// rls::ggx_D for two microfacet normals; yax, yay = FpFast::shared_rcp of g.ax, g.ay
RLS_DEV F2 ggx_D(Fp2 &fp, const Ggx &g, float yax, float yay, V2 m)
{
    F2 MdotU = dot(m, bc3(g.U));
    F2 MdotV = dot(m, bc3(g.V));
    F2 MdotN2 = sqr(dot(bc3(g.N), m));
    F2 denominator = bc(g.ax * g.ay) * sqr(sqr(fp.div_by(MdotU, g.ax, yax)) + sqr(fp.div_by(MdotV, g.ay, yay)) + MdotN2);
    return fp.div(bc(kInvPi), denominator);
}

// rls::dielectric_unit (rls_fused.cuh) with the reflection evaluation at L (lane 0) and the
// refraction evaluation at T (lane 1) paired.
RLS_DEV Dielectric dielectric_unit_paired(FpFast &fp, const Shading &sh, float ior, float rough, float aniso,
                                          float rx, float ry, bool ndf)
{
    Dielectric r;
    Ggx g;
    ggx_init(fp, g, sh, rls::mk3(1.0f, 1.0f, 1.0f), ior, rough, aniso);
    g.ndf = ndf;
    const GgxShared s = ggx_shared(fp, g);
    const f3 m = ggx_sample_normal(fp, g, rx, ry);
    const float Vm = dot(g.wo, m);
    r.wi_r = m * (2.0f * fp.abs_nz(Vm)) - g.wo;
    r.F = rls::ggx_fresnel_c(fp, s.ratio2, fabsf(dot(r.wi_r, m)));
    const f3 L = r.wi_r;
    const float LdotN = dot(L, g.N);

    // getRefractDirection(m, V): src/rlGgx.h:277-291.  Under total internal reflection the
    // refraction lane runs on L (the sampled direction becomes the reflection, :232-236).
    const float cosThetaTSqr = 1.0f + s.eta * (rls::sqr(Vm) - 1.0f);
    const bool tir = cosThetaTSqr < 0.0f;
    f3 T = L;
    float TdotN = LdotN;
    if (!tir) {
        const float sc = s.eta * Vm - s.sgnV * fp.sqrt(cosThetaTSqr);
        T = m * sc - g.wo * s.eta;
        TdotN = dot(T, g.N);
    }
    float yax, yay;
    fp.shared_rcp2(g.ax, g.ay, yax, yay);

    // ---- lanes: (reflection at L, refraction at T)
    Fp2 f2(fp);
    const F2 XdotN = mk(LdotN, TdotN);
    const F2 G1x = ggx_G1_value2(f2, rls::sqr(g.rough), XdotN);               // (G1l, G1t)
    // half vectors: hr' = normalize(o + i) (lane 0), ht = -normalize(i * iorIn + o * iorOut) (lane 1).
    // The negation is applied to the vector, as the reference does: -(a + b) and (-a) + (-b) differ
    // in the sign of an exactly cancelling sum, which would surface in the sign of a zero f_t.
    const F2 pm = mk(1.0f, -1.0f);
    const V2 Hn = normalize(f2, zip3(L + g.wo, g.wo * g.iorIn + T * g.iorOut)) * pm;
    const F2 A = dot(bc3(g.wo), Hn);                                          // (V.H, I.ht)
    const F2 B = dot(zip3(L, T), Hn);                                         // (L.H, O.ht)
    const F2 Dx = ggx_D(f2, g, yax, yay, Hn);
    // hr = sgn(V.N) * H: a multiplication by +-1 is exact, dot(v, s*h) == s*dot(v, h) up to the
    // sign of an exact zero, which neither |.|, the comparison with 0 nor the square below sees
    const F2 sg = mk(s.sgnV, 1.0f);
    const F2 Ar = A * sg;                                                     // (V.hr, I.ht)
    const F2 Br = B * sg;                                                     // (L.hr, O.ht)
    const F2 Fx = ggx_fresnel_c(f2, bc(s.ratio2), abs_m(Ar));                 // (F(V,hr), F(V,ht))
    const F2 G1i = sel(lt(Ar * s.VdotN, 0.0f), 0.0f, bc(s.G1v));
    const F2 G1o = sel(lt(Br * XdotN, 0.0f), 0.0f, G1x);
    const float IdotH = hi(Ar), OdotH = hi(Br);
    // numerators: F * G * D * 0.25   |   |O.h I.h| * iorOut^2 * (1 - F) * G * D
    const float lead_t = rls::abs_m(OdotH * IdotH) * rls::sqr(g.iorOut) * (1.0f - hi(Fx));
    const F2 num = mk(lo(Fx), lead_t) * (G1i * G1o) * Dx * mk(0.25f, 1.0f);   // x * 1 is exact
    // denominators: |L.N| |V.N|   |   |T.N| |V.N| (iorIn I.h + iorOut O.h)^2
    const F2 den = abs_m(XdotN) * s.absVdotN * mk(1.0f, rls::sqr(g.iorIn * IdotH + g.iorOut * OdotH));
    B2 live; live.a = s.sgnV != 0.0f; live.b = true;                          // D(0 vector) = 1/0 in the scalar code
    f2.require(live);
    const F2 res = f2.div_pz(num, den);                                       // (reflection, refraction)
    f2.give(fp);

    // evalPdf(L): VNDFKernel::evalPdf (src/rlGgx.h:72-80) or NDFKernel::evalPdf (:45-50)
    const float VH = lo(A), D_H = lo(Dx);
    if (g.ndf) {
        const f3 H = rls::mk3(lo(Hn.x), lo(Hn.y), lo(Hn.z));
        r.pdf_r = fp.div_pz(D_H * rls::abs_m(dot(H, g.N)) * 0.25f, rls::abs_m(VH));
    } else {
        const float G1_pdf = (VH * s.VdotN < 0.0f) ? 0.0f : s.G1v;
        r.pdf_r = rls::max_m(fp.div_pz(D_H * G1_pdf, s.absVdotN) * 0.25f, kEps);
    }
    const bool zeroL = rls::is_zero(L);
    r.f_r = zeroL ? 0.0f : lo(res) * LdotN;
    uint32_t fl = 0;
    if (zeroL) fl |= 0x0001u;
    if (LdotN <= 0.0f) fl |= 0x0002u;
    if (r.pdf_r == 0.0f) fl |= 0x0004u;
    if (r.f_r == 0.0f) fl |= 0x0008u;
    if (r.pdf_r == kEps) fl |= 0x0040u;
    if (g.entering) fl |= 0x0010u;
    if (tir) fl |= 0x0020u;
    r.wi_t = T;
    r.f_t = tir ? 0.0f : hi(res);
    // getSampleWeight(V, wi_t, m): src/rlGgx.h:294-301
    {
        const float mN = dot(m, g.N);
        const float G1i_w = (Vm * s.VdotN < 0.0f) ? 0.0f : s.G1v;
        const float G1o_w = (dot(T, m) * TdotN < 0.0f) ? 0.0f : hi(G1x);
        r.w_t = (G1i_w * G1o_w) * fp.abs_nz(fp.div(Vm, s.absVdotN * fp.abs_nz(mN)));
    }
    r.flags = fl;
    return r;
}

} // namespace pk
} // namespace rls
