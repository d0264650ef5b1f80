// rls_ggx.cuh -- device restatement of the rlGgx sampler (reference src/rlGgx.h,
// src/rlGgx.cpp:14-99) and of the rlUtil helpers it uses (src/rlUtil.h:21-39,
// src/rlUtil.cpp:3-27).  Register-resident, one sample per thread; every expression
// keeps the reference's operation order (see rls_math.cuh for the numerical contract).
#pragma once
#include "rls_math.cuh"
#include "rls_libm.cuh"

namespace rls {

// ---------------------------------------------------------------- rlUtil helpers
// src/rlUtil.h:21-29
template <class Fp>
RLS_DEV f3 spherical_direction(Fp &fp, float cosTheta, float phi)
{
    float s, c;
    rlm::sincosf_(fp, phi, &s, &c);
    float r = fp.sqrt(1.0f - sqr(cosTheta));
    return mk3(r * c, r * s, cosTheta);
}
// src/rlUtil.h:31-34 -- 2*ABS(i.n)*n - i
RLS_DEV f3 reflect_direction(f3 i, f3 n)
{
    float s = 2.0f * abs_m(dot(i, n));
    return n * s - i;
}
// src/rlUtil.h:36-39
RLS_DEV float color_to_luminance(f3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }
// src/rlUtil.cpp:3-27 (x, y only; callers overwrite z)
template <class Fp>
RLS_DEV f2 concentric_disk_sample(Fp &fp, float rx, float ry)
{
    rx = rx * 2.0f - 1.0f;
    ry = ry * 2.0f - 1.0f;
    f2 o; o.x = 0.0f; o.y = 0.0f;
    if (rx == 0.0f && ry == 0.0f) return o;
    // rlUtil.cpp:16-22: the two sides divide different operands, so the operands are selected and ONE quotient is
    // formed (a 50 / 50 branch would make every warp run both)
    const bool wide = abs_m(rx) > abs_m(ry);
    const float r = wide ? rx : ry;
    const float q = fp.div(wide ? kHalfPi * 0.5f * ry : 0.5f * rx, r);
    const float phi = wide ? q : kHalfPi * (1.0f - q);
    float s, c;
    rlm::sincosf_(fp, phi, &s, &c);
    o.x = r * c;
    o.y = r * s;
    return o;
}

// ------------------------------------------- visible-normal sampling (Heitz-d'Eon)
// src/rlGgx.cpp:18-25
template <class Fp>
RLS_DEV f2 uniform_slope(Fp &fp, float rx, float ry)
{
    float r = fp.sqrt(fp.div(rx, 1.0f - rx));
    float phi = kTwoPi * ry;
    float s, c;
    rlm::sincosf_(fp, phi, &s, &c);
    f2 o; o.x = r * c; o.y = r * s;
    return o;
}
// src/rlGgx.cpp:14-61 (VNDFKernel::sampleSlope) for theta >= AI_EPSILON; rlDisney.cpp:416-463 is
// the same code.  The theta < AI_EPSILON early-out (:27) is taken by the caller.
// *early (optional) = the |A^2 - 1| < AI_EPSILON early-out (:38) was taken (RLS_FLAG_SLOPE_EARLY_OUT).
template <class Fp>
RLS_DEV f2 sample_slope(Fp &fp, float theta, float rx, float ry, bool *early = nullptr)
{
    float B = rlm::tanf_(fp, theta);
    float B2 = sqr(B);
    // == 2/(..) bitwise: the divisor 1 + sqrt(t) is in [1, 2^60 + 1] once t is tracked
    float G1 = 2.0f * fp.rcp_in_window(1.0f + fp.sqrt(1.0f + B2));

    float A = fp.div(2.0f * rx, G1) - 1.0f;
    float A2 = sqr(A);
    if (abs_m(A2 - 1.0f) < kEps) {
        if (early) *early = true;
        return uniform_slope(fp, rx, ry);
    }

    float tmp = fp.rcp(A2 - 1.0f);
    float D = fp.sqrt(max_m(0.0f, B2 * sqr(tmp) - (A2 - B2) * tmp));
    float slopeX1 = B * tmp - D;
    float slopeX2 = B * tmp + D;
    f2 slope;
    slope.x = (A < 0.0f || slopeX2 > fp.rcp(B)) ? slopeX1 : slopeX2;

    float sign = 1.0f;
    if (ry > 0.5f) {
        ry = 2.0f * (ry - 0.5f);
    } else {
        sign = -1.0f;
        ry = 2.0f * (0.5f - ry);
    }
    float z = fp.div(ry * (ry * (ry * 0.27385f - 0.73369f) + 0.46341f),
                     ry * (ry * (ry * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    slope.y = sign * z * fp.sqrt(1.0f + sqr(slope.x));
    return slope;
}
// src/rlGgx.cpp:63-99 (VNDFKernel::evalSample); rlDisney.cpp:467-502 is the same code.
//
// Control flow is arranged so that a warp runs ONE sincosf for the slope/rotation stage: when
// the stretched view is (nearly) the normal the reference leaves theta = phi = 0 (:82), takes
// the uniform-slope early-out (sincosf(2 pi ry), :18-25) and then rotates by phi = 0, for which
// sincosf returns exactly (0, 1); otherwise it rotates by sincosf(phi).  theta = acosf(V.z) with
// V.z < 1 - 1e-4 is >= 0.0141 (or NaN), so the early-out is taken exactly when theta was left 0.
// The sampler in three stages, so that a caller with other lobes in the same warp (rlDisney's
// GTR1 clearcoat) can share the one sincosf and the rotate/normalize tail:
//   vndf_prepare  view -> stretched polar angles (theta, phi), or `along_normal`
//   vndf_angle    the argument of the slope/rotation sincosf
//   vndf_omega    (sin, cos) of that angle -> un-normalised local microfacet normal
struct VndfState { bool along_normal; float theta, phi; };

template <class Fp>
RLS_DEV VndfState vndf_prepare(Fp &fp, f3 view, f3 U, f3 Vax, f3 N, float ax, float ay)
{
    float cosThetaV = clamp_m(dot(N, view), -1.0f, 1.0f);
    float phiV = rlm::atan2f_(fp, dot(Vax, view), dot(U, view));
    f3 V = spherical_direction(fp, cosThetaV, phiV);

    V.x *= ax;
    V.y *= ay;
    V = normalize(fp, V);

    VndfState st;
    st.along_normal = !(V.z < (1.0f - kEps));
    st.theta = 0.0f; st.phi = 0.0f;
    if (!st.along_normal) {
        st.theta = rlm::acosf_(fp, V.z);
        st.phi = rlm::atan2f_(fp, V.y, V.x);
    }
    return st;
}
RLS_DEV float vndf_angle(const VndfState &st, float ry) { return st.along_normal ? kTwoPi * ry : st.phi; }

template <class Fp>
RLS_DEV f3 vndf_omega(Fp &fp, const VndfState &st, float s, float c, float ax, float ay, float rx, float ry,
                      bool *early = nullptr)
{
    if (early) *early = st.along_normal;         // theta left 0: the :27 early-out
    f2 slope;
    float sinPhi, cosPhi;
    if (st.along_normal) {                       // uniform_slope(rx, ry), then a rotation by phi = 0
        float r = fp.sqrt(fp.div(rx, 1.0f - rx));
        slope.x = r * c;
        slope.y = r * s;
        sinPhi = 0.0f;
        cosPhi = 1.0f;
    } else {
        slope = sample_slope(fp, st.theta, rx, ry, early);
        sinPhi = s;
        cosPhi = c;
    }
    f3 omega;
    omega.x = -(cosPhi * slope.x - sinPhi * slope.y) * ax;
    omega.y = -(sinPhi * slope.x + cosPhi * slope.y) * ay;
    omega.z = 1.0f;
    return omega;
}

template <class Fp>
RLS_DEV f3 sample_visible_normal(Fp &fp, f3 view, f3 U, f3 Vax, f3 N, float ax, float ay, float rx, float ry,
                                 bool *early = nullptr)
{
    const VndfState st = vndf_prepare(fp, view, U, Vax, N, ax, ay);
    float s, c;
    rlm::sincosf_(fp, vndf_angle(st, ry), &s, &c);
    return normalize(fp, rotate_to_frame(vndf_omega(fp, st, s, c, ax, ay, rx, ry, early), U, Vax, N));
}

// src/rlGgx.h:33-41 (NDFKernel::evalSample, Burley Eq.14): plain NDF sampling; also
// DisneySampler::sampleGTR2AnisoDirection (src/rlDisney.cpp:406-414) with (rx, ry) swapped.
template <class Fp>
RLS_DEV f3 sample_ndf_normal(Fp &fp, f3 U, f3 Vax, f3 N, float ax, float ay, float rx, float ry)
{
    float g = fp.sqrt(fp.div(rx, 1.0f - rx));
    float phi = kTwoPi * ry;
    float s, c;
    rlm::sincosf_(fp, phi, &s, &c);
    f3 omega = mk3(g * ax * c, g * ay * s, 1.0f);
    return normalize(fp, rotate_to_frame(omega, U, Vax, N));
}

// ------------------------------------------------------------------------ rlGgx
struct Ggx {
    f3 U, V, N, wo, ks;
    float iorIn, iorOut, rough, ax, ay;
    float iorB;      // max(ior, 1e-4): the side of the interface that is not the unit index (:138-142)
    bool entering;
    bool ndf;        // GgxSamplerT<NDFKernel> instead of the shipped GgxSamplerT<VNDFKernel>
};
template <class Fp>
RLS_DEV f3 ggx_sample_normal(Fp &fp, const Ggx &g, float rx, float ry, bool *early = nullptr)
{
    if (early) *early = false;                   // NDFKernel has no early-out
    if (g.ndf) return sample_ndf_normal(fp, g.U, g.V, g.N, g.ax, g.ay, rx, ry);
    return sample_visible_normal(fp, g.wo, g.U, g.V, g.N, g.ax, g.ay, rx, ry, early);
}

// src/rlGgx.h:130-156 (GgxSamplerT ctor)
template <class Fp>
RLS_DEV void ggx_init(Fp &fp, Ggx &g, const Shading &sh, f3 ks, float ior, float roughness, float aniso)
{
    // :137  dot(sg->N, sg->Rd) with sg->N = +-N, sg->Rd = -wo: the products are those of dot(wo, N)
    // with the signs applied exactly, the sum is +-dot(wo, N) up to the sign of an exact zero, which
    // the comparison does not see.  The fused units reuse the value as V.N (same products, same order).
    const float woN = dot(sh.wo, sh.N);
    g.entering = (sh.backfacing ? woN : -woN) < kEps;
    float a = 1.0f, b = max_m(ior, 1e-4f);       // :138-139
    g.iorIn = g.entering ? a : b;                // :140-142 (swap)
    g.iorOut = g.entering ? b : a;
    g.iorB = b;
    g.wo = sh.wo;                                // :144  -(-wo) is exact
    g.U = sh.U; g.V = sh.V; g.N = sh.N;          // :145-146, explicit frame
    if (aniso == 0.0f) {
        // The isotropic node (a uniform branch when `anisotropic` is a uniform parameter): aspect =
        // sqrt(1 - 0) = 1 and r^2/1 = r^2*1 = r^2 exactly -- no root, no quotient.
        g.ax = g.ay = max_m(1e-4f, sqr(roughness));
    } else {
        float aspect = fp.sqrt(1.0f - aniso * 0.9f); // :148
        g.ax = max_m(1e-4f, fp.div_pz(sqr(roughness), aspect));   // roughness 0 is a legal parameter
        g.ay = max_m(1e-4f, sqr(roughness) * aspect);
    }
    g.rough = max_m(1e-5f, sqr(roughness));      // :155
    g.ks = ks;
    g.ndf = false;
}
// src/rlGgx.h:249-270
// Walter Eq.22 given c = |i.m| and ratio2 = SQR(mIorOut / mIorIn) (:258).
// c = -0 (ABS keeps the sign of a zero) and c = +0 give the same bits: c*c = +0, g -+ (-0) = g -+ 0
// (also for g = 0), (-0)*x -+ 1 = -+1.  The fused units therefore pass fabsf(i.m), an operand modifier.
template <class Fp>
RLS_DEV float ggx_fresnel_c(Fp &fp, float ratio2, float c)
{
    float gSqr = ratio2 - 1.0f + c * c;
    // :260-263 returns 1 for g^2 < 0.  Select instead of branch (a warp almost always holds both
    // kinds of lane): those lanes run the formula on the in-window dummy g^2 = 1 and drop the result.
    const bool tir = gSqr < 0.0f;
    float gg = fp.sqrt(tir ? 1.0f : gSqr);
    float gmc = gg - c;
    float gpc = gg + c;
    // gmc == 0 for ior 1 (a legal parameter): zero-tolerant numerator over gpc > 0
    float v = 0.5f * sqr(fp.div_pz(gmc, gpc)) * (1.0f + sqr(fp.div(c * gpc - 1.0f, c * gmc + 1.0f)));
    return tir ? 1.0f : v;
}
template <class Fp>
RLS_DEV float ggx_fresnel(Fp &fp, const Ggx &g, f3 i, f3 m)
{
    return ggx_fresnel_c(fp, sqr(fp.div(g.iorOut, g.iorIn)), abs_m(dot(i, m)));
}
// src/rlGgx.h:343-357, split so the view-only half can be shared: the masking value
// depends on (v.n) only, the zero test on sign(v.m * v.n).
template <class Fp>
RLS_DEV float ggx_G1_value(Fp &fp, const Ggx &g, float VdotN)
{
    float cosSqr = sqr(VdotN);
    float tanSqr = fp.rcp(cosSqr) - 1.0f;
    float denominator = 1.0f + fp.sqrt(1.0f + sqr(g.rough) * tanSqr);
    return fp.div(2.0f, denominator);
}
template <class Fp>
RLS_DEV float ggx_G1(Fp &fp, const Ggx &g, f3 v, f3 m, f3 n)
{
    float VdotM = dot(v, m);
    float VdotN = dot(v, n);
    if (VdotM * VdotN < 0.0f) return 0.0f;
    return ggx_G1_value(fp, g, VdotN);
}
// src/rlGgx.h:332-340
template <class Fp>
RLS_DEV float ggx_D(Fp &fp, const Ggx &g, f3 m)
{
    float MdotU = dot(m, g.U);
    float MdotV = dot(m, g.V);
    float MdotN2 = sqr(dot(g.N, m));
    float denominator = g.ax * g.ay * sqr(sqr(fp.div(MdotU, g.ax)) + sqr(fp.div(MdotV, g.ay)) + MdotN2);
    return fp.div(kInvPi, denominator);
}
// src/rlGgx.h:304-313
template <class Fp>
RLS_DEV float ggx_reflection(Fp &fp, const Ggx &g, f3 i, f3 o, f3 n)
{
    f3 hr = normalize(fp, o + i) * sgn_m(dot(i, n));
    float reflectWeight = ggx_fresnel(fp, g, i, hr);
    float LdotN = abs_m(dot(o, n));
    float VdotN = abs_m(dot(i, n));
    float G = ggx_G1(fp, g, i, hr, n) * ggx_G1(fp, g, o, hr, n);
    return fp.div_pz(reflectWeight * G * ggx_D(fp, g, hr) * 0.25f, LdotN * VdotN);   // F or G may be 0
}
// src/rlGgx.h:316-328
template <class Fp>
RLS_DEV float ggx_refraction(Fp &fp, const Ggx &g, f3 i, f3 o, f3 n)
{
    f3 ht = -normalize(fp, i * g.iorIn + o * g.iorOut);
    float refractWeight = 1.0f - ggx_fresnel(fp, g, i, ht);
    float OdotN = abs_m(dot(o, n));
    float IdotN = abs_m(dot(i, n));
    float OdotH = dot(o, ht);
    float IdotH = dot(i, ht);
    float denominator = OdotN * IdotN * sqr(g.iorIn * IdotH + g.iorOut * OdotH);
    float G = ggx_G1(fp, g, i, ht, n) * ggx_G1(fp, g, o, ht, n);
    return fp.div_pz(abs_m(OdotH * IdotH) * sqr(g.iorOut) * refractWeight * G * ggx_D(fp, g, ht), denominator);
}
// src/rlGgx.h:277-291 (eta is NOT squared, as in the reference)
template <class Fp>
RLS_DEV bool ggx_refract_direction(Fp &fp, const Ggx &g, f3 m, f3 i, f3 &dir)
{
    float sign = sgn_m(dot(i, g.N));
    float IdotM = dot(i, m);
    float eta = fp.div(g.iorIn, g.iorOut);
    float cosThetaTSqr = 1.0f + eta * (sqr(IdotM) - 1.0f);
    if (cosThetaTSqr < 0.0f) return false;
    float s = eta * IdotM - sign * fp.sqrt(cosThetaTSqr);
    dir = m * s - i * eta;
    return true;
}
// src/rlGgx.h:294-301
template <class Fp>
RLS_DEV float ggx_sample_weight(Fp &fp, const Ggx &g, f3 i, f3 o, f3 m)
{
    float IdotH = dot(i, m);
    float MdotN = abs_m(dot(m, g.N));
    float IdotN = abs_m(dot(i, g.N));
    float G = ggx_G1(fp, g, i, m, g.N) * ggx_G1(fp, g, o, m, g.N);
    return G * abs_m(fp.div(IdotH, IdotN * MdotN));
}
// src/rlGgx.h:110-119,158-165
template <class Fp>
RLS_DEV f3 ggx_eval_brdf(Fp &fp, const Ggx &g, f3 L)
{
    if (is_zero(L)) return mk3(0.0f, 0.0f, 0.0f);
    if (abs_m(g.ks.x) < kEps && abs_m(g.ks.y) < kEps && abs_m(g.ks.z) < kEps) return mk3(0.0f, 0.0f, 0.0f);
    float refl = ggx_reflection(fp, g, g.wo, L, g.N);
    return g.ks * refl * dot(L, g.N);
}
// src/rlGgx.h:121-127 with VNDFKernel::evalPdf :72-80 (floored at AI_EPSILON, no zero-L guard)
template <class Fp>
RLS_DEV float ggx_eval_pdf(Fp &fp, const Ggx &g, f3 L)
{
    f3 H = normalize(fp, g.wo + L);
    if (g.ndf) {                                  // NDFKernel::evalPdf src/rlGgx.h:45-50, no floor
        float IdotM = abs_m(dot(g.wo, H));
        float MdotN = abs_m(dot(H, g.N));
        return fp.div_pz(ggx_D(fp, g, H) * MdotN * 0.25f, IdotM);
    }
    float IdotN = abs_m(dot(g.wo, g.N));
    float pdf = fp.div_pz(ggx_D(fp, g, H) * ggx_G1(fp, g, g.wo, H, g.N), IdotN) * 0.25f;
    return max_m(pdf, kEps);
}

RLS_DEV uint32_t bsdf_flags(f3 L, f3 N, f3 f, float pdf)
{
    uint32_t fl = 0;
    if (is_zero(L)) fl |= 0x0001u;                 // RLS_FLAG_ZERO_L
    if (dot(L, N) <= 0.0f) fl |= 0x0002u;          // RLS_FLAG_BELOW_HORIZON
    if (pdf == 0.0f) fl |= 0x0004u;                // RLS_FLAG_PDF_ZERO
    if (is_zero(f)) fl |= 0x0008u;                 // RLS_FLAG_F_BLACK
    if (pdf == kEps) fl |= 0x0040u;                // RLS_FLAG_PDF_FLOORED
    return fl;
}

} // namespace rls
