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
RLS_DEV f3 spherical_direction(float cosTheta, float phi)
{
    float s, c;
    rlm::sincosf_(phi, &s, &c);
    float r = sqrtf(1.0f - sqr(cosTheta));
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
RLS_DEV f2 concentric_disk_sample(float rx, float ry)
{
    rx = rx * 2.0f - 1.0f;
    ry = ry * 2.0f - 1.0f;
    f2 o; o.x = 0.0f; o.y = 0.0f;
    if (rx == 0.0f && ry == 0.0f) return o;
    float r, phi;
    if (abs_m(rx) > abs_m(ry)) {
        r = rx;
        phi = kHalfPi * 0.5f * ry / rx;
    } else {
        r = ry;
        phi = kHalfPi * (1.0f - 0.5f * rx / ry);
    }
    float s, c;
    rlm::sincosf_(phi, &s, &c);
    o.x = r * c;
    o.y = r * s;
    return o;
}

// ------------------------------------------- visible-normal sampling (Heitz-d'Eon)
// src/rlGgx.cpp:18-25
RLS_DEV f2 uniform_slope(float rx, float ry)
{
    float r = sqrtf(rx / (1.0f - rx));
    float phi = kTwoPi * ry;
    float s, c;
    rlm::sincosf_(phi, &s, &c);
    f2 o; o.x = r * c; o.y = r * s;
    return o;
}
// src/rlGgx.cpp:14-61 (VNDFKernel::sampleSlope); rlDisney.cpp:416-463 is the same code.
RLS_DEV f2 sample_slope(float theta, float rx, float ry)
{
    if (theta < kEps) return uniform_slope(rx, ry);

    float B = rlm::tanf_(theta);
    float B2 = sqr(B);
    float G1 = 2.0f * (1.0f / (1.0f + sqrtf(1.0f + B2)));   // == 2/(..) bitwise: the divisor is in [2, 2^64]

    float A = 2.0f * rx / G1 - 1.0f;
    float A2 = sqr(A);
    if (abs_m(A2 - 1.0f) < kEps) return uniform_slope(rx, ry);

    float tmp = 1.0f / (A2 - 1.0f);
    float D = sqrtf(max_m(0.0f, B2 * sqr(tmp) - (A2 - B2) * tmp));
    float slopeX1 = B * tmp - D;
    float slopeX2 = B * tmp + D;
    f2 slope;
    slope.x = (A < 0.0f || slopeX2 > 1.0f / B) ? slopeX1 : slopeX2;

    float sign = 1.0f;
    if (ry > 0.5f) {
        ry = 2.0f * (ry - 0.5f);
    } else {
        sign = -1.0f;
        ry = 2.0f * (0.5f - ry);
    }
    float z = (ry * (ry * (ry * 0.27385f - 0.73369f) + 0.46341f))
            / (ry * (ry * (ry * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    slope.y = sign * z * sqrtf(1.0f + sqr(slope.x));
    return slope;
}
// src/rlGgx.cpp:63-99 (VNDFKernel::evalSample); rlDisney.cpp:467-502 is the same code.
RLS_DEV f3 sample_visible_normal(f3 view, f3 U, f3 Vax, f3 N, float ax, float ay, float rx, float ry)
{
    float cosThetaV = clamp_m(dot(N, view), -1.0f, 1.0f);
    float phiV = rlm::atan2f_(dot(Vax, view), dot(U, view));
    f3 V = spherical_direction(cosThetaV, phiV);

    V.x *= ax;
    V.y *= ay;
    V = normalize(V);

    float theta = 0.0f, phi = 0.0f;
    if (V.z < (1.0f - kEps)) {
        theta = rlm::acosf_(V.z);
        phi = rlm::atan2f_(V.y, V.x);
    }
    f2 slope = sample_slope(theta, rx, ry);

    float sinPhi, cosPhi;
    rlm::sincosf_(phi, &sinPhi, &cosPhi);
    f3 omega;
    omega.x = -(cosPhi * slope.x - sinPhi * slope.y) * ax;
    omega.y = -(sinPhi * slope.x + cosPhi * slope.y) * ay;
    omega.z = 1.0f;
    return normalize(rotate_to_frame(omega, U, Vax, N));
}

// src/rlGgx.h:33-41 (NDFKernel::evalSample, Burley Eq.14): plain NDF sampling; also
// DisneySampler::sampleGTR2AnisoDirection (src/rlDisney.cpp:406-414) with (rx, ry) swapped.
RLS_DEV f3 sample_ndf_normal(f3 U, f3 Vax, f3 N, float ax, float ay, float rx, float ry)
{
    float g = sqrtf(rx / (1.0f - rx));
    float phi = kTwoPi * ry;
    float s, c;
    rlm::sincosf_(phi, &s, &c);
    f3 omega = mk3(g * ax * c, g * ay * s, 1.0f);
    return normalize(rotate_to_frame(omega, U, Vax, N));
}

// ------------------------------------------------------------------------ rlGgx
struct Ggx {
    f3 U, V, N, wo, ks;
    float iorIn, iorOut, rough, ax, ay;
    bool entering;
    bool ndf;        // GgxSamplerT<NDFKernel> instead of the shipped GgxSamplerT<VNDFKernel>
};
RLS_DEV f3 ggx_sample_normal(const Ggx &g, float rx, float ry)
{
    if (g.ndf) return sample_ndf_normal(g.U, g.V, g.N, g.ax, g.ay, rx, ry);
    return sample_visible_normal(g.wo, g.U, g.V, g.N, g.ax, g.ay, rx, ry);
}

// src/rlGgx.h:130-156 (GgxSamplerT ctor)
RLS_DEV void ggx_init(Ggx &g, const Shading &sh, f3 ks, float ior, float roughness, float aniso)
{
    f3 Ngeo = sh.backfacing ? -sh.N : sh.N;      // sg->N
    f3 Rd = -sh.wo;                              // sg->Rd
    g.entering = dot(Ngeo, Rd) < kEps;           // :137
    float a = 1.0f, b = max_m(ior, 1e-4f);       // :138-139
    g.iorIn = g.entering ? a : b;                // :140-142 (swap)
    g.iorOut = g.entering ? b : a;
    g.wo = sh.wo;                                // :144  -(-wo) is exact
    g.U = sh.U; g.V = sh.V; g.N = sh.N;          // :145-146, explicit frame
    float aspect = sqrtf(1.0f - aniso * 0.9f);   // :148
    g.ax = max_m(1e-4f, sqr(roughness) / aspect);
    g.ay = max_m(1e-4f, sqr(roughness) * aspect);
    g.rough = max_m(1e-5f, sqr(roughness));      // :155
    g.ks = ks;
    g.ndf = false;
}
// src/rlGgx.h:249-270
RLS_DEV float ggx_fresnel(const Ggx &g, f3 i, f3 m)
{
    float c = abs_m(dot(i, m));
    float gSqr = sqr(g.iorOut / g.iorIn) - 1.0f + c * c;
    if (gSqr < 0.0f) return 1.0f;
    float gg = sqrtf(gSqr);
    float gmc = gg - c;
    float gpc = gg + c;
    return 0.5f * sqr(gmc / gpc) * (1.0f + sqr((c * gpc - 1.0f) / (c * gmc + 1.0f)));
}
// src/rlGgx.h:343-357, split so the view-only half can be shared: the masking value
// depends on (v.n) only, the zero test on sign(v.m * v.n).
RLS_DEV float ggx_G1_value(const Ggx &g, float VdotN)
{
    float cosSqr = sqr(VdotN);
    float tanSqr = 1.0f / cosSqr - 1.0f;
    float denominator = 1.0f + sqrtf(1.0f + sqr(g.rough) * tanSqr);
    return 2.0f / denominator;
}
RLS_DEV float ggx_G1(const Ggx &g, f3 v, f3 m, f3 n)
{
    float VdotM = dot(v, m);
    float VdotN = dot(v, n);
    if (VdotM * VdotN < 0.0f) return 0.0f;
    return ggx_G1_value(g, VdotN);
}
// src/rlGgx.h:332-340
RLS_DEV float ggx_D(const Ggx &g, f3 m)
{
    float MdotU = dot(m, g.U);
    float MdotV = dot(m, g.V);
    float MdotN2 = sqr(dot(g.N, m));
    float denominator = g.ax * g.ay * sqr(sqr(MdotU / g.ax) + sqr(MdotV / g.ay) + MdotN2);
    return kInvPi / denominator;
}
// src/rlGgx.h:304-313
RLS_DEV float ggx_reflection(const Ggx &g, f3 i, f3 o, f3 n)
{
    f3 hr = normalize(o + i) * sgn_m(dot(i, n));
    float reflectWeight = ggx_fresnel(g, i, hr);
    float LdotN = abs_m(dot(o, n));
    float VdotN = abs_m(dot(i, n));
    float G = ggx_G1(g, i, hr, n) * ggx_G1(g, o, hr, n);
    return reflectWeight * G * ggx_D(g, hr) * 0.25f / (LdotN * VdotN);
}
// src/rlGgx.h:316-328
RLS_DEV float ggx_refraction(const Ggx &g, f3 i, f3 o, f3 n)
{
    f3 ht = -normalize(i * g.iorIn + o * g.iorOut);
    float refractWeight = 1.0f - ggx_fresnel(g, i, ht);
    float OdotN = abs_m(dot(o, n));
    float IdotN = abs_m(dot(i, n));
    float OdotH = dot(o, ht);
    float IdotH = dot(i, ht);
    float denominator = OdotN * IdotN * sqr(g.iorIn * IdotH + g.iorOut * OdotH);
    float G = ggx_G1(g, i, ht, n) * ggx_G1(g, o, ht, n);
    return abs_m(OdotH * IdotH) * sqr(g.iorOut) * refractWeight * G * ggx_D(g, ht) / denominator;
}
// src/rlGgx.h:277-291 (eta is NOT squared, as in the reference)
RLS_DEV bool ggx_refract_direction(const Ggx &g, f3 m, f3 i, f3 &dir)
{
    float sign = sgn_m(dot(i, g.N));
    float IdotM = dot(i, m);
    float eta = g.iorIn / g.iorOut;
    float cosThetaTSqr = 1.0f + eta * (sqr(IdotM) - 1.0f);
    if (cosThetaTSqr < 0.0f) return false;
    float s = eta * IdotM - sign * sqrtf(cosThetaTSqr);
    dir = m * s - i * eta;
    return true;
}
// src/rlGgx.h:294-301
RLS_DEV float ggx_sample_weight(const Ggx &g, f3 i, f3 o, f3 m)
{
    float IdotH = dot(i, m);
    float MdotN = abs_m(dot(m, g.N));
    float IdotN = abs_m(dot(i, g.N));
    float G = ggx_G1(g, i, m, g.N) * ggx_G1(g, o, m, g.N);
    return G * abs_m(IdotH / (IdotN * MdotN));
}
// src/rlGgx.h:110-119,158-165
RLS_DEV f3 ggx_eval_brdf(const Ggx &g, f3 L)
{
    if (is_zero(L)) return mk3(0.0f, 0.0f, 0.0f);
    if (abs_m(g.ks.x) < kEps && abs_m(g.ks.y) < kEps && abs_m(g.ks.z) < kEps) return mk3(0.0f, 0.0f, 0.0f);
    float refl = ggx_reflection(g, g.wo, L, g.N);
    return g.ks * refl * dot(L, g.N);
}
// src/rlGgx.h:121-127 with VNDFKernel::evalPdf :72-80 (floored at AI_EPSILON, no zero-L guard)
RLS_DEV float ggx_eval_pdf(const Ggx &g, f3 L)
{
    f3 H = normalize(g.wo + L);
    if (g.ndf) {                                  // NDFKernel::evalPdf src/rlGgx.h:45-50, no floor
        float IdotM = abs_m(dot(g.wo, H));
        float MdotN = abs_m(dot(H, g.N));
        return ggx_D(g, H) * MdotN * 0.25f / IdotM;
    }
    float IdotN = abs_m(dot(g.wo, g.N));
    float pdf = ggx_D(g, H) * ggx_G1(g, g.wo, H, g.N) / IdotN * 0.25f;
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
