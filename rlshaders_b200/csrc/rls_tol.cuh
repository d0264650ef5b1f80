// rls_tol.cuh -- the TOLERANCE arithmetic policy (RLS_ARITH_TOLERANT) of the fused units.
//
// The two bit-exact policies (rls_fp.cuh) reproduce the reference's binary32 operations one by one and are
// instruction-issue bound at ~0.4 of the HBM roofline.  This policy computes the SAME functions
// (src/rlGgx.h:130-357, src/rlGgx.cpp:14-99, src/rlDisney.cpp:155-577, src/rlSss.cpp:20-106) to within a stated
// tolerance instead of to the bit:
//   * multiply-adds are fused (the translation unit rls_tol.cu is compiled with FMA contraction);
//   * 1/x, sqrt, 1/sqrt, exp2 are the MUFU approximations (<= 1-2 ulp), log / sin / cos are short FMA polynomials
//     (<= 1-2 ulp); no binary64 anywhere;
//   * visible-normal sampling is evaluated ALGEBRAICALLY: the reference goes view -> atan2f -> sincosf -> stretch ->
//     acosf / atan2f -> tanf -> ... -> sincosf (src/rlGgx.cpp:63-99), i.e. it converts directions to angles and back
//     four times.  cos / sin of an atan2 are the normalised components, tan(acos(z)) = sqrt(1 - z^2) / z: six
//     transcendentals become three reciprocal square roots.  Likewise the half vector of a direction obtained by
//     reflecting about m IS m (src/rlGgx.h:304-313 re-derives it with a normalize), and the refraction half vector
//     -normalize(eta_i i + eta_o o) of a direction refracted about m is -+m (src/rlGgx.h:316-328).
//   * quantities the reference forms by CANCELLATION from raw inputs are formed by the same unfused operations in the
//     same order (mul_rn / add_rn below), so that the reference's own rounding noise is reproduced where that is
//     cheap: N.wo, 1 - cos^2, the slope_y rational polynomial, 1 - x C of the profile's inverse CDF.
//
// FLAGS STAY BIT-EXACT.  Every comparison that decides a flag bit, a lobe or a discontinuous choice registers its
// comparand with a Bands tracker: when the comparand lies within a band of its threshold (a band wider than the
// error this policy can have there) the sample is appended to a re-run list and a second kernel evaluates it with the
// bit-exact policy (rls_b200.cu: k_*_rerun).  Comparands formed from raw inputs by the reference's exact operations
// need no band.  tests/native/tol_check.cpp runs these very functions on the CPU against the oracle (the header is
// __host__ __device__) and prints flag mismatches, error percentiles and the re-run fraction.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RLT_HD __host__ __device__ __forceinline__
#define RLT_M  __host__ __device__ __forceinline__      /* member functions */
#else
#define RLT_HD static inline
#define RLT_M  inline
#endif

namespace rls {
namespace tol {

constexpr float kEps = 1.0e-4f;                       // AI_EPSILON
constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kHalfPi = 1.57079632679489661923f;
constexpr float kInvPi = 0.31830988618379067154f;

// ------------------------------------------------------------------ bit casts
RLT_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RLT_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// ------------------------------------------------------------------ primitive operations
// Host build (tests/native/tol_check.cpp): the approximations are emulated by the correctly rounded result, optionally
// moved by a pseudo-random -1 / 0 / +1 ulp (RLS_TOL_EMULATE_ULP) so that the bands are exercised with MUFU-sized errors.
#if !defined(__CUDA_ARCH__)
RLT_HD float emu_(float y)
{
#if defined(RLS_TOL_EMULATE_ULP)
    uint32_t u = f2u(y);
    if ((u & 0x7f800000u) == 0x7f800000u || (u & 0x7fffffffu) == 0u) return y;
    uint32_t h = u * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const uint32_t k = h % 3u;
    return u2f(u + k - 1u);
#else
    return y;
#endif
}
#endif
RLT_HD float rcp(float x)
{
#if defined(__CUDA_ARCH__)
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return emu_(1.0f / x);
#endif
}
RLT_HD float rsq(float x)
{
#if defined(__CUDA_ARCH__)
    float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return emu_((float)(1.0 / sqrt((double)x)));
#endif
}
RLT_HD float sqrt_(float x)
{
#if defined(__CUDA_ARCH__)
    float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return emu_(sqrtf(x));
#endif
}
RLT_HD float ex2(float x)
{
#if defined(__CUDA_ARCH__)
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return emu_((float)exp2((double)x));
#endif
}
RLT_HD float fma_(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
// One IEEE operation that the compiler may NOT contract with its neighbours (the reference's own rounding).
RLT_HD float mul_rn(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    float p = a * b; __asm__ volatile("" : "+x"(p)); return p;
#endif
}
RLT_HD float add_rn(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    float p = a + b; __asm__ volatile("" : "+x"(p)); return p;
#endif
}
RLT_HD float sub_rn(float a, float b) { return add_rn(a, -b); }
// Correctly rounded sqrt and reciprocal (the reference's sqrtf and 1.0f / x), for the few comparands that are formed on
// the reference's own operations next to a threshold.
RLT_HD float sqrt_rn(float x)
{
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    float p = sqrtf(x); __asm__ volatile("" : "+x"(p)); return p;
#endif
}
RLT_HD float rcp_rn(float x)
{
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    float p = 1.0f / x; __asm__ volatile("" : "+x"(p)); return p;
#endif
}
RLT_HD float div(float a, float b) { return a * rcp(b); }
RLT_HD float sqr(float a) { return a * a; }
RLT_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
RLT_HD float lerp(float t, float a, float b) { return fma_(b, t, (1.0f - t) * a); }        // LERP(t, a, b)

// sin and cos of |x| <~ 100 (the path's arguments are angles in [-2pi, 2pi]): Cody-Waite reduction by pi/2 in two
// FMA steps, degree-7 / degree-8 minimax polynomials (Cephes sinf / cosf coefficients), <= 1.5 ulp.
RLT_HD void sincos_(float x, float *sp, float *cp)
{
    const float j = rintf(x * 0.63661977236758134308f);
    float r = fma_(j, -1.57079637050628662109375f, x);
    r = fma_(j, 4.37113900018624283e-8f, r);
    const float z = r * r;
    float s = fma_(-1.9515295891e-4f, z, 8.3321608736e-3f);
    s = fma_(s, z, -1.6666654611e-1f);
    s = fma_(s * z, r, r);
    float c = fma_(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    c = fma_(c, z, 4.166664568298827e-2f);
    c = fma_(c * z, z, fma_(-0.5f, z, 1.0f));
    const int q = (int)j;
    const float ss = (q & 1) ? c : s, cc = (q & 1) ? s : c;
    *sp = (q & 2) ? -ss : ss;
    *cp = ((q + 1) & 2) ? -cc : cc;
}
// log(x) for positive normal x: x = 2^e m, m in [sqrt(1/2), sqrt(2)), Cephes logf polynomial in m - 1 (<= 1 ulp,
// accurate RELATIVE to log(x) near x = 1, where lg2.approx is only accurate absolutely).
RLT_HD float log_(float x)
{
    uint32_t ix = f2u(x);
    const uint32_t off = ix - 0x3f3504f3u;                       // bits(sqrt(1/2))
    const int e = (int)off >> 23;
    const float m = u2f(ix - ((uint32_t)e << 23));               // = (off & 0x007fffff) + bits(sqrt(1/2))
    const float t = m - 1.0f;
    const float z = t * t;
    float p = fma_(7.0376836292e-2f, t, -1.1514610310e-1f);
    p = fma_(p, t, 1.1676998740e-1f);
    p = fma_(p, t, -1.2420140846e-1f);
    p = fma_(p, t, 1.4249322787e-1f);
    p = fma_(p, t, -1.6668057665e-1f);
    p = fma_(p, t, 2.0000714765e-1f);
    p = fma_(p, t, -2.4999993993e-1f);
    p = fma_(p, t, 3.3333331174e-1f);
    float y = p * t * z;
    const float fe = (float)e;
    y = fma_(fe, -2.12194440e-4f, y);
    y = fma_(-0.5f, z, y);
    return fma_(fe, 0.693359375f, t + y);
}
RLT_HD float exp_(float x) { return ex2(x * 1.44269504088896340736f); }
// x^y for x > 0: exp2(y log2 x) with log2 from log_ (relative accuracy also near x = 1)
RLT_HD float pow_(float x, float y) { return ex2(y * (log_(x) * 1.44269504088896340736f)); }
RLT_HD float pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }

// ------------------------------------------------------------------ vectors
struct v3 { float x, y, z; };
RLT_HD v3 mk(float x, float y, float z) { v3 v; v.x = x; v.y = y; v.z = z; return v; }
RLT_HD v3 operator+(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
RLT_HD v3 operator-(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
RLT_HD v3 operator*(v3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
RLT_HD float dot(v3 a, v3 b) { return fma_(a.x, b.x, fma_(a.y, b.y, a.z * b.z)); }
// AiV3Dot in the reference's own operation order: (ax bx + ay by) + az bz, every operation rounded
RLT_HD float dot_rn(v3 a, v3 b) { return add_rn(add_rn(mul_rn(a.x, b.x), mul_rn(a.y, b.y)), mul_rn(a.z, b.z)); }
RLT_HD v3 normalize(v3 a) { const float i = rsq(dot(a, a)); return a * i; }
RLT_HD v3 to_frame(float x, float y, float z, v3 u, v3 v, v3 w)      // AiV3RotateToFrame
{
    return mk(fma_(x, u.x, fma_(y, v.x, z * w.x)), fma_(x, u.y, fma_(y, v.y, z * w.y)), fma_(x, u.z, fma_(y, v.z, z * w.z)));
}

// Base half-width of the bands on comparands that follow the sampled direction (N.L, V.m, cos^2 theta_t, ...); the
// sample's own estimate of the reference's rounding noise is added to it.
#ifndef RLS_TOL_BAND_BASE
#define RLS_TOL_BAND_BASE 1e-4f      /* 1e-3 until round 2: 0.22 % of the samples re-run; 1e-4: 0.08 % (0 flag mismatches in
                                        2 x 2.7e8 samples on the host build, plain and ulp-perturbed; 3e-5 was clean too) */
#endif
// 1 + B^2 - A^2 (the radicand of the slope root) cancels as rx -> 1.  Below RLS_TOL_U2_GUARD x (1 + B^2) the sample is
// re-run (there the reference's own radicand can round negative: NaN against a number); above it the reference's noise
// in the root, 6e-8 B^2 |tmp| / (|A| u) from the difference B^2 tmp^2 - (A^2 - B^2) tmp, joins the band width with a
// factor 4.  (Guard 1e-3 and no noise term until late in round 2: this one test listed 0.027 % of all samples.)
#ifndef RLS_TOL_U2_GUARD
#define RLS_TOL_U2_GUARD 1e-5f
#endif
#ifndef RLS_TOL_U2_NOISE
#define RLS_TOL_U2_NOISE 2.4e-7f
#endif
// The reference re-derives the half vector as normalize(V + L) with L = 2 |V.m| m - V stored in binary32: V + L cancels to
// 2 |V.m| m plus ~1e-7 of rounding residue, so ITS V.h = V.m +- 1e-7 / (2 |V.m|) -- the sign test on V.h (the masking
// terms, src/rlGgx.h:348) is decided by rounding once |V.m| falls below ~5e-4.  Half-width to add to a band on V.m.
// (Found by tests/hunts/tol_stress_hunt.py with rx = 1 - 2^-24, which puts m in the tangent plane; the 1e-3 base width of
// the round's first half had covered it by accident.)
#define RLS_TOL_HALF_VECTOR_NOISE(vm) (4e-7f * rcp(fmaxf(fabsf(vm), 1e-4f)))
// ------------------------------------------------------------------ the band tracker
#if defined(RLS_TOL_BAND_STATS) && !defined(__CUDACC__)
// Host-only diagnostics (tests/native/tol_host.cpp -DRLS_TOL_BAND_STATS): which band sends a sample to the re-run FIRST,
// by source line of the near() / require() call.
extern "C" void rls_tol_band_hit(int line);
struct Bands {
    bool rerun;
    Bands() : rerun(false) {}
    void near(float x, float t, float width, int line = __builtin_LINE())
    {
        const bool hit = !(fabsf(x - t) > width);
        if (hit && !rerun) rls_tol_band_hit(line);
        rerun = rerun || hit;
    }
    void require(bool cond, int line = __builtin_LINE())
    {
        if (!cond && !rerun) rls_tol_band_hit(line);
        rerun = rerun || !cond;
    }
};
#else
struct Bands {
    bool rerun;
    RLT_M Bands() : rerun(false) {}
    // comparand x decides a flag / lobe / discontinuous choice at threshold t; `width` bounds this policy's error there
    RLT_M void near(float x, float t, float width) { rerun = rerun || !(fabsf(x - t) > width); }   // NaN -> re-run
    RLT_M void require(bool cond) { rerun = rerun || !cond; }
};
#endif

// Flag bits (include/rls_b200.h)
constexpr uint32_t kFlagZeroL = 0x0001u, kFlagBelowHorizon = 0x0002u, kFlagPdfZero = 0x0004u, kFlagBlack = 0x0008u,
                   kFlagEntering = 0x0010u, kFlagTir = 0x0020u, kFlagPdfFloored = 0x0040u, kFlagSlopeEarlyOut = 0x0080u;

// ================================================================== visible-normal sampling
// VNDFKernel::evalSample + sampleSlope (src/rlGgx.cpp:14-99; src/rlDisney.cpp:416-502 is the same code).
// vz = N.wo in the reference's operation order.  Returns the world-space microfacet normal; `early` = the
// uniform-slope early-out was taken (src/rlGgx.cpp:27,38); `noise` = an estimate of the REFERENCE's own rounding noise
// in the sampled direction (its two roots of :42-46 are differences of terms ~ 1 / (A^2 - 1): 1 ulp becomes
// ~1e-7 / |A^2 - 1|), which the bands on direction-derived comparands add to their width.
RLT_HD v3 sample_visible_normal(Bands &bd, v3 wo, v3 U, v3 Vax, v3 N, float vz, float ax, float ay, float rx, float ry,
                                bool &early, float &noise)
{
    noise = 0.0f;
    // :66-75  view -> local polar -> sphericalDirection: (r cos phiV, r sin phiV, cz), r = sqrt(1 - cz^2)
    const float vx = dot(U, wo), vy = dot(Vax, wo);
    const float cz = clampf(vz, -1.0f, 1.0f);
    const float r = sqrt_(sub_rn(1.0f, mul_rn(cz, cz)));
    const float h2 = fma_(vx, vx, vy * vy);
    const bool pole = !(h2 > 0.0f);                             // atan2f(0, 0) = 0
    const float ih = pole ? 0.0f : rsq(h2);
    const float cph = pole ? 1.0f : vx * ih, sph = vy * ih;
    // :77-79  stretch, normalize
    const float sx = ax * (r * cph), sy = ay * (r * sph);
    const float q2 = fma_(sx, sx, sy * sy);
    const float in = rsq(fma_(cz, cz, q2));
    float Vz = cz * in;
    // :82  theta = phi = 0 unless V.z < 1 - eps.  The approximate V.z above is within ~3 ulps of the reference's; next to
    // the threshold the comparand is formed again on the reference's own operations (AiV3Normalize: len = sqrtf(x x +
    // y y + z z), inv = 1 / len, z inv).  x and y still carry this policy's error, but there x^2 + y^2 = 2e-4 z^2, so a
    // few ulps in them move len^2 by 1e-10 relative: the result is the reference's V.z but for a rounding boundary
    // crossed in ~0.1 % of these samples, by one ulp, which the 2.5-ulp band covers.  (With the approximate comparand
    // and its 10-ulp band this test listed 0.03-0.04 % of all samples, a third of the re-run list.)
    float vband = 6e-7f;
    if (fabsf(Vz - (1.0f - kEps)) < 1e-5f) {
        const float len = sqrt_rn(add_rn(add_rn(mul_rn(sx, sx), mul_rn(sy, sy)), mul_rn(cz, cz)));
        Vz = mul_rn(cz, rcp_rn(len));
        vband = 1.5e-7f;
    }
    const bool along = !(Vz < 1.0f - kEps);
    bd.near(Vz, 1.0f - kEps, vband);
    bd.require(cz > kEps);          // views at / below the horizon: tanf(acosf(.)) of the reference near its pole
    float slx, sly, cosPhi, sinPhi;
    bool uniform = along;
    float A = 0.0f, B = 0.0f, S = 0.0f, A2 = 0.0f;
    if (!along) {
        // :83-84  tan(theta) = |V.xy| / V.z, (cos phi, sin phi) = V.xy / |V.xy|
        const float iq = rsq(q2);
        cosPhi = sx * iq; sinPhi = sy * iq;
        B = (q2 * iq) * rcp(cz);
        // sampleSlope :29-38
        S = sqrt_(fma_(B, B, 1.0f));
        A = fma_(rx, 1.0f + S, -1.0f);                          // 2 rx / G1 - 1,  G1 = 2 / (1 + S)
        A2 = A * A;
        const float dA = fabsf(A2 - 1.0f);
        uniform = dA < kEps;
        bd.near(dA, kEps, 4e-6f * (1.0f + S));
    } else {
        cosPhi = 1.0f; sinPhi = 0.0f;
    }
    early = uniform;
    if (uniform) {                                              // :18-25
        const float ru = sqrt_(div(rx, 1.0f - rx));
        float s, c;
        sincos_(kTwoPi * ry, &s, &c);
        slx = ru * c; sly = ru * s;
    } else {
        // :40-46.  With tmp = 1 / (A^2 - 1), u = sqrt(1 + B^2 - A^2): D = sqrt(B^2 tmp^2 - (A^2 - B^2) tmp) = |A| u |tmp|,
        // and slopeX2 > 1/B  <=>  A > 1 (for 0 <= A < 1 the second root stays below 1/B, for A > 1 above it; they meet
        // only where D = 0).  The root the reference picks is therefore tmp (B - A u) for every A -- for A >= 0 a
        // difference that cancels as A -> 1, evaluated here in its conjugate form (A^2 - B^2) / (B + A u), which is
        // smooth through A = 1.
        const float tmp = rcp(A2 - 1.0f);
        const float u2 = fma_(B, B, 1.0f - A2);                 // 1 + B^2 - A^2 >= 0; cancels as rx -> 1 (A -> S)
        bd.require(u2 > RLS_TOL_U2_GUARD * (S * S));
        const float uu = sqrt_(fmaxf(0.0f, u2));
        const float Au = A * uu;
        slx = (A < 0.0f) ? (B - Au) * tmp : (A2 - B * B) * rcp(B + Au);
        noise = 4e-7f * fabsf(tmp) * (1.0f + B) + RLS_TOL_U2_NOISE * (B * B) * fabsf(tmp) * rcp(fabsf(A) * uu);
        // :48-58  slope_y: the rational fit on the reference's own operations (its denominator cancels to 5e-4)
        const bool up = ry > 0.5f;
        const float t = up ? 2.0f * (ry - 0.5f) : 2.0f * (0.5f - ry);
        const float num = mul_rn(t, add_rn(mul_rn(t, sub_rn(mul_rn(t, 0.27385f), 0.73369f)), 0.46341f));
        const float den = add_rn(mul_rn(t, sub_rn(mul_rn(t, add_rn(mul_rn(t, 0.093073f), 0.309420f)), 1.0f)), 0.597999f);
        const float z = div(num, den);
        sly = (up ? z : -z) * sqrt_(fma_(slx, slx, 1.0f));
    }
    // :91-98  rotate by phi, unstretch, to the world frame, normalize
    const float ox = -(cosPhi * slx - sinPhi * sly) * ax;
    const float oy = -(sinPhi * slx + cosPhi * sly) * ay;
    return normalize(to_frame(ox, oy, 1.0f, U, Vax, N));
}

// ================================================================== rlGgx
struct GgxT {
    v3 U, V, N, wo;
    float vz;             // N.wo, reference operation order
    float b;              // max(ior, 1e-4)
    float ax, ay, a2g;    // alpha_x, alpha_y; mRoughness^2 (the isotropic alpha G1 uses, src/rlGgx.h:155,355)
    bool entering;
};
// src/rlGgx.h:130-156
RLT_HD void ggx_init(GgxT &g, v3 U, v3 V, v3 N, v3 wo, bool backfacing, float ior, float roughness, float aniso)
{
    g.U = U; g.V = V; g.N = N; g.wo = wo;
    g.vz = dot_rn(wo, N);                                       // :137 dot(sg->N, sg->Rd) = -+ this, exactly
    g.entering = (backfacing ? g.vz : -g.vz) < kEps;
    g.b = ior > 1e-4f ? ior : 1e-4f;
    const float r2 = roughness * roughness;
    if (aniso == 0.0f) {
        g.ax = g.ay = fmaxf(1e-4f, r2);
    } else {
        const float aspect = sqrt_(fma_(aniso, -0.9f, 1.0f));
        g.ax = fmaxf(1e-4f, div(r2, aspect));
        g.ay = fmaxf(1e-4f, r2 * aspect);
    }
    const float rg = fmaxf(1e-5f, r2);
    g.a2g = rg * rg;
}
// src/rlGgx.h:249-270 given c = |i.m| and ratio2 = (mIorOut / mIorIn)^2
RLT_HD float fresnel_c(float ratio2, float c)
{
    const float gSqr = fma_(c, c, ratio2 - 1.0f);
    const float gg = sqrt_(fmaxf(gSqr, 0.0f));
    const float gmc = gg - c, gpc = gg + c;
    const float a = div(gmc, gpc), bq = div(fma_(c, gpc, -1.0f), fma_(c, gmc, 1.0f));
    const float v = 0.5f * (a * a) * fma_(bq, bq, 1.0f);
    return gSqr < 0.0f ? 1.0f : v;                              // continuous at g^2 = 0: no band
}
// src/rlGgx.h:343-357, the value part: 2 / (1 + sqrt(1 + alpha^2 tan^2))
RLT_HD float G1_value(float a2g, float cosv)
{
    const float t2 = rcp(cosv * cosv) - 1.0f;
    return 2.0f * rcp(1.0f + sqrt_(fma_(a2g, t2, 1.0f)));
}
// src/rlGgx.h:332-340
RLT_HD float ggx_D(const GgxT &g, v3 m, float mN)
{
    const float a = dot(m, g.U) * rcp(g.ax), b = dot(m, g.V) * rcp(g.ay);
    const float t = fma_(a, a, fma_(b, b, mN * mN));
    return kInvPi * rcp(g.ax * g.ay * (t * t));
}

struct DielectricT { float F, f_r, pdf_r, f_t, w_t; v3 wi_r, wi_t; uint32_t flags; };

// The rough-dielectric unit (same composition as rls_fused.cuh dielectric_unit / both oracles).
// ior_band = false (the albedo sweep, whose table reads neither F_BLACK nor f_t): no band around ior = 1.
RLT_HD DielectricT dielectric_unit(Bands &bd, v3 U, v3 V, v3 N, v3 wo, bool backfacing, float ior, float roughness,
                                   float aniso, float rx, float ry, bool ior_band = true)
{
    DielectricT r;
    GgxT g;
    ggx_init(g, U, V, N, wo, backfacing, ior, roughness, aniso);
    const float VdotN = g.vz, absVN = fabsf(VdotN);
    const float sgnV = VdotN < 0.0f ? -1.0f : (VdotN > 0.0f ? 1.0f : 0.0f);
    const float invB = rcp(g.b);
    const float ratio = g.entering ? g.b : invB;                // mIorOut / mIorIn
    const float eta = g.entering ? invB : g.b;                  // mIorIn / mIorOut
    const float iorIn = g.entering ? 1.0f : g.b, iorOut = g.entering ? g.b : 1.0f;
    const float ratio2 = ratio * ratio;
    if (ior_band) bd.near(g.b, 1.0f, 1e-4f);                    // ior == 1: F == 0 and the zero half vector are rounding-decided

    bool early;
    float noise;
    const v3 m = sample_visible_normal(bd, wo, U, V, N, g.vz, g.ax, g.ay, rx, ry, early, noise);
    const float Vm = dot(wo, m), aVm = fabsf(Vm);
    const float mN = dot(m, N);
    const float band = RLS_TOL_BAND_BASE + noise;                           // on comparands that follow the sampled direction
    bd.near(Vm, 0.0f, band + RLS_TOL_HALF_VECTOR_NOISE(Vm));    // sign of V.m decides the masking terms
    // reflectDirection(V, m) = 2|V.m| m - V; its half vector with V is m, V.H = V.m, L.H = 2|V.m| - V.m
    r.wi_r = m * (2.0f * aVm) - wo;
    const float LH = 2.0f * aVm - Vm;
    r.F = fresnel_c(ratio2, fabsf(LH));                         // fresnel(L, m), c = |L.m|
    const float LdotN = dot(r.wi_r, N);
    bd.near(LdotN, 0.0f, band);
    const float Dm = ggx_D(g, m, mN);
    const float G1v = G1_value(g.a2g, VdotN), G1l = G1_value(g.a2g, LdotN);
    // evalPdf (src/rlGgx.h:121-127, :72-80): max(D G1(V, H) / (4 |V.N|), eps)
    const float i4vn = 0.25f * rcp(absVN);
    const float G1p = (Vm * VdotN < 0.0f) ? 0.0f : G1v;
    const float pdf_raw = Dm * G1p * i4vn;
    bd.near(pdf_raw, kEps, 2e-2f * kEps);
    r.pdf_r = fmaxf(pdf_raw, kEps);
    // evalBrdf (src/rlGgx.h:304-313) x L.N: hr = sgn(V.N) H; F G D / (4 |L.N| |V.N|) * (L.N)
    const float Fh = fresnel_c(ratio2, fabsf(Vm));              // fresnel(V, hr): c = |V.hr|
    const float G1i = ((Vm * sgnV) * VdotN < 0.0f) ? 0.0f : G1v;
    const float G1o = ((LH * sgnV) * LdotN < 0.0f) ? 0.0f : G1l;
    const float sL = LdotN < 0.0f ? -1.0f : (LdotN > 0.0f ? 1.0f : 0.0f);
    r.f_r = (sgnV != 0.0f) ? Fh * (G1i * G1o) * Dm * i4vn * sL : 0.0f;
    uint32_t fl = 0;
    if (LdotN <= 0.0f) fl |= kFlagBelowHorizon;
    if (r.f_r == 0.0f) fl |= kFlagBlack;
    if (r.pdf_r == kEps) fl |= kFlagPdfFloored;
    if (g.entering) fl |= kFlagEntering;
    if (early) fl |= kFlagSlopeEarlyOut;

    // getRefractDirection(m, V) (src/rlGgx.h:277-291; eta is not squared, as in the reference)
    const float cT2 = fma_(eta, fma_(Vm, Vm, -1.0f), 1.0f);
    bd.near(cT2, 0.0f, band * (1.0f + eta));
    float TdotN, G1t, Tm;
    if (cT2 < 0.0f) {                                           // total internal reflection: reflect about m
        fl |= kFlagTir;
        r.wi_t = r.wi_r;
        r.f_t = 0.0f;
        TdotN = LdotN; G1t = G1l; Tm = LH;
    } else {
        const float cT = sqrt_(cT2);
        const float sc = fma_(eta, Vm, -sgnV * cT);
        r.wi_t = m * sc - wo * eta;
        TdotN = dot(r.wi_t, N);
        G1t = G1_value(g.a2g, TdotN);
        Tm = sc - eta * Vm;                                     // T.m = -sgn(V.N) sqrt(cT2)
        // refraction(V, T, N) (src/rlGgx.h:316-328): ht = -normalize(eta_i V + eta_o T) = -sgn(sc) m
        if (ior_band) bd.near(sc, 0.0f, 1e-3f);                 // ht is the reference's rounding residue when sc -> 0
        const float sh = sc < 0.0f ? 1.0f : -1.0f;
        const float IdotH = sh * Vm, OdotH = sh * Tm;
        const float w = fma_(iorIn, IdotH, iorOut * OdotH);
        const float G1ti = (IdotH * VdotN < 0.0f) ? 0.0f : G1v;
        const float G1to = (OdotH * TdotN < 0.0f) ? 0.0f : G1t;
        bd.near(TdotN, 0.0f, band);
        r.f_t = fabsf(OdotH * IdotH) * (iorOut * iorOut) * (1.0f - Fh) * (G1ti * G1to) * Dm *
                rcp(fabsf(TdotN) * absVN * (w * w));
    }
    // getSampleWeight(V, wi_t, m) (src/rlGgx.h:294-301)
    {
        const float G1i2 = (Vm * VdotN < 0.0f) ? 0.0f : G1v;
        const float G1o2 = (Tm * TdotN < 0.0f) ? 0.0f : G1t;
        r.w_t = (G1i2 * G1o2) * fabsf(Vm * rcp(absVN * fabsf(mN)));
    }
    r.flags = fl;
    return r;
}

// Fused rlGgx unit with a KsColor (config 1): ctor + evalSample + evalBrdf + evalPdf (+ the Fresnel term).
struct GgxBsdfT { v3 L, f; float pdf, fresnel; uint32_t flags; };
RLT_HD GgxBsdfT ggx_unit(Bands &bd, v3 U, v3 V, v3 N, v3 wo, bool backfacing, v3 ks, float ior, float roughness, float aniso,
                         float rx, float ry)
{
    GgxBsdfT o;
    GgxT g;
    ggx_init(g, U, V, N, wo, backfacing, ior, roughness, aniso);
    const float VdotN = g.vz, absVN = fabsf(VdotN);
    const float ratio = g.entering ? g.b : rcp(g.b);
    const float ratio2 = ratio * ratio;
    bd.near(g.b, 1.0f, 1e-4f);
    bool early;
    float noise;
    const v3 m = sample_visible_normal(bd, wo, U, V, N, g.vz, g.ax, g.ay, rx, ry, early, noise);   // requires V.N > eps
    const float Vm = dot(wo, m), aVm = fabsf(Vm);
    const float band = RLS_TOL_BAND_BASE + noise;
    bd.near(Vm, 0.0f, band + RLS_TOL_HALF_VECTOR_NOISE(Vm));
    o.L = m * (2.0f * aVm) - wo;
    const float LH = 2.0f * aVm - Vm;
    o.fresnel = fresnel_c(ratio2, fabsf(LH));
    const float LdotN = dot(o.L, N);
    bd.near(LdotN, 0.0f, band);
    const float Dm = ggx_D(g, m, dot(m, N));
    const float G1v = G1_value(g.a2g, VdotN), G1l = G1_value(g.a2g, LdotN);
    const float i4vn = 0.25f * rcp(absVN);
    const float pdf_raw = Dm * ((Vm * VdotN < 0.0f) ? 0.0f : G1v) * i4vn;
    bd.near(pdf_raw, kEps, 2e-2f * kEps);
    o.pdf = fmaxf(pdf_raw, kEps);
    const float Fh = fresnel_c(ratio2, aVm);
    const float G1o = (LH * LdotN < 0.0f) ? 0.0f : G1l;
    const float sL = LdotN < 0.0f ? -1.0f : (LdotN > 0.0f ? 1.0f : 0.0f);
    const float refl_cos = Fh * (((Vm < 0.0f) ? 0.0f : G1v) * G1o) * Dm * i4vn * sL;
    const bool black = fabsf(ks.x) < kEps && fabsf(ks.y) < kEps && fabsf(ks.z) < kEps;       // AiColorIsSmall, src/rlGgx.h:160
    o.f = black ? mk(0.0f, 0.0f, 0.0f) : ks * refl_cos;
    uint32_t fl = 0;
    if (LdotN <= 0.0f) fl |= kFlagBelowHorizon;
    if (o.f.x == 0.0f && o.f.y == 0.0f && o.f.z == 0.0f) fl |= kFlagBlack;
    if (o.pdf == kEps) fl |= kFlagPdfFloored;
    if (g.entering) fl |= kFlagEntering;
    if (early) fl |= kFlagSlopeEarlyOut;
    o.flags = fl;
    return o;
}

// ================================================================== rlDisney
struct DisneyIn {
    v3 base;
    float subsurface, metallic, specular, specular_tint, roughness, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss;
};
struct DisneyT { v3 Ls, fs, Ld, fd; float ps, pd; uint32_t flags; };

// src/rlDisney.cpp:570-577
RLT_HD float smithG(float NdotV, float alphaG)
{
    const float a = alphaG * alphaG, b = NdotV * NdotV;
    return rcp(NdotV + sqrt_(fma_(-a, b, a + b)));
}
// The fused rlDisney unit (same composition as rls_fused.cuh disney_unit / both oracles): ctor (src/rlDisney.cpp:155-192),
// glossy triple on (rx_s, ry_s), diffuse triple on (rx_d, ry_d).  visible = mSampleFromVisibleNormal (:191).
RLT_HD DisneyT disney_unit(Bands &bd, v3 U, v3 V, v3 N, v3 wo, const DisneyIn &p, bool visible, float rx_s, float ry_s,
                           float rx_d, float ry_d)
{
    DisneyT o;
    const v3 white = mk(1.0f, 1.0f, 1.0f), zero = mk(0.0f, 0.0f, 0.0f);
    // ---- ctor
    const float specular = p.specular * 0.08f;
    const float clearcoat = p.clearcoat * 0.25f;                  // exact (power of two)
    const float r2 = p.roughness * p.roughness;
    const float aspect = sqrt_(fma_(p.anisotropic, -0.9f, 1.0f));
    const float ax = fmaxf(1e-2f, div(r2, aspect)), ay = fmaxf(1e-2f, r2 * aspect);
    const float lum = fma_(p.base.x, 0.212671f, fma_(p.base.y, 0.715160f, p.base.z * 0.072169f));
    const float il = rcp(lum);
    const v3 tint = lum > 0.0f ? p.base * il : white;
    const float ost = 1.0f - p.specular_tint, osh = 1.0f - p.sheen_tint, om = 1.0f - p.metallic;
    const v3 F0 = mk(fma_(p.base.x, p.metallic, om * (specular * fma_(tint.x, p.specular_tint, ost))),
                     fma_(p.base.y, p.metallic, om * (specular * fma_(tint.y, p.specular_tint, ost))),
                     fma_(p.base.z, p.metallic, om * (specular * fma_(tint.z, p.specular_tint, ost))));
    const v3 sheenColor = mk(p.sheen * fma_(tint.x, p.sheen_tint, osh), p.sheen * fma_(tint.y, p.sheen_tint, osh),
                             p.sheen * fma_(tint.z, p.sheen_tint, osh));
    const float VdotN = dot_rn(wo, N);                            // reference order: the eps tests on it stay exact
    // D_GTR1 (:545-551): alpha = lerp(gloss, 0.1, 0.001), (a2 - 1) / (pi ln(a2) (1 + (a2 - 1) x))
    const float g1a = fma_(0.001f, p.clearcoat_gloss, (1.0f - p.clearcoat_gloss) * 0.1f);
    const float g1a2 = g1a * g1a, g1k = (g1a2 - 1.0f) * kInvPi * rcp(log_(g1a2));

    // ---- specular sample (:367-390).  The lobe choice is one IEEE quotient on raw inputs: exact.
#if defined(__CUDA_ARCH__)
    const float gtr2Weight = __fdiv_rn(1.0f, add_rn(clearcoat, 1.0f));
#else
    const float gtr2Weight = 1.0f / add_rn(clearcoat, 1.0f);
#endif
    const uint32_t lobe = rx_s < gtr2Weight ? 0u : 1u;
    bool early = false;
    float noise = 0.0f;
    v3 M;
    if (lobe == 0u) {
        // The reference's own quotient, correctly rounded: rx -> 1 makes sqrt(rx / (1 - rx)) of the uniform-slope path amplify
        // one ulp of rx by 1 / (1 - rx) (3e4 in the sample the ulp-perturbed host build found), so rx must be ITS bits.
#if defined(__CUDA_ARCH__)
        const float rx = __fdiv_rn(rx_s, gtr2Weight);
#else
        float rx = rx_s / gtr2Weight; __asm__ volatile("" : "+x"(rx));
#endif
        if (visible) {
            M = sample_visible_normal(bd, wo, U, V, N, VdotN, ax, ay, rx, ry_s, early, noise);
        } else {                                                  // sampleGTR2AnisoDirection (:406-414)
            const float g = sqrt_(div(ry_s, 1.0f - ry_s));
            float s, c;
            sincos_(kTwoPi * rx, &s, &c);
            M = normalize(to_frame(g * ax * c, g * ay * s, 1.0f, U, V, N));
        }
    } else {                                                      // sampleGTR1Direction (:393-404), a2 = roughness^2
        const float rx = clampf(div(rx_s - gtr2Weight, 1.0f - gtr2Weight), 0.0f, 1.0f);
        float s, c;
        sincos_(kTwoPi * rx, &s, &c);
        // a2 = 0 (roughness exactly 0 is a legal node value): powf(0, y > 0) = 0, cos(theta) = 1; log_ takes normal x only
        bd.require(r2 == 0.0f || r2 >= 1.2e-38f);
        const float pw = (r2 > 0.0f) ? pow_(r2, 1.0f - ry_s) : 0.0f;
        const float ct2 = (r2 == 1.0f) ? 1.0f - ry_s : div(1.0f - pw, 1.0f - r2);
        bd.near(r2, 1.0f, 1e-3f);                                 // (1 - a2^(1-ry)) / (1 - a2) cancels as a2 -> 1
        const float ct = sqrt_(fmaxf(ct2, 0.0f)), st = sqrt_(fmaxf(1.0f - ct2, 0.0f));
        M = normalize(to_frame(st * c, st * s, ct, U, V, N));
        // The reference's own rounding noise in cos(theta): 1 - powf(a2, 1 - ry) cancels as ry -> 1 (a few ulps of 1 over
        // (1 - a2) in cos^2, i.e. e2 / cos in cos, and sqrt(e2) once cos^2 itself is below e2: there the reference's N.M is
        // 0 or not by rounding alone).  Found by the ulp-perturbed host build at ry = 1 - 2^-24.
        // ... and the same absolute error sits in sin^2 = 1 - cos^2 as ry -> 0 (cos -> 1): the smaller of the two decides.
        const float e2 = 2.4e-7f * rcp(fabsf(1.0f - r2));
        noise = e2 * rcp(fmaxf(fminf(ct, st), sqrt_(e2)));
    }
    const float band = RLS_TOL_BAND_BASE + noise;                             // on comparands that follow the sampled direction
    const float NM = dot(N, M);
    bd.near(NM, 0.5f * kEps, 0.5f * kEps + band);                 // N.M < 0 and N.M < eps
    const bool zeroS = NM < 0.0f;
    const float VM = dot(wo, M), aVM = fabsf(VM);
    bd.near(VM, 0.5f * kEps, 0.5f * kEps + band + RLS_TOL_HALF_VECTOR_NOISE(VM));   // L ~ -V: the reference's half vector normalize(L + V) is rounding noise; L.M < eps
    uint32_t fls = lobe << 8;
    if (early) fls |= kFlagSlopeEarlyOut;
    if (zeroS) {
        o.Ls = zero; o.fs = zero; o.ps = 0.0f;
        fls |= kFlagZeroL | kFlagBelowHorizon | kFlagPdfZero | kFlagBlack;
    } else {
        o.Ls = M * (2.0f * aVM) - wo;
        // the half vector of Ls and V is M: L.M = 2|V.M| - V.M, N.M as sampled
        const float LdotN = dot(o.Ls, N);
        const float LdotM = 2.0f * aVM - VM;
        const float NM2 = NM * NM;
        const float hu = dot(M, U) * rcp(ax), hv = dot(M, V) * rcp(ay);
        const float tD = fma_(hu, hu, fma_(hv, hv, NM2));
        const float Ds = kInvPi * rcp(ax * ay * (tD * tD));       // D_GTR2Aniso (:561-568)
        const float Dr = g1k * rcp(fma_(g1a2 - 1.0f, NM2, 1.0f)); // D_GTR1
        // pdf (:520-543)
        const float cw = div(clearcoat, clearcoat + 1.0f);
        if (visible) {
            const float Vn = fmaxf(1e-4f, VdotN);
            const float Dw = smithG(LdotM, r2) * Ds * 2.0f * LdotM * rcp(Vn);
            o.ps = fma_(Dr * NM * rcp(LdotM), cw, (1.0f - cw) * Dw) * 0.25f;
        } else {
            o.ps = fma_(Dr, cw, (1.0f - cw) * Ds) * NM * 0.25f * rcp(LdotM);
        }
        // eval (:318-356) x N.L (:136)
        bd.near(LdotN, 0.5f * kEps, 0.5f * kEps + band);          // L.N <= 0 and L.N < eps
        if (LdotN < kEps || VdotN < kEps || NM < kEps || LdotM < kEps) {
            o.fs = zero;
        } else {
            // L.h within an ulp or two of 1 (m along the view): the reference's (1 - L.h)^5 is exactly 0 or ~1e-36 by rounding
            // alone, and with F0 = 0 and no clearcoat that alone decides whether f is black (tests/hunts/tol_stress_hunt.py --ulp)
            bd.near(LdotM, 1.0f, 5e-7f);
            const float FH = pow5(clampf(1.0f - LdotM, 0.0f, 1.0f));
            const float Gs = smithG(LdotN, r2) * smithG(VdotN, r2);
            const float Gr = smithG(LdotN, 0.25f) * smithG(VdotN, 0.25f);
            const float coat = clearcoat * Dr * fma_(FH, 0.96f, 0.04f) * Gr;          // Fr = lerp(FH, 0.04, 1)
            const float sh = FH * om, DG = Ds * Gs, oF = 1.0f - FH;
            o.fs = mk((fma_(fma_(F0.x, oF, FH), DG, coat) + sheenColor.x * sh) * LdotN,
                      (fma_(fma_(F0.y, oF, FH), DG, coat) + sheenColor.y * sh) * LdotN,
                      (fma_(fma_(F0.z, oF, FH), DG, coat) + sheenColor.z * sh) * LdotN);
        }
        if (LdotN <= 0.0f) fls |= kFlagBelowHorizon;
        if (o.ps == 0.0f) fls |= kFlagPdfZero;
        if (o.fs.x == 0.0f && o.fs.y == 0.0f && o.fs.z == 0.0f) fls |= kFlagBlack;
    }

    // ---- diffuse triple (:359-365, 199-236, 515-518)
    uint32_t fld = 0;
    {
        const float ux = fma_(rx_d, 2.0f, -1.0f), uy = fma_(ry_d, 2.0f, -1.0f);   // exact products
        float kx = 0.0f, ky = 0.0f, rr = 0.0f;
        if (!(ux == 0.0f && uy == 0.0f)) {                        // src/rlUtil.cpp:3-27
            const bool wide = fabsf(ux) > fabsf(uy);
            rr = wide ? ux : uy;
            const float q = div(wide ? (kHalfPi * 0.5f) * uy : 0.5f * ux, rr);
            const float phi = wide ? q : kHalfPi * (1.0f - q);
            float s, c;
            sincos_(phi, &s, &c);
            kx = rr * c; ky = rr * s;
        }
        const float z2 = fma_(-rr, rr, 1.0f);                     // 1 - k.x^2 - k.y^2 with cos^2 + sin^2 = 1
        bd.near(z2, 0.0f, 2e-6f);                                 // the rim: the reference's z is rounding noise there
        const float z = sqrt_(fmaxf(z2, 0.0f));
        o.Ld = to_frame(kx, ky, z, U, V, N);
        const float LdotN = dot(o.Ld, N);
        bd.near(LdotN, 1.6e-4f, 1.9e-4f);                         // L.N <= 0, L.N < eps, L.N / pi <= eps
        const float pdr = LdotN * kInvPi;
        o.pd = fmaxf(1e-4f, pdr);
        // evalDiffuse: H = normalize(L + V); L.H = V.H = sqrt((1 + L.V) / 2) -- only its square is used
        const float LH2 = fma_(0.5f, dot(o.Ld, wo), 0.5f);
        bd.near(LH2, 0.0f, 1e-6f);                                // L.H < eps  <=>  L ~ -V
        if (LdotN < kEps || VdotN < kEps || LH2 < kEps * kEps) {
            o.fd = zero;
        } else {
            const float FL = pow5(clampf(1.0f - LdotN, 0.0f, 1.0f)), FV = pow5(clampf(1.0f - VdotN, 0.0f, 1.0f));
            const float F90 = fma_(2.0f * p.roughness, LH2, 0.5f);
            const float dF = fma_(F90, FL, 1.0f - FL) * fma_(F90, FV, 1.0f - FV);          // lerp(FL, 1, F90) ...
            const float Fss90 = p.roughness * LH2;
            const float Fss = fma_(Fss90, FL, 1.0f - FL) * fma_(Fss90, FV, 1.0f - FV);
            const float ss = 1.25f * fma_(Fss, rcp(LdotN + VdotN) - 0.5f, 0.5f);
            const float k = kInvPi * fma_(ss, p.subsurface, (1.0f - p.subsurface) * dF) * om * LdotN;
            o.fd = p.base * k;
        }
        if (LdotN <= 0.0f) fld |= kFlagBelowHorizon;
        if (o.fd.x == 0.0f && o.fd.y == 0.0f && o.fd.z == 0.0f) fld |= kFlagBlack;
        if (o.pd == 1e-4f) fld |= kFlagPdfFloored;
    }
    o.flags = fls | (fld << 16);
    return o;
}

// ================================================================== NDProfile (rlSss / rlSkin)
struct ProfileT { float r, pdf; v3 Rd; uint32_t flags; };
// setDistance + getRadius + getPdf + evalProfile (src/rlSss.cpp:20-106, src/rlSss.h:30-42) for one sample.
RLT_HD ProfileT skin_profile_unit(Bands &bd, v3 dist, float rx)
{
    ProfileT o;
    const float d[3] = { dist.x, dist.y, dist.z };
    const float R = mul_rn(fmaxf(dist.x, fmaxf(dist.y, dist.z)), 3.0f);
    float id[3], C1[3], C2[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        id[i] = rcp(d[i]);
        const float q = -R * id[i] * 1.44269504088896340736f;    // -R/d in base 2
        C1[i] = 1.0f - ex2(q);
        C2[i] = 1.0f - ex2(q * 0.33333333333f);
    }
    // getRadius (:36-66): channel by thirds (raw input: exact), then the exponential lobe
    const int ch = rx < 0.3333f ? 0 : (rx > 0.6666f ? 2 : 1);
    const float lo = ch == 0 ? 0.0f : (ch == 2 ? 0.6666f : 0.3333f), hi = ch == 0 ? 0.3333f : (ch == 2 ? 1.0f : 0.6666f);
    float x = clampf((rx - lo) * rcp(hi - lo), 0.0f, 1.0f);
    uint32_t fl = (uint32_t)ch << 8;
    const float dc = ch == 0 ? d[0] : (ch == 1 ? d[1] : d[2]);
    const bool degenerate = R < kEps || dc < kEps;               // raw inputs: exact
    const float w1 = ch == 0 ? C1[0] : (ch == 1 ? C1[1] : C1[2]), w2 = ch == 0 ? C2[0] : (ch == 1 ? C2[1] : C2[2]);
    const float w = w1 * rcp(fma_(w2, 3.0f, w1));
    const bool wide = x > w;
    bd.near(x, w, 2e-6f);
    x = clampf(wide ? (x - w) * rcp(1.0f - w) : x * rcp(w), 0.0f, 1.0f);
    // 1 - x C on the reference's own operations (it cancels for small radii), then the log relative to it
    const float arg = sub_rn(1.0f, mul_rn(x, wide ? w2 : w1));
    float r = log_(arg) * (wide ? -3.0f * dc : -dc);
    if (degenerate) { fl |= 0x0800u; r = 0.0f; } else if (wide) fl |= 0x0400u;
    bd.require(d[0] >= kEps && d[1] >= kEps && d[2] >= kEps);    // getPdf floors d at eps, evalProfile returns 1: exact path
    bd.near(r, kEps, 1e-8f);                                     // evalProfile is white below eps
    // getPdf (:68-84) and evalProfile (:86-106) at r
    const float ir = rcp(r), nr = -r * 1.44269504088896340736f;
    float pdf = 0.0f, rd[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float t = nr * id[i];
        const float s12 = ex2(t) + ex2(t * 0.33333333333f);
        pdf = fma_(s12 * id[i], rcp(fma_(C2[i], 3.0f, C1[i])), pdf);
        rd[i] = s12 * ir * id[i] * (0.125f * kInvPi);
    }
    o.r = r;
    o.pdf = pdf * ir * (kInvPi * (1.0f / 6.0f));
    o.Rd = mk(rd[0], rd[1], rd[2]);
    if (R < kEps) { o.pdf = 1.0f; o.Rd = mk(0.0f, 0.0f, 0.0f); }
    else if (r < kEps) o.Rd = mk(1.0f, 1.0f, 1.0f);
    o.flags = fl;
    return o;
}

} // namespace tol
} // namespace rls
