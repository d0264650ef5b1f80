/*
 * oracle/rls_oracle.c -- TEST INFRASTRUCTURE ONLY (kind "port").
 *
 * A plain-C, scalar restatement of the reference's BSDF hot path.  Every function cites
 * the reference lines it follows (paths relative to /root/reference/).  Operation order
 * is kept identical to the reference so that, compiled with the same pinned flags
 * (-O2 -ffp-contract=off -fno-fast-math), it is BIT-IDENTICAL to oracle/_ref/librls_ref.so
 * (the reference's own sources behind oracle/shim/ai.h); tests/test_oracle_pinning.py
 * asserts exactly that, and tests/golden/ holds vectors generated from the reference
 * library by tests/golden/make_golden.py.
 *
 * Third-party arithmetic below the reference: glibc 2.39 libm binary32 functions
 * (sqrtf sincosf atan2f acosf tanf powf logf expf) and the Arnold 4.2.11 inline vector
 * maths, whose semantics are fixed by oracle/shim/ai.h (normative, see its header).
 *
 * Parity status: pinned against the reference compiled here; the reference's own test
 * suite holds no function-level vectors for this path.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_common.h"

#define EPS        1.0e-4f                     /* AI_EPSILON */
#define PI_F       3.14159265358979323846f     /* AI_PI */
#define TWO_PI_F   6.28318530717958647692f     /* AI_PITIMES2 */
#define HALF_PI_F  1.57079632679489661923f     /* AI_PIOVER2 */
#define INV_PI_F   0.31830988618379067154f     /* AI_ONEOVERPI */

#define SQRF(a)    ((a) * (a))
#define ABSF(a)    (((a) < 0) ? -(a) : (a))
#define MAXF(a, b) (((a) > (b)) ? (a) : (b))
#define CLAMPF(v, lo, hi) (((v) < (lo)) ? (lo) : (((v) > (hi)) ? (hi) : (v)))

typedef struct { float x, y, z; } v3;

static inline v3 mk3(float x, float y, float z) { v3 v; v.x = x; v.y = y; v.z = z; return v; }
static inline v3 add3(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 neg3(v3 a) { return mk3(-a.x, -a.y, -a.z); }
static inline v3 scale3(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline v3 mul3(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          /* AiV3Dot */
static inline int   iszero3(v3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
static inline v3 normalize3(v3 a)                                                              /* AiV3Normalize */
{
    float len = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len != 0.0f) {
        float inv = 1.0f / len;
        return mk3(a.x * inv, a.y * inv, a.z * inv);
    }
    return a;
}
static inline v3 rotate_to_frame(v3 a, v3 u, v3 v, v3 w)                                       /* AiV3RotateToFrame */
{
    return mk3(a.x * u.x + a.y * v.x + a.z * w.x,
               a.x * u.y + a.y * v.y + a.z * w.y,
               a.x * u.z + a.y * v.z + a.z * w.z);
}
static inline int sgn(float a) { return (a < 0) ? -1 : ((a > 0) ? 1 : 0); }                    /* SGN */
static inline float lerpf(float t, float a, float b) { return (1.0f - t) * a + b * t; }        /* LERP */
static inline v3 lerp3(float t, v3 a, v3 b) { return add3(scale3(a, 1.0f - t), scale3(b, t)); }
static inline float linearstep(float lo, float hi, float t)                                    /* LINEARSTEP */
{
    float v = (t - lo) / (hi - lo);
    return CLAMPF(v, 0.0f, 1.0f);
}

/* ------------------------------------------------------------------ rlUtil */
/* src/rlUtil.h:21-29 */
static inline v3 spherical_direction(float cosTheta, float phi)
{
    v3 o;
    o.z = cosTheta;
    float r = sqrtf(1.0f - SQRF(o.z));
    o.x = r * cosf(phi);
    o.y = r * sinf(phi);
    return o;
}
/* src/rlUtil.h:31-34 (note the ABS) */
static inline v3 reflect_direction(v3 i, v3 n)
{
    float d = dot3(i, n);
    float s = 2.0f * ABSF(d);
    return sub3(scale3(n, s), i);
}
/* src/rlUtil.h:36-39 */
static inline float color_to_luminance(v3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }
/* src/rlUtil.cpp:3-27; z is left unset by the reference and overwritten by callers */
static inline v3 concentric_disk_sample(float rx, float ry)
{
    rx = rx * 2.0f - 1.0f;
    ry = ry * 2.0f - 1.0f;
    v3 result = mk3(0.0f, 0.0f, 0.0f);
    if (rx == 0.0f && ry == 0.0f) return result;
    float r, phi;
    if (ABSF(rx) > ABSF(ry)) {
        r = rx;
        phi = HALF_PI_F * 0.5f * ry / rx;
    } else {
        r = ry;
        phi = HALF_PI_F * (1.0f - 0.5f * rx / ry);
    }
    result.x = r * cosf(phi);
    result.y = r * sinf(phi);
    return result;
}

/* ------------------------------------------------- visible-normal sampling */
typedef struct { float x, y; } v2;

/* src/rlGgx.cpp:18-25 / src/rlDisney.cpp:420-427 */
static inline v2 uniform_slope(float rx, float ry)
{
    v2 s;
    float r = sqrtf(rx / (1.0f - rx));
    float phi = TWO_PI_F * ry;
    s.x = r * cosf(phi);
    s.y = r * sinf(phi);
    return s;
}
/* src/rlGgx.cpp:14-61 (VNDFKernel::sampleSlope) == src/rlDisney.cpp:416-463 */
static _Thread_local int g_slope_early_out;    /* RLS_FLAG_SLOPE_EARLY_OUT of the last sample_slope call on this thread */
static v2 sample_slope(float theta, float rx, float ry)
{
    v2 slope;
    g_slope_early_out = 1;
    if (theta < EPS) return uniform_slope(rx, ry);

    float B = tanf(theta);
    float B2 = SQRF(B);
    float G1 = 2.0f / (1.0f + sqrtf(1.0f + B2));

    float A = 2.0f * rx / G1 - 1.0f;
    float A2 = SQRF(A);
    if (ABSF(A2 - 1.0f) < EPS) return uniform_slope(rx, ry);
    g_slope_early_out = 0;

    float tmp = 1.0f / (A2 - 1.0f);
    float D = sqrtf(MAXF(0.0f, B2 * SQRF(tmp) - (A2 - B2) * tmp));
    float slopeX1 = B * tmp - D;
    float slopeX2 = B * tmp + D;
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
    slope.y = sign * z * sqrtf(1.0f + SQRF(slope.x));
    return slope;
}
/* src/rlGgx.cpp:63-99 (VNDFKernel::evalSample) == src/rlDisney.cpp:467-502 */
static v3 sample_visible_normal(v3 view, v3 U, v3 Vax, v3 N, float ax, float ay, float rx, float ry)
{
    v3 V = view;
    float d = dot3(N, V);
    float cosThetaV = CLAMPF(d, -1.0f, 1.0f);
    float phiV = atan2f(dot3(Vax, V), dot3(U, V));
    V = spherical_direction(cosThetaV, phiV);

    V.x *= ax;
    V.y *= ay;
    V = normalize3(V);

    float theta = 0.0f, phi = 0.0f;
    if (V.z < (1.0f - EPS)) {
        theta = acosf(V.z);
        phi = atan2f(V.y, V.x);
    }
    v2 slope = sample_slope(theta, rx, ry);

    float cosPhi = cosf(phi);
    float sinPhi = sinf(phi);
    v3 omega;
    omega.x = -(cosPhi * slope.x - sinPhi * slope.y) * ax;
    omega.y = -(sinPhi * slope.x + cosPhi * slope.y) * ay;
    omega.z = 1.0f;
    omega = rotate_to_frame(omega, U, Vax, N);
    return normalize3(omega);
}

/* ------------------------------------------------------------------- rlGgx */
typedef struct {
    v3 U, V, N, wo, ks;
    float iorIn, iorOut, rough, ax, ay;
    int entering;
    int kernel;      /* RLS_GGX_SAMPLER_VNDF (shipped, src/rlGgx.h:375) or RLS_GGX_SAMPLER_NDF */
} ggx_t;

static inline void load_shading(const rls_shading_soa *s, size_t i, v3 *U, v3 *V, v3 *N, v3 *wo, int *back)
{
    *U = mk3(s->U.x[i], s->U.y[i], s->U.z[i]);
    *V = mk3(s->V.x[i], s->V.y[i], s->V.z[i]);
    *N = mk3(s->N.x[i], s->N.y[i], s->N.z[i]);
    *wo = mk3(s->wo.x[i], s->wo.y[i], s->wo.z[i]);
    *back = s->backfacing && s->backfacing[i];
}

/* src/rlGgx.h:130-156 (GgxSamplerT ctor) */
static void ggx_init(ggx_t *g, v3 U, v3 V, v3 Nf, v3 wo, int backfacing, v3 ks, float ior, float roughness, float aniso)
{
    v3 Ngeo = backfacing ? neg3(Nf) : Nf;      /* sg->N */
    v3 Rd = neg3(wo);                          /* sg->Rd */
    g->entering = dot3(Ngeo, Rd) < EPS;        /* :137 */
    g->iorIn = 1.0f;
    g->iorOut = MAXF(ior, 1e-4f);              /* :139 */
    if (!g->entering) { float t = g->iorIn; g->iorIn = g->iorOut; g->iorOut = t; }
    g->wo = neg3(Rd);                          /* :144 */
    g->U = U; g->V = V; g->N = Nf;             /* :145-146, explicit frame */
    float aspect = sqrtf(1.0f - aniso * 0.9f); /* :148 */
    g->ax = MAXF(1e-4f, SQRF(roughness) / aspect);
    g->ay = MAXF(1e-4f, SQRF(roughness) * aspect);
    g->rough = MAXF(1e-5f, SQRF(roughness));   /* :155 */
    g->ks = ks;
    g->kernel = RLS_GGX_SAMPLER_VNDF;
}
/* src/rlGgx.h:33-41 (NDFKernel::evalSample, [2] Eq.14) */
static v3 sample_ndf_normal(v3 U, v3 Vax, v3 N, float ax, float ay, float rx, float ry)
{
    float g = sqrtf(rx / (1.0f - rx));
    float phi = TWO_PI_F * ry;
    v3 omega = mk3(g * ax * cosf(phi), g * ay * sinf(phi), 1.0f);
    omega = rotate_to_frame(omega, U, Vax, N);
    return normalize3(omega);
}
static v3 ggx_sample_normal(const ggx_t *g, float rx, float ry)
{
    g_slope_early_out = 0;                       /* NDFKernel has no early-out */
    if (g->kernel == RLS_GGX_SAMPLER_NDF) return sample_ndf_normal(g->U, g->V, g->N, g->ax, g->ay, rx, ry);
    return sample_visible_normal(g->wo, g->U, g->V, g->N, g->ax, g->ay, rx, ry);
}
/* src/rlGgx.h:249-270 */
static float ggx_fresnel(const ggx_t *g, v3 i, v3 m)
{
    float d = dot3(i, m);
    float c = ABSF(d);
    float gSqr = SQRF(g->iorOut / g->iorIn) - 1.0f + c * c;
    if (gSqr < 0.0f) return 1.0f;
    float gg = sqrtf(gSqr);
    float gmc = gg - c;
    float gpc = gg + c;
    return 0.5f * SQRF(gmc / gpc) * (1.0f + SQRF((c * gpc - 1.0f) / (c * gmc + 1.0f)));
}
/* src/rlGgx.h:343-357 (isotropic mRoughness even when anisotropic) */
static float ggx_G1(const ggx_t *g, v3 v, v3 m, v3 n)
{
    float VdotM = dot3(v, m);
    float VdotN = dot3(v, n);
    if (VdotM * VdotN < 0.0f) return 0.0f;
    float cosSqr = SQRF(VdotN);
    float tanSqr = 1.0f / cosSqr - 1.0f;
    float denominator = 1.0f + sqrtf(1.0f + SQRF(g->rough) * tanSqr);
    return 2.0f / denominator;
}
/* src/rlGgx.h:272-275 */
static float ggx_G(const ggx_t *g, v3 i, v3 o, v3 m, v3 n) { return ggx_G1(g, i, m, n) * ggx_G1(g, o, m, n); }
/* src/rlGgx.h:332-340 */
static float ggx_D(const ggx_t *g, v3 m)
{
    float MdotU = dot3(m, g->U);
    float MdotV = dot3(m, g->V);
    float MdotN2 = SQRF(dot3(g->N, m));
    float denominator = g->ax * g->ay * SQRF(SQRF(MdotU / g->ax) + SQRF(MdotV / g->ay) + MdotN2);
    return INV_PI_F / denominator;
}
/* src/rlGgx.h:304-313 */
static float ggx_reflection(const ggx_t *g, v3 i, v3 o, v3 n)
{
    v3 hr = scale3(normalize3(add3(o, i)), (float)sgn(dot3(i, n)));
    float reflectWeight = ggx_fresnel(g, i, hr);
    float dl = dot3(o, n), dv = dot3(i, n);
    float LdotN = ABSF(dl);
    float VdotN = ABSF(dv);
    return reflectWeight * ggx_G(g, i, o, hr, n) * ggx_D(g, hr) * 0.25f / (LdotN * VdotN);
}
/* src/rlGgx.h:316-328 */
static float ggx_refraction(const ggx_t *g, v3 i, v3 o, v3 n)
{
    v3 ht = neg3(normalize3(add3(scale3(i, g->iorIn), scale3(o, g->iorOut))));
    float refractWeight = 1.0f - ggx_fresnel(g, i, ht);
    float don = dot3(o, n), din = dot3(i, n);
    float OdotN = ABSF(don);
    float IdotN = ABSF(din);
    float OdotH = dot3(o, ht);
    float IdotH = dot3(i, ht);
    float denominator = OdotN * IdotN * SQRF(g->iorIn * IdotH + g->iorOut * OdotH);
    float num = OdotH * IdotH;
    return ABSF(num) * SQRF(g->iorOut) * refractWeight * ggx_G(g, i, o, ht, n) * ggx_D(g, ht) / denominator;
}
/* src/rlGgx.h:277-291 (eta not squared, as in the reference) */
static int ggx_refract_direction(const ggx_t *g, v3 m, v3 i, v3 *dir)
{
    int sign = sgn(dot3(i, g->N));
    float IdotM = dot3(i, m);
    float eta = g->iorIn / g->iorOut;
    float cosThetaTSqr = 1.0f + eta * (SQRF(IdotM) - 1.0f);
    if (cosThetaTSqr < 0.0f) return 0;
    float s = eta * IdotM - sign * sqrtf(cosThetaTSqr);
    *dir = sub3(scale3(m, s), scale3(i, eta));
    return 1;
}
/* src/rlGgx.h:294-301 */
static float ggx_sample_weight(const ggx_t *g, v3 i, v3 o, v3 m)
{
    float IdotH = dot3(i, m);
    float dm = dot3(m, g->N), di = dot3(i, g->N);
    float MdotN = ABSF(dm);
    float IdotN = ABSF(di);
    float q = IdotH / (IdotN * MdotN);
    return ggx_G(g, i, o, m, g->N) * ABSF(q);
}
/* src/rlGgx.h:110-119,158-165 */
static v3 ggx_eval_brdf(const ggx_t *g, v3 L)
{
    if (iszero3(L)) return mk3(0, 0, 0);
    if (ABSF(g->ks.x) < EPS && ABSF(g->ks.y) < EPS && ABSF(g->ks.z) < EPS) return mk3(0, 0, 0);
    float refl = ggx_reflection(g, g->wo, L, g->N);
    float d = dot3(L, g->N);
    return scale3(scale3(g->ks, refl), d);
}
/* src/rlGgx.h:121-127 + VNDFKernel::evalPdf :72-80 */
static float ggx_eval_pdf(const ggx_t *g, v3 L)
{
    v3 V = g->wo;
    v3 H = normalize3(add3(V, L));
    float din = dot3(V, g->N);
    float IdotN = ABSF(din);
    if (g->kernel == RLS_GGX_SAMPLER_NDF) {
        /* NDFKernel::evalPdf src/rlGgx.h:45-50: [1] Eq.38, no floor */
        float dim = dot3(V, H), dmn = dot3(H, g->N);
        float IdotM = ABSF(dim);
        float MdotN = ABSF(dmn);
        return ggx_D(g, H) * MdotN * 0.25f / IdotM;
    }
    float pdf = ggx_D(g, H) * ggx_G1(g, V, H, g->N) / IdotN * 0.25f;
    return MAXF(pdf, EPS);
}
/* src/rlGgx.h:97-107 */
static v3 ggx_eval_sample(const ggx_t *g, float rx, float ry, float *fresnel)
{
    v3 M = ggx_sample_normal(g, rx, ry);
    v3 L = reflect_direction(g->wo, M);
    if (fresnel) *fresnel = (0.0f + ggx_fresnel(g, L, M)) / 1.0f;   /* :103-104,181-184 */
    return L;
}

static inline uint32_t bsdf_flags(v3 L, v3 N, v3 f, float pdf)
{
    uint32_t fl = 0;
    if (iszero3(L)) fl |= RLS_FLAG_ZERO_L;
    if (dot3(L, N) <= 0.0f) fl |= RLS_FLAG_BELOW_HORIZON;
    if (pdf == 0.0f) fl |= RLS_FLAG_PDF_ZERO;
    if (iszero3(f)) fl |= RLS_FLAG_F_BLACK;
    if (pdf == EPS) fl |= RLS_FLAG_PDF_FLOORED;
    return fl;
}

static inline void ggx_from_params(ggx_t *g, const rls_shading_soa *sg, const rls_ggx_params *p, size_t i)
{
    v3 U, V, N, wo; int back;
    load_shading(sg, i, &U, &V, &N, &wo, &back);
    float c[3];
    orc_p3(&p->KsColor, i, c);
    ggx_init(g, U, V, N, wo, back, mk3(c[0], c[1], c[2]), orc_p1(&p->ior, i),
             orc_p1(&p->specularRoughness, i), orc_p1(&p->anisotropic, i));
    g->kernel = p->normal_sampler;
}

typedef struct { float F, f_r, pdf_r, f_t, w_t; v3 wi_r, wi_t; uint32_t flags; } dielectric_t;

/* The rough-dielectric unit: src/rlGgx.h:228-243 loop body with the in-tree refraction
 * restatement (getRefractDirection) standing in for Arnold's AiRefractRay. */
static dielectric_t dielectric_unit(v3 U, v3 V, v3 N, v3 wo, int back, float ior, float rough, float aniso, float rx, float ry, int kernel)
{
    dielectric_t r;
    ggx_t g;
    ggx_init(&g, U, V, N, wo, back, mk3(1.0f, 1.0f, 1.0f), ior, rough, aniso);
    g.kernel = kernel;
    v3 m = ggx_sample_normal(&g, rx, ry);
    r.wi_r = reflect_direction(g.wo, m);
    r.F = ggx_fresnel(&g, r.wi_r, m);
    v3 fr = ggx_eval_brdf(&g, r.wi_r);
    r.f_r = fr.x;
    r.pdf_r = ggx_eval_pdf(&g, r.wi_r);
    r.flags = bsdf_flags(r.wi_r, g.N, fr, r.pdf_r);
    if (g.entering) r.flags |= RLS_FLAG_ENTERING;
    if (g_slope_early_out) r.flags |= RLS_FLAG_SLOPE_EARLY_OUT;
    v3 t;
    if (ggx_refract_direction(&g, m, g.wo, &t)) {
        r.wi_t = t;
        r.f_t = ggx_refraction(&g, g.wo, t, g.N);
    } else {
        r.wi_t = reflect_direction(g.wo, m);
        r.f_t = 0.0f;
        r.flags |= RLS_FLAG_TIR;
    }
    r.w_t = ggx_sample_weight(&g, g.wo, r.wi_t, m);
    return r;
}

/* ---------------------------------------------------------------- rlDisney */
typedef struct {
    v3 U, V, N, wo;
    v3 base, F0, sheenColor;
    float roughness, subsurface, metallic, clearcoat, clearcoatGloss;
    float specRough, ax, ay;
    int visibleNormal;   /* mSampleFromVisibleNormal, src/rlDisney.cpp:191 */
} disney_t;

/* src/rlDisney.cpp:155-192 */
static void disney_init(disney_t *d, const rls_shading_soa *sg, const rls_disney_params *p, size_t i)
{
    int back;
    load_shading(sg, i, &d->U, &d->V, &d->N, &d->wo, &back);
    d->wo = neg3(neg3(d->wo));                                   /* mViewDir = -sg->Rd */
    float c[3];
    orc_p3(&p->base_color, i, c);
    d->base = mk3(c[0], c[1], c[2]);
    d->roughness = orc_p1(&p->roughness, i);
    d->subsurface = orc_p1(&p->subsurface, i);
    float specular = orc_p1(&p->specular, i) * 0.08f;            /* :163 */
    float specularTint = orc_p1(&p->specular_tint, i);
    d->metallic = orc_p1(&p->metallic, i);
    float sheen = orc_p1(&p->sheen, i);
    float sheenTint = orc_p1(&p->sheen_tint, i);
    float anisotropic = orc_p1(&p->anisotropic, i);
    d->clearcoat = orc_p1(&p->clearcoat, i) * 0.25f;             /* :169 */
    d->clearcoatGloss = orc_p1(&p->clearcoat_gloss, i);

    float aspect = sqrtf(1.0f - anisotropic * 0.9f);             /* :177 */
    d->ax = MAXF(1e-2f, SQRF(d->roughness) / aspect);
    d->ay = MAXF(1e-2f, SQRF(d->roughness) * aspect);
    d->specRough = SQRF(d->roughness);                           /* :181 */

    float luminance = color_to_luminance(d->base);
    v3 white = mk3(1.0f, 1.0f, 1.0f);
    v3 tint = luminance > 0.0f ? mk3(d->base.x / luminance, d->base.y / luminance, d->base.z / luminance) : white;
    v3 metallicColor = scale3(lerp3(specularTint, white, tint), specular);   /* :187 */
    d->F0 = lerp3(d->metallic, metallicColor, d->base);                      /* :188 */
    d->sheenColor = scale3(lerp3(sheenTint, white, tint), sheen);            /* :190 */
    d->visibleNormal = p->sample_from_visible_normal != 0;
}
/* src/rlDisney.cpp:406-414 */
static v3 disney_sample_gtr2_aniso(const disney_t *d, float rx, float ry)
{
    float g = sqrtf(ry / (1.0f - ry));
    float phi = TWO_PI_F * rx;
    v3 omega = mk3(g * d->ax * cosf(phi), g * d->ay * sinf(phi), 1.0f);
    omega = rotate_to_frame(omega, d->U, d->V, d->N);
    return normalize3(omega);
}
/* src/rlDisney.cpp:570-577 */
static inline float smithG_GGX(float NdotV, float alphaG)
{
    float a = alphaG * alphaG;
    float b = NdotV * NdotV;
    return 1.0f / (NdotV + sqrtf(a + b - a * b));
}
/* src/rlDisney.cpp:545-551 */
static inline float D_GTR1(const disney_t *d, float MdotN2)
{
    float alpha = lerpf(d->clearcoatGloss, 0.1f, 0.001f);
    float a2 = SQRF(alpha);
    float denominator = logf(a2) * (1.0f + (a2 - 1.0f) * MdotN2);
    return (a2 - 1.0f) * INV_PI_F / denominator;
}
/* src/rlDisney.cpp:561-568 */
static inline float D_GTR2Aniso(const disney_t *d, v3 m, float MdotN2)
{
    float HdotU = dot3(m, d->U);
    float HdotV = dot3(m, d->V);
    float denominator = d->ax * d->ay * SQRF(SQRF(HdotU / d->ax) + SQRF(HdotV / d->ay) + MdotN2);
    return INV_PI_F / denominator;
}
/* src/rlDisney.cpp:199-236 */
static v3 disney_eval_diffuse(const disney_t *d, v3 L)
{
    float LdotN = dot3(L, d->N);
    float VdotN = dot3(d->wo, d->N);
    if (LdotN < EPS || VdotN < EPS) return mk3(0, 0, 0);
    v3 H = normalize3(add3(L, d->wo));
    float LdotH = dot3(L, H);
    float NdotH = dot3(d->wo, H);   /* sic: V.H, :210 */
    if (NdotH < EPS || LdotH < EPS) return mk3(0, 0, 0);
    float LdotH2 = SQRF(LdotH);
    float tl = 1.0f - LdotN, tv = 1.0f - VdotN;
    float FL = powf(CLAMPF(tl, 0.0f, 1.0f), 5.0f);
    float FV = powf(CLAMPF(tv, 0.0f, 1.0f), 5.0f);
    float F90 = 0.5f + 2.0f * d->roughness * LdotH2;
    float diffuseFactor = lerpf(FL, 1.0f, F90) * lerpf(FV, 1.0f, F90);
    float Fss90 = d->roughness * LdotH2;
    float Fss = lerpf(FL, 1.0f, Fss90) * lerpf(FV, 1.0f, Fss90);
    float ssFactor = 1.25f * (Fss * (1.0f / (LdotN + VdotN) - 0.5f) + 0.5f);
    v3 diffuse = scale3(scale3(d->base, INV_PI_F), lerpf(d->subsurface, diffuseFactor, ssFactor));
    return scale3(diffuse, 1.0f - d->metallic);
}
/* src/rlDisney.cpp:318-356 */
static v3 disney_eval_specular(const disney_t *d, v3 L)
{
    float LdotN = dot3(L, d->N);
    float VdotN = dot3(d->wo, d->N);
    if (LdotN < EPS || VdotN < EPS) return mk3(0, 0, 0);
    v3 M = normalize3(add3(L, d->wo));
    float LdotM = dot3(L, M);
    float NdotM = dot3(d->N, M);
    if (NdotM < EPS || LdotM < EPS) return mk3(0, 0, 0);
    float NdotM2 = SQRF(NdotM);
    float Ds = D_GTR2Aniso(d, M, NdotM2);
    float th = 1.0f - LdotM;
    float FH = powf(CLAMPF(th, 0.0f, 1.0f), 5.0f);
    v3 Fs = lerp3(FH, d->F0, mk3(1.0f, 1.0f, 1.0f));
    float Gs = smithG_GGX(LdotN, d->specRough) * smithG_GGX(VdotN, d->specRough);
    float Dr = D_GTR1(d, NdotM2);
    float Fr = lerpf(FH, 0.04f, 1.0f);
    float Gr = smithG_GGX(LdotN, 0.25f) * smithG_GGX(VdotN, 0.25f);
    v3 Fsheen = scale3(scale3(d->sheenColor, FH), 1.0f - d->metallic);
    v3 spec = scale3(scale3(Fs, Ds), Gs);
    float coat = d->clearcoat * Dr * Fr * Gr;
    return add3(mk3(spec.x + coat, spec.y + coat, spec.z + coat), Fsheen);
}
/* src/rlDisney.cpp:120-137 */
static v3 disney_eval_brdf(const disney_t *d, int type, v3 L)
{
    if (iszero3(L)) return mk3(0, 0, 0);
    float NdotL = dot3(d->N, L);
    if (type == RLS_RAY_DIFFUSE) return scale3(disney_eval_diffuse(d, L), NdotL);
    return scale3(disney_eval_specular(d, L), NdotL);
}
/* src/rlDisney.cpp:359-365 */
static v3 disney_sample_diffuse(const disney_t *d, float rx, float ry)
{
    v3 omega = concentric_disk_sample(rx, ry);
    float t = 1.0f - SQRF(omega.x) - SQRF(omega.y);
    omega.z = sqrtf(MAXF(0.0f, t));
    return rotate_to_frame(omega, d->U, d->V, d->N);
}
/* src/rlDisney.cpp:393-404 (a2 = roughness^2, not the clearcoat-gloss alpha) */
static v3 disney_sample_gtr1(const disney_t *d, float rx, float ry)
{
    float phiH = TWO_PI_F * rx;
    float a2 = SQRF(d->roughness);
    float cosThetaH = a2 == 1.0f ? sqrtf(1.0f - ry)
                                 : sqrtf((1.0f - powf(a2, 1.0f - ry)) / (1.0f - a2));
    v3 omega = spherical_direction(cosThetaH, phiH);
    omega = rotate_to_frame(omega, d->U, d->V, d->N);
    return normalize3(omega);
}
/* src/rlDisney.cpp:367-390; *lobe: 0 = GTR2, 1 = GTR1 */
static v3 disney_sample_specular(const disney_t *d, float rx, float ry, uint32_t *lobe)
{
    g_slope_early_out = 0;                       /* set by the GTR2 visible-normal lobe only */
    v3 M;
    float gtr2Weight = 1.0f / (d->clearcoat + 1.0f);
    if (rx < gtr2Weight) {
        rx /= gtr2Weight;
        M = d->visibleNormal ? sample_visible_normal(d->wo, d->U, d->V, d->N, d->ax, d->ay, rx, ry)
                             : disney_sample_gtr2_aniso(d, rx, ry);
        *lobe = 0;
    } else {
        rx = (rx - gtr2Weight) / (1.0f - gtr2Weight);
        M = disney_sample_gtr1(d, rx, ry);
        *lobe = 1;
    }
    if (dot3(d->N, M) < 0.0f) return mk3(0, 0, 0);
    return reflect_direction(d->wo, M);
}
/* src/rlDisney.cpp:515-518 */
static float disney_diffuse_pdf(const disney_t *d, v3 i) { float p = dot3(i, d->N) * INV_PI_F; return MAXF(1e-4f, p); }
/* src/rlDisney.cpp:520-543 (mSampleFromVisibleNormal == true) */
static float disney_specular_pdf(const disney_t *d, v3 i)
{
    v3 m = normalize3(add3(i, d->wo));
    float dim = dot3(i, m);
    float IdotM = ABSF(dim);
    float MdotN = dot3(m, d->N);
    if (MdotN < 0.0f) return 0.0f;
    float MdotN2 = SQRF(MdotN);
    float clearcoatWeight = d->clearcoat / (d->clearcoat + 1.0f);
    if (!d->visibleNormal) {                      /* :541-542 */
        float D0 = lerpf(clearcoatWeight, D_GTR2Aniso(d, m, MdotN2), D_GTR1(d, MdotN2));
        return D0 * ABSF(MdotN) * 0.25f / IdotM;
    }
    float dvn = dot3(d->wo, d->N);
    float VdotN = MAXF(1e-4f, dvn);
    float Dw = smithG_GGX(IdotM, d->specRough) * D_GTR2Aniso(d, m, MdotN2) * 2.0f * IdotM / VdotN;
    float D = lerpf(clearcoatWeight, Dw, D_GTR1(d, MdotN2) * ABSF(MdotN) / IdotM);
    return D * 0.25f;
}
/* src/rlDisney.cpp:139-152 */
static float disney_eval_pdf(const disney_t *d, int type, v3 L)
{
    if (iszero3(L)) return 0.0f;
    if (type == RLS_RAY_DIFFUSE) return disney_diffuse_pdf(d, L);
    return disney_specular_pdf(d, L);
}

/* ------------------------------------------------------------------- rlSss */
typedef struct { float d[3], C1[3], C2[3], R; } ndprofile_t;

/* src/rlSss.cpp:20-34 (the unused `s` of :23 is dead code) */
static void nd_set_distance(ndprofile_t *p, v3 dist)
{
    p->d[0] = dist.x; p->d[1] = dist.y; p->d[2] = dist.z;
    p->R = MAXF(dist.x, MAXF(dist.y, dist.z)) * 3.0f;
    for (int i = 0; i < 3; i++) {
        float d = p->d[i];
        p->C1[i] = 1.0f - expf(-p->R / d);
        p->C2[i] = 1.0f - expf(-p->R / d / 3.0f);
    }
}
/* src/rlSss.h:30-42 */
static int nd_select_dist_lobe(float *x)
{
    if (*x < 0.3333f) { *x = linearstep(0.0f, 0.3333f, *x); return 0; }
    else if (*x > 0.6666f) { *x = linearstep(0.6666f, 1.0f, *x); return 2; }
    *x = linearstep(0.3333f, 0.6666f, *x);
    return 1;
}
/* src/rlSss.cpp:36-66 */
static float nd_get_radius(const ndprofile_t *p, float rx)
{
    if (p->R < EPS) return 0.0f;
    int distIdx = nd_select_dist_lobe(&rx);
    float d = p->d[distIdx];
    if (d < EPS) return 0.0f;
    float w1 = p->C1[distIdx];
    float w2 = p->C2[distIdx];
    float w = w1 / (w1 + w2 * 3.0f);
    float r;
    if (rx > w) {
        rx = linearstep(w, 1.0f, rx);
        r = logf(1.0f - rx * w2) * (-d * 3.0f);
    } else {
        rx = linearstep(0.0f, w, rx);
        r = logf(1.0f - rx * w1) * (-d);
    }
    return r;
}
/* src/rlSss.cpp:68-84 */
static float nd_get_pdf(const ndprofile_t *p, float r)
{
    if (p->R < EPS) return 1.0f;
    float pdf = 0.0f;
    for (unsigned i = 0; i < 3; i++) {
        float d = MAXF(p->d[i], EPS);
        float p1 = expf(-r / d);
        float p2 = expf(-r / d / 3.0f);
        pdf += (p1 + p2) / d / (p->C1[i] + p->C2[i] * 3.0f);
    }
    return pdf / (TWO_PI_F * r * 3.0f);
}
/* src/rlSss.cpp:86-106 */
static v3 nd_eval_profile(const ndprofile_t *p, float r)
{
    if (p->R < EPS) return mk3(0, 0, 0);
    else if (r < EPS) return mk3(1.0f, 1.0f, 1.0f);
    float denom = 8.0f * PI_F * r;
    float out[3];
    for (unsigned i = 0; i < 3; i++) {
        float d = p->d[i];
        out[i] = d < EPS ? 1.0f : (expf(-r / d) + expf(-r / (3.0f * d))) / (denom * d);
    }
    return mk3(out[0], out[1], out[2]);
}
static uint32_t nd_flags(const ndprofile_t *p, float rx)
{
    float x = rx;
    int ch = nd_select_dist_lobe(&x);
    uint32_t fl = (uint32_t)ch << RLS_FLAG_LOBE_SHIFT;
    float d = p->d[ch];
    if (p->R < EPS || d < EPS) {
        fl |= RLS_FLAG_DEGENERATE;
    } else {
        float w1 = p->C1[ch], w2 = p->C2[ch];
        float w = w1 / (w1 + w2 * 3.0f);
        if (x > w) fl |= RLS_FLAG_EXP_LOBE;
    }
    return fl;
}
static inline void nd_load(const rls_ndprofile_soa *s, size_t i, ndprofile_t *p)
{
    p->d[0] = s->distance.x[i]; p->d[1] = s->distance.y[i]; p->d[2] = s->distance.z[i];
    p->C1[0] = s->C1.x[i]; p->C1[1] = s->C1.y[i]; p->C1[2] = s->C1.z[i];
    p->C2[0] = s->C2.x[i]; p->C2[1] = s->C2.y[i]; p->C2[2] = s->C2.z[i];
    p->R = s->max_radius[i];
}
static inline v3 skin_scatter_dist(const rls_skin_params *sp, size_t i)
{
    float d[3];
    orc_p3(&sp->sss_scatter_dist, i, d);
    float scale = orc_p1(&sp->sss_dist_multiplier, i);
    return scale3(mk3(d[0], d[1], d[2]), scale);    /* src/rlSkin.cpp:236 */
}

/* ================================================================ exports */
const char *oracle_kind(void) { return "port"; }
int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_flag_probe(int on) { (void)on; }   /* the port derives the flag from its own code path: free */
void oracle_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}

static inline void st3(rls_vec3 o, size_t i, v3 v) { o.x[i] = v.x; o.y[i] = v.y; o.z[i] = v.z; }

void oracle_ggx_eval_sample(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                            const float *rx, const float *ry, rls_vec3 out_wi, float *out_fresnel)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        float F;
        v3 L = ggx_eval_sample(&g, rx[i], ry[i], &F);
        st3(out_wi, i, L);
        if (out_fresnel) out_fresnel[i] = F;
    }
}
void oracle_ggx_eval_brdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                          rls_cvec3 wi, rls_vec3 out_f)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        st3(out_f, i, ggx_eval_brdf(&g, mk3(wi.x[i], wi.y[i], wi.z[i])));
    }
}
void oracle_ggx_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                         rls_cvec3 wi, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        out_pdf[i] = ggx_eval_pdf(&g, mk3(wi.x[i], wi.y[i], wi.z[i]));
    }
}
/* src/rlGgx.h:277-291, :316-328, :294-301 at caller-supplied directions */
void oracle_ggx_refract_direction(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                  rls_cvec3 m, rls_vec3 out_wi, uint32_t *out_flags)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        v3 t = mk3(0.0f, 0.0f, 0.0f);
        int ok = ggx_refract_direction(&g, mk3(m.x[i], m.y[i], m.z[i]), g.wo, &t);
        st3(out_wi, i, ok ? t : mk3(0.0f, 0.0f, 0.0f));
        if (out_flags) out_flags[i] = (ok ? 0u : RLS_FLAG_TIR) | (g.entering ? RLS_FLAG_ENTERING : 0u);
    }
}
void oracle_ggx_eval_btdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 wi, float *out_ft)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        out_ft[i] = ggx_refraction(&g, g.wo, mk3(wi.x[i], wi.y[i], wi.z[i]), g.N);
    }
}
void oracle_ggx_sample_weight(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                              rls_cvec3 wi, rls_cvec3 m, float *out_weight)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        out_weight[i] = ggx_sample_weight(&g, g.wo, mk3(wi.x[i], wi.y[i], wi.z[i]), mk3(m.x[i], m.y[i], m.z[i]));
    }
}
void oracle_ggx_sample_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                const float *rx, const float *ry, const rls_bsdf_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        float F;
        v3 L = ggx_eval_sample(&g, rx[i], ry[i], &F);
        v3 f = ggx_eval_brdf(&g, L);
        float pdf = ggx_eval_pdf(&g, L);
        st3(out->wi, i, L);
        st3(out->f, i, f);
        out->pdf[i] = pdf;
        if (out->fresnel) out->fresnel[i] = F;
        uint32_t fl = bsdf_flags(L, g.N, f, pdf);
        if (g.entering) fl |= RLS_FLAG_ENTERING;
        if (g_slope_early_out) fl |= RLS_FLAG_SLOPE_EARLY_OUT;      /* set by ggx_eval_sample above */
        out->flags[i] = fl;
    }
}
void oracle_ggx_dielectric_sample_eval_pdf(size_t n, const rls_shading_soa *sg,
                                           const rls_ggx_params *p, const float *rx,
                                           const float *ry, const rls_ggx_dielectric_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        v3 U, V, N, wo; int back;
        load_shading(sg, i, &U, &V, &N, &wo, &back);
        dielectric_t r = dielectric_unit(U, V, N, wo, back, orc_p1(&p->ior, i),
                                         orc_p1(&p->specularRoughness, i), orc_p1(&p->anisotropic, i),
                                         rx[i], ry[i], p->normal_sampler);
        out->fresnel[i] = r.F;
        st3(out->wi_r, i, r.wi_r);
        out->f_r[i] = r.f_r;
        out->pdf_r[i] = r.pdf_r;
        st3(out->wi_t, i, r.wi_t);
        out->f_t[i] = r.f_t;
        out->weight_t[i] = r.w_t;
        out->flags[i] = r.flags;
    }
}

void oracle_disney_eval_sample(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                               int sample_type, const float *rx, const float *ry,
                               rls_vec3 out_wi, uint32_t *out_flags)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        disney_t d; disney_init(&d, sg, p, i);
        uint32_t lobe = 0;
        v3 L = sample_type == RLS_RAY_DIFFUSE ? disney_sample_diffuse(&d, rx[i], ry[i])
                                              : disney_sample_specular(&d, rx[i], ry[i], &lobe);
        st3(out_wi, i, L);
        if (out_flags) {
            uint32_t fl = 0;
            if (iszero3(L)) fl |= RLS_FLAG_ZERO_L;
            if (dot3(L, d.N) <= 0.0f) fl |= RLS_FLAG_BELOW_HORIZON;
            fl |= lobe << RLS_FLAG_LOBE_SHIFT;
            if (sample_type != RLS_RAY_DIFFUSE && g_slope_early_out) fl |= RLS_FLAG_SLOPE_EARLY_OUT;
            out_flags[i] = fl;
        }
    }
}
void oracle_disney_eval_brdf(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                             int sample_type, rls_cvec3 wi, rls_vec3 out_f)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        disney_t d; disney_init(&d, sg, p, i);
        st3(out_f, i, disney_eval_brdf(&d, sample_type, mk3(wi.x[i], wi.y[i], wi.z[i])));
    }
}
void oracle_disney_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                            int sample_type, rls_cvec3 wi, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        disney_t d; disney_init(&d, sg, p, i);
        out_pdf[i] = disney_eval_pdf(&d, sample_type, mk3(wi.x[i], wi.y[i], wi.z[i]));
    }
}
void oracle_disney_sample_eval_pdf(size_t n, const rls_shading_soa *sg,
                                   const rls_disney_params *p, const float *rx_s,
                                   const float *ry_s, const float *rx_d, const float *ry_d,
                                   const rls_disney_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        disney_t d; disney_init(&d, sg, p, i);
        uint32_t lobe = 0;
        v3 Ls = disney_sample_specular(&d, rx_s[i], ry_s[i], &lobe);
        const int early = g_slope_early_out;
        v3 fs = disney_eval_brdf(&d, RLS_RAY_GLOSSY, Ls);
        float ps = disney_eval_pdf(&d, RLS_RAY_GLOSSY, Ls);
        v3 Ld = disney_sample_diffuse(&d, rx_d[i], ry_d[i]);
        v3 fd = disney_eval_brdf(&d, RLS_RAY_DIFFUSE, Ld);
        float pd = disney_eval_pdf(&d, RLS_RAY_DIFFUSE, Ld);
        st3(out->wi_s, i, Ls); st3(out->f_s, i, fs); out->pdf_s[i] = ps;
        st3(out->wi_d, i, Ld); st3(out->f_d, i, fd); out->pdf_d[i] = pd;
        uint32_t fls = bsdf_flags(Ls, d.N, fs, ps) & ~RLS_FLAG_PDF_FLOORED;
        fls |= lobe << RLS_FLAG_LOBE_SHIFT;
        if (early) fls |= RLS_FLAG_SLOPE_EARLY_OUT;
        uint32_t fld = bsdf_flags(Ld, d.N, fd, pd);
        out->flags[i] = fls | (fld << RLS_FLAG_DIFFUSE_SHIFT);
    }
}

void oracle_ndprofile_set_distance(size_t n, rls_cvec3 dist, rls_cvec3 albedo,
                                   const rls_ndprofile_soa *o)
{
    (void)albedo;   /* only feeds the dead `s` of src/rlSss.cpp:22-23 */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ndprofile_t p;
        nd_set_distance(&p, mk3(dist.x[i], dist.y[i], dist.z[i]));
        st3(o->distance, i, mk3(p.d[0], p.d[1], p.d[2]));
        st3(o->C1, i, mk3(p.C1[0], p.C1[1], p.C1[2]));
        st3(o->C2, i, mk3(p.C2[0], p.C2[1], p.C2[2]));
        o->max_radius[i] = p.R;
    }
}
void oracle_ndprofile_get_radius(size_t n, const rls_ndprofile_soa *profile, const float *rx,
                                 float *out_r, uint32_t *out_flags)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ndprofile_t p; nd_load(profile, i, &p);
        out_r[i] = nd_get_radius(&p, rx[i]);
        if (out_flags) out_flags[i] = nd_flags(&p, rx[i]);
    }
}
void oracle_ndprofile_get_pdf(size_t n, const rls_ndprofile_soa *profile, const float *r, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ndprofile_t p; nd_load(profile, i, &p);
        out_pdf[i] = nd_get_pdf(&p, r[i]);
    }
}
void oracle_ndprofile_eval_profile(size_t n, const rls_ndprofile_soa *profile, const float *r, rls_vec3 out_rd)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ndprofile_t p; nd_load(profile, i, &p);
        st3(out_rd, i, nd_eval_profile(&p, r[i]));
    }
}
/* ------------------------------------------------------- GaussianProfile
 * src/rlSss.h:63-97 (alternative Profile argument of SssSampler; fast_exp = expf per the shim). */
typedef struct { float var, R, norm; } gaussprofile_t;
/* :71-76 */
static void gauss_set_distance(gaussprofile_t *p, float dist_x)
{
    p->R = dist_x;
    p->var = (p->R * p->R) / 12.46f;
    p->norm = 1.0f - expf(-(p->R * p->R) * 0.5f / p->var);
}
/* :78-81 */
static float gauss_get_radius(const gaussprofile_t *p, float rx)
{
    return sqrtf(-2.0f * p->var * logf(1.0f - rx * p->norm));
}
/* :88-91 */
static float gauss_eval_profile(const gaussprofile_t *p, float r)
{
    return 0.15915494309189533577f / p->var * expf(-r * r * 0.5f / p->var);
}
/* :83-86 */
static float gauss_get_pdf(const gaussprofile_t *p, float r) { return gauss_eval_profile(p, r) / p->norm; }
static inline void gauss_load(const rls_gaussprofile_soa *s, size_t i, gaussprofile_t *p)
{
    p->var = s->variance[i]; p->R = s->max_radius[i]; p->norm = s->norm[i];
}
void oracle_gaussprofile_set_distance(size_t n, rls_cvec3 dist, rls_cvec3 albedo, const rls_gaussprofile_soa *o)
{
    (void)albedo;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        gaussprofile_t p; gauss_set_distance(&p, dist.x[i]);
        o->variance[i] = p.var; o->max_radius[i] = p.R; o->norm[i] = p.norm;
    }
}
void oracle_gaussprofile_get_radius(size_t n, const rls_gaussprofile_soa *profile, const float *rx, float *out_r)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { gaussprofile_t p; gauss_load(profile, i, &p); out_r[i] = gauss_get_radius(&p, rx[i]); }
}
void oracle_gaussprofile_get_pdf(size_t n, const rls_gaussprofile_soa *profile, const float *r, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { gaussprofile_t p; gauss_load(profile, i, &p); out_pdf[i] = gauss_get_pdf(&p, r[i]); }
}
void oracle_gaussprofile_eval_profile(size_t n, const rls_gaussprofile_soa *profile, const float *r, float *out_rd)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { gaussprofile_t p; gauss_load(profile, i, &p); out_rd[i] = gauss_eval_profile(&p, r[i]); }
}
void oracle_gaussprofile_sample_eval_pdf(size_t n, const float *dist_x, const float *rx, float *out_r,
                                         float *out_pdf, float *out_rd)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        gaussprofile_t p; gauss_set_distance(&p, dist_x[i]);
        float r = gauss_get_radius(&p, rx[i]);
        out_r[i] = r; out_pdf[i] = gauss_get_pdf(&p, r); out_rd[i] = gauss_eval_profile(&p, r);
    }
}

void oracle_skin_profile_sample_eval_pdf(size_t n, const rls_skin_params *sp, const float *rx,
                                         const rls_profile_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ndprofile_t p;
        nd_set_distance(&p, skin_scatter_dist(sp, i));
        float r = nd_get_radius(&p, rx[i]);
        out->r[i] = r;
        out->pdf[i] = nd_get_pdf(&p, r);
        st3(out->Rd, i, nd_eval_profile(&p, r));
        out->flags[i] = nd_flags(&p, rx[i]);
    }
}
/* src/rlSkin.cpp:191,204,214,228,231,238 */
void oracle_skin_layer_weights(size_t n, const rls_skin_params *sp, const float *avg_f_sheen,
                               const float *avg_f_spec, float *out_spec_scale, float *out_sss_weight)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float sheenWeight = orc_p1(&sp->sheen_weight, i);
        float specularWeight = orc_p1(&sp->specular_weight, i);
        float sssWeight = orc_p1(&sp->sss_weight, i);
        float sheenFresnel = 0.0f, specularFresnel = 0.0f;
        if (sheenWeight > EPS) sheenFresnel = avg_f_sheen[i] * sheenWeight;
        if (specularWeight > EPS) specularFresnel = avg_f_spec[i] * specularWeight;
        out_spec_scale[i] = specularWeight * (1.0f - sheenFresnel);
        sssWeight *= 1.0f - specularFresnel * (1.0f - sheenFresnel);
        out_sss_weight[i] = sssWeight;
    }
}
/* src/rlSss.h:487-533 (getProbeRay), origin = 0 */
void oracle_skin_probe_ray(size_t n, const rls_shading_soa *sg, const rls_skin_params *sp,
                           const float *rx_in, const float *ry_in, const rls_probe_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        v3 U, V, N, wo; int back;
        load_shading(sg, i, &U, &V, &N, &wo, &back);
        ndprofile_t p;
        nd_set_distance(&p, skin_scatter_dist(sp, i));
        float rx = rx_in[i], ry = ry_in[i];
        int idx;
        if (rx < 0.5f) { idx = 0; rx = linearstep(0.0f, 0.5f, rx); }
        else if (rx < 0.75f) { idx = 2; rx = linearstep(0.5f, 0.75f, rx); }
        else { idx = 3; rx = linearstep(0.75f, 1.0f, rx); }
        float r = nd_get_radius(&p, rx);
        float rmax = p.R;
        float phi = TWO_PI_F * ry;
        v3 offset;
        offset.x = cosf(phi) * r;
        offset.z = sinf(phi) * r;
        offset.y = sqrtf(rmax * rmax - r * r);
        float maxdist = (float)((double)(offset.y * 2.0f));
        v3 dir;
        if ((idx & 0x03) < 2) {
            dir = neg3(N);
            offset = rotate_to_frame(offset, U, neg3(dir), V);
        } else if ((idx & 0x03) == 2) {
            dir = (idx & 0x04) > 0 ? neg3(U) : U;
            offset = rotate_to_frame(offset, V, neg3(dir), N);
        } else {
            dir = (idx & 0x04) > 0 ? neg3(V) : V;
            offset = rotate_to_frame(offset, N, neg3(dir), U);
        }
        out->r[i] = r;
        st3(out->origin, i, add3(mk3(0.0f, 0.0f, 0.0f), offset));
        st3(out->dir, i, dir);
        out->maxdist[i] = maxdist;
        out->flags[i] = nd_flags(&p, rx) | ((uint32_t)idx << RLS_FLAG_PROBE_AXIS_SHIFT);
    }
}

/* src/rlSss.h:246-266: world->local offset (shim AiM4Frame), three projected radii, MIS pdf */
void oracle_skin_probe_mis_pdf(size_t n, const rls_shading_soa *sg, const rls_skin_params *sp,
                               rls_cvec3 disp, rls_cvec3 hit_normal, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        v3 U, V, N, wo; int back;
        load_shading(sg, i, &U, &V, &N, &wo, &back);
        ndprofile_t p;
        nd_set_distance(&p, skin_scatter_dist(sp, i));
        v3 dp = mk3(disp.x[i], disp.y[i], disp.z[i]);
        v3 hn = mk3(hit_normal.x[i], hit_normal.y[i], hit_normal.z[i]);
        v3 offset = mk3(dot3(dp, U), dot3(dp, V), dot3(dp, N));
        offset = mul3(offset, offset);
        float rr0 = sqrtf(offset.y + offset.z);
        float rr1 = sqrtf(offset.x + offset.z);
        float rr2 = sqrtf(offset.x + offset.y);
        float du = dot3(U, hn), dv = dot3(V, hn), dn = dot3(N, hn);
        out_pdf[i] = nd_get_pdf(&p, rr0) * ABSF(du) * 0.25f
                   + nd_get_pdf(&p, rr1) * ABSF(dv) * 0.25f
                   + nd_get_pdf(&p, rr2) * ABSF(dn) * 0.5f;
    }
}

/* ---- rlSkin's glossy layers (src/rlSkin.cpp:184-238) */
static void skin_layer(v3 U, v3 V, v3 N, v3 wo, int back, const float *c, float ior, float rough, uint32_t K, size_t P,
                       size_t p, const float *rx, const float *ry, rls_cvec3 li, float *avg, float *est)
{
    ggx_t g;
    v3 color = mk3(c[0], c[1], c[2]);
    ggx_init(&g, U, V, N, wo, back, color, ior, rough, 0.0f);          /* src/rlSkin.cpp:192,215 */
    float reflectWeight = 0.0f, count = 0.0f;                          /* src/rlGgx.h:371-372 */
    float acc[3] = { 0.0f, 0.0f, 0.0f };
    int small = ABSF(color.x) < EPS && ABSF(color.y) < EPS && ABSF(color.z) < EPS;   /* src/rlGgx.h:174-176 */
    if (!small) {
        for (uint32_t k = 0; k < K; k++) {
            size_t idx = (size_t)k * P + p;
            v3 M = ggx_sample_normal(&g, rx[idx], ry[idx]);            /* src/rlGgx.h:97-107 */
            v3 L = reflect_direction(g.wo, M);
            reflectWeight += ggx_fresnel(&g, L, M);
            count += 1.0f;
            v3 f = ggx_eval_brdf(&g, L);
            float pdf = ggx_eval_pdf(&g, L);
            float w[3] = { f.x / pdf, f.y / pdf, f.z / pdf };
            if (li.x) { w[0] *= li.x[idx]; w[1] *= li.y[idx]; w[2] *= li.z[idx]; }
            acc[0] += w[0]; acc[1] += w[1]; acc[2] += w[2];
        }
    }
    *avg = count > 0.0f ? reflectWeight / count : 1.0f;                /* src/rlGgx.h:181-184 */
    float invK = 1.0f / (float)K;
    est[0] = acc[0] * invK; est[1] = acc[1] * invK; est[2] = acc[2] * invK;
}

void oracle_skin_glossy_layers(size_t n, uint32_t K, const rls_shading_soa *sg, const rls_skin_params *sp,
                               const float *rx_a, const float *ry_a, const float *rx_b, const float *ry_b,
                               rls_cvec3 li_a, rls_cvec3 li_b, const rls_skin_layers_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        v3 U, V, N, wo; int back;
        load_shading(sg, i, &U, &V, &N, &wo, &back);
        uint32_t flags = 0;
        float sheenFresnel = 0.0f, specularFresnel = 0.0f;
        float sheen[3] = { 0, 0, 0 }, specular[3] = { 0, 0, 0 }, c[3], avg;
        float sheenWeight = orc_p1(&sp->sheen_weight, i);
        if (sheenWeight > EPS) {                                                           /* :191 */
            orc_p3(&sp->sheen_color, i, c);
            skin_layer(U, V, N, wo, back, c, orc_p1(&sp->sheen_ior, i), orc_p1(&sp->sheen_roughness, i), K, n, i,
                       rx_a, ry_a, li_a, &avg, sheen);
            sheenFresnel = avg * sheenWeight;                                              /* :204 */
            flags |= RLS_SKIN_SHEEN_EVALUATED;
        }
        for (int j = 0; j < 3; j++) sheen[j] *= sheenWeight;                               /* :207 */
        float specularWeight = orc_p1(&sp->specular_weight, i);
        if (specularWeight > EPS) {                                                        /* :214 */
            orc_p3(&sp->specular_color, i, c);
            skin_layer(U, V, N, wo, back, c, orc_p1(&sp->specular_ior, i), orc_p1(&sp->specular_roughness, i), K, n, i,
                       rx_b, ry_b, li_b, &avg, specular);
            specularFresnel = avg * specularWeight;                                        /* :228 */
            flags |= RLS_SKIN_SPECULAR_EVALUATED;
        }
        float scale = specularWeight * (1.0f - sheenFresnel);                              /* :231 */
        for (int j = 0; j < 3; j++) specular[j] *= scale;
        float sssWeight = orc_p1(&sp->sss_weight, i);
        sssWeight *= 1.0f - specularFresnel * (1.0f - sheenFresnel);                       /* :238 */
        if (sssWeight < EPS) flags |= RLS_SKIN_SSS_SKIPPED;                                /* :244 */
        st3(out->sheen, i, mk3(sheen[0], sheen[1], sheen[2]));
        st3(out->specular, i, mk3(specular[0], specular[1], specular[2]));
        out->sheen_fresnel[i] = sheenFresnel;
        out->specular_fresnel[i] = specularFresnel;
        out->sss_weight[i] = sssWeight;
        out->flags[i] = flags;
    }
}

/* ---- 8(f) f3: one MIS light sample (two-sample power heuristic; include/rls_b200.h) */
static inline float power_heuristic(float a, float b) { float a2 = a * a; return a2 / (a2 + b * b); }
typedef struct { v3 rgb; float wl, wb; } mis_t;
static mis_t mis_combine(v3 Ld, v3 Li, float pl, v3 fl, float pbl, int have, v3 L, v3 fb, float pb, v3 Lib, float plb)
{
    mis_t o; o.rgb = mk3(0.0f, 0.0f, 0.0f); o.wl = 0.0f; o.wb = 0.0f;
    if (!iszero3(Ld) && pl > 0.0f) {
        o.wl = power_heuristic(pl, pbl);
        float s = o.wl / pl;
        o.rgb = mk3(fl.x * Li.x * s, fl.y * Li.y * s, fl.z * Li.z * s);
    }
    if (have && !iszero3(L) && pb > 0.0f) {
        o.wb = power_heuristic(pb, plb);
        float s = o.wb / pb;
        o.rgb = add3(o.rgb, mk3(fb.x * Lib.x * s, fb.y * Lib.y * s, fb.z * Lib.z * s));
    }
    return o;
}
static inline v3 ld3(rls_cvec3 v, size_t i) { return mk3(v.x[i], v.y[i], v.z[i]); }

void oracle_ggx_evaluate_light_sample(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                      const rls_light_sample *light, const float *rx, const float *ry,
                                      const rls_light_sample *at_l, rls_vec3 out_rgb, float *wl, float *wb)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        ggx_t g; ggx_from_params(&g, sg, p, i);
        v3 Ld = ld3(light->dir, i);
        v3 fl = ggx_eval_brdf(&g, Ld);
        float pbl = ggx_eval_pdf(&g, Ld);
        v3 L = mk3(0.0f, 0.0f, 0.0f), fb = L, Lib = L;
        float pb = 0.0f, plb = 0.0f;
        if (at_l) {
            L = ggx_eval_sample(&g, rx[i], ry[i], NULL);
            fb = ggx_eval_brdf(&g, L);
            pb = ggx_eval_pdf(&g, L);
            Lib = ld3(at_l->radiance, i);
            plb = at_l->pdf[i];
        }
        mis_t m = mis_combine(Ld, ld3(light->radiance, i), light->pdf[i], fl, pbl, at_l != NULL, L, fb, pb, Lib, plb);
        st3(out_rgb, i, m.rgb);
        if (wl) wl[i] = m.wl;
        if (wb) wb[i] = m.wb;
    }
}

void oracle_disney_evaluate_light_sample(size_t n, const rls_shading_soa *sg, const rls_disney_params *p, int sample_type,
                                         const rls_light_sample *light, const float *rx, const float *ry,
                                         const rls_light_sample *at_l, rls_vec3 out_rgb, float *wl, float *wb)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        disney_t d; disney_init(&d, sg, p, i);
        v3 Ld = ld3(light->dir, i);
        v3 fl = disney_eval_brdf(&d, sample_type, Ld);
        float pbl = disney_eval_pdf(&d, sample_type, Ld);
        v3 L = mk3(0.0f, 0.0f, 0.0f), fb = L, Lib = L;
        float pb = 0.0f, plb = 0.0f;
        if (at_l) {
            uint32_t lobe = 0;
            L = sample_type == RLS_RAY_DIFFUSE ? disney_sample_diffuse(&d, rx[i], ry[i])
                                               : disney_sample_specular(&d, rx[i], ry[i], &lobe);
            fb = disney_eval_brdf(&d, sample_type, L);
            pb = disney_eval_pdf(&d, sample_type, L);
            Lib = ld3(at_l->radiance, i);
            plb = at_l->pdf[i];
        }
        mis_t m = mis_combine(Ld, ld3(light->radiance, i), light->pdf[i], fl, pbl, at_l != NULL, L, fb, pb, Lib, plb);
        st3(out_rgb, i, m.rgb);
        if (wl) wl[i] = m.wl;
        if (wb) wb[i] = m.wb;
    }
}

/* ---- 8(f) f4: SampleWriter (src/rlUtil.h:98-163) */
static inline void writer_pixel(float *image, int W, int H, int x, int y, v3 rgb)        /* writePixel :158-163 */
{
    size_t stride = (size_t)W * H, at = (size_t)x + (size_t)y * W;
    image[at] = rgb.z; image[at + stride] = rgb.y; image[at + stride * 2] = rgb.x;
}
typedef struct { int node, type; ggx_t g; disney_t d; } writer_brdf_t;
static void writer_setup(writer_brdf_t *b, int node, const rls_shading_soa *sg, const void *params, size_t point, int type)
{
    b->node = node; b->type = type;
    if (node == RLS_NODE_GGX) ggx_from_params(&b->g, sg, (const rls_ggx_params *)params, point);
    else disney_init(&b->d, sg, (const rls_disney_params *)params, point);
}
void oracle_sample_writer_radiance(int node, const rls_shading_soa *sg, const void *params, size_t point, int sample_type,
                                   int W, int H, float *image)
{
    writer_brdf_t b; writer_setup(&b, node, sg, params, point, sample_type);
    for (int j = 0; j < H; j++) {                                                        /* :103-113 */
        float theta = HALF_PI_F * j / H;
        for (int i = 0; i < W; i++) {
            float phi = TWO_PI_F * i / W;
            v3 dir = spherical_direction(cosf(theta), phi);
            v3 c = node == RLS_NODE_GGX ? ggx_eval_brdf(&b.g, dir) : disney_eval_brdf(&b.d, sample_type, dir);
            writer_pixel(image, W, H, i, j, c);
        }
    }
}
void oracle_sample_writer_scatter(int node, const rls_shading_soa *sg, const void *params, size_t point, int sample_type,
                                  size_t n, const float *rx, const float *ry, int W, int H, float *image, uint32_t *missing)
{
    writer_brdf_t b; writer_setup(&b, node, sg, params, point, sample_type);
    uint32_t missingCount = 0;
    for (size_t k = 0; k < n; k++) {                                                     /* :127-150 */
        uint32_t lobe = 0;
        v3 dir = node == RLS_NODE_GGX ? ggx_eval_sample(&b.g, rx[k], ry[k], NULL)
               : (sample_type == RLS_RAY_DIFFUSE ? disney_sample_diffuse(&b.d, rx[k], ry[k])
                                                 : disney_sample_specular(&b.d, rx[k], ry[k], &lobe));
        if (iszero3(dir)) continue;
        float theta = acosf(dir.z);
        float phi = atan2f(dir.y, dir.x);
        if (phi < 0.0f) phi += TWO_PI_F;
        int i = (int)(phi * 0.15915494309189533577f * W);
        int j = (int)(theta / HALF_PI_F * H);
        i = i < 0 ? 0 : (i > W - 1 ? W - 1 : i);
        j = j < 0 ? 0 : (j > H - 1 ? H - 1 : j);
        if (theta > HALF_PI_F) { writer_pixel(image, W, H, i, j, mk3(1.0f, 0.0f, 0.0f)); missingCount++; }
        else writer_pixel(image, W, H, i, j, mk3(0.0f, 1.0f, 0.0f));
    }
    if (missing) *missing = missingCount;
}

void oracle_albedo_sweep(const rls_sweep_grid *g, uint64_t seed, uint32_t spp_begin,
                         uint32_t spp_end, double *table)
{
    const uint32_t cells = (uint32_t)(g->n_rough * g->n_cos * g->n_ior);
#pragma omp parallel for schedule(dynamic, 16)
    for (uint32_t cell = 0; cell < cells; cell++) {
        float rough, cosv, ior;
        orc_sweep_cell(g, cell, &rough, &cosv, &ior);
        v3 U = mk3(1.0f, 0.0f, 0.0f), V = mk3(0.0f, 1.0f, 0.0f), N = mk3(0.0f, 0.0f, 1.0f);
        v3 wo = mk3(sqrtf(1.0f - cosv * cosv), 0.0f, cosv);
        double acc[RLS_SWEEP_VALUES_PER_CELL] = { 0, 0, 0, 0, 0 };
        for (uint32_t k = spp_begin; k < spp_end; k++) {
            uint64_t idx = ((uint64_t)cell << 32) | (uint64_t)k;
            float rx = orc_uniform(seed, 0u, idx);
            float ry = orc_uniform(seed, 1u, idx);
            dielectric_t r = dielectric_unit(U, V, N, wo, 0, ior, rough, 0.0f, rx, ry, RLS_GGX_SAMPLER_VNDF);
            int valid = !(r.flags & (RLS_FLAG_ZERO_L | RLS_FLAG_BELOW_HORIZON));
            if (valid) { acc[0] += (double)(r.f_r / r.pdf_r); acc[3] += 1.0; }
            if (r.flags & RLS_FLAG_TIR) acc[4] += 1.0; else acc[1] += (double)r.w_t;
            acc[2] += (double)r.F;
        }
        for (int j = 0; j < RLS_SWEEP_VALUES_PER_CELL; j++) table[(size_t)cell * RLS_SWEEP_VALUES_PER_CELL + j] = acc[j];
    }
}

void oracle_synth_uniform(size_t n, uint64_t seed, uint32_t stream, uint64_t first_index,
                          float lo, float hi, float *out)
{
    for (size_t i = 0; i < n; i++) out[i] = lo + (hi - lo) * orc_uniform(seed, stream, first_index + i);
}
