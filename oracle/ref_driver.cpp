/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (kind "reference").
 *
 * Drives the reference's own, unmodified C++ (included / linked from
 * /root/reference/src where it lies; nothing is copied into this repository) through
 * the batched oracle_api.h surface.  One sampler object is constructed per sample,
 * exactly as the reference constructs one per shading point (src/rlGgx.cpp:261,
 * src/rlDisney.cpp:690, src/rlSkin.cpp:241), including its make_shared
 * (src/rlGgx.h:152).
 *
 * `#define private public` is applied to the reference headers in THIS translation
 * unit only, to reach mNormalSampler (src/rlGgx.h:360), NDProfile::mC1/mC2
 * (src/rlSss.h:59) and SssSampler::getProbeRay (src/rlSss.h:487).
 */
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <ai.h>

#define private public
#include "rlUtil.h"
#include "rlGgx.h"
#include "rlSss.h"
#include "rlDisney.cpp"   /* DisneySampler is file-local (src/rlDisney.cpp:105) */
#undef private

#include "oracle_common.h"

namespace {

struct Shading {
    AtShaderGlobals sg;
    AtVector U, V;
};

inline void loadShading(const rls_shading_soa *s, size_t i, Shading &o)
{
    std::memset(&o.sg, 0, sizeof(o.sg));
    AiV3Create(o.U, s->U.x[i], s->U.y[i], s->U.z[i]);
    AiV3Create(o.V, s->V.x[i], s->V.y[i], s->V.z[i]);
    AiV3Create(o.sg.Nf, s->N.x[i], s->N.y[i], s->N.z[i]);
    bool back = s->backfacing && s->backfacing[i];
    o.sg.N = back ? -o.sg.Nf : o.sg.Nf;
    o.sg.Ns = o.sg.Nf;
    o.sg.Ng = o.sg.N;
    o.sg.Ngf = o.sg.Nf;
    AtVector wo;
    AiV3Create(wo, s->wo.x[i], s->wo.y[i], s->wo.z[i]);
    o.sg.Rd = -wo;
    rls_shim_set_frame(&o.U, &o.V);
}

inline void store3(rls_vec3 o, size_t i, float a, float b, float c) { o.x[i] = a; o.y[i] = b; o.z[i] = c; }

struct GgxArgs { AtColor ks; float ior, rough, aniso; };
inline GgxArgs ggxArgs(const rls_ggx_params *p, size_t i)
{
    GgxArgs a;
    float c[3];
    orc_p3(&p->KsColor, i, c);
    a.ks = rls_shim_rgb(c[0], c[1], c[2]);
    a.ior = orc_p1(&p->ior, i);
    a.rough = orc_p1(&p->specularRoughness, i);
    a.aniso = orc_p1(&p->anisotropic, i);
    return a;
}

inline uint32_t bsdfFlags(const AtVector &L, const AtVector &N, const AtColor &f, float pdf)
{
    uint32_t fl = 0;
    if (L == AI_V3_ZERO) fl |= RLS_FLAG_ZERO_L;
    if (AiV3Dot(L, N) <= 0.0f) fl |= RLS_FLAG_BELOW_HORIZON;
    if (pdf == 0.0f) fl |= RLS_FLAG_PDF_ZERO;
    if (f == AI_RGB_BLACK) fl |= RLS_FLAG_F_BLACK;
    if (pdf == AI_EPSILON) fl |= RLS_FLAG_PDF_FLOORED;
    return fl;
}

/* RLS_FLAG_SLOPE_EARLY_OUT: see orc_vndf_early_out (oracle_common.h).  Only VNDFKernel has the early-outs. */
template <typename Sampler> struct UsesVndf { static const bool value = false; };
template <> struct UsesVndf<rls::GgxSampler> { static const bool value = true; };
/* The probe re-evaluates part of the sampler: it is builder code, so a TIMED run of the reference (bench.py
 * --impl reference, cpu_baseline) switches it off with oracle_set_flag_probe(0) and the bit is then simply absent. */
static int g_flag_probe = 1;
inline uint32_t slopeEarlyOutFlag(const Shading &sh, float ax, float ay, float rx)
{
    if (!g_flag_probe) return 0u;
    const float view[3] = { -sh.sg.Rd.x, -sh.sg.Rd.y, -sh.sg.Rd.z };
    const float U[3] = { sh.U.x, sh.U.y, sh.U.z }, V[3] = { sh.V.x, sh.V.y, sh.V.z };
    const float N[3] = { sh.sg.Nf.x, sh.sg.Nf.y, sh.sg.Nf.z };
    return orc_vndf_early_out(view, U, V, N, ax, ay, rx) ? RLS_FLAG_SLOPE_EARLY_OUT : 0u;
}

/* The dielectric unit of work, composed from the reference's own members. */
struct DielectricResult {
    float F, f_r, pdf_r, f_t, w_t;
    AtVector wi_r, wi_t;
    uint32_t flags;
};
template <typename Sampler>
inline DielectricResult dielectricUnitT(Shading &sh, float ior, float rough, float aniso, float rx, float ry)
{
    DielectricResult r;
    Sampler s(&sh.sg, AI_RGB_WHITE, ior, rough, aniso);
    const AtVector V = s.mViewDir;
    const AtVector N = s.mAxisN;
    AtVector m = s.mNormalSampler->evalSample(rx, ry);
    r.wi_r = rls::reflectDirection(V, m);
    r.F = s.fresnel(r.wi_r, m);
    AtColor fr = Sampler::evalBrdf(&s, &r.wi_r);
    r.f_r = fr.r;
    r.pdf_r = Sampler::evalPdf(&s, &r.wi_r);
    r.flags = bsdfFlags(r.wi_r, N, fr, r.pdf_r);
    if (AiV3Dot(sh.sg.N, sh.sg.Rd) < AI_EPSILON) r.flags |= RLS_FLAG_ENTERING;
    if (UsesVndf<Sampler>::value) r.flags |= slopeEarlyOutFlag(sh, s.mAlphaX, s.mAlphaY, rx);
    AtVector t;
    if (s.getRefractDirection(m, V, t)) {
        r.wi_t = t;
        r.f_t = s.refraction(V, t, N);
    } else {
        r.wi_t = rls::reflectDirection(V, m);
        r.f_t = 0.0f;
        r.flags |= RLS_FLAG_TIR;
    }
    r.w_t = s.getSampleWeight(V, r.wi_t, m);
    return r;
}

typedef rls::GgxSamplerT<rls::NDFKernel> GgxNdfSampler;   /* src/rlGgx.h:24-56 as the template argument */
inline DielectricResult dielectricUnit(Shading &sh, float ior, float rough, float aniso, float rx, float ry, int kernel = 0)
{
    return kernel == RLS_GGX_SAMPLER_NDF ? dielectricUnitT<GgxNdfSampler>(sh, ior, rough, aniso, rx, ry)
                                         : dielectricUnitT<rls::GgxSampler>(sh, ior, rough, aniso, rx, ry);
}

inline void disneyTable(const rls_disney_params *p, size_t i, float *t /* [64*3] */)
{
    float c[3];
    orc_p3(&p->base_color, i, c);
    t[p_base_color * 3 + 0] = c[0]; t[p_base_color * 3 + 1] = c[1]; t[p_base_color * 3 + 2] = c[2];
    t[p_subsurface * 3] = orc_p1(&p->subsurface, i);
    t[p_metallic * 3] = orc_p1(&p->metallic, i);
    t[p_Ks * 3] = orc_p1(&p->specular, i);
    t[p_specular_tint * 3] = orc_p1(&p->specular_tint, i);
    t[p_roughness * 3] = orc_p1(&p->roughness, i);
    t[p_anisotropic * 3] = orc_p1(&p->anisotropic, i);
    t[p_sheen * 3] = orc_p1(&p->sheen, i);
    t[p_sheen_tint * 3] = orc_p1(&p->sheen_tint, i);
    t[p_clearcoat * 3] = orc_p1(&p->clearcoat, i);
    t[p_clearcoat_gloss * 3] = orc_p1(&p->clearcoat_gloss, i);
}

inline uint32_t disneyLobe(const rls_disney_params *p, size_t i, float rx)
{
    /* Same float expression as src/rlDisney.cpp:169,373-375. */
    float clearcoat = orc_p1(&p->clearcoat, i) * 0.25f;
    float gtr2Weight = 1.0f / (clearcoat + 1.0f);
    return rx < gtr2Weight ? 0u : 1u;
}

/* RLS_FLAG_SLOPE_EARLY_OUT of the rlDisney glossy sample: GTR2 lobe, visible-normal sampling, rx rescaled as
 * src/rlDisney.cpp:376 before sampleGTR2AnisoDirectionFromSlope. */
inline uint32_t disneySlopeEarlyOut(const rls_disney_params *p, size_t i, const Shading &sh, const DisneySampler &s, float rx)
{
    if (disneyLobe(p, i, rx) != 0u || !s.mSampleFromVisibleNormal) return 0u;
    float clearcoat = orc_p1(&p->clearcoat, i) * 0.25f;
    float gtr2Weight = 1.0f / (clearcoat + 1.0f);
    return slopeEarlyOutFlag(sh, s.mAlphaX, s.mAlphaY, rx / gtr2Weight);
}

inline void loadProfile(const rls_ndprofile_soa *s, size_t i, rls::NDProfile &p)
{
    AiV3Create(p.mDistance, s->distance.x[i], s->distance.y[i], s->distance.z[i]);
    AiV3Create(p.mC1, s->C1.x[i], s->C1.y[i], s->C1.z[i]);
    AiV3Create(p.mC2, s->C2.x[i], s->C2.y[i], s->C2.z[i]);
    p.mMaxRadius = s->max_radius[i];
}

inline uint32_t profileFlags(const rls::NDProfile &p, float rx)
{
    float x = rx;
    int ch = rls::NDProfile::selectDistLobe(x);
    uint32_t fl = (uint32_t)ch << RLS_FLAG_LOBE_SHIFT;
    float d = p.mDistance[ch];
    if (p.mMaxRadius < AI_EPSILON || d < AI_EPSILON) {
        fl |= RLS_FLAG_DEGENERATE;
    } else {
        float w1 = p.mC1[ch], w2 = p.mC2[ch];
        float w = w1 / (w1 + w2 * 3.0f);
        if (x > w) fl |= RLS_FLAG_EXP_LOBE;
    }
    return fl;
}

} // namespace

template <typename Sampler>
void ggxEvalSampleT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, const float *rx, const float *ry,
                    rls_vec3 out_wi, float *out_fresnel)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        AtVector L = Sampler::evalSample(&s, rx[i], ry[i]);
        store3(out_wi, i, L.x, L.y, L.z);
        if (out_fresnel) out_fresnel[i] = s.getAvgReflectWeight();
    }
}
template <typename Sampler>
void ggxEvalBrdfT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 wi, rls_vec3 out_f)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        AtVector L; AiV3Create(L, wi.x[i], wi.y[i], wi.z[i]);
        AtColor f = Sampler::evalBrdf(&s, &L);
        store3(out_f, i, f.r, f.g, f.b);
    }
}
template <typename Sampler>
void ggxEvalPdfT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 wi, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        AtVector L; AiV3Create(L, wi.x[i], wi.y[i], wi.z[i]);
        out_pdf[i] = Sampler::evalPdf(&s, &L);
    }
}
template <typename Sampler>
void ggxSampleEvalPdfT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, const float *rx, const float *ry,
                       const rls_bsdf_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        AtVector L = Sampler::evalSample(&s, rx[i], ry[i]);
        AtColor f = Sampler::evalBrdf(&s, &L);
        float pdf = Sampler::evalPdf(&s, &L);
        store3(out->wi, i, L.x, L.y, L.z);
        store3(out->f, i, f.r, f.g, f.b);
        out->pdf[i] = pdf;
        if (out->fresnel) out->fresnel[i] = s.getAvgReflectWeight();
        uint32_t fl = bsdfFlags(L, sh.sg.Nf, f, pdf);
        if (AiV3Dot(sh.sg.N, sh.sg.Rd) < AI_EPSILON) fl |= RLS_FLAG_ENTERING;
        if (UsesVndf<Sampler>::value) fl |= slopeEarlyOutFlag(sh, s.mAlphaX, s.mAlphaY, rx[i]);
        out->flags[i] = fl;
    }
}
/* The refraction half at caller-supplied directions: the reference's own members (src/rlGgx.h:277-328). */
template <typename Sampler>
void ggxRefractDirectionT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 m, rls_vec3 out_wi, uint32_t *out_flags)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        AtVector mm, t = AI_V3_ZERO; AiV3Create(mm, m.x[i], m.y[i], m.z[i]);
        bool ok = s.getRefractDirection(mm, s.mViewDir, t);
        if (!ok) t = AI_V3_ZERO;
        store3(out_wi, i, t.x, t.y, t.z);
        if (out_flags) out_flags[i] = (ok ? 0u : RLS_FLAG_TIR) | (AiV3Dot(sh.sg.N, sh.sg.Rd) < AI_EPSILON ? RLS_FLAG_ENTERING : 0u);
    }
}
template <typename Sampler>
void ggxEvalBtdfT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 wi, float *out_ft)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        AtVector o; AiV3Create(o, wi.x[i], wi.y[i], wi.z[i]);
        out_ft[i] = s.refraction(s.mViewDir, o, s.mAxisN);
    }
}
template <typename Sampler>
void ggxSampleWeightT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 wi, rls_cvec3 m, float *out_w)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        AtVector o, mm; AiV3Create(o, wi.x[i], wi.y[i], wi.z[i]); AiV3Create(mm, m.x[i], m.y[i], m.z[i]);
        out_w[i] = s.getSampleWeight(s.mViewDir, o, mm);
    }
}
#define GGX_DISPATCH(p, fn, ...) \
    do { if ((p)->normal_sampler == RLS_GGX_SAMPLER_NDF) fn<GgxNdfSampler>(__VA_ARGS__); else fn<rls::GgxSampler>(__VA_ARGS__); } while (0)

extern "C" {

const char *oracle_kind(void) { return "reference"; }
int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_flag_probe(int on) { g_flag_probe = on; }
void oracle_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}

void oracle_ggx_eval_sample(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                            const float *rx, const float *ry, rls_vec3 out_wi, float *out_fresnel)
{
    GGX_DISPATCH(p, ggxEvalSampleT, n, sg, p, rx, ry, out_wi, out_fresnel);
}

void oracle_ggx_eval_brdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                          rls_cvec3 wi, rls_vec3 out_f)
{
    GGX_DISPATCH(p, ggxEvalBrdfT, n, sg, p, wi, out_f);
}

void oracle_ggx_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                         rls_cvec3 wi, float *out_pdf)
{
    GGX_DISPATCH(p, ggxEvalPdfT, n, sg, p, wi, out_pdf);
}

void oracle_ggx_refract_direction(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                  rls_cvec3 m, rls_vec3 out_wi, uint32_t *out_flags)
{
    GGX_DISPATCH(p, ggxRefractDirectionT, n, sg, p, m, out_wi, out_flags);
}
void oracle_ggx_eval_btdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 wi, float *out_ft)
{
    GGX_DISPATCH(p, ggxEvalBtdfT, n, sg, p, wi, out_ft);
}
void oracle_ggx_sample_weight(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                              rls_cvec3 wi, rls_cvec3 m, float *out_weight)
{
    GGX_DISPATCH(p, ggxSampleWeightT, n, sg, p, wi, m, out_weight);
}

void oracle_ggx_sample_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                const float *rx, const float *ry, const rls_bsdf_out *out)
{
    GGX_DISPATCH(p, ggxSampleEvalPdfT, n, sg, p, rx, ry, out);
}

void oracle_ggx_dielectric_sample_eval_pdf(size_t n, const rls_shading_soa *sg,
                                           const rls_ggx_params *p, const float *rx,
                                           const float *ry, const rls_ggx_dielectric_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        DielectricResult r = dielectricUnit(sh, a.ior, a.rough, a.aniso, rx[i], ry[i], p->normal_sampler);
        out->fresnel[i] = r.F;
        store3(out->wi_r, i, r.wi_r.x, r.wi_r.y, r.wi_r.z);
        out->f_r[i] = r.f_r;
        out->pdf_r[i] = r.pdf_r;
        store3(out->wi_t, i, r.wi_t.x, r.wi_t.y, r.wi_t.z);
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
        Shading sh; loadShading(sg, i, sh);
        float table[64 * 3] = { 0 };
        disneyTable(p, i, table);
        rls_shim_set_param_table(table);
        DisneySampler s(nullptr, &sh.sg);
        s.mSampleFromVisibleNormal = p->sample_from_visible_normal != 0;   /* src/rlDisney.cpp:191 */
        s.setSampleType((AtUInt16)sample_type);
        AtVector L = DisneySampler::evalSample(&s, rx[i], ry[i]);
        store3(out_wi, i, L.x, L.y, L.z);
        if (out_flags) {
            uint32_t fl = 0;
            if (AiV3IsZero(L)) fl |= RLS_FLAG_ZERO_L;
            if (AiV3Dot(L, sh.sg.Nf) <= 0.0f) fl |= RLS_FLAG_BELOW_HORIZON;
            if (sample_type != AI_RAY_DIFFUSE) fl |= disneyLobe(p, i, rx[i]) << RLS_FLAG_LOBE_SHIFT;
            if (sample_type != AI_RAY_DIFFUSE) fl |= disneySlopeEarlyOut(p, i, sh, s, rx[i]);
            out_flags[i] = fl;
        }
    }
}

void oracle_disney_eval_brdf(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                             int sample_type, rls_cvec3 wi, rls_vec3 out_f)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        float table[64 * 3] = { 0 };
        disneyTable(p, i, table);
        rls_shim_set_param_table(table);
        DisneySampler s(nullptr, &sh.sg);
        s.mSampleFromVisibleNormal = p->sample_from_visible_normal != 0;   /* src/rlDisney.cpp:191 */
        s.setSampleType((AtUInt16)sample_type);
        AtVector L; AiV3Create(L, wi.x[i], wi.y[i], wi.z[i]);
        AtColor f = DisneySampler::evalBrdf(&s, &L);
        store3(out_f, i, f.r, f.g, f.b);
    }
}

void oracle_disney_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                            int sample_type, rls_cvec3 wi, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        float table[64 * 3] = { 0 };
        disneyTable(p, i, table);
        rls_shim_set_param_table(table);
        DisneySampler s(nullptr, &sh.sg);
        s.mSampleFromVisibleNormal = p->sample_from_visible_normal != 0;   /* src/rlDisney.cpp:191 */
        s.setSampleType((AtUInt16)sample_type);
        AtVector L; AiV3Create(L, wi.x[i], wi.y[i], wi.z[i]);
        out_pdf[i] = DisneySampler::evalPdf(&s, &L);
    }
}

void oracle_disney_sample_eval_pdf(size_t n, const rls_shading_soa *sg,
                                   const rls_disney_params *p, const float *rx_s,
                                   const float *ry_s, const float *rx_d, const float *ry_d,
                                   const rls_disney_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        float table[64 * 3] = { 0 };
        disneyTable(p, i, table);
        rls_shim_set_param_table(table);
        DisneySampler s(nullptr, &sh.sg);
        s.mSampleFromVisibleNormal = p->sample_from_visible_normal != 0;   /* src/rlDisney.cpp:191 */

        s.setSampleType(AI_RAY_GLOSSY);
        AtVector Ls = DisneySampler::evalSample(&s, rx_s[i], ry_s[i]);
        AtColor fs = DisneySampler::evalBrdf(&s, &Ls);
        float ps = DisneySampler::evalPdf(&s, &Ls);

        s.setSampleType(AI_RAY_DIFFUSE);
        AtVector Ld = DisneySampler::evalSample(&s, rx_d[i], ry_d[i]);
        AtColor fd = DisneySampler::evalBrdf(&s, &Ld);
        float pd = DisneySampler::evalPdf(&s, &Ld);

        store3(out->wi_s, i, Ls.x, Ls.y, Ls.z);
        store3(out->f_s, i, fs.r, fs.g, fs.b);
        out->pdf_s[i] = ps;
        store3(out->wi_d, i, Ld.x, Ld.y, Ld.z);
        store3(out->f_d, i, fd.r, fd.g, fd.b);
        out->pdf_d[i] = pd;

        uint32_t fls = bsdfFlags(Ls, sh.sg.Nf, fs, ps) & ~RLS_FLAG_PDF_FLOORED;
        fls |= disneyLobe(p, i, rx_s[i]) << RLS_FLAG_LOBE_SHIFT;
        fls |= disneySlopeEarlyOut(p, i, sh, s, rx_s[i]);
        uint32_t fld = bsdfFlags(Ld, sh.sg.Nf, fd, pd);
        out->flags[i] = fls | (fld << RLS_FLAG_DIFFUSE_SHIFT);
    }
}

void oracle_ndprofile_set_distance(size_t n, rls_cvec3 dist, rls_cvec3 albedo,
                                   const rls_ndprofile_soa *o)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        rls::NDProfile p;
        AtVector d; AiV3Create(d, dist.x[i], dist.y[i], dist.z[i]);
        AtColor a = rls_shim_rgb(albedo.x[i], albedo.y[i], albedo.z[i]);
        p.setDistance(d, a);
        store3(o->distance, i, p.mDistance.x, p.mDistance.y, p.mDistance.z);
        store3(o->C1, i, p.mC1.x, p.mC1.y, p.mC1.z);
        store3(o->C2, i, p.mC2.x, p.mC2.y, p.mC2.z);
        o->max_radius[i] = p.mMaxRadius;
    }
}

void oracle_ndprofile_get_radius(size_t n, const rls_ndprofile_soa *profile, const float *rx,
                                 float *out_r, uint32_t *out_flags)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        rls::NDProfile p; loadProfile(profile, i, p);
        out_r[i] = p.getRadius(rx[i]);
        if (out_flags) out_flags[i] = profileFlags(p, rx[i]);
    }
}

void oracle_ndprofile_get_pdf(size_t n, const rls_ndprofile_soa *profile, const float *r, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        rls::NDProfile p; loadProfile(profile, i, p);
        out_pdf[i] = p.getPdf(r[i]);
    }
}

void oracle_ndprofile_eval_profile(size_t n, const rls_ndprofile_soa *profile, const float *r, rls_vec3 out_rd)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        rls::NDProfile p; loadProfile(profile, i, p);
        AtRGB c = p.evalProfile(r[i]);
        store3(out_rd, i, c.r, c.g, c.b);
    }
}

/* GaussianProfile: the reference class itself (src/rlSss.h:63-97). */
static inline void loadGauss(const rls_gaussprofile_soa *s, size_t i, rls::GaussianProfile &p)
{
    p.mVariance = s->variance[i]; p.mMaxRadius = s->max_radius[i]; p.mNorm = s->norm[i];
}
void oracle_gaussprofile_set_distance(size_t n, rls_cvec3 dist, rls_cvec3 albedo, const rls_gaussprofile_soa *o)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        rls::GaussianProfile p;
        AtVector d; AiV3Create(d, dist.x[i], dist.y ? dist.y[i] : 0.0f, dist.z ? dist.z[i] : 0.0f);
        AtColor a = albedo.x ? rls_shim_rgb(albedo.x[i], albedo.y[i], albedo.z[i]) : rls_shim_rgb(1.0f, 1.0f, 1.0f);
        p.setDistance(d, a);
        o->variance[i] = p.mVariance; o->max_radius[i] = p.mMaxRadius; o->norm[i] = p.mNorm;
    }
}
void oracle_gaussprofile_get_radius(size_t n, const rls_gaussprofile_soa *profile, const float *rx, float *out_r)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { rls::GaussianProfile p; loadGauss(profile, i, p); out_r[i] = p.getRadius(rx[i]); }
}
void oracle_gaussprofile_get_pdf(size_t n, const rls_gaussprofile_soa *profile, const float *r, float *out_pdf)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { rls::GaussianProfile p; loadGauss(profile, i, p); out_pdf[i] = p.getPdf(r[i]); }
}
void oracle_gaussprofile_eval_profile(size_t n, const rls_gaussprofile_soa *profile, const float *r, float *out_rd)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { rls::GaussianProfile p; loadGauss(profile, i, p); out_rd[i] = p.evalProfile(r[i]); }
}
void oracle_gaussprofile_sample_eval_pdf(size_t n, const float *dist_x, const float *rx, float *out_r,
                                         float *out_pdf, float *out_rd)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        rls::GaussianProfile p;
        AtVector d; AiV3Create(d, dist_x[i], 0.0f, 0.0f);
        p.setDistance(d, rls_shim_rgb(1.0f, 1.0f, 1.0f));
        float r = p.getRadius(rx[i]);
        out_r[i] = r; out_pdf[i] = p.getPdf(r); out_rd[i] = p.evalProfile(r);
    }
}

void oracle_skin_profile_sample_eval_pdf(size_t n, const rls_skin_params *sp, const float *rx,
                                         const rls_profile_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float c[3], d[3];
        orc_p3(&sp->sss_color, i, c);
        orc_p3(&sp->sss_scatter_dist, i, d);
        float scale = orc_p1(&sp->sss_dist_multiplier, i);
        AtColor albedo = rls_shim_rgb(c[0], c[1], c[2]);
        /* src/rlSkin.cpp:236: AiShaderEvalParamVec(p_scatter_distance) * distanceScale */
        AtVector dist = rls_shim_v3(d[0], d[1], d[2]) * scale;
        rls::NDProfile p;
        p.setDistance(dist, albedo);
        float r = p.getRadius(rx[i]);
        out->r[i] = r;
        out->pdf[i] = p.getPdf(r);
        AtRGB rd = p.evalProfile(r);
        store3(out->Rd, i, rd.r, rd.g, rd.b);
        out->flags[i] = profileFlags(p, rx[i]);
    }
}

void oracle_skin_layer_weights(size_t n, const rls_skin_params *sp, const float *avg_f_sheen,
                               const float *avg_f_spec, float *out_spec_scale, float *out_sss_weight)
{
    /* src/rlSkin.cpp:174-238 lives inside shader_evaluate (needs Arnold's light loop), so
     * this six-flop hand-off is restated here rather than called. */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float sheenWeight = orc_p1(&sp->sheen_weight, i);
        float specularWeight = orc_p1(&sp->specular_weight, i);
        float sssWeight = orc_p1(&sp->sss_weight, i);
        float sheenFresnel = 0.0f, specularFresnel = 0.0f;
        if (sheenWeight > AI_EPSILON) sheenFresnel = avg_f_sheen[i] * sheenWeight;       /* :191,204 */
        if (specularWeight > AI_EPSILON) specularFresnel = avg_f_spec[i] * specularWeight; /* :214,228 */
        out_spec_scale[i] = specularWeight * (1.0f - sheenFresnel);                        /* :231 */
        sssWeight *= 1.0f - specularFresnel * (1.0f - sheenFresnel);                       /* :238 */
        out_sss_weight[i] = sssWeight;
    }
}

void oracle_skin_probe_ray(size_t n, const rls_shading_soa *sg, const rls_skin_params *sp,
                           const float *rx, const float *ry, const rls_probe_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        float c[3], d[3];
        orc_p3(&sp->sss_color, i, c);
        orc_p3(&sp->sss_scatter_dist, i, d);
        float scale = orc_p1(&sp->sss_dist_multiplier, i);
        AtVector dist = rls_shim_v3(d[0], d[1], d[2]) * scale;
        rls::SssSampler<rls::NDProfile> s(&sh.sg, rls_shim_rgb(c[0], c[1], c[2]), dist);
        AtRay ray; std::memset(&ray, 0, sizeof(ray));
        float r = s.getProbeRay(rx[i], ry[i], AI_V3_ZERO, ray);
        out->r[i] = r;
        store3(out->origin, i, ray.origin.x, ray.origin.y, ray.origin.z);
        store3(out->dir, i, ray.dir.x, ray.dir.y, ray.dir.z);
        out->maxdist[i] = (float)ray.maxdist;
        /* axis pick restated from src/rlSss.h:491-500 (the idx local is not observable) */
        float x = rx[i];
        uint32_t axis;
        if (x < 0.5f) { axis = 0; x = LINEARSTEP(0.0f, 0.5f, x); }
        else if (x < 0.75f) { axis = 2; x = LINEARSTEP(0.5f, 0.75f, x); }
        else { axis = 3; x = LINEARSTEP(0.75f, 1.0f, x); }
        out->flags[i] = profileFlags(s.mProfile, x) | (axis << RLS_FLAG_PROBE_AXIS_SHIFT);
    }
}

void oracle_skin_probe_mis_pdf(size_t n, const rls_shading_soa *sg, const rls_skin_params *sp,
                               rls_cvec3 disp, rls_cvec3 hit_normal, float *out_pdf)
{
    /* src/rlSss.h:246-266 sits inside integrateScatter (needs Arnold's probe tracer), so the
     * combine is restated around the reference's own NDProfile::getPdf and the shim's
     * AiM4Frame / AiM4VectorByMatrixMult. */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        float c[3], d[3];
        orc_p3(&sp->sss_color, i, c);
        orc_p3(&sp->sss_scatter_dist, i, d);
        float scale = orc_p1(&sp->sss_dist_multiplier, i);
        AtVector dist = rls_shim_v3(d[0], d[1], d[2]) * scale;
        rls::SssSampler<rls::NDProfile> s(&sh.sg, rls_shim_rgb(c[0], c[1], c[2]), dist);
        AtVector dp; AiV3Create(dp, disp.x[i], disp.y[i], disp.z[i]);
        AtVector hn; AiV3Create(hn, hit_normal.x[i], hit_normal.y[i], hit_normal.z[i]);
        AtVector offset;
        AiM4VectorByMatrixMult(&offset, s.mWorldToLocalMat, &dp);
        offset *= offset;
        float rr[3];
        rr[0] = sqrt(offset[1] + offset[2]);
        rr[1] = sqrt(offset[0] + offset[2]);
        rr[2] = sqrt(offset[0] + offset[1]);
        out_pdf[i] = s.mProfile.getPdf(rr[0]) * ABS(AiV3Dot(s.mAxisU, hn)) * 0.25f
                   + s.mProfile.getPdf(rr[1]) * ABS(AiV3Dot(s.mAxisV, hn)) * 0.25f
                   + s.mProfile.getPdf(rr[2]) * ABS(AiV3Dot(s.mAxisN, hn)) * 0.5f;
    }
}

/* ---- rlSkin's glossy layers.  src/rlSkin.cpp:184-238 lives inside
 * shader_evaluate, so the node's statements are restated around the reference's own GgxSampler
 * (its evalSample accumulates the Fresnel average exactly as in the plugin). */
static void skinLayerRef(Shading &sh, const float *c, float ior, float rough, uint32_t K, size_t P, size_t p,
                         const float *rx, const float *ry, rls_cvec3 li, float *avg, float *est)
{
    AtColor color = rls_shim_rgb(c[0], c[1], c[2]);
    rls::GgxSampler s(&sh.sg, color, ior, rough);                    /* src/rlSkin.cpp:192,215 */
    float acc[3] = { 0.0f, 0.0f, 0.0f };
    if (!AiColorIsSmall(color)) {                                    /* integrateGlossy, src/rlGgx.h:174-176 */
        for (uint32_t k = 0; k < K; k++) {
            size_t idx = (size_t)k * P + p;
            AtVector L = rls::GgxSampler::evalSample(&s, rx[idx], ry[idx]);
            AtColor f = rls::GgxSampler::evalBrdf(&s, &L);
            float pdf = rls::GgxSampler::evalPdf(&s, &L);
            float w[3] = { f.r / pdf, f.g / pdf, f.b / pdf };
            if (li.x) { w[0] *= li.x[idx]; w[1] *= li.y[idx]; w[2] *= li.z[idx]; }
            acc[0] += w[0]; acc[1] += w[1]; acc[2] += w[2];
        }
    }
    *avg = s.getAvgReflectWeight();
    float invK = 1.0f / (float)K;
    est[0] = acc[0] * invK; est[1] = acc[1] * invK; est[2] = acc[2] * invK;
}

void oracle_skin_glossy_layers(size_t n, uint32_t K, const rls_shading_soa *sg, const rls_skin_params *sp,
                               const float *rx_a, const float *ry_a, const float *rx_b, const float *ry_b,
                               rls_cvec3 li_a, rls_cvec3 li_b, const rls_skin_layers_out *out)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        uint32_t flags = 0;
        float sheenFresnel = 0.0f, specularFresnel = 0.0f;
        float sheen[3] = { 0, 0, 0 }, specular[3] = { 0, 0, 0 }, c[3], avg;
        float sheenWeight = orc_p1(&sp->sheen_weight, i);
        if (sheenWeight > AI_EPSILON) {
            orc_p3(&sp->sheen_color, i, c);
            skinLayerRef(sh, c, orc_p1(&sp->sheen_ior, i), orc_p1(&sp->sheen_roughness, i), K, n, i, rx_a, ry_a, li_a, &avg, sheen);
            sheenFresnel = avg * sheenWeight;
            flags |= RLS_SKIN_SHEEN_EVALUATED;
        }
        for (int j = 0; j < 3; j++) sheen[j] *= sheenWeight;
        float specularWeight = orc_p1(&sp->specular_weight, i);
        if (specularWeight > AI_EPSILON) {
            orc_p3(&sp->specular_color, i, c);
            skinLayerRef(sh, c, orc_p1(&sp->specular_ior, i), orc_p1(&sp->specular_roughness, i), K, n, i, rx_b, ry_b, li_b, &avg, specular);
            specularFresnel = avg * specularWeight;
            flags |= RLS_SKIN_SPECULAR_EVALUATED;
        }
        float scale = specularWeight * (1.0f - sheenFresnel);
        for (int j = 0; j < 3; j++) specular[j] *= scale;
        float sssWeight = orc_p1(&sp->sss_weight, i);
        sssWeight *= 1.0f - specularFresnel * (1.0f - sheenFresnel);
        if (sssWeight < AI_EPSILON) flags |= RLS_SKIN_SSS_SKIPPED;
        store3(out->sheen, i, sheen[0], sheen[1], sheen[2]);
        store3(out->specular, i, specular[0], specular[1], specular[2]);
        out->sheen_fresnel[i] = sheenFresnel;
        out->specular_fresnel[i] = specularFresnel;
        out->sss_weight[i] = sssWeight;
        out->flags[i] = flags;
    }
}

/* ---- 8(f) f3: one MIS light sample (two-sample power heuristic, include/rls_b200.h). */
extern "C++" {
namespace {
inline float powerHeuristic(float a, float b) { float a2 = a * a; return a2 / (a2 + b * b); }
struct MisOut { float rgb[3], wl, wb; };
inline MisOut misCombine(const AtVector &Ld, const float *Li, float pl, const AtColor &fl, float pbl, bool have,
                         const AtVector &L, const AtColor &fb, float pb, const float *Lib, float plb)
{
    MisOut o = { { 0.0f, 0.0f, 0.0f }, 0.0f, 0.0f };
    if (!(Ld == AI_V3_ZERO) && pl > 0.0f) {
        o.wl = powerHeuristic(pl, pbl);
        float s = o.wl / pl;
        o.rgb[0] = fl.r * Li[0] * s; o.rgb[1] = fl.g * Li[1] * s; o.rgb[2] = fl.b * Li[2] * s;
    }
    if (have && !(L == AI_V3_ZERO) && pb > 0.0f) {
        o.wb = powerHeuristic(pb, plb);
        float s = o.wb / pb;
        o.rgb[0] = o.rgb[0] + fb.r * Lib[0] * s; o.rgb[1] = o.rgb[1] + fb.g * Lib[1] * s; o.rgb[2] = o.rgb[2] + fb.b * Lib[2] * s;
    }
    return o;
}
template <typename Brdf, typename SampleFn>
inline void lightSampleOne(Brdf &brdf, SampleFn sample, size_t i, const rls_light_sample *light, const rls_light_sample *at_l,
                           rls_vec3 out_rgb, float *wl, float *wb)
{
    AtVector Ld; AiV3Create(Ld, light->dir.x[i], light->dir.y[i], light->dir.z[i]);
    float Li[3] = { light->radiance.x[i], light->radiance.y[i], light->radiance.z[i] };
    AtColor fl = Brdf::evalBrdf(&brdf, &Ld);
    float pbl = Brdf::evalPdf(&brdf, &Ld);
    AtVector L = AI_V3_ZERO; AtColor fb = AI_RGB_BLACK; float pb = 0.0f, plb = 0.0f, Lib[3] = { 0, 0, 0 };
    if (at_l) {
        L = sample();
        fb = Brdf::evalBrdf(&brdf, &L);
        pb = Brdf::evalPdf(&brdf, &L);
        Lib[0] = at_l->radiance.x[i]; Lib[1] = at_l->radiance.y[i]; Lib[2] = at_l->radiance.z[i];
        plb = at_l->pdf[i];
    }
    MisOut m = misCombine(Ld, Li, light->pdf[i], fl, pbl, at_l != nullptr, L, fb, pb, Lib, plb);
    store3(out_rgb, i, m.rgb[0], m.rgb[1], m.rgb[2]);
    if (wl) wl[i] = m.wl;
    if (wb) wb[i] = m.wb;
}
template <typename Sampler>
void ggxLightSampleT(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, const rls_light_sample *light,
                     const float *rx, const float *ry, const rls_light_sample *at_l, rls_vec3 out_rgb, float *wl, float *wb)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        GgxArgs a = ggxArgs(p, i);
        Sampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso);
        lightSampleOne(s, [&]() { return Sampler::evalSample(&s, rx[i], ry[i]); }, i, light, at_l, out_rgb, wl, wb);
    }
}
} // namespace
} // extern "C++"

void oracle_ggx_evaluate_light_sample(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                      const rls_light_sample *light, const float *rx, const float *ry,
                                      const rls_light_sample *at_l, rls_vec3 out_rgb, float *wl, float *wb)
{
    GGX_DISPATCH(p, ggxLightSampleT, n, sg, p, light, rx, ry, at_l, out_rgb, wl, wb);
}

void oracle_disney_evaluate_light_sample(size_t n, const rls_shading_soa *sg, const rls_disney_params *p, int sample_type,
                                         const rls_light_sample *light, const float *rx, const float *ry,
                                         const rls_light_sample *at_l, rls_vec3 out_rgb, float *wl, float *wb)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Shading sh; loadShading(sg, i, sh);
        float table[64 * 3];
        disneyTable(p, i, table);
        rls_shim_set_param_table(table);
        DisneySampler s(nullptr, &sh.sg);
        s.mSampleFromVisibleNormal = p->sample_from_visible_normal != 0;
        s.setSampleType((AtUInt16)sample_type);
        lightSampleOne(s, [&]() { return DisneySampler::evalSample(&s, rx[i], ry[i]); }, i, light, at_l, out_rgb, wl, wb);
    }
}

/* ---- 8(f) f4: SampleWriter (src/rlUtil.h:44-171).  The class itself needs tinyexr and an
 * Arnold sampler; its two loops are restated around the reference's own triple. */
extern "C++" {
namespace {
inline void writerPixel(float *image, int W, int H, int x, int y, const AtColor &c)     /* writePixel :158-163 */
{
    size_t stride = (size_t)W * H, at = (size_t)x + (size_t)y * W;
    image[at] = c.b; image[at + stride] = c.g; image[at + stride * 2] = c.r;
}
template <typename Brdf>
void writerRadiance(Brdf &brdf, int W, int H, float *image)                                /* :98-114 */
{
    for (int j = 0; j < H; j++) {
        float theta = AI_PIOVER2 * j / H;
        for (int i = 0; i < W; i++) {
            float phi = AI_PITIMES2 * i / W;
            AtVector dir = rls::sphericalDirection(cosf(theta), phi);
            writerPixel(image, W, H, i, j, Brdf::evalBrdf(&brdf, &dir));
        }
    }
}
template <typename Brdf>
void writerScatter(Brdf &brdf, size_t n, const float *rx, const float *ry, int W, int H, float *image, uint32_t *missing)   /* :116-156 */
{
    uint32_t missingCount = 0;
    for (size_t k = 0; k < n; k++) {
        AtVector dir = Brdf::evalSample(&brdf, rx[k], ry[k]);
        if (AiV3IsZero(dir)) continue;
        float theta = acosf(dir.z);
        float phi = atan2f(dir.y, dir.x);
        if (phi < 0.0f) phi += AI_PITIMES2;
        int i = CLAMP(static_cast<int>(phi * AI_ONEOVER2PI * W), 0, W - 1);
        int j = CLAMP(static_cast<int>(theta / AI_PIOVER2 * H), 0, H - 1);
        if (theta > AI_PIOVER2) { writerPixel(image, W, H, i, j, AI_RGB_RED); missingCount++; }
        else writerPixel(image, W, H, i, j, AI_RGB_GREEN);
    }
    if (missing) *missing = missingCount;
}
} // namespace
} // extern "C++"

void oracle_sample_writer_radiance(int node, const rls_shading_soa *sg, const void *params, size_t point, int sample_type,
                                   int W, int H, float *image)
{
    Shading sh; loadShading(sg, point, sh);
    if (node == RLS_NODE_GGX) {
        const rls_ggx_params *p = (const rls_ggx_params *)params;
        GgxArgs a = ggxArgs(p, point);
        if (p->normal_sampler == RLS_GGX_SAMPLER_NDF) { GgxNdfSampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso); writerRadiance(s, W, H, image); }
        else { rls::GgxSampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso); writerRadiance(s, W, H, image); }
    } else {
        const rls_disney_params *p = (const rls_disney_params *)params;
        float table[64 * 3];
        disneyTable(p, point, table);
        rls_shim_set_param_table(table);
        DisneySampler s(nullptr, &sh.sg);
        s.mSampleFromVisibleNormal = p->sample_from_visible_normal != 0;
        s.setSampleType((AtUInt16)sample_type);
        writerRadiance(s, W, H, image);
    }
}

void oracle_sample_writer_scatter(int node, const rls_shading_soa *sg, const void *params, size_t point, int sample_type,
                                  size_t n, const float *rx, const float *ry, int W, int H, float *image, uint32_t *missing)
{
    Shading sh; loadShading(sg, point, sh);
    if (node == RLS_NODE_GGX) {
        const rls_ggx_params *p = (const rls_ggx_params *)params;
        GgxArgs a = ggxArgs(p, point);
        if (p->normal_sampler == RLS_GGX_SAMPLER_NDF) { GgxNdfSampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso); writerScatter(s, n, rx, ry, W, H, image, missing); }
        else { rls::GgxSampler s(&sh.sg, a.ks, a.ior, a.rough, a.aniso); writerScatter(s, n, rx, ry, W, H, image, missing); }
    } else {
        const rls_disney_params *p = (const rls_disney_params *)params;
        float table[64 * 3];
        disneyTable(p, point, table);
        rls_shim_set_param_table(table);
        DisneySampler s(nullptr, &sh.sg);
        s.mSampleFromVisibleNormal = p->sample_from_visible_normal != 0;
        s.setSampleType((AtUInt16)sample_type);
        writerScatter(s, n, rx, ry, W, H, image, missing);
    }
}

void oracle_albedo_sweep(const rls_sweep_grid *g, uint64_t seed, uint32_t spp_begin,
                         uint32_t spp_end, double *table)
{
    const uint32_t cells = (uint32_t)(g->n_rough * g->n_cos * g->n_ior);
#pragma omp parallel for schedule(dynamic, 16)
    for (uint32_t cell = 0; cell < cells; cell++) {
        float rough, cosv, ior;
        orc_sweep_cell(g, cell, &rough, &cosv, &ior);
        Shading sh;
        std::memset(&sh.sg, 0, sizeof(sh.sg));
        AiV3Create(sh.U, 1.0f, 0.0f, 0.0f);
        AiV3Create(sh.V, 0.0f, 1.0f, 0.0f);
        AiV3Create(sh.sg.Nf, 0.0f, 0.0f, 1.0f);
        sh.sg.N = sh.sg.Nf;
        AtVector wo;
        AiV3Create(wo, sqrtf(1.0f - cosv * cosv), 0.0f, cosv);
        sh.sg.Rd = -wo;
        rls_shim_set_frame(&sh.U, &sh.V);
        double acc[RLS_SWEEP_VALUES_PER_CELL] = { 0, 0, 0, 0, 0 };
        for (uint32_t k = spp_begin; k < spp_end; k++) {
            uint64_t idx = ((uint64_t)cell << 32) | (uint64_t)k;
            float rx = orc_uniform(seed, 0u, idx);
            float ry = orc_uniform(seed, 1u, idx);
            DielectricResult r = dielectricUnit(sh, ior, rough, 0.0f, rx, ry);
            bool valid = !(r.flags & (RLS_FLAG_ZERO_L | RLS_FLAG_BELOW_HORIZON));
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

} // extern "C"
