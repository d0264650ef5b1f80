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
        out->flags[i] = fl;
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
