// tests/native/tol_host.cpp -- TEST INFRASTRUCTURE.  Runs the tolerance-policy units of
// rlshaders_b200/csrc/rls_tol.cuh (the header is __host__ __device__) on the CPU, sample by sample, behind the ABI
// structs of include/rls_b200.h, so that the bands and the error percentiles of RLS_ARITH_TOLERANT can be checked
// against the oracle without a GPU (tests/test_tol_host.py).  Built by tests/native/build_tol_host.sh with
// -mfma -ffp-contract=fast (the device build contracts too); -DRLS_TOL_EMULATE_ULP moves every emulated MUFU result by a
// pseudo-random -1/0/+1 ulp.  `rerun[i]` = 1 where the band tracker sends the sample to the bit-exact re-run.
#include <stddef.h>
#include <stdint.h>
#include "../../include/rls_b200.h"
#include "../../rlshaders_b200/csrc/rls_tol.cuh"

using namespace rls::tol;

static inline float p1(const rls_param1 &p, size_t i) { return p.array ? p.array[i] : p.value; }
static inline v3 p3(const rls_param3 &p, size_t i)
{
    return mk(p.array.x ? p.array.x[i] : p.value[0], p.array.y ? p.array.y[i] : p.value[1], p.array.z ? p.array.z[i] : p.value[2]);
}
static inline v3 ld(const rls_cvec3 &v, size_t i) { return mk(v.x[i], v.y[i], v.z[i]); }
static inline void st(const rls_vec3 &v, size_t i, v3 a) { v.x[i] = a.x; v.y[i] = a.y; v.z[i] = a.z; }

extern "C" {

void tol_ggx_dielectric(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, const float *rx, const float *ry,
                        const rls_ggx_dielectric_out *o, uint8_t *rerun)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Bands bd;
        const DielectricT r = dielectric_unit(bd, ld(sg->U, i), ld(sg->V, i), ld(sg->N, i), ld(sg->wo, i),
                                              sg->backfacing ? sg->backfacing[i] != 0 : false, p1(p->ior, i),
                                              p1(p->specularRoughness, i), p1(p->anisotropic, i), rx[i], ry[i]);
        o->fresnel[i] = r.F; st(o->wi_r, i, r.wi_r); o->f_r[i] = r.f_r; o->pdf_r[i] = r.pdf_r;
        st(o->wi_t, i, r.wi_t); o->f_t[i] = r.f_t; o->weight_t[i] = r.w_t; o->flags[i] = r.flags;
        rerun[i] = bd.rerun ? 1 : 0;
    }
}

void tol_ggx_sample_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, const float *rx, const float *ry,
                             const rls_bsdf_out *o, uint8_t *rerun)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Bands bd;
        const GgxBsdfT r = ggx_unit(bd, ld(sg->U, i), ld(sg->V, i), ld(sg->N, i), ld(sg->wo, i),
                                    sg->backfacing ? sg->backfacing[i] != 0 : false, p3(p->KsColor, i), p1(p->ior, i),
                                    p1(p->specularRoughness, i), p1(p->anisotropic, i), rx[i], ry[i]);
        st(o->wi, i, r.L); st(o->f, i, r.f); o->pdf[i] = r.pdf; if (o->fresnel) o->fresnel[i] = r.fresnel; o->flags[i] = r.flags;
        rerun[i] = bd.rerun ? 1 : 0;
    }
}

void tol_disney(size_t n, const rls_shading_soa *sg, const rls_disney_params *p, const float *rx_s, const float *ry_s,
                const float *rx_d, const float *ry_d, const rls_disney_out *o, uint8_t *rerun)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Bands bd;
        DisneyIn in;
        in.base = p3(p->base_color, i); in.subsurface = p1(p->subsurface, i); in.metallic = p1(p->metallic, i);
        in.specular = p1(p->specular, i); in.specular_tint = p1(p->specular_tint, i); in.roughness = p1(p->roughness, i);
        in.anisotropic = p1(p->anisotropic, i); in.sheen = p1(p->sheen, i); in.sheen_tint = p1(p->sheen_tint, i);
        in.clearcoat = p1(p->clearcoat, i); in.clearcoat_gloss = p1(p->clearcoat_gloss, i);
        const DisneyT r = disney_unit(bd, ld(sg->U, i), ld(sg->V, i), ld(sg->N, i), ld(sg->wo, i), in,
                                      p->sample_from_visible_normal != 0, rx_s[i], ry_s[i], rx_d[i], ry_d[i]);
        st(o->wi_s, i, r.Ls); st(o->f_s, i, r.fs); o->pdf_s[i] = r.ps;
        st(o->wi_d, i, r.Ld); st(o->f_d, i, r.fd); o->pdf_d[i] = r.pd;
        o->flags[i] = r.flags;
        rerun[i] = bd.rerun ? 1 : 0;
    }
}

void tol_skin_profile(size_t n, const rls_skin_params *p, const float *rx, const rls_profile_out *o, uint8_t *rerun)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        Bands bd;
        const float mult = p1(p->sss_dist_multiplier, i);
        const v3 d = p3(p->sss_scatter_dist, i);
        const ProfileT r = skin_profile_unit(bd, mk(mul_rn(d.x, mult), mul_rn(d.y, mult), mul_rn(d.z, mult)), rx[i]);
        o->r[i] = r.r; o->pdf[i] = r.pdf; st(o->Rd, i, r.Rd); o->flags[i] = r.flags;
        rerun[i] = bd.rerun ? 1 : 0;
    }
}

} // extern "C"
