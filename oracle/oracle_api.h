/*
 * oracle/oracle_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The CPU checker's entry points.  Two shared libraries export exactly this surface:
 *
 *   oracle/_ref/librls_ref.so   kind "reference": the reference's own C++ sources from
 *                               /root/reference/src compiled UNMODIFIED against
 *                               oracle/shim/ai.h, driven by oracle/ref_driver.cpp.
 *   oracle/librls_oracle.so     kind "port": a plain-C restatement (oracle/rls_oracle.c)
 *                               that follows the reference line by line; pinned
 *                               bit-for-bit against the "reference" library and against
 *                               the golden vectors in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load these libraries.  The product (rlshaders_b200/) never does.
 *
 * Argument structs are the product's own ABI structs (include/rls_b200.h) with HOST
 * pointers, so a parity test feeds identical descriptors to both sides.
 */
#ifndef RLS_ORACLE_API_H
#define RLS_ORACLE_API_H

#include "../include/rls_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *oracle_kind(void);             /* "reference" or "port"            */
int  oracle_max_threads(void);             /* OpenMP threads available          */
void oracle_set_threads(int n);            /* 0 = all                           */
/* RLS_FLAG_SLOPE_EARLY_OUT costs the reference library a partial re-evaluation of the sampler (oracle_common.h
 * orc_vndf_early_out): 0 switches that probe off for timed runs (the bit is then absent), 1 (default) on. */
void oracle_set_flag_probe(int on);

void oracle_ggx_eval_sample(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                            const float *rx, const float *ry, rls_vec3 out_wi, float *out_fresnel);
void oracle_ggx_eval_brdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                          rls_cvec3 wi, rls_vec3 out_f);
void oracle_ggx_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                         rls_cvec3 wi, float *out_pdf);
void oracle_ggx_refract_direction(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                  rls_cvec3 m, rls_vec3 out_wi, uint32_t *out_flags);
void oracle_ggx_eval_btdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p, rls_cvec3 wi, float *out_ft);
void oracle_ggx_sample_weight(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                              rls_cvec3 wi, rls_cvec3 m, float *out_weight);
void oracle_ggx_sample_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                const float *rx, const float *ry, const rls_bsdf_out *out);
void oracle_ggx_dielectric_sample_eval_pdf(size_t n, const rls_shading_soa *sg,
                                           const rls_ggx_params *p, const float *rx,
                                           const float *ry, const rls_ggx_dielectric_out *out);

void oracle_disney_eval_sample(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                               int sample_type, const float *rx, const float *ry,
                               rls_vec3 out_wi, uint32_t *out_flags);
void oracle_disney_eval_brdf(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                             int sample_type, rls_cvec3 wi, rls_vec3 out_f);
void oracle_disney_eval_pdf(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                            int sample_type, rls_cvec3 wi, float *out_pdf);
void oracle_disney_sample_eval_pdf(size_t n, const rls_shading_soa *sg,
                                   const rls_disney_params *p, const float *rx_s,
                                   const float *ry_s, const float *rx_d, const float *ry_d,
                                   const rls_disney_out *out);

void oracle_ndprofile_set_distance(size_t n, rls_cvec3 dist, rls_cvec3 albedo,
                                   const rls_ndprofile_soa *out_profile);
void oracle_ndprofile_get_radius(size_t n, const rls_ndprofile_soa *profile, const float *rx,
                                 float *out_r, uint32_t *out_flags);
void oracle_ndprofile_get_pdf(size_t n, const rls_ndprofile_soa *profile, const float *r,
                              float *out_pdf);
void oracle_ndprofile_eval_profile(size_t n, const rls_ndprofile_soa *profile, const float *r,
                                   rls_vec3 out_rd);
/* GaussianProfile (src/rlSss.h:63-97) */
void oracle_gaussprofile_set_distance(size_t n, rls_cvec3 dist, rls_cvec3 albedo,
                                      const rls_gaussprofile_soa *out_profile);
void oracle_gaussprofile_get_radius(size_t n, const rls_gaussprofile_soa *profile, const float *rx, float *out_r);
void oracle_gaussprofile_get_pdf(size_t n, const rls_gaussprofile_soa *profile, const float *r, float *out_pdf);
void oracle_gaussprofile_eval_profile(size_t n, const rls_gaussprofile_soa *profile, const float *r, float *out_rd);
void oracle_gaussprofile_sample_eval_pdf(size_t n, const float *dist_x, const float *rx, float *out_r,
                                         float *out_pdf, float *out_rd);
void oracle_skin_profile_sample_eval_pdf(size_t n, const rls_skin_params *p, const float *rx,
                                         const rls_profile_out *out);
void oracle_skin_layer_weights(size_t n, const rls_skin_params *p, const float *avg_f_sheen,
                               const float *avg_f_spec, float *out_spec_scale, float *out_sss_weight);
void oracle_skin_probe_ray(size_t n, const rls_shading_soa *sg, const rls_skin_params *p,
                           const float *rx, const float *ry, const rls_probe_out *out);

void oracle_skin_probe_mis_pdf(size_t n, const rls_shading_soa *sg, const rls_skin_params *p,
                               rls_cvec3 disp, rls_cvec3 hit_normal, float *out_pdf);

/* the callers of the triple; definitions in include/rls_b200.h. */
void oracle_skin_glossy_layers(size_t n_points, uint32_t k, const rls_shading_soa *sg,
                               const rls_skin_params *p, const float *rx_sheen, const float *ry_sheen,
                               const float *rx_specular, const float *ry_specular,
                               rls_cvec3 li_sheen, rls_cvec3 li_specular, const rls_skin_layers_out *out);
void oracle_ggx_evaluate_light_sample(size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                      const rls_light_sample *light, const float *rx, const float *ry,
                                      const rls_light_sample *light_at_l, rls_vec3 out_rgb,
                                      float *out_w_light, float *out_w_brdf);
void oracle_disney_evaluate_light_sample(size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                                         int sample_type, const rls_light_sample *light, const float *rx,
                                         const float *ry, const rls_light_sample *light_at_l,
                                         rls_vec3 out_rgb, float *out_w_light, float *out_w_brdf);
void oracle_sample_writer_radiance(int node, const rls_shading_soa *sg, const void *params, size_t point,
                                   int sample_type, int width, int height, float *image);
void oracle_sample_writer_scatter(int node, const rls_shading_soa *sg, const void *params, size_t point,
                                  int sample_type, size_t n_samples, const float *rx, const float *ry,
                                  int width, int height, float *image, uint32_t *out_missing);

void oracle_albedo_sweep(const rls_sweep_grid *grid, uint64_t seed, uint32_t spp_begin,
                         uint32_t spp_end, double *table);

/* The integer hash of the synthetic generators, restated on the host. */
void oracle_synth_uniform(size_t n, uint64_t seed, uint32_t stream, uint64_t first_index,
                          float lo, float hi, float *out);

#ifdef __cplusplus
}
#endif
#endif
