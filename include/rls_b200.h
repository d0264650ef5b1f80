/*
 * rls_b200.h -- C ABI of the B200-native rlShaders BSDF hot path.
 *
 * This is the drop-in boundary for the per-shading-sample path of shihchinw/rlShaders:
 * the Arnold BRDF-callback triple
 *
 *     static AtVector evalSample(const void *brdfData, float rx, float ry);
 *     static AtColor  evalBrdf  (const void *brdfData, const AtVector *indir);
 *     static float    evalPdf   (const void *brdfData, const AtVector *indir);
 *
 * (reference src/rlGgx.h:97,110,121 and src/rlDisney.cpp:109,120,139), the sampler
 * constructors that build `brdfData` from AtShaderGlobals + node parameters
 * (src/rlGgx.h:130-156, src/rlDisney.cpp:155-192) and the diffusion-profile surface
 * NDProfile::{setDistance,getRadius,getPdf,evalProfile} (src/rlSss.h:49-55) -- all
 * restated as BATCHED, structure-of-arrays entry points.  One array element = one
 * shading sample; the per-shading-point `brdfData` object is replaced by the SoA
 * shading inputs (`rls_shading_soa`) plus a parameter block whose field names are the
 * node parameter names of src/rlShaders.mtd / node_parameters verbatim.
 *
 * Conventions
 *   - Plain C: pointers, sizes, PODs.  No torch / CUDA types in any signature
 *     (the stream is passed as an opaque void* = cudaStream_t).
 *   - Every array pointer is a DEVICE pointer unless the entry point ends in `_host`,
 *     in which case every array pointer is a (preferably pinned) HOST pointer and the
 *     library stages chunks through its own device scratch (H2D, kernel, D2H overlapped).
 *   - The caller owns all buffers.  Device entry points do not allocate and do not
 *     synchronise: work is enqueued on the context's stream.
 *   - Return value: RLS_OK (0) or a negative RLS_ERR_* code; rls_last_error_string()
 *     gives the detail.  There is no CPU fallback: without a usable sm_100 device
 *     rls_init fails.
 *   - Error semantics of the reference are preserved in-band: an invalid sample is the
 *     zero vector (src/rlDisney.cpp:385-387), a zero `indir` evaluates to black
 *     (src/rlGgx.h:112-115, src/rlDisney.cpp:124-127) and to pdf 0 in rlDisney
 *     (src/rlDisney.cpp:141-144) but NOT in rlGgx (src/rlGgx.h:121-127).
 */
#ifndef RLS_B200_H
#define RLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLS_B200_ABI_VERSION 1

/* ------------------------------------------------------------------ errors */
#define RLS_OK                    0
#define RLS_ERR_INVALID_ARGUMENT (-1)
#define RLS_ERR_CUDA             (-2)
#define RLS_ERR_NO_DEVICE        (-3)
#define RLS_ERR_OUT_OF_MEMORY    (-4)
#define RLS_ERR_NCCL             (-5)   /* rls_multi_*: libnccl.so.2 missing or an NCCL call failed */

/* Sample type of the rlDisney triple; values mirror AI_RAY_DIFFUSE / AI_RAY_GLOSSY as
 * passed to DisneySampler::setSampleType (src/rlDisney.cpp:194,242,268,275,281). */
#define RLS_RAY_DIFFUSE 0x20
#define RLS_RAY_GLOSSY  0x40

/* ------------------------------------------------- per-sample flags (u32) */
/* Bit-exact parity target.  For the fused rlDisney op the low half-word describes the
 * specular (glossy) triple and the high half-word the diffuse triple (same layout,
 * shifted by RLS_FLAG_DIFFUSE_SHIFT). */
#define RLS_FLAG_ZERO_L         0x0001u /* sampled direction is the zero vector (invalid sample)     */
#define RLS_FLAG_BELOW_HORIZON  0x0002u /* dot(L, N) <= 0 for the sampled direction                  */
#define RLS_FLAG_PDF_ZERO       0x0004u /* evalPdf returned exactly 0                                */
#define RLS_FLAG_F_BLACK        0x0008u /* evalBrdf returned exactly (0,0,0)                         */
#define RLS_FLAG_ENTERING       0x0010u /* dot(sg->N, sg->Rd) < 1e-4, src/rlGgx.h:137                */
#define RLS_FLAG_TIR            0x0020u /* refraction failed, sample reflected, src/rlGgx.h:232-236  */
#define RLS_FLAG_PDF_FLOORED    0x0040u /* pdf floor 1e-4 was taken, src/rlGgx.h:79, rlDisney.cpp:517 */
#define RLS_FLAG_SLOPE_EARLY_OUT 0x0080u /* visible-normal sampling took the uniform-slope early-out: stretched view
                                           along the normal (theta left 0) or |A^2 - 1| < 1e-4, src/rlGgx.cpp:27,38,82
                                           (rlDisney.cpp:416-463,485): fused units and rls_disney_eval_sample        */
#define RLS_FLAG_LOBE_SHIFT     8       /* bits 8-9: rlDisney 0 = GTR2, 1 = GTR1 (rlDisney.cpp:375); */
#define RLS_FLAG_LOBE_MASK      0x0300u /*           profile: colour channel 0..2 (rlSss.h:30-42)    */
#define RLS_FLAG_EXP_LOBE       0x0400u /* profile: second exponential (3d) chosen, rlSss.cpp:57     */
#define RLS_FLAG_DEGENERATE     0x0800u /* profile: maxRadius or d below 1e-4, rlSss.cpp:38,46       */
#define RLS_FLAG_PROBE_AXIS_SHIFT 12    /* bits 12-13: probe axis 0 = N, 2 = U, 3 = V (rlSss.h:491-500) */
#define RLS_FLAG_PROBE_AXIS_MASK  0x3000u
#define RLS_FLAG_DIFFUSE_SHIFT  16

/* ------------------------------------------------------------ basic types */
typedef struct rls_context rls_context;

typedef struct rls_cvec3 { const float *x, *y, *z; } rls_cvec3;   /* SoA input vector / colour  */
typedef struct rls_vec3  { float *x, *y, *z; } rls_vec3;          /* SoA output vector / colour */

/* A node parameter: uniform `value`, or one value per sample when `array` is non-NULL
 * (Arnold evaluates linked parameters per shading point: AiShaderEvalParam*). */
typedef struct rls_param1 { float value;    const float *array; } rls_param1;
typedef struct rls_param3 { float value[3]; rls_cvec3    array; } rls_param3;

/* The AtShaderGlobals fields the path reads, one entry per sample.
 *   U,V,N : the orthonormal shading frame mBasis / mAxisU,V,N with N = sg->Nf
 *           (src/rlGgx.h:145-146, src/rlDisney.cpp:173-174).  Supplied explicitly
 *           because AiBuildLocalFramePolar is proprietary.
 *   wo    : view direction mViewDir = -sg->Rd (src/rlGgx.h:144, src/rlDisney.cpp:172).
 *   backfacing : optional (may be NULL = all 0).  1 where sg->N == -sg->Nf, which makes
 *           the rlGgx constructor swap the IORs (src/rlGgx.h:137-142). */
typedef struct rls_shading_soa {
    rls_cvec3      U, V, N;
    rls_cvec3      wo;
    const uint8_t *backfacing;
} rls_shading_soa;

/* rlGgx node parameters (src/rlGgx.cpp:172-186, src/rlShaders.mtd:7-29).  The sampler
 * path reads KsColor, specularRoughness, ior, anisotropic; the others are accepted so a
 * parameter block can be filled straight from the node and are ignored here (they scale
 * results outside the BRDF triple, src/rlGgx.cpp:297-305). */
typedef struct rls_ggx_params {
    rls_param3 KsColor;
    rls_param1 Ks;
    rls_param1 specularRoughness;
    rls_param1 ior;
    rls_param1 anisotropic;
    rls_param3 KdColor;
    rls_param1 Kd;
    rls_param1 diffuseRoughness;
    rls_param3 KtColor;
    rls_param1 Kt;
    rls_param1 opacity;
    rls_param3 opacity_color;
    /* Not a node parameter: which microfacet-normal sampler the GgxSamplerT template is
     * instantiated with.  0 = VNDFKernel (the shipped `using GgxSampler`, src/rlGgx.h:375),
     * 1 = NDFKernel (src/rlGgx.h:24-56: Burley Eq.14 sampling, pdf D|m.n|/(4|i.m|), no floor). */
    int32_t    normal_sampler;
} rls_ggx_params;
#define RLS_GGX_SAMPLER_VNDF 0
#define RLS_GGX_SAMPLER_NDF  1

/* rlDisney node parameters (src/rlDisney.cpp:606-625). */
typedef struct rls_disney_params {
    rls_param3 base_color;
    rls_param1 subsurface;
    rls_param1 metallic;
    rls_param1 specular;
    rls_param1 specular_tint;
    rls_param1 roughness;
    rls_param1 anisotropic;
    rls_param1 sheen;
    rls_param1 sheen_tint;
    rls_param1 clearcoat;
    rls_param1 clearcoat_gloss;
    rls_param3 opacity;               /* accepted, ignored */
    rls_param1 indirectDiffuseScale;  /* accepted, ignored */
    rls_param1 indirectSpecularScale; /* accepted, ignored */
    /* Not a node parameter: DisneySampler::mSampleFromVisibleNormal (src/rlDisney.cpp:191 sets it
     * true).  0 selects sampleGTR2AnisoDirection (:406-414) and the pdf branch of :541-542. */
    int32_t    sample_from_visible_normal;
} rls_disney_params;

/* rlSkin node parameters (src/rlSkin.cpp:109-131, src/rlShaders.mtd:43-64).  The
 * profile path reads sss_color, sss_dist_multiplier, sss_scatter_dist; the layer-weight
 * op reads sss_weight, specular_weight, sheen_weight. */
typedef struct rls_skin_params {
    rls_param3 sss_color;
    rls_param1 sss_weight;
    rls_param1 sss_dist_multiplier;
    rls_param3 sss_scatter_dist;
    int32_t    sss_cavity_fadeout;
    rls_param3 specular_color;
    rls_param1 specular_weight;
    rls_param1 specular_roughness;
    rls_param1 specular_ior;
    rls_param3 sheen_color;
    rls_param1 sheen_weight;
    rls_param1 sheen_roughness;
    rls_param1 sheen_ior;
    rls_param1 opacity;
    rls_param3 opacity_color;
} rls_skin_params;

/* Result of one fused {construct, evalSample, evalBrdf, evalPdf} unit. */
typedef struct rls_bsdf_out {
    rls_vec3  wi;       /* L = evalSample(rx, ry); zero vector = invalid sample             */
    rls_vec3  f;        /* evalBrdf(L): BRDF x cosine, RGB                                   */
    float    *pdf;      /* evalPdf(L)                                                        */
    float    *fresnel;  /* optional (NULL): the fresnel(L, M) term evalSample accumulates
                           for getAvgReflectWeight, src/rlGgx.h:103,181-184                  */
    uint32_t *flags;    /* RLS_FLAG_*                                                        */
} rls_bsdf_out;

/* Result of the rough-dielectric unit (Walter'07): one visible-normal sample m, then
 * BOTH branches, as in integrateRefract's loop body (src/rlGgx.h:228-243). */
typedef struct rls_ggx_dielectric_out {
    float    *fresnel;   /* F = fresnel(wi_r, m), src/rlGgx.h:103,249-270                     */
    rls_vec3  wi_r;      /* reflectDirection(wo, m), src/rlUtil.h:31-34                       */
    float    *f_r;       /* reflection(wo, wi_r, N) * dot(wi_r, N)  (evalBrdf with KsColor 1) */
    float    *pdf_r;     /* evalPdf(wi_r), src/rlGgx.h:121-127                                */
    rls_vec3  wi_t;      /* getRefractDirection(m, wo), src/rlGgx.h:277-291; on TIR the
                            reflected direction (src/rlGgx.h:232-236)                         */
    float    *f_t;       /* refraction(wo, wi_t, N), src/rlGgx.h:316-328; 0 on TIR            */
    float    *weight_t;  /* getSampleWeight(wo, wi_t, m), src/rlGgx.h:294-301                 */
    uint32_t *flags;
} rls_ggx_dielectric_out;

/* Result of the fused rlDisney unit: specular (glossy) triple + diffuse triple. */
typedef struct rls_disney_out {
    rls_vec3  wi_s, f_s;  float *pdf_s;
    rls_vec3  wi_d, f_d;  float *pdf_d;
    uint32_t *flags;
} rls_disney_out;

/* NDProfile state produced by setDistance (src/rlSss.cpp:20-34), one per sample. */
typedef struct rls_ndprofile_soa {
    rls_vec3 distance;    /* mDistance        */
    rls_vec3 C1, C2;      /* mC1, mC2         */
    float   *max_radius;  /* mMaxRadius       */
} rls_ndprofile_soa;

/* GaussianProfile state produced by its setDistance (src/rlSss.h:71-76), one per sample.  The
 * alternative `Profile` argument of SssSampler<Profile> (src/rlSss.h:63-97); rlSkin instantiates
 * SssSampler<NDProfile> (src/rlSkin.cpp:236), this is the variant. */
typedef struct rls_gaussprofile_soa {
    float *variance;      /* mVariance  = maxRadius^2 / 12.46          */
    float *max_radius;    /* mMaxRadius = dist.x                       */
    float *norm;          /* mNorm      = 1 - exp(-maxRadius^2 / (2 v)) */
} rls_gaussprofile_soa;

/* Result of the fused skin-profile unit: setDistance + getRadius + getPdf + evalProfile. */
typedef struct rls_profile_out {
    float    *r;      /* getRadius(rx), src/rlSss.cpp:36-66    */
    float    *pdf;    /* getPdf(r), src/rlSss.cpp:68-84         */
    rls_vec3  Rd;     /* evalProfile(r), src/rlSss.cpp:86-106   */
    uint32_t *flags;
} rls_profile_out;

/* Probe-ray geometry of SssSampler::getProbeRay (src/rlSss.h:487-533). */
typedef struct rls_probe_out {
    float    *r;         /* sampled radius                        */
    rls_vec3  origin;    /* ray.origin - sg->P (offset only)      */
    rls_vec3  dir;       /* ray.dir                               */
    float    *maxdist;   /* ray.maxdist = 2 sqrt(rmax^2 - r^2)    */
    uint32_t *flags;     /* channel, exp lobe, probe axis         */
} rls_probe_out;

/* ------------------------------------------------------------ life cycle */
/* Creates a context on CUDA device `device`.  `stream` is an existing cudaStream_t to
 * enqueue on (NULL = the library creates its own non-blocking stream; pass
 * cudaStreamLegacy / cudaStreamPerThread to name a default stream).  Fails with
 * RLS_ERR_NO_DEVICE when no sm_100 device is usable -- there is no CPU path.
 * Replaces the plugin entry NodeLoader (src/_PluginMain.cpp:16-47) + node_initialize. */
int  rls_init(int device, void *stream, rls_context **out_ctx);
int  rls_shutdown(rls_context *ctx);                      /* node_finish analogue          */
int  rls_synchronize(rls_context *ctx);                   /* waits for the context stream  */
const char *rls_last_error_string(const rls_context *ctx);/* ctx may be NULL: init errors  */
int  rls_abi_version(void);
/* The cudaStream_t (as void*) this context enqueues on. */
void *rls_stream(const rls_context *ctx);
/* Device memory on the context's device for a host driver that links against this C ABI only
 * (rlshaders_b200/host/rls_driver.cpp); rls_memcpy_to_host waits for the context's stream first. */
int rls_device_alloc(rls_context *ctx, size_t bytes, void **out_ptr);
int rls_device_free(rls_context *ctx, void *ptr);
int rls_memcpy_to_host(rls_context *ctx, void *host_dst, const void *device_src, size_t bytes);
/* Number of kernels this context has launched since creation (bench bookkeeping). */
uint64_t rls_kernel_launch_count(const rls_context *ctx);
/* Arithmetic policy of the fused *_sample_eval_pdf kernels (rlshaders_b200/csrc/rls_fp.cuh).
 * Both produce the same bits for every input; RLS_ARITH_FAST (default) runs guard-free
 * division / sqrt / reciprocal sequences and re-runs a sample with the guarded IEEE operators
 * when an operand left the window in which the two agree; RLS_ARITH_EXACT always uses the
 * guarded operators (A/B testing, tests).  rls_fallback_count reports how many samples were
 * re-run since creation / the last reset (synchronises the context's streams). */
#define RLS_ARITH_FAST  0
#define RLS_ARITH_EXACT 1
/* RLS_ARITH_TOLERANT (opt-in): the four fused *_sample_eval_pdf units and the albedo sweep computed to a stated
 * TOLERANCE instead of to the bit (rlshaders_b200/csrc/rls_tol.cuh: fused multiply-adds, MUFU reciprocals / roots / exp2,
 * polynomial log / sin / cos, visible-normal sampling without its angle round trips) -- about 40 % of the instructions
 * of the bit-exact kernels, which moves them from issue bound to HBM bound (config 2: 0.94 of the measured HBM peak).
 * FLAGS, LOBE CHOICES AND DISCONTINUOUS BRANCHES STAY BIT-EXACT: a sample whose deciding comparand lies within a band of
 * its threshold -- the band includes the sample's own estimate of the REFERENCE's rounding noise where its algorithm is
 * ill-conditioned -- is re-evaluated by the bit-exact policy in a second kernel on the same stream (rls_fallback_count
 * counts them, ~3e-4 of the samples).  Inputs are assumed finite.  Value contract, measured against the reference
 * compiled on the host (tests/test_tolerant_policy.py): directions >= 95 % within 1e-6 absolute and
 * >= 99.9 % within 1e-4; f / pdf / radii >= 90 % within 1e-5 relative and >= 99.9 % within 1e-3 -- the spread is the
 * reference's own rounding noise (its angle round trips and cancellations amplify 1 ulp to more than 1e-6 in several
 * per cent of the samples), not an error of this policy: against a binary64 evaluation of the same algorithm the
 * tolerance results are closer than the reference's own.  The entry points without a tolerance form (triples with
 * explicit wi, profiles, callers) run RLS_ARITH_FAST under this setting. */
#define RLS_ARITH_TOLERANT 2
int rls_set_arith_policy(rls_context *ctx, int policy);
int rls_get_arith_policy(const rls_context *ctx);
int rls_fallback_count(rls_context *ctx, uint64_t *out_count, int reset);
/* The node names this library stands in for: "rlGgx", "rlDisney", "rlSkin"; NULL past
 * the end -- same enumeration contract as NodeLoader(i, ...). */
const char *rls_node_name(int i);

/* -------------------------------------------------------------- rlGgx */
/* GgxSampler ctor + evalSample (src/rlGgx.h:97-107,130-156; src/rlGgx.cpp:14-99). */
int rls_ggx_eval_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                        const rls_ggx_params *params, const float *rx, const float *ry,
                        rls_vec3 out_wi, float *out_fresnel /* may be NULL */);
/* GgxSampler ctor + evalBrdf at a caller-supplied indir (src/rlGgx.h:110-119,158-165). */
int rls_ggx_eval_brdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                      const rls_ggx_params *params, rls_cvec3 wi, rls_vec3 out_f);
/* GgxSampler ctor + evalPdf at a caller-supplied indir (src/rlGgx.h:121-127,72-80). */
int rls_ggx_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                     const rls_ggx_params *params, rls_cvec3 wi, float *out_pdf);
/* The refraction half at caller-supplied directions (the pieces the rough-dielectric unit below composes):
 *   rls_ggx_refract_direction  getRefractDirection(m, V) (src/rlGgx.h:277-291; eta not squared, as written) for a
 *                              microfacet normal m: out_wi = the refracted direction, or the zero vector with
 *                              RLS_FLAG_TIR in out_flags (may be NULL) when it fails; RLS_FLAG_ENTERING as the ctor decides;
 *   rls_ggx_eval_btdf          refraction(V, wi, N) (src/rlGgx.h:316-328, Walter Eq.21) at a transmitted direction wi;
 *   rls_ggx_sample_weight      getSampleWeight(V, wi, m) (src/rlGgx.h:294-301, Walter Eq.41). */
int rls_ggx_refract_direction(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                              const rls_ggx_params *params, rls_cvec3 m, rls_vec3 out_wi, uint32_t *out_flags);
int rls_ggx_eval_btdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                      const rls_ggx_params *params, rls_cvec3 wi, float *out_ft);
int rls_ggx_sample_weight(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                          const rls_ggx_params *params, rls_cvec3 wi, rls_cvec3 m, float *out_weight);
/* The fused unit of work: ctor + evalSample + evalBrdf(L) + evalPdf(L). */
int rls_ggx_sample_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                            const rls_ggx_params *params, const float *rx, const float *ry,
                            const rls_bsdf_out *out);
/* Rough dielectric: reflection + refraction branches from one visible-normal sample. */
int rls_ggx_dielectric_sample_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                       const rls_ggx_params *params, const float *rx,
                                       const float *ry, const rls_ggx_dielectric_out *out);

/* ------------------------------------------------------------ rlDisney */
/* sample_type: RLS_RAY_DIFFUSE or RLS_RAY_GLOSSY (DisneySampler::setSampleType). */
int rls_disney_eval_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                           const rls_disney_params *params, int sample_type,
                           const float *rx, const float *ry, rls_vec3 out_wi,
                           uint32_t *out_flags /* may be NULL */);
int rls_disney_eval_brdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                         const rls_disney_params *params, int sample_type, rls_cvec3 wi,
                         rls_vec3 out_f);
int rls_disney_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                        const rls_disney_params *params, int sample_type, rls_cvec3 wi,
                        float *out_pdf);
/* Fused: ctor + glossy triple on (rx_s, ry_s) + diffuse triple on (rx_d, ry_d). */
int rls_disney_sample_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                               const rls_disney_params *params, const float *rx_s,
                               const float *ry_s, const float *rx_d, const float *ry_d,
                               const rls_disney_out *out);

/* ------------------------------------------------- rlSss / rlSkin profile */
/* `dist` is the already-multiplied scatter distance (src/rlSkin.cpp:236). */
int rls_ndprofile_set_distance(rls_context *ctx, size_t n, rls_cvec3 dist, rls_cvec3 albedo,
                               const rls_ndprofile_soa *out_profile);
int rls_ndprofile_get_radius(rls_context *ctx, size_t n, const rls_ndprofile_soa *profile,
                             const float *rx, float *out_r, uint32_t *out_flags /* may be NULL */);
int rls_ndprofile_get_pdf(rls_context *ctx, size_t n, const rls_ndprofile_soa *profile,
                          const float *r, float *out_pdf);
int rls_ndprofile_eval_profile(rls_context *ctx, size_t n, const rls_ndprofile_soa *profile,
                               const float *r, rls_vec3 out_rd);
/* Fused skin-profile unit: scatterDist = sss_scatter_dist * sss_dist_multiplier,
 * setDistance(scatterDist, sss_color), r = getRadius(rx), getPdf(r), evalProfile(r). */
int rls_skin_profile_sample_eval_pdf(rls_context *ctx, size_t n, const rls_skin_params *params,
                                     const float *rx, const rls_profile_out *out);
/* rlSkin layer hand-off (src/rlSkin.cpp:204,228,231,238): from the two average Fresnel
 * weights to the specular scale and the SSS weight. */
int rls_skin_layer_weights(rls_context *ctx, size_t n, const rls_skin_params *params,
                           const float *avg_fresnel_sheen, const float *avg_fresnel_specular,
                           float *out_specular_scale, float *out_sss_weight);

/* GaussianProfile (src/rlSss.h:63-97), single channel.  Arnold's `fast_exp` is proprietary: the
 * oracle shim defines it as expf (normative, oracle/shim/ai.h) and so does the library.
 * set_distance reads dist.x only (src/rlSss.h:73); `albedo` is accepted and unused as in the reference.
 * No guards, exactly as written: max_radius = 0 gives variance = 0 and NaN / Inf downstream. */
int rls_gaussprofile_set_distance(rls_context *ctx, size_t n, rls_cvec3 dist, rls_cvec3 albedo,
                                  const rls_gaussprofile_soa *out_profile);
int rls_gaussprofile_get_radius(rls_context *ctx, size_t n, const rls_gaussprofile_soa *profile,
                                const float *rx, float *out_r);
int rls_gaussprofile_get_pdf(rls_context *ctx, size_t n, const rls_gaussprofile_soa *profile,
                             const float *r, float *out_pdf);
int rls_gaussprofile_eval_profile(rls_context *ctx, size_t n, const rls_gaussprofile_soa *profile,
                                  const float *r, float *out_rd);
/* Fused unit, the GaussianProfile counterpart of rls_skin_profile_sample_eval_pdf:
 * setDistance((dist_x, ., .)), r = getRadius(rx), getPdf(r), evalProfile(r). */
int rls_gaussprofile_sample_eval_pdf(rls_context *ctx, size_t n, const float *dist_x, const float *rx,
                                     float *out_r, float *out_pdf, float *out_rd);

/* Probe-ray geometry of SssSampler::getProbeRay (src/rlSss.h:487-533) for one (rx, ry) pair per
 * sample: axis pick 50/25/25 % N/U/V, r = getRadius(rx'), disc offset, chord length.  `sg`
 * supplies the probe frame (N = sg->Ns, U/V explicit); origins are relative to sg->P. */
int rls_skin_probe_ray(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_skin_params *params,
                       const float *rx, const float *ry, const rls_probe_out *out);
/* The 3-axis MIS pdf of one probe hit (src/rlSss.h:252-263): `disp` = hit position - sg->P,
 * `hit_normal` = the hit's shading normal;
 * pdf = 1/4 p(r_U)|U.n| + 1/4 p(r_V)|V.n| + 1/2 p(r_N)|N.n| with p = NDProfile::getPdf. */
int rls_skin_probe_mis_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_skin_params *params,
                           rls_cvec3 disp, rls_cvec3 hit_normal, float *out_pdf);

/* -------------------------------------------- host-buffer (end-to-end) forms */
/* Same contracts as the device forms above but every array pointer is a HOST pointer.
 * `chunk` = samples per staged chunk (0 = library default).  Synchronous: the call returns when every output is in the
 * host buffers (on an error, too: the library's copy streams are drained first).  The work runs on the context's own
 * staging streams, ordered AFTER whatever was queued on the context's stream when the call was made.  A context allocates
 * its staging buffers (3 x chunk x the entry point's bytes per sample, e.g. 3 x 2^21 x 144 B = 0.9 GB for the dielectric
 * unit) at the first such call and keeps them: use one context per device for the host forms, not one per thread. */
int rls_ggx_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                 const rls_ggx_params *params, const float *rx, const float *ry,
                                 const rls_bsdf_out *out, size_t chunk);
int rls_ggx_dielectric_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                            const rls_ggx_params *params, const float *rx,
                                            const float *ry, const rls_ggx_dielectric_out *out,
                                            size_t chunk);
int rls_disney_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                    const rls_disney_params *params, const float *rx_s,
                                    const float *ry_s, const float *rx_d, const float *ry_d,
                                    const rls_disney_out *out, size_t chunk);
int rls_skin_profile_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_skin_params *params,
                                          const float *rx, const rls_profile_out *out, size_t chunk);

/* Compact shading frames for the host forms (PCIe is what bounds them: 65 -> 45 B uploaded per rlGgx sample).
 * An orthonormal frame has three degrees of freedom; shipping U, V, N costs nine floats.  `rls_shading_quat_soa` carries
 * a unit quaternion q = (x, y, z, w) per sample instead, and the frame is DEFINED as the columns of q's rotation matrix,
 * every entry evaluated in binary32 with one rounding per operation, in exactly this order (x2 = x + x etc.):
 *     x2 = x + x, y2 = y + y, z2 = z + z
 *     xx = x * x2, yy = y * y2, zz = z * z2, xy = x * y2, xz = x * z2, yz = y * z2, wx = w * x2, wy = w * y2, wz = w * z2
 *     U = (1 - (yy + zz),  xy + wz,        xz - wy)
 *     V = (xy - wz,        1 - (xx + zz),  yz + wx)
 *     N = (xz + wy,        yz - wx,        1 - (xx + yy))
 * (tests/oracle_lib.py frame_from_quaternion restates it in numpy; the decoded U, V, N then feed the reference's own
 * constructors, src/rlGgx.h:145-146, src/rlDisney.cpp:173-174, so parity is checked on identical frame bits).
 * rls_frame_from_quaternion is the decode as a device entry point; the *_hostq forms upload q, wo, backfacing, decode
 * on the device into the staging buffers and run the same fused kernels as the *_host forms. */
typedef struct rls_shading_quat_soa {
    const float   *qx, *qy, *qz, *qw;
    rls_cvec3      wo;
    const uint8_t *backfacing;          /* optional, as in rls_shading_soa */
} rls_shading_quat_soa;
int rls_frame_from_quaternion(rls_context *ctx, size_t n, const float *qx, const float *qy, const float *qz,
                              const float *qw, rls_vec3 out_U, rls_vec3 out_V, rls_vec3 out_N);
int rls_ggx_sample_eval_pdf_hostq(rls_context *ctx, size_t n, const rls_shading_quat_soa *sq,
                                  const rls_ggx_params *params, const float *rx, const float *ry,
                                  const rls_bsdf_out *out, size_t chunk);
int rls_ggx_dielectric_sample_eval_pdf_hostq(rls_context *ctx, size_t n, const rls_shading_quat_soa *sq,
                                             const rls_ggx_params *params, const float *rx,
                                             const float *ry, const rls_ggx_dielectric_out *out,
                                             size_t chunk);
int rls_disney_sample_eval_pdf_hostq(rls_context *ctx, size_t n, const rls_shading_quat_soa *sq,
                                     const rls_disney_params *params, const float *rx_s,
                                     const float *ry_s, const float *rx_d, const float *ry_d,
                                     const rls_disney_out *out, size_t chunk);
/* Pinned host allocation helpers for the _host forms. */
int rls_host_alloc(rls_context *ctx, size_t bytes, void **out_ptr);
int rls_host_free(rls_context *ctx, void *ptr);

/* --------------------------------------- directional-albedo / furnace sweep */
/* Table cell (i_r, i_c, i_e) has roughness = lerp over [roughness_lo, roughness_hi],
 * cos(theta_v) = (i_c + 1) / n_cos, ior = lerp over [ior_lo, ior_hi].  For sample k in
 * [spp_begin, spp_end) of every cell the rlGgx unit is evaluated with KsColor = 1 on the
 * canonical frame and five sums are accumulated per cell, in FP64:
 *   [0] sum f_r / pdf_r (valid samples)   [1] sum weight_t (refracted samples)
 *   [2] sum fresnel                       [3] valid-sample count   [4] TIR count
 * `table` is a device array of n_rough * n_cos * n_ior * 5 doubles and is OVERWRITTEN.
 * Sharding spp ranges across GPUs and summing the tables (NCCL all-reduce) gives the
 * full sweep; sums are order-independent up to FP64 rounding. */
typedef struct rls_sweep_grid {
    int32_t n_rough, n_cos, n_ior;
    float   roughness_lo, roughness_hi;
    float   ior_lo, ior_hi;
} rls_sweep_grid;
#define RLS_SWEEP_VALUES_PER_CELL 5
int rls_albedo_sweep(rls_context *ctx, const rls_sweep_grid *grid, uint64_t seed,
                     uint32_t spp_begin, uint32_t spp_end, double *table);

/* ---------------------------------------------- several devices, one process */
/* The multi-GPU half of the host driver that stands where Arnold's bucket / thread parallelism around shader_evaluate
 * stood (src/rlGgx.cpp:248-327).  rls_multi owns one context (own stream) per device.
 *   configs 1-4: samples are independent -- the caller gives every device a contiguous index range through the ordinary
 *     entry points on rls_multi_context(m, k) and times the slowest device with rls_multi_timer_begin / _end; no collective.
 *   config 5:    rls_multi_albedo_sweep splits the spp range [0, spp) of every cell over the devices (device k takes
 *     [spp k / G, spp (k+1) / G)), every device writes its partial table into tables[k] (a device pointer ON device k), and one
 *     ncclAllReduce(sum, FP64) per device leaves the full table in every tables[k].  Single process (ncclCommInitAll), one
 *     stream per device; from the second call with the same arguments on, kernel(s) + all-reduce are replayed from one
 *     CUDA graph per device.  libnccl.so.2 is loaded at the first such call (no link-time dependency).
 * n_devices = 0: every visible device; devices = NULL: 0 .. n_devices - 1. */
typedef struct rls_multi rls_multi;
#define RLS_MULTI_NO_GRAPH  1   /* flags of rls_multi_albedo_sweep: enqueue kernel + collective every call, no CUDA graph */
#define RLS_MULTI_NO_REDUCE 2   /*                                  leave the partial tables (no collective)              */
int  rls_multi_init(int n_devices, const int *devices, rls_multi **out);
int  rls_multi_shutdown(rls_multi *m);
int  rls_multi_device_count(const rls_multi *m);
rls_context *rls_multi_context(rls_multi *m, int k);
int  rls_multi_synchronize(rls_multi *m);
const char *rls_multi_last_error_string(const rls_multi *m);   /* m may be NULL: init errors */
int  rls_multi_timer_begin(rls_multi *m);
int  rls_multi_timer_end(rls_multi *m, float *out_ms_slowest_device, float *out_ms_per_device /* n floats, may be NULL */);
int  rls_multi_albedo_sweep(rls_multi *m, const rls_sweep_grid *grid, uint64_t seed, uint32_t spp,
                            double *const *tables, int flags);
uint64_t rls_multi_graph_replays(const rls_multi *m);           /* how many sweeps ran from the captured graphs */
int  rls_multi_set_nccl_library(const char *path);              /* optional: where libnccl.so.2 is (process-wide) */
/* The partition every multi-device path of this library uses (and the one process-per-GPU harnesses reuse, so that both
 * forms shard identically): part k of `parts` owns the contiguous items [k * total / parts, (k + 1) * total / parts) --
 * index ranges of the sample stream for configs 1-4, the spp range of every cell for the sweep.  Needs no device. */
int  rls_multi_partition(uint64_t total, int parts, int k, uint64_t *out_begin, uint64_t *out_end);

/* ------------------------- callers of the triple */
/* f2 -- rlSkin's two glossy layers for P shading points with K BRDF samples each
 * (src/rlSkin.cpp:184-238).  Per point and per layer (sheen first, then specular) exactly the
 * calls the node makes: GgxSampler(sg, color, ior, roughness) (:192,:215; anisotropic 0), then for
 * k = 0..K-1 the triple AiBRDFIntegrate drives (src/rlGgx.h:172-179): L = evalSample(rx, ry) --
 * which accumulates fresnel(L, m) and the sample count (src/rlGgx.h:100-104) --, evalBrdf(L),
 * evalPdf(L); getAvgReflectWeight() (src/rlGgx.h:181-184) and the layer hand-off
 * (src/rlSkin.cpp:204,:228,:231,:238,:244).  A layer whose weight is <= 1e-4 is skipped (:191,:214).
 * Arnold's integrator itself is proprietary; the layer estimate is DEFINED here as
 *     (1/K) sum_k evalBrdf(L_k) / evalPdf(L_k) * Li_k        (Li = 1 when li_* is all-NULL)
 * accumulated in binary32 in sample order, then scaled as the node does.
 * Sample arrays are sample-major: element [k * n_points + p]. */
typedef struct rls_skin_layers_out {
    rls_vec3  sheen;             /* sheen estimate * sheen_weight                          (:207) */
    rls_vec3  specular;          /* specular estimate * specular_weight * (1 - sheenFresnel) (:231) */
    float    *sheen_fresnel;     /* getAvgReflectWeight() * sheen_weight, 0 if skipped      (:204) */
    float    *specular_fresnel;  /* getAvgReflectWeight() * specular_weight, 0 if skipped   (:228) */
    float    *sss_weight;        /* sss_weight * (1 - specularFresnel * (1 - sheenFresnel)) (:238) */
    uint32_t *flags;             /* RLS_SKIN_* */
} rls_skin_layers_out;
#define RLS_SKIN_SHEEN_EVALUATED    0x1u   /* sheen_weight > 1e-4     */
#define RLS_SKIN_SPECULAR_EVALUATED 0x2u   /* specular_weight > 1e-4  */
#define RLS_SKIN_SSS_SKIPPED        0x4u   /* sss weight < 1e-4 (:244) */
int rls_skin_glossy_layers(rls_context *ctx, size_t n_points, uint32_t k, const rls_shading_soa *sg,
                           const rls_skin_params *params,
                           const float *rx_sheen, const float *ry_sheen,
                           const float *rx_specular, const float *ry_specular,
                           rls_cvec3 li_sheen, rls_cvec3 li_specular,     /* all-NULL = radiance 1 */
                           const rls_skin_layers_out *out);

/* f3 -- one light sample evaluated against a BRDF with multiple importance sampling: the shape
 * of AiEvaluateLightSample(sg, brdf, evalSample, evalBrdf, evalPdf) (src/rlGgx.h:167-170,
 * src/rlDisney.cpp:266-277; called from the light loops src/rlGgx.cpp:286-295,
 * src/rlDisney.cpp:695-704).  Arnold's estimator is proprietary; DEFINED here as the two-sample
 * power heuristic (beta = 2):
 *     light half : f(Ld) * Li * w(p_l, p_b(Ld)) / p_l          f = evalBrdf, p_b = evalPdf
 *     BRDF half  : L = evalSample(rx, ry);  f(L) * Li_b * w(p_b(L), p_lb) / p_b(L)
 *     w(a, b) = a^2 / (a^2 + b^2)
 * `light` describes the light sample (sg->Ld, sg->Li, its pdf); `light_at_l` (may be NULL: no
 * BRDF half) the same light evaluated along the sampled direction L, which the caller obtains
 * from rls_*_eval_sample with the same (rx, ry).  A half whose pdf is 0 or whose direction is
 * the zero vector contributes 0. */
typedef struct rls_light_sample {
    rls_cvec3    dir;       /* sg->Ld (unit); ignored for light_at_l (the direction is L)   */
    rls_cvec3    radiance;  /* sg->Li                                                       */
    const float *pdf;       /* light pdf w.r.t. solid angle                                 */
} rls_light_sample;
int rls_ggx_evaluate_light_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                  const rls_ggx_params *params, const rls_light_sample *light,
                                  const float *rx, const float *ry, const rls_light_sample *light_at_l,
                                  rls_vec3 out_rgb, float *out_w_light /* may be NULL */,
                                  float *out_w_brdf /* may be NULL */);
int rls_disney_evaluate_light_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                     const rls_disney_params *params, int sample_type,
                                     const rls_light_sample *light, const float *rx, const float *ry,
                                     const rls_light_sample *light_at_l, rls_vec3 out_rgb,
                                     float *out_w_light, float *out_w_brdf);

/* f4 -- the developer's visual check SampleWriter (src/rlUtil.h:44-171; driven from the
 * commented sweep src/rlGgx.cpp:202-224, src/rlDisney.cpp:642-653) for ONE shading point
 * (element `point` of sg / params):
 *   radiance image (writeRadiance :98-114): pixel (i, j) = evalBrdf(sphericalDirection(
 *       cosf(pi/2 * j / height), 2 pi * i / width)) -- the direction is used as is, i.e. the
 *       image is in the frame of the shading normal only when the frame is the identity;
 *   sample scatter (writeSample :116-156): for each uniform pair L = evalSample(rx, ry); a
 *       zero L is skipped; pixel (clamp(int(phi / 2pi * width)), clamp(int(theta / (pi/2) *
 *       height))) is painted green, or red when theta > pi/2; where several samples hit a
 *       pixel the LAST one wins, as in the sequential loop.
 * `image` is a device array of 3 * width * height floats in the writer's plane order B, G, R
 * (writePixel :158-163).  The radiance pass overwrites it; the scatter pass paints over it, so
 * calling both reproduces the file the writer saves.  out_missing (device, may be NULL)
 * receives the count of samples with theta > pi/2 (:152). */
#define RLS_NODE_GGX    0
#define RLS_NODE_DISNEY 1
int rls_sample_writer_radiance(rls_context *ctx, int node, const rls_shading_soa *sg, const void *params,
                               size_t point, int sample_type, int width, int height, float *image);
int rls_sample_writer_scatter(rls_context *ctx, int node, const rls_shading_soa *sg, const void *params,
                              size_t point, int sample_type, size_t n_samples, const float *rx,
                              const float *ry, int width, int height, float *image,
                              uint32_t *scratch /* width * height, device */, uint32_t *out_missing);

/* ------------------------------------------- synthetic workload generators */
/* Counter-based integer hash h(seed, stream, index) -> 24-bit uniform in
 * [2^-24, 1 - 2^-24]; out[i] = lo + (hi - lo) * u(first_index + i).  The integer part is
 * bit-reproducible on the host (tests restate it in numpy). */
int rls_synth_uniform(rls_context *ctx, size_t n, uint64_t seed, uint32_t stream,
                      uint64_t first_index, float lo, float hi, float *out);
/* Random orthonormal frames (N uniform on the sphere) and view vectors with
 * cos(theta_v) ~ U[cos_lo, cos_hi], phi ~ U[0, 2pi) about N.  Writes the 12 arrays of
 * `sg` (cast away const) and, if sg->backfacing != NULL, 1 with probability
 * `backfacing_fraction`. */
int rls_synth_shading(rls_context *ctx, size_t n, uint64_t seed, uint64_t first_index,
                      float cos_lo, float cos_hi, float backfacing_fraction,
                      const rls_shading_soa *sg);

/* ------------------------------------------------------------ diagnostics */
/* Evaluates one of the library's own binary32 transcendentals (the device restatement of the
 * host libm the reference links against, rlshaders_b200/csrc/rls_libm.cuh) element-wise, so a
 * test can compare the DEVICE results with the host C library bit for bit.
 * fn: 0 sincosf(a) -> out0 = sin, out1 = cos;  1 tanf(a);  2 atanf(a);  3 acosf(a);  4 expf(a);
 *     5 logf(a);  6 atan2f(a, b);  7 powf(a, b).  `b` / `out1` may be NULL when unused. */
int rls_debug_libm(rls_context *ctx, int fn, size_t n, const float *a, const float *b,
                   float *out0, float *out1);

/* Fast arithmetic policy == exact policy over a RANGE OF BIT PATTERNS (rlshaders_b200/csrc/rls_fp.cuh):
 * argument k is the binary32 with bits first_bits + k * stride, k < count.  For every argument whose
 * fast-policy evaluation leaves the operand tracker satisfied the result must equal the guarded
 * IEEE evaluation bit for bit.  counts (device, 3 x uint64, caller-zeroed): [0] arguments accepted
 * by the tracker, [1] mismatches among them (must stay 0), [2] arguments sent to the exact re-run.
 * fn: 0 sqrt(a); 1 1/a; 2 a/b; 3 tanf(a); 4 acosf(a); 5 atan2f(a, b); 6 atan2f(b, a);
 *     7 a/b with a zero-tolerant numerator and b > 0; 8 b/a; 9 a/3 with the literal reciprocal;
 *     10 a/b through the shared refined reciprocal of b; 11 expf(a); 12 sincosf(a) (sine and cosine). */
int rls_debug_policy_check(rls_context *ctx, int fn, uint32_t first_bits, uint64_t count, uint32_t stride,
                           float b, unsigned long long *counts);

#ifdef __cplusplus
}
#endif
#endif /* RLS_B200_H */
