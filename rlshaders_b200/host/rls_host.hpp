// rls_host.hpp -- C++ host-side mirror of the reference's sampler interface over the C ABI.
//
// The reference gives Arnold, per shading point, a sampler object plus the static callback
// triple evalSample / evalBrdf / evalPdf (src/rlGgx.h:97-127, src/rlDisney.cpp:109-152) and
// Arnold loops over samples.  Here the loop is the batch: a sampler is constructed over a
// whole batch of shading points (SoA, host memory) and each call evaluates every sample on the
// GPU through include/rls_b200.h.  Names, argument meaning and in-band error behaviour (zero
// vector = invalid sample, black / pdf 0 on a zero direction) follow the reference.
//
// Header-only; link with librls_b200.so.  No CUDA headers are needed by the client.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rls_b200.h"

namespace rls {
namespace host {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// Owns an rls_context (one per host thread / device, like one Arnold render thread).
class Context {
public:
    explicit Context(int device = 0)
    {
        if (rls_init(device, nullptr, &mCtx) != RLS_OK)
            throw Error(std::string("rls_init: ") + rls_last_error_string(nullptr));
    }
    ~Context() { if (mCtx) rls_shutdown(mCtx); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    rls_context *get() const { return mCtx; }
    void check(int rc) const { if (rc != RLS_OK) throw Error(rls_last_error_string(mCtx)); }
    // Pinned host arrays: the *_host entry points overlap H2D / kernel / D2H on them.
    template <typename T> T *alloc(size_t n) const
    {
        void *p = nullptr;
        check(rls_host_alloc(mCtx, n * sizeof(T), &p));
        return static_cast<T *>(p);
    }
    void free(void *p) const { rls_host_free(mCtx, p); }
private:
    rls_context *mCtx = nullptr;
};

// A batch of pinned SoA float arrays with one owner.
class Arena {
public:
    explicit Arena(const Context &ctx) : mCtx(ctx) {}
    ~Arena() { for (void *p : mBlocks) mCtx.free(p); }
    float *floats(size_t n) { float *p = mCtx.alloc<float>(n); mBlocks.push_back(p); return p; }
    uint32_t *words(size_t n) { uint32_t *p = mCtx.alloc<uint32_t>(n); mBlocks.push_back(p); return p; }
    uint8_t *bytes(size_t n) { uint8_t *p = mCtx.alloc<uint8_t>(n); mBlocks.push_back(p); return p; }
    rls_vec3 vec3(size_t n) { rls_vec3 v = { floats(n), floats(n), floats(n) }; return v; }
private:
    const Context &mCtx;
    std::vector<void *> mBlocks;
};

inline rls_cvec3 as_const(const rls_vec3 &v) { rls_cvec3 c = { v.x, v.y, v.z }; return c; }
inline rls_param1 uniform(float v) { rls_param1 p = { v, nullptr }; return p; }
inline rls_param3 uniform(float r, float g, float b) { rls_param3 p = { { r, g, b }, { nullptr, nullptr, nullptr } }; return p; }
inline rls_param1 varying(const float *a) { rls_param1 p = { 0.0f, a }; return p; }

// Node-parameter defaults of the reference (src/rlGgx.cpp:172-186, rlDisney.cpp:606-628,
// rlSkin.cpp:109-131).
inline rls_ggx_params ggx_defaults()
{
    rls_ggx_params p = {};
    p.KsColor = uniform(1, 1, 1); p.Ks = uniform(0.5f); p.specularRoughness = uniform(0.0f);
    p.ior = uniform(1.0f); p.anisotropic = uniform(0.0f);
    p.KdColor = uniform(1, 1, 1); p.Kd = uniform(0.5f); p.diffuseRoughness = uniform(0.0f);
    p.KtColor = uniform(1, 1, 1); p.Kt = uniform(0.0f); p.opacity = uniform(1.0f); p.opacity_color = uniform(1, 1, 1);
    return p;
}
inline rls_disney_params disney_defaults()
{
    rls_disney_params p = {};
    p.base_color = uniform(1, 1, 1); p.opacity = uniform(1, 1, 1);
    p.indirectDiffuseScale = uniform(1.0f); p.indirectSpecularScale = uniform(1.0f);
    p.sample_from_visible_normal = 1;   // src/rlDisney.cpp:191
    return p;   // the ten scalar parameters default to 0
}
inline rls_skin_params skin_defaults()
{
    rls_skin_params p = {};
    p.sss_color = uniform(1, 1, 1); p.sss_weight = uniform(1.0f); p.sss_dist_multiplier = uniform(1.0f);
    p.sss_scatter_dist = uniform(1, 1, 1); p.sss_cavity_fadeout = 1;
    p.specular_color = uniform(1, 1, 1); p.specular_weight = uniform(0.6f); p.specular_roughness = uniform(0.5f);
    p.specular_ior = uniform(1.44f); p.sheen_color = uniform(1, 1, 1); p.sheen_weight = uniform(0.0f);
    p.sheen_roughness = uniform(0.35f); p.sheen_ior = uniform(1.44f); p.opacity = uniform(1.0f);
    p.opacity_color = uniform(1, 1, 1);
    return p;
}

// Compact frames (include/rls_b200.h rls_shading_quat_soa): the unit quaternion of the rotation whose columns are
// U, V, N, for hosts that hold their frames as vectors.  Computed in double (largest-component branch) and rounded once;
// the library's decode of q DEFINES the frame the kernels see, so encode once and use q from then on.
inline void quaternion_from_frame(const float U[3], const float V[3], const float N[3], float q[4])
{
    const double m00 = U[0], m10 = U[1], m20 = U[2], m01 = V[0], m11 = V[1], m21 = V[2], m02 = N[0], m12 = N[1], m22 = N[2];
    double c[4][4] = { { 1 + m00 - m11 - m22, m01 + m10, m02 + m20, m21 - m12 },
                       { m01 + m10, 1 - m00 + m11 - m22, m12 + m21, m02 - m20 },
                       { m02 + m20, m12 + m21, 1 - m00 - m11 + m22, m10 - m01 },
                       { m21 - m12, m02 - m20, m10 - m01, 1 + m00 + m11 + m22 } };
    int best = 0;
    for (int k = 1; k < 4; k++) if (c[k][k] > c[best][best]) best = k;
    double s = 0.0;
    for (int k = 0; k < 4; k++) s += c[best][k] * c[best][k];
    s = 1.0 / std::sqrt(s);
    for (int k = 0; k < 4; k++) q[k] = (float)(c[best][k] * s);
}

// Batched rls::GgxSampler.  `sg` and every array are pinned host memory of n entries; the second constructor takes the
// compact shading batch (quaternion frames) and routes to the *_hostq forms.
class GgxSampler {
public:
    GgxSampler(const Context &ctx, size_t n, const rls_shading_soa &sg, const rls_ggx_params &params)
        : mCtx(ctx), mN(n), mSg(sg), mSq(), mQuat(false), mParams(params) {}
    GgxSampler(const Context &ctx, size_t n, const rls_shading_quat_soa &sq, const rls_ggx_params &params)
        : mCtx(ctx), mN(n), mSg(), mSq(sq), mQuat(true), mParams(params) {}
    // evalSample + evalBrdf + evalPdf for every sample (the fused unit of work).
    void sampleEvalPdf(const float *rx, const float *ry, const rls_bsdf_out &out, size_t chunk = 0) const
    {
        mCtx.check(mQuat ? rls_ggx_sample_eval_pdf_hostq(mCtx.get(), mN, &mSq, &mParams, rx, ry, &out, chunk)
                         : rls_ggx_sample_eval_pdf_host(mCtx.get(), mN, &mSg, &mParams, rx, ry, &out, chunk));
    }
    // Rough dielectric: reflection and refraction branches (src/rlGgx.h:228-243).
    void dielectricSampleEvalPdf(const float *rx, const float *ry, const rls_ggx_dielectric_out &out, size_t chunk = 0) const
    {
        mCtx.check(mQuat ? rls_ggx_dielectric_sample_eval_pdf_hostq(mCtx.get(), mN, &mSq, &mParams, rx, ry, &out, chunk)
                         : rls_ggx_dielectric_sample_eval_pdf_host(mCtx.get(), mN, &mSg, &mParams, rx, ry, &out, chunk));
    }
private:
    const Context &mCtx;
    size_t mN;
    rls_shading_soa mSg;
    rls_shading_quat_soa mSq;
    bool mQuat;
    rls_ggx_params mParams;
};

class DisneySampler {
public:
    DisneySampler(const Context &ctx, size_t n, const rls_shading_soa &sg, const rls_disney_params &params)
        : mCtx(ctx), mN(n), mSg(sg), mSq(), mQuat(false), mParams(params) {}
    DisneySampler(const Context &ctx, size_t n, const rls_shading_quat_soa &sq, const rls_disney_params &params)
        : mCtx(ctx), mN(n), mSg(), mSq(sq), mQuat(true), mParams(params) {}
    void sampleEvalPdf(const float *rx_s, const float *ry_s, const float *rx_d, const float *ry_d,
                       const rls_disney_out &out, size_t chunk = 0) const
    {
        mCtx.check(mQuat ? rls_disney_sample_eval_pdf_hostq(mCtx.get(), mN, &mSq, &mParams, rx_s, ry_s, rx_d, ry_d, &out, chunk)
                         : rls_disney_sample_eval_pdf_host(mCtx.get(), mN, &mSg, &mParams, rx_s, ry_s, rx_d, ry_d, &out, chunk));
    }
private:
    const Context &mCtx;
    size_t mN;
    rls_shading_soa mSg;
    rls_shading_quat_soa mSq;
    bool mQuat;
    rls_disney_params mParams;
};

class SkinProfile {
public:
    SkinProfile(const Context &ctx, size_t n, const rls_skin_params &params) : mCtx(ctx), mN(n), mParams(params) {}
    void sampleEvalPdf(const float *rx, const rls_profile_out &out, size_t chunk = 0) const
    {
        mCtx.check(rls_skin_profile_sample_eval_pdf_host(mCtx.get(), mN, &mParams, rx, &out, chunk));
    }
private:
    const Context &mCtx;
    size_t mN;
    rls_skin_params mParams;
};

} // namespace host
} // namespace rls
