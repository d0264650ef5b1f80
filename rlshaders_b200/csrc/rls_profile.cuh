// rls_profile.cuh -- device restatement of the normalized-diffusion profile NDProfile
// (reference src/rlSss.cpp:20-106, src/rlSss.h:30-42), the probe-ray geometry of
// SssSampler::getProbeRay (src/rlSss.h:487-533) and the rlSkin layer hand-off
// (src/rlSkin.cpp:191-238).
#pragma once
#include "rls_ggx.cuh"

namespace rls {

struct NdProfile {
    float d[3], C1[3], C2[3], R;
    float yd[3];     // fused unit only: the refined reciprocal of d[i] that its quotients by d[i] share (Fp::shared_rcp)
};

// src/rlSss.cpp:20-34.  The `s` of :23 (powf of the albedo luminance) is dead code in the
// reference and is not evaluated; the albedo therefore does not enter the profile.
template <class Fp>
RLS_DEV void nd_set_distance(Fp &fp, NdProfile &p, f3 dist)
{
    p.d[0] = dist.x; p.d[1] = dist.y; p.d[2] = dist.z;
    p.R = max_m(dist.x, max_m(dist.y, dist.z)) * 3.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float d = p.d[i];
        p.yd[i] = fp.shared_rcp(d);
        const float q = fp.div_by(-p.R, d, p.yd[i]);         // -R/d, evaluated once for :31 and :32
        p.C1[i] = 1.0f - rlm::expf_(fp, q);
        p.C2[i] = 1.0f - rlm::expf_(fp, fp.div3(q));
    }
}
// src/rlSss.h:30-42 (thirds at 0.3333f / 0.6666f)
template <class Fp>
RLS_DEV int nd_select_dist_lobe(Fp &fp, float &x)
{
    // one LINEARSTEP on selected bounds instead of three divergent ones (x - 0 is exact; hi - lo is the same
    // binary32 difference whether the compiler or the FADD forms it)
    const int ch = x < 0.3333f ? 0 : (x > 0.6666f ? 2 : 1);
    const float lo = ch == 0 ? 0.0f : (ch == 2 ? 0.6666f : 0.3333f);
    const float hi = ch == 0 ? 0.3333f : (ch == 2 ? 1.0f : 0.6666f);
    x = linearstep_m(fp, lo, hi, x);
    return ch;
}
RLS_DEV float pick3(const float (&a)[3], int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }

// src/rlSss.cpp:36-66; also reports the channel / exponential-lobe / degenerate flags.
template <class Fp>
RLS_DEV float nd_get_radius(Fp &fp, const NdProfile &p, float rx, uint32_t &flags)
{
    float x = rx;
    int ch = nd_select_dist_lobe(fp, x);
    flags = (uint32_t)ch << 8;                       // RLS_FLAG_LOBE_SHIFT
    float d = pick3(p.d, ch);
    if (p.R < kEps || d < kEps) { flags |= 0x0800u; return 0.0f; }   // RLS_FLAG_DEGENERATE
    float w1 = pick3(p.C1, ch);
    float w2 = pick3(p.C2, ch);
    float w = fp.div(w1, w1 + w2 * 3.0f);
    // :52-65.  The two exponential lobes run the same operations on different operands, so the operands are
    // selected and ONE linearstep + logf is evaluated (as a branch, almost every warp ran both sides: two quotients
    // and two logf for one); x - 0 and w - 0 are exact, so LINEARSTEP(0, w, x) is unchanged.
    const bool wide = x > w;                         // the exp(-r/3d) lobe
    flags |= wide ? 0x0400u : 0u;                    // RLS_FLAG_EXP_LOBE
    x = linearstep_m(fp, wide ? w : 0.0f, wide ? 1.0f : w, x);
    return rlm::logf_(fp, 1.0f - x * (wide ? w2 : w1)) * (wide ? -d * 3.0f : -d);
}
// src/rlSss.cpp:68-84
template <class Fp>
RLS_DEV float nd_get_pdf(Fp &fp, const NdProfile &p, float r)
{
    if (p.R < kEps) return 1.0f;
    float pdf = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float d = max_m(p.d[i], kEps);
        float p1 = rlm::expf_(fp, fp.div(-r, d));
        float p2 = rlm::expf_(fp, fp.div(fp.div(-r, d), 3.0f));
        pdf += fp.div(fp.div(p1 + p2, d), p.C1[i] + p.C2[i] * 3.0f);
    }
    return fp.div(pdf, kTwoPi * r * 3.0f);
}
// src/rlSss.cpp:86-106
template <class Fp>
RLS_DEV f3 nd_eval_profile(Fp &fp, const NdProfile &p, float r)
{
    if (p.R < kEps) return mk3(0.0f, 0.0f, 0.0f);
    else if (r < kEps) return mk3(1.0f, 1.0f, 1.0f);
    float denom = 8.0f * kPi * r;
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float d = p.d[i];
        o[i] = 1.0f;
        if (!(d < kEps)) o[i] = fp.div(rlm::expf_(fp, fp.div(-r, d)) + rlm::expf_(fp, fp.div(-r, 3.0f * d)), denom * d);
    }
    return mk3(o[0], o[1], o[2]);
}

// getPdf(r) and evalProfile(r) at the same radius (the fused skin unit): exp(-r/d) of a channel
// is the same IEEE operation in both (src/rlSss.cpp:78 and :102) whenever d >= AI_EPSILON, where
// getPdf's floored distance MAX(d, AI_EPSILON) is d itself -- evaluate it once.
template <class Fp>
RLS_DEV void nd_pdf_and_profile(Fp &fp, const NdProfile &p, float r, float &pdf_out, f3 &rd_out)
{
    if (p.R < kEps) { pdf_out = 1.0f; rd_out = mk3(0.0f, 0.0f, 0.0f); return; }
    const bool white = r < kEps;                 // evalProfile: white for r < AI_EPSILON
    const float denom = 8.0f * kPi * r;
    float pdf = 0.0f;
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float d = p.d[i];
        const float dm = max_m(d, kEps);
        // dm == d unless d < AI_EPSILON (or NaN).  The fast policy sends those samples to the exact
        // re-run (fp.require) and forms the quotients by dm with setDistance's reciprocal of d.
        const bool same = dm == d;
        fp.require(same);
        const float q1 = Fp::kFast ? fp.div_by(-r, d, p.yd[i]) : fp.div(-r, dm);
        const float p1 = rlm::expf_(fp, q1);
        const float p2 = rlm::expf_(fp, fp.div3(q1));
        const float s12 = p1 + p2;
        pdf += fp.div(Fp::kFast ? fp.div_by(s12, d, p.yd[i]) : fp.div(s12, dm), p.C1[i] + p.C2[i] * 3.0f);
        o[i] = 1.0f;
        if (!white && !(d < kEps)) {             // dm == d here (unless d is NaN): exp(-r/d) == p1
            const float e1 = (Fp::kFast || same) ? p1 : rlm::expf_(fp, fp.div(-r, d));
            o[i] = fp.div(e1 + rlm::expf_(fp, fp.div(-r, 3.0f * d)), denom * d);
        }
    }
    pdf_out = fp.div(pdf, kTwoPi * r * 3.0f);
    rd_out = mk3(o[0], o[1], o[2]);
}

// ------------------------------------------------------------------ GaussianProfile
// src/rlSss.h:63-97, the alternative `Profile` argument of SssSampler (never instantiated by the
// plugin: rlSkin uses SssSampler<NDProfile>).  `fast_exp` is Arnold's; the shim (normative) and this
// restatement use expf.  No guards, as written: R = 0 -> variance 0 -> NaN / Inf downstream.
struct GaussProfile { float var, R, norm; };
constexpr float kInvTwoPi = 0.15915494309189533577f;   // AI_ONEOVER2PI

// :71-76  (`albedo` unused; only dist.x is read)
template <class Fp>
RLS_DEV void gauss_set_distance(Fp &fp, GaussProfile &p, float dist_x)
{
    p.R = dist_x;
    p.var = fp.div(p.R * p.R, 12.46f);
    p.norm = 1.0f - rlm::expf_(fp, fp.div(-(p.R * p.R) * 0.5f, p.var));
}
// :78-81
template <class Fp>
RLS_DEV float gauss_get_radius(Fp &fp, const GaussProfile &p, float rx)
{
    return fp.sqrt(-2.0f * p.var * rlm::logf_(fp, 1.0f - rx * p.norm));
}
// :88-91
template <class Fp>
RLS_DEV float gauss_eval_profile(Fp &fp, const GaussProfile &p, float r)
{
    return fp.div(kInvTwoPi, p.var) * rlm::expf_(fp, fp.div(-r * r * 0.5f, p.var));
}
// :83-86
template <class Fp>
RLS_DEV float gauss_get_pdf(Fp &fp, const GaussProfile &p, float r)
{
    return fp.div(gauss_eval_profile(fp, p, r), p.norm);
}

// The fused unit: setDistance((dist_x, ., .)), r = getRadius(rx), evalProfile(r), getPdf(r).  The three
// quotients by mVariance share one refined reciprocal under the fast policy (plain divisions under FpExact);
// getPdf's evalProfile(r) is the same IEEE value as the evalProfile(r) result and is formed once.
struct Gauss1 { float r, pdf, rd; };
template <class Fp>
RLS_DEV Gauss1 gauss_profile_unit(Fp &fp, float dist_x, float rx)
{
    Gauss1 o;
    const float R2 = dist_x * dist_x;
    const float var = fp.div(R2, 12.46f);
    const float yv = fp.shared_rcp(var);
    const float norm = 1.0f - rlm::expf_(fp, fp.div_by(-R2 * 0.5f, var, yv));
    o.r = fp.sqrt(-2.0f * var * rlm::logf_(fp, 1.0f - rx * norm));
    o.rd = fp.div_by(kInvTwoPi, var, yv) * rlm::expf_(fp, fp.div_by(-o.r * o.r * 0.5f, var, yv));
    o.pdf = fp.div(o.rd, norm);
    return o;
}

struct SkinParamsDev {
    P3 sss_color, sss_scatter_dist;
    P1 sss_weight, sss_dist_multiplier, specular_weight, sheen_weight;
};
// src/rlSkin.cpp:236
template <bool kReload = false>
RLS_DEV f3 skin_scatter_dist(const SkinParamsDev &p, uint32_t i)
{
    return fetch<kReload>(p.sss_scatter_dist, i) * fetch<kReload>(p.sss_dist_multiplier, i);
}

} // namespace rls
