// rls_math.cuh -- scalar/vector device maths shared by every kernel of the path.
//
// Numerical contract (DESIGN.md "Numerics"):
//   * this translation unit is compiled with -fmad=false -prec-div=true -prec-sqrt=true
//     -ftz=false, so every + - * / sqrt below is ONE correctly rounded IEEE-754 binary32
//     operation, in the order written -- the same value the reference's host build
//     (-ffp-contract=off) produces.  Operation order therefore follows the reference
//     expression by expression (cited per function in rls_ggx.cuh / rls_disney.cuh /
//     rls_profile.cuh).
//   * transcendentals go through rlm:: (rls_libm.cuh), which reproduces the host
//     libm's binary32 results.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rls_fp.cuh"

#define RLS_DEV __device__ __forceinline__

namespace rls {

constexpr float kEps      = 1.0e-4f;                    // AI_EPSILON
constexpr float kPi       = 3.14159265358979323846f;    // AI_PI
constexpr float kTwoPi    = 6.28318530717958647692f;    // AI_PITIMES2
constexpr float kHalfPi   = 1.57079632679489661923f;    // AI_PIOVER2
constexpr float kInvPi    = 0.31830988618379067154f;    // AI_ONEOVERPI

struct f3 { float x, y, z; };
struct f2 { float x, y; };

RLS_DEV f3 mk3(float x, float y, float z) { f3 v; v.x = x; v.y = y; v.z = z; return v; }
RLS_DEV f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RLS_DEV f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RLS_DEV f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RLS_DEV f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RLS_DEV float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RLS_DEV bool  is_zero(f3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

RLS_DEV float sqr(float a) { return a * a; }
// The reference's macros: ABS(a) = a < 0 ? -a : a (keeps -0 and NaN as they are),
// MAX(a,b) = a > b ? a : b (NaN -> b).  fabsf/fmaxf differ on those inputs.
RLS_DEV float abs_m(float a) { return (a < 0.0f) ? -a : a; }
RLS_DEV float max_m(float a, float b) { return (a > b) ? a : b; }
RLS_DEV float clamp_m(float v, float lo, float hi) { return (v < lo) ? lo : ((v > hi) ? hi : v); }
RLS_DEV float sgn_m(float a) { return (a < 0.0f) ? -1.0f : ((a > 0.0f) ? 1.0f : 0.0f); }
RLS_DEV float lerp_m(float t, float a, float b) { return (1.0f - t) * a + b * t; }
RLS_DEV f3    lerp_m(float t, f3 a, f3 b) { return a * (1.0f - t) + b * t; }
template <class Fp>
RLS_DEV float linearstep_m(Fp &fp, float lo, float hi, float t) { return clamp_m(fp.div_pz(t - lo, hi - lo), 0.0f, 1.0f); }

// AiV3Normalize: reciprocal, then three multiplies; the zero vector stays zero.
template <class Fp>
RLS_DEV f3 normalize(Fp &fp, f3 a)
{
    float len = fp.sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    if (Fp::kFast || len != 0.0f) {              // fast policy: a zero length has already left the window
        float inv = fp.rcp_in_window(len);       // len = sqrt(t), t in [2^-60, 2^120] once tracked
        return mk3(a.x * inv, a.y * inv, a.z * inv);
    }
    return a;
}
// AiV3RotateToFrame(a, u, v, w)
RLS_DEV f3 rotate_to_frame(f3 a, f3 u, f3 v, f3 w)
{
    return mk3(a.x * u.x + a.y * v.x + a.z * w.x,
               a.x * u.y + a.y * v.y + a.z * w.y,
               a.x * u.z + a.y * v.z + a.z * w.z);
}

// ---- parameter fetch: uniform value or per-sample array (include/rls_b200.h rls_param1/3)
struct P1 { float value; const float *array; };
struct P3 { float value[3]; const float *x, *y, *z; };
// kReload = true: ld.global.cg instead of ld.global.nc -- used by the exact re-run of the fused
// kernels, which reloads its inputs so that the fast path need not keep them in registers (the
// compiler cannot merge the two kinds of load).
template <bool kReload = false> RLS_DEV float ld_in(const float *p) { return kReload ? __ldcg(p) : __ldg(p); }
template <bool kReload = false> RLS_DEV float fetch(const P1 &p, uint32_t i) { return p.array ? ld_in<kReload>(p.array + i) : p.value; }
template <bool kReload = false> RLS_DEV f3 fetch(const P3 &p, uint32_t i)
{
    return mk3(p.x ? ld_in<kReload>(p.x + i) : p.value[0], p.y ? ld_in<kReload>(p.y + i) : p.value[1],
               p.z ? ld_in<kReload>(p.z + i) : p.value[2]);
}
// kArrays = true: the launch site has checked that the parameter IS a per-sample array, so the
// kernel skips the per-parameter pointer tests (a dozen of them per rlDisney sample).
template <bool kArrays, bool kReload = false> RLS_DEV float fetch_t(const P1 &p, uint32_t i)
{
    return kArrays ? ld_in<kReload>(p.array + i) : fetch<kReload>(p, i);
}
template <bool kArrays, bool kReload = false> RLS_DEV f3 fetch_t(const P3 &p, uint32_t i)
{
    return kArrays ? mk3(ld_in<kReload>(p.x + i), ld_in<kReload>(p.y + i), ld_in<kReload>(p.z + i)) : fetch<kReload>(p, i);
}
struct CV3 { const float *x, *y, *z; };
struct V3  { float *x, *y, *z; };
template <bool kReload = false> RLS_DEV f3 load3(const CV3 &v, uint32_t i)
{
    return mk3(ld_in<kReload>(v.x + i), ld_in<kReload>(v.y + i), ld_in<kReload>(v.z + i));
}
RLS_DEV void store3(const V3 &v, uint32_t i, f3 a) { v.x[i] = a.x; v.y[i] = a.y; v.z[i] = a.z; }

// Shading inputs of one sample (include/rls_b200.h rls_shading_soa).
struct ShadingSoA { CV3 U, V, N, wo; const uint8_t *backfacing; };
struct Shading { f3 U, V, N, wo; bool backfacing; };
template <bool kReload = false> RLS_DEV Shading load_shading(const ShadingSoA &s, uint32_t i)
{
    Shading o;
    o.U = load3<kReload>(s.U, i); o.V = load3<kReload>(s.V, i); o.N = load3<kReload>(s.N, i); o.wo = load3<kReload>(s.wo, i);
    o.backfacing = s.backfacing ? ((kReload ? __ldcg(s.backfacing + i) : __ldg(s.backfacing + i)) != 0) : false;
    return o;
}

} // namespace rls
