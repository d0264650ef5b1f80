// rls_f2.cuh -- two-lane (packed f32x2) vectors, predicates, the reference's macros lane-wise and
// the two-lane form of the fast arithmetic policy (rls_fp.cuh FpFast).  Used by the paired
// evaluation of the fused units (rls_pair.cuh: two same-shaped evaluations of ONE sample share an
// instruction stream) and by the two-samples-per-thread experiment (rls_packed.cuh).
// The f32x2 primitives themselves (F2, mk, fma2, +, -, *) live in rls_fp.cuh.
#pragma once
#include "../rls_math.cuh"

namespace rls {
// ------------------------------------------------------------------ packed f32x2 primitives
// One FADD2 / FFMA2 performs the same IEEE-754 binary32 operation on two lanes for ONE issue slot
// (two FMA-pipe cycles: profiles/r01_ffma2_microbench.txt).  ptxas 12.9 contracts
// mul.rn.f32x2 -> add.rn.f32x2 into a fused FFMA2 even with --fmad=false
// (tools/microbench/f32x2_exact_test.cu), so every packed multiply is spelled
// fma.rn.f32x2(a, b, {-0, -0}) with the addend read from a __constant__ the HOST fills at
// rls_init: x*y + (-0) is x*y for every x*y (signed zeros included) and ptxas can neither fold a
// value it does not know nor contract an fma into the add that follows.  mk(a, a) lowers to the
// SASS broadcast operand form (R.F32), so a scalar shared by both lanes costs no extra register.
namespace pk {

__constant__ unsigned long long c_negzero2;      // {-0.0f, -0.0f}; written by rls_init (opaque to ptxas)

struct F2 { unsigned long long v; };
struct B2 { bool a, b; };                        // per-lane predicate

RLS_FP_D F2 mk(float a, float b) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
RLS_FP_D float lo(F2 p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); (void)b; return a; }
RLS_FP_D float hi(F2 p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); (void)a; return b; }
RLS_FP_D F2 bc(float a) { return mk(a, a); }
RLS_FP_D F2 nz() { F2 r; r.v = c_negzero2; return r; }
RLS_FP_D F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
RLS_FP_D F2 fma2_rd(F2 a, F2 b, F2 c) { F2 r; asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
RLS_FP_D F2 operator+(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
RLS_FP_D F2 operator-(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
RLS_FP_D F2 operator*(F2 a, F2 b) { return fma2(a, b, nz()); }
RLS_FP_D F2 operator-(F2 a) { return nz() - a; }                  // (-0) - x == -x for every x
RLS_FP_D F2 operator*(F2 a, float s) { return a * bc(s); }
RLS_FP_D F2 operator+(F2 a, float s) { return a + bc(s); }
RLS_FP_D F2 operator-(F2 a, float s) { return a - bc(s); }
RLS_FP_D F2 sqr(F2 a) { return a * a; }
RLS_FP_D F2 and_bits(F2 a, uint32_t m) { return mk(__uint_as_float(__float_as_uint(lo(a)) & m), __uint_as_float(__float_as_uint(hi(a)) & m)); }

} // namespace pk
} // namespace rls

namespace rls {
namespace pk {

RLS_DEV B2 operator!(B2 m) { B2 r; r.a = !m.a; r.b = !m.b; return r; }
RLS_DEV B2 operator&&(B2 m, B2 n) { B2 r; r.a = m.a && n.a; r.b = m.b && n.b; return r; }
RLS_DEV B2 operator||(B2 m, B2 n) { B2 r; r.a = m.a || n.a; r.b = m.b || n.b; return r; }
RLS_DEV B2 lt(F2 x, F2 y) { B2 r; r.a = lo(x) < lo(y); r.b = hi(x) < hi(y); return r; }
RLS_DEV B2 lt(F2 x, float y) { B2 r; r.a = lo(x) < y; r.b = hi(x) < y; return r; }
RLS_DEV B2 gt(F2 x, F2 y) { B2 r; r.a = lo(x) > lo(y); r.b = hi(x) > hi(y); return r; }
RLS_DEV B2 gt(F2 x, float y) { B2 r; r.a = lo(x) > y; r.b = hi(x) > y; return r; }
RLS_DEV B2 ne(F2 x, float y) { B2 r; r.a = lo(x) != y; r.b = hi(x) != y; return r; }
RLS_DEV B2 eq(F2 x, float y) { B2 r; r.a = lo(x) == y; r.b = hi(x) == y; return r; }
RLS_DEV B2 le(F2 x, float y) { B2 r; r.a = lo(x) <= y; r.b = hi(x) <= y; return r; }
RLS_DEV F2 sel(B2 m, F2 x, F2 y) { return mk(m.a ? lo(x) : lo(y), m.b ? hi(x) : hi(y)); }
RLS_DEV F2 sel(B2 m, float x, F2 y) { return mk(m.a ? x : lo(y), m.b ? x : hi(y)); }
RLS_DEV F2 sel(B2 m, F2 x, float y) { return mk(m.a ? lo(x) : y, m.b ? hi(x) : y); }
RLS_DEV F2 sel(B2 m, float x, float y) { return mk(m.a ? x : y, m.b ? x : y); }

// the reference's macros, lane-wise (rls_math.cuh)
RLS_DEV F2 abs_m(F2 a) { return mk(rls::abs_m(lo(a)), rls::abs_m(hi(a))); }
RLS_DEV F2 max_m(float a, F2 b) { return mk(rls::max_m(a, lo(b)), rls::max_m(a, hi(b))); }
RLS_DEV F2 max_m(F2 a, float b) { return mk(rls::max_m(lo(a), b), rls::max_m(hi(a), b)); }
RLS_DEV F2 clamp_m(F2 v, float l, float h) { return mk(rls::clamp_m(lo(v), l, h), rls::clamp_m(hi(v), l, h)); }
RLS_DEV F2 sgn_m(F2 a) { return mk(rls::sgn_m(lo(a)), rls::sgn_m(hi(a))); }

struct V2 { F2 x, y, z; };
RLS_DEV V2 mk3(F2 x, F2 y, F2 z) { V2 v; v.x = x; v.y = y; v.z = z; return v; }
RLS_DEV V2 operator+(V2 a, V2 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RLS_DEV V2 operator-(V2 a, V2 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RLS_DEV V2 operator-(V2 a) { return mk3(-a.x, -a.y, -a.z); }
RLS_DEV V2 operator*(V2 a, F2 s) { return mk3(a.x * s, a.y * s, a.z * s); }
RLS_DEV F2 dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RLS_DEV V2 sel(B2 m, V2 a, V2 b) { return mk3(sel(m, a.x, b.x), sel(m, a.y, b.y), sel(m, a.z, b.z)); }
RLS_DEV B2 is_zero(V2 a) { B2 r; r.a = lo(a.x) == 0.0f && lo(a.y) == 0.0f && lo(a.z) == 0.0f; r.b = hi(a.x) == 0.0f && hi(a.y) == 0.0f && hi(a.z) == 0.0f; return r; }
RLS_DEV V2 rotate_to_frame(V2 a, V2 u, V2 v, V2 w)
{
    return mk3(a.x * u.x + a.y * v.x + a.z * w.x, a.x * u.y + a.y * v.y + a.z * w.y, a.x * u.z + a.y * v.z + a.z * w.z);
}
RLS_DEV f3 lane0(V2 a) { return rls::mk3(lo(a.x), lo(a.y), lo(a.z)); }
RLS_DEV f3 lane1(V2 a) { return rls::mk3(hi(a.x), hi(a.y), hi(a.z)); }

// ------------------------------------------------------------ the fast policy, two lanes
// Same sequences and the same window as FpFast (rls_fp.cuh); ONE tracker for both lanes.
struct Fp2 {
    float lo_, hi_;
    uint32_t ilo_;
    RLS_DEV Fp2() : lo_(1.0f), hi_(1.0f), ilo_(0xffffffffu) {}
    // hand-over of the tracker state from / to the scalar policy of the same sample
    RLS_DEV explicit Fp2(const FpFast &f) : lo_(f.lo), hi_(f.hi), ilo_(f.ilo) {}
    RLS_DEV void give(FpFast &f) const { f.lo = lo_; f.hi = hi_; f.ilo = ilo_; }
    static RLS_DEV F2 mufu_rcp(F2 x) { return mk(FpFast::mufu_rcp(lo(x)), FpFast::mufu_rcp(hi(x))); }
    static RLS_DEV F2 mufu_rsq(F2 x) { return mk(FpFast::mufu_rsq(lo(x)), FpFast::mufu_rsq(hi(x))); }
    static RLS_DEV F2 rcp_refined(F2 b, F2 nb)
    {
        F2 y = mufu_rcp(b);
        F2 e = fma2(y, nb, bc(1.0f));
        return fma2(y, e, y);
    }
    RLS_DEV void track_abs(F2 a, F2 b)
    {
        lo_ = fminf(fminf(lo_, fabsf(lo(a))), fabsf(lo(b))); lo_ = fminf(fminf(lo_, fabsf(hi(a))), fabsf(hi(b)));
        hi_ = fmaxf(fmaxf(hi_, fabsf(lo(a))), fabsf(lo(b))); hi_ = fmaxf(fmaxf(hi_, fabsf(hi(a))), fabsf(hi(b)));
    }
    RLS_DEV void track_zero_ok(F2 a)
    {
        ilo_ = min(min(ilo_, (__float_as_uint(lo(a)) & 0x7fffffffu) - 1u), (__float_as_uint(hi(a)) & 0x7fffffffu) - 1u);
    }
    RLS_DEV F2 div(F2 a, F2 b)
    {
        F2 nb = -b;
        F2 y = rcp_refined(b, nb);
        F2 q = fma2(a, y, bc(0.0f));
        F2 r = fma2(q, nb, a);
        track_abs(a, b);
        return fma2(y, r, q);
    }
    // a / b for a scalar divisor b shared by both lanes, y = FpFast::shared_rcp(b) (b tracked there)
    RLS_DEV F2 div_by(F2 a, float b, float y)
    {
        F2 q = fma2(a, bc(y), bc(0.0f));
        F2 r = fma2(q, bc(-b), a);
        lo_ = fminf(fminf(lo_, fabsf(lo(a))), fabsf(hi(a)));
        hi_ = fmaxf(fmaxf(hi_, fabsf(lo(a))), fabsf(hi(a)));
        return fma2(bc(y), r, q);
    }
    RLS_DEV F2 div_z(F2 a, F2 b)                 // zero numerator allowed, sign of a zero quotient unspecified
    {
        F2 nb = -b;
        F2 y = rcp_refined(b, nb);
        F2 q = fma2(a, y, bc(0.0f));
        F2 r = fma2(q, nb, a);
        lo_ = fminf(fminf(lo_, fabsf(lo(b))), fabsf(hi(b)));
        hi_ = fmaxf(fmaxf(hi_, fabsf(lo(a))), fabsf(lo(b))); hi_ = fmaxf(fmaxf(hi_, fabsf(hi(a))), fabsf(hi(b)));
        track_zero_ok(a);
        return fma2(y, r, q);
    }
    RLS_DEV F2 div_pz(F2 a, F2 b)                // b > 0; zero numerator keeps its IEEE sign
    {
        F2 nb = -b;
        F2 y = rcp_refined(b, nb);
        F2 q = a * y;
        F2 r = fma2_rd(q, nb, a);
        lo_ = fminf(fminf(lo_, lo(b)), hi(b));
        hi_ = fmaxf(fmaxf(hi_, fabsf(lo(a))), lo(b)); hi_ = fmaxf(fmaxf(hi_, fabsf(hi(a))), hi(b));
        track_zero_ok(a);
        return fma2(y, r, q);
    }
    RLS_DEV F2 rcp(F2 x)
    {
        F2 y = mufu_rcp(x);
        F2 t = fma2(x, y, bc(-1.0f));
        lo_ = fminf(fminf(lo_, fabsf(lo(x))), fabsf(lo(y))); lo_ = fminf(fminf(lo_, fabsf(hi(x))), fabsf(hi(y)));
        return fma2(y, -t, y);
    }
    RLS_DEV F2 rcp_in_window(F2 x)
    {
        F2 y = mufu_rcp(x);
        F2 t = fma2(x, y, bc(-1.0f));
        return fma2(y, -t, y);
    }
    RLS_DEV F2 sqrt(F2 x)
    {
        F2 y = mufu_rsq(x);
        F2 g = x * y, h = y * 0.5f;
        F2 r = fma2(-g, g, x);
        lo_ = fminf(fminf(lo_, lo(x)), lo(y)); lo_ = fminf(fminf(lo_, hi(x)), hi(y));
        return fma2(r, h, g);
    }
    RLS_DEV void require(B2 c) { lo_ = (c.a && c.b) ? lo_ : 0.0f; }
    RLS_DEV bool ok() const { return lo_ >= 0x1p-60f && hi_ <= 0x1p60f && ilo_ >= 0x217fffffu; }
};

RLS_DEV V2 normalize(Fp2 &fp, V2 a)
{
    F2 len = fp.sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    F2 inv = fp.rcp_in_window(len);
    return mk3(a.x * inv, a.y * inv, a.z * inv);
}

} // namespace pk
} // namespace rls
