// rls_packed.cuh -- the rlGgx path with TWO samples per thread and every binary32 add / multiply /
// fma issued as ONE packed f32x2 instruction (FADD2 / FFMA2 on sm_100).
//
// Why: the fused kernels are instruction-issue bound with the FMA pipe half idle
// (profiles/r01_ncu_summary.md); FP32 arithmetic is 52 % of the issue slots because bit-exactness
// forbids FMA contraction.  A packed instruction costs one issue slot for two lanes (and two FMA
// pipe cycles: profiles/r01_ffma2_microbench.txt), so pairing samples removes a third of the slots.
//
// Bit-exactness: a packed lane operation is the same IEEE-754 operation as the scalar one.  The
// one trap is ptxas 12.9 contracting mul.rn.f32x2 -> add.rn.f32x2 into a fused FFMA2 even with
// --fmad=false (tools/microbench/f32x2_exact_test.cu).  Every multiply is therefore spelled
// fma.rn.f32x2(a, b, NZ) with NZ = {-0, -0} read from a __constant__ the HOST fills at rls_init:
// x*y + (-0) is x*y for every x*y (signed zeros included) and ptxas cannot fold a value it does
// not know, nor contract an fma into the add that follows.
//
// This header mirrors rls_ggx.cuh / rls_fused.cuh statement by statement for the shipped sampler
// (visible normals) under the FAST arithmetic policy only: control flow is select-based (both
// lanes of a thread share one instruction stream), masked-off lanes are fed in-window dummy
// operands, and every condition the stream does not carry is folded into the operand tracker;
// the kernel re-runs flagged pairs with the scalar FpExact code.  tests/test_gpu_parity.py checks
// packed == scalar == oracle, bit for bit.
#pragma once
#include "rls_pair.cuh"

namespace rls {
namespace pk {

// ------------------------------------------------------------------ libm, two lanes
RLS_DEV void sincosf2(F2 a, F2 &s, F2 &c)
{
    float s0, c0, s1, c1;
    rlm::sincosf_(lo(a), &s0, &c0);
    rlm::sincosf_(hi(a), &s1, &c1);
    s = mk(s0, s1); c = mk(c0, c1);
}

// rlm::atanf_nonneg_ (fast policy), lane-wise regimes, packed arithmetic
RLS_DEV F2 atanf_nonneg2(Fp2 &fp, F2 ax)
{
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
                aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
                aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
                aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const uint32_t i0 = __float_as_uint(lo(ax)), i1 = __float_as_uint(hi(ax));
    B2 ok; ok.a = i0 < 0x4c000000u; ok.b = i1 < 0x4c000000u;
    fp.require(ok);
    // candidates of every regime, packed
    const F2 n1 = ax * 2.0f - 1.0f, n2 = ax - 1.0f, n3 = ax - 1.5f;
    const F2 d1 = bc(2.0f) + ax, d2 = ax + 1.0f, d3 = bc(1.0f) + ax * 1.5f;
    float num[2], den[2], hiv[2], lov[2];
    bool r0[2], tiny[2];
#pragma unroll
    for (int l = 0; l < 2; l++) {
        const int32_t ix = (int32_t)(l ? i1 : i0);
        const float x = l ? hi(ax) : lo(ax);
        const bool a0 = ix < 0x3ee00000, a1 = ix < 0x3f300000, a2 = ix < 0x3f980000, a3 = ix < 0x401c0000;
        r0[l] = a0; tiny[l] = ix < 0x31000000;
        num[l] = a0 ? x : (a1 ? (l ? hi(n1) : lo(n1)) : (a2 ? (l ? hi(n2) : lo(n2)) : (a3 ? (l ? hi(n3) : lo(n3)) : -1.0f)));
        den[l] = a0 ? 1.0f : (a1 ? (l ? hi(d1) : lo(d1)) : (a2 ? (l ? hi(d2) : lo(d2)) : (a3 ? (l ? hi(d3) : lo(d3)) : x)));
        hiv[l] = a1 ? 4.6364760399e-01f : (a2 ? 7.8539812565e-01f : (a3 ? 9.8279368877e-01f : 1.5707962513e+00f));
        lov[l] = a1 ? 5.0121582440e-09f : (a2 ? 3.7748947079e-08f : (a3 ? 3.4473217170e-08f : 7.5497894159e-08f));
    }
    F2 t = fp.div_pz(mk(num[0], num[1]), mk(den[0], den[1]));
    F2 z = t * t;
    F2 w = z * z;
    F2 s1 = z * (bc(aT0) + w * (bc(aT2) + w * (bc(aT4) + w * (bc(aT6) + w * (bc(aT8) + w * aT10)))));
    F2 s2 = w * (bc(aT1) + w * (bc(aT3) + w * (bc(aT5) + w * (bc(aT7) + w * aT9))));
    F2 ts = t * (s1 + s2);
    F2 small = t - ts;
    F2 big = mk(hiv[0], hiv[1]) - ((ts - mk(lov[0], lov[1])) - t);
    return mk(r0[0] ? (tiny[0] ? lo(ax) : lo(small)) : lo(big), r0[1] ? (tiny[1] ? hi(ax) : hi(small)) : hi(big));
}
// rlm::atan2f_ fast path
RLS_DEV F2 atan2f2(Fp2 &fp, F2 y, F2 x)
{
    const float pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    F2 q = fp.div_z(y, x);
    F2 z0 = atanf_nonneg2(fp, and_bits(q, 0x7fffffffu));
    F2 alt = bc(pi) - (z0 - pi_lo);
    B2 xneg; xneg.a = (int32_t)__float_as_uint(lo(x)) < 0; xneg.b = (int32_t)__float_as_uint(hi(x)) < 0;
    F2 w0 = sel(xneg, alt, z0);
    return mk(__uint_as_float(__float_as_uint(lo(w0)) | (__float_as_uint(lo(y)) & 0x80000000u)),
              __uint_as_float(__float_as_uint(hi(w0)) | (__float_as_uint(hi(y)) & 0x80000000u)));
}
// rlm::acosf_ for |x| < 1, branch-free (all three regimes evaluated; inactive lanes get x = 0.75)
RLS_DEV F2 acosf2(Fp2 &fp, B2 active, F2 xin)
{
    const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f,
                pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f,
                qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f,
                qS4 = 7.7038154006e-02f;
    const F2 x = sel(active, xin, 0.75f);
    const int32_t h0 = (int32_t)__float_as_uint(lo(x)), h1 = (int32_t)__float_as_uint(hi(x));
    const int32_t i0 = h0 & 0x7fffffff, i1 = h1 & 0x7fffffff;
    B2 dom; dom.a = i0 < 0x3f800000 && i0 > 0x32800000; dom.b = i1 < 0x3f800000 && i1 > 0x32800000;
    fp.require(dom);                             // |x| >= 1 and |x| <= 2^-26 take special returns in the original
    B2 small; small.a = i0 < 0x3f000000; small.b = i1 < 0x3f000000;
    B2 negx; negx.a = h0 < 0; negx.b = h1 < 0;
    const F2 ax = and_bits(x, 0x7fffffffu);
    const F2 z = sel(small, x * x, (bc(1.0f) - ax) * 0.5f);
    const F2 p = z * (bc(pS0) + z * (bc(pS1) + z * (bc(pS2) + z * (bc(pS3) + z * (bc(pS4) + z * pS5)))));
    const F2 q = bc(1.0f) + z * (bc(qS1) + z * (bc(qS2) + z * (bc(qS3) + z * qS4)));
    const F2 r = fp.div_pz(p, q);
    const F2 res_small = bc(pio2_hi) - (x - (bc(pio2_lo) - x * r));
    const F2 zs = sel(small, 0.25f, z);          // in-window dummy for the lanes that do not take a root
    const F2 s = fp.sqrt(zs);
    const F2 res_neg = bc(pi) - (s + (r * s - pio2_lo)) * 2.0f;
    const F2 df = and_bits(s, 0xfffff000u);
    const F2 c = fp.div_pz(zs - df * df, s + df);
    const F2 res_pos = (df + (r * s + c)) * 2.0f;
    return sel(small, res_small, sel(negx, res_neg, res_pos));
}
// rlm::tanf_ for |x| < 120: binary64 reduction per lane, fdlibm kernel packed
RLS_DEV F2 tanf2(Fp2 &fp, F2 xin)
{
    const float pio4 = 7.8539812565e-01f, pio4lo = 3.7748947079e-08f;
    const float T0 = 3.3333334327e-01f, T1 = 1.3333334029e-01f, T2 = 5.3968254477e-02f,
                T3 = 2.1869488060e-02f, T4 = 8.8632395491e-03f, T5 = 3.5920790397e-03f,
                T6 = 1.4562094584e-03f, T7 = 5.8804126456e-04f, T8 = 2.4646313977e-04f,
                T9 = 7.8179444245e-05f, T10 = 7.1407252108e-05f, T11 = -1.8558637748e-05f,
                T12 = 2.5907305826e-05f;
    float y0[2], y1[2], fiy[2], sgn[2];
    bool big[2], iy1[2], neg[2];
#pragma unroll
    for (int l = 0; l < 2; l++) {
        const float xf = l ? hi(xin) : lo(xin);
        double dx = (double)xf;
        double r = dx * rlm::kSinCosC[7];
        int n = ((int32_t)r + 0x800000) >> 24;
        dx = dx - (double)n * rlm::kSinCosC[8];
        y0[l] = (float)dx;
        y1[l] = (float)(dx - (double)y0[l]);
        const int iy = 1 - ((n & 1) << 1);
        const int32_t hx = (int32_t)__float_as_uint(y0[l]);
        big[l] = (hx & 0x7fffffff) >= 0x3f2ca140;
        neg[l] = hx < 0;
        sgn[l] = (float)(1 - ((hx >> 30) & 2));
        fiy[l] = (float)iy;
        iy1[l] = iy == 1;
    }
    B2 range = lt(and_bits(xin, 0x7fffffffu), 120.0f);
    fp.require(range);
    B2 bigm; bigm.a = big[0]; bigm.b = big[1];
    B2 negm; negm.a = neg[0]; negm.b = neg[1];
    F2 x = mk(y0[0], y0[1]), y = mk(y1[0], y1[1]);
    const F2 FIY = mk(fiy[0], fiy[1]), SGN = mk(sgn[0], sgn[1]);
    {
        F2 xa = sel(negm, -x, x), ya = sel(negm, -y, y);
        F2 xb = (bc(pio4) - xa) + (bc(pio4lo) - ya);
        x = sel(bigm, xb, x);
        y = sel(bigm, 0.0f, y);
    }
    // |x| < 2^-13 (tiny reduced argument, or pi/4 - |x| tiny) takes special returns in the original
    fp.require(!lt(and_bits(x, 0x7fffffffu), 0x1p-13f));
    F2 z = x * x;
    F2 w = z * z;
    F2 r = bc(T1) + w * (bc(T3) + w * (bc(T5) + w * (bc(T7) + w * (bc(T9) + w * T11))));
    F2 v = z * (bc(T2) + w * (bc(T4) + w * (bc(T6) + w * (bc(T8) + w * (bc(T10) + w * T12)))));
    F2 s = z * x;
    r = y + z * (s * (r + v) + y);
    r = r + s * T0;
    w = x + r;
    F2 q = fp.div(sel(bigm, w * w, -1.0f), sel(bigm, w + FIY, w));
    F2 res_big = SGN * (FIY - (x - (q - r)) * 2.0f);
    F2 zt = and_bits(w, 0xfffff000u);
    F2 vt = r - (zt - x);
    F2 t = and_bits(q, 0xfffff000u);
    F2 st = bc(1.0f) + t * zt;
    F2 res_inv = t + q * (st + t * vt);
    B2 iym; iym.a = iy1[0]; iym.b = iy1[1];
    return sel(bigm, res_big, sel(iym, w, res_inv));
}

// --------------------------------------------------------------------------- rlGgx, two lanes
struct Ggx2 {
    V2 U, V, N, wo;
    F2 iorIn, iorOut, rough, ax, ay;
    B2 entering;
};
// rls::ggx_init
RLS_DEV void ggx_init(Fp2 &fp, Ggx2 &g, V2 U, V2 V, V2 N, V2 wo, B2 backfacing, F2 ior, F2 roughness, F2 aniso)
{
    V2 Ngeo = sel(backfacing, -N, N);
    V2 Rd = -wo;
    g.entering = lt(dot(Ngeo, Rd), kEps);
    F2 b = max_m(ior, 1e-4f);
    g.iorIn = sel(g.entering, 1.0f, b);
    g.iorOut = sel(g.entering, b, 1.0f);
    g.wo = wo; g.U = U; g.V = V; g.N = N;
    F2 aspect = fp.sqrt(bc(1.0f) - aniso * 0.9f);
    g.ax = max_m(1e-4f, fp.div_pz(sqr(roughness), aspect));
    g.ay = max_m(1e-4f, sqr(roughness) * aspect);
    g.rough = max_m(1e-5f, sqr(roughness));
}
RLS_DEV F2 ggx_G1_value2(Fp2 &fp, const Ggx2 &g, F2 VdotN)
{
    F2 cosSqr = sqr(VdotN);
    F2 tanSqr = fp.rcp(cosSqr) - 1.0f;
    F2 denominator = bc(1.0f) + fp.sqrt(bc(1.0f) + sqr(g.rough) * tanSqr);
    return fp.rcp_in_window(denominator) * 2.0f;
}
RLS_DEV F2 ggx_D(Fp2 &fp, const Ggx2 &g, V2 m)
{
    F2 MdotU = dot(m, g.U);
    F2 MdotV = dot(m, g.V);
    F2 MdotN2 = sqr(dot(g.N, m));
    F2 denominator = g.ax * g.ay * sqr(sqr(fp.div(MdotU, g.ax)) + sqr(fp.div(MdotV, g.ay)) + MdotN2);
    return fp.div(bc(kInvPi), denominator);
}
struct GgxShared2 { F2 VdotN, absVdotN, sgnV, G1v, ratio2; };
RLS_DEV GgxShared2 ggx_shared(Fp2 &fp, const Ggx2 &g)
{
    GgxShared2 s;
    s.VdotN = dot(g.wo, g.N);
    s.absVdotN = abs_m(s.VdotN);
    s.sgnV = sgn_m(s.VdotN);
    s.G1v = ggx_G1_value2(fp, g, s.VdotN);
    s.ratio2 = sqr(fp.div(g.iorOut, g.iorIn));
    return s;
}
// rls::ggx_reflect_eval_pdf (VNDFKernel)
RLS_DEV void ggx_reflect_eval_pdf(Fp2 &fp, const Ggx2 &g, const GgxShared2 &s, V2 L, F2 LdotN, F2 G1l, F2 &refl, F2 &pdf)
{
    V2 H = normalize(fp, L + g.wo);
    F2 VH = dot(g.wo, H);
    F2 LH = dot(L, H);
    F2 D_H = ggx_D(fp, g, H);
    F2 G1_pdf = sel(lt(VH * s.VdotN, 0.0f), 0.0f, s.G1v);
    pdf = max_m(fp.div_pz(D_H * G1_pdf, s.absVdotN) * 0.25f, kEps);
    F2 VHr = VH * s.sgnV, LHr = LH * s.sgnV;
    F2 F = ggx_fresnel_c(fp, s.ratio2, abs_m(VHr));
    F2 G1i = sel(lt(VHr * s.VdotN, 0.0f), 0.0f, s.G1v);
    F2 G1o = sel(lt(LHr * LdotN, 0.0f), 0.0f, G1l);
    fp.require(ne(s.sgnV, 0.0f));                                 // D(0 vector) = 1/0 in the scalar code
    refl = fp.div_pz(F * (G1i * G1o) * D_H * 0.25f, abs_m(LdotN) * s.absVdotN);
}

// sample_slope for the lanes in `active` (the others run on theta = 1)
RLS_DEV void sample_slope2(Fp2 &fp, B2 active, F2 theta_in, F2 rx, F2 ry, F2 &sx, F2 &sy)
{
    const F2 theta = sel(active, theta_in, 1.0f);
    F2 B = tanf2(fp, theta);
    F2 Bsq = sqr(B);
    F2 G1 = fp.rcp_in_window(bc(1.0f) + fp.sqrt(bc(1.0f) + Bsq)) * 2.0f;
    F2 A = fp.div(rx * 2.0f, G1) - 1.0f;
    F2 A2 = sqr(A);
    fp.require(!(active && lt(abs_m(A2 - 1.0f), kEps)));          // uniform-slope fallback of :38 not carried
    F2 tmp = fp.rcp(A2 - 1.0f);
    F2 D = fp.sqrt(max_m(0.0f, Bsq * sqr(tmp) - (A2 - Bsq) * tmp));
    F2 slopeX1 = B * tmp - D;
    F2 slopeX2 = B * tmp + D;
    sx = sel(lt(A, 0.0f) || gt(slopeX2, fp.rcp(B)), slopeX1, slopeX2);
    B2 up = gt(ry, 0.5f);
    F2 sign = sel(up, 1.0f, -1.0f);
    F2 u = sel(up, (ry - 0.5f) * 2.0f, (bc(0.5f) - ry) * 2.0f);
    F2 z = fp.div(u * (u * (u * 0.27385f - 0.73369f) + 0.46341f),
                  u * (u * (u * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    sy = sign * z * fp.sqrt(bc(1.0f) + sqr(sx));
}
// rls::sample_visible_normal
RLS_DEV V2 sample_visible_normal(Fp2 &fp, const Ggx2 &g, F2 rx, F2 ry)
{
    F2 cosThetaV = clamp_m(dot(g.N, g.wo), -1.0f, 1.0f);
    F2 phiV = atan2f2(fp, dot(g.V, g.wo), dot(g.U, g.wo));
    F2 s, c;
    sincosf2(phiV, s, c);
    F2 r = fp.sqrt(bc(1.0f) - sqr(cosThetaV));
    V2 V = normalize(fp, mk3(r * c * g.ax, r * s * g.ay, cosThetaV));
    const B2 along = !lt(V.z, 1.0f - kEps);
    const B2 act = !along;
    F2 theta = sel(along, 0.0f, acosf2(fp, act, V.z));
    F2 phi = sel(along, 0.0f, atan2f2(fp, sel(act, V.y, 1.0f), sel(act, V.x, 1.0f)));
    sincosf2(sel(along, ry * kTwoPi, phi), s, c);
    // uniform slope (all lanes; its operands are in-window for every rx) and the main path
    F2 ru = fp.sqrt(fp.div(rx, bc(1.0f) - rx));
    F2 mx, my;
    sample_slope2(fp, act, theta, rx, ry, mx, my);
    F2 slx = sel(along, ru * c, mx), sly = sel(along, ru * s, my);
    F2 sinPhi = sel(along, 0.0f, s), cosPhi = sel(along, 1.0f, c);
    V2 omega;
    omega.x = -(cosPhi * slx - sinPhi * sly) * g.ax;
    omega.y = -(sinPhi * slx + cosPhi * sly) * g.ay;
    omega.z = bc(1.0f);
    return normalize(fp, rotate_to_frame(omega, g.U, g.V, g.N));
}

struct Dielectric2 { F2 F, f_r, pdf_r, f_t, w_t; V2 wi_r, wi_t; uint32_t flags0, flags1; };
// rls::dielectric_unit (VNDF sampler), both lanes
RLS_DEV Dielectric2 dielectric_unit(Fp2 &fp, V2 U, V2 V, V2 N, V2 wo, B2 backfacing, F2 ior, F2 rough, F2 aniso, F2 rx, F2 ry)
{
    Dielectric2 r;
    Ggx2 g;
    ggx_init(fp, g, U, V, N, wo, backfacing, ior, rough, aniso);
    const GgxShared2 s = ggx_shared(fp, g);
    V2 m = sample_visible_normal(fp, g, rx, ry);
    F2 Vm = dot(g.wo, m);
    r.wi_r = m * (abs_m(Vm) * 2.0f) - g.wo;
    r.F = ggx_fresnel_c(fp, s.ratio2, abs_m(dot(r.wi_r, m)));
    const V2 L = r.wi_r;
    const F2 LdotN = dot(L, g.N);
    const F2 G1l = ggx_G1_value2(fp, g, LdotN);
    F2 refl;
    ggx_reflect_eval_pdf(fp, g, s, L, LdotN, G1l, refl, r.pdf_r);
    const B2 zeroL = is_zero(L);
    r.f_r = sel(zeroL, 0.0f, refl * LdotN);

    const F2 eta = fp.div(g.iorIn, g.iorOut);
    const F2 cosThetaTSqr = bc(1.0f) + eta * (sqr(Vm) - 1.0f);
    const F2 mN = dot(m, g.N);
    const B2 tir = lt(cosThetaTSqr, 0.0f);
    // refracted branch for every lane (TIR lanes run it on the in-window dummy 1)
    F2 sc = eta * Vm - s.sgnV * fp.sqrt(sel(tir, 1.0f, cosThetaTSqr));
    V2 T = m * sc - g.wo * eta;
    F2 TdotN_t = dot(T, g.N);
    F2 G1t_t = ggx_G1_value2(fp, g, TdotN_t);
    V2 ht = -normalize(fp, g.wo * g.iorIn + T * g.iorOut);
    F2 IdotH = dot(g.wo, ht);
    F2 OdotH = dot(T, ht);
    F2 refractWeight = bc(1.0f) - ggx_fresnel_c(fp, s.ratio2, abs_m(IdotH));
    F2 denominator = abs_m(TdotN_t) * s.absVdotN * sqr(g.iorIn * IdotH + g.iorOut * OdotH);
    F2 G1i_t = sel(lt(IdotH * s.VdotN, 0.0f), 0.0f, s.G1v);
    F2 G1o_t = sel(lt(OdotH * TdotN_t, 0.0f), 0.0f, G1t_t);
    F2 f_t = fp.div_pz(abs_m(OdotH * IdotH) * sqr(g.iorOut) * refractWeight * (G1i_t * G1o_t) * ggx_D(fp, g, ht), denominator);
    r.wi_t = sel(tir, r.wi_r, T);
    r.f_t = sel(tir, 0.0f, f_t);
    const F2 TdotN = sel(tir, LdotN, TdotN_t);
    const F2 G1t = sel(tir, G1l, G1t_t);
    {
        F2 G1i = sel(lt(Vm * s.VdotN, 0.0f), 0.0f, s.G1v);
        F2 G1o = sel(lt(dot(r.wi_t, m) * TdotN, 0.0f), 0.0f, G1t);
        r.w_t = (G1i * G1o) * abs_m(fp.div(Vm, s.absVdotN * abs_m(mN)));
    }
#pragma unroll
    for (int l = 0; l < 2; l++) {
        const float pr = l ? hi(r.pdf_r) : lo(r.pdf_r), fr = l ? hi(r.f_r) : lo(r.f_r), ln = l ? hi(LdotN) : lo(LdotN);
        uint32_t fl = 0;
        if (l ? zeroL.b : zeroL.a) fl |= 0x0001u;
        if (ln <= 0.0f) fl |= 0x0002u;
        if (pr == 0.0f) fl |= 0x0004u;
        if (fr == 0.0f) fl |= 0x0008u;
        if (pr == kEps) fl |= 0x0040u;
        if (l ? g.entering.b : g.entering.a) fl |= 0x0010u;
        if (l ? tir.b : tir.a) fl |= 0x0020u;
        if (l) r.flags1 = fl; else r.flags0 = fl;
    }
    return r;
}

// ------------------------------------------------------------------- pair loads / stores
// Thread t owns samples 2t and 2t + 1; `tail` marks the last thread of an odd batch, which
// owns one sample only (its second lane duplicates the first and is not stored).
RLS_DEV F2 ld2(const float *p, uint32_t t, bool tail)
{
    if (tail) { float v = __ldg(p + 2u * t); return mk(v, v); }
    float2 v = __ldg(reinterpret_cast<const float2 *>(p) + t);
    return mk(v.x, v.y);
}
RLS_DEV V2 ld2(const CV3 &v, uint32_t t, bool tail) { return mk3(ld2(v.x, t, tail), ld2(v.y, t, tail), ld2(v.z, t, tail)); }
RLS_DEV F2 fetch2(const P1 &p, uint32_t t, bool tail) { return p.array ? ld2(p.array, t, tail) : bc(p.value); }
RLS_DEV void st2(float *p, uint32_t t, bool tail, F2 v)
{
    if (tail) p[2u * t] = lo(v);
    else reinterpret_cast<float2 *>(p)[t] = make_float2(lo(v), hi(v));
}
RLS_DEV void st2(const V3 &o, uint32_t t, bool tail, V2 v) { st2(o.x, t, tail, v.x); st2(o.y, t, tail, v.y); st2(o.z, t, tail, v.z); }

} // namespace pk
} // namespace rls
