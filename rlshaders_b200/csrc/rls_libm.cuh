// rls_libm.cuh -- binary32 transcendentals that reproduce the HOST C library bit for bit.
//
// Why: the reference's hot path calls sincosf, atan2f, acosf, tanf, powf, logf, expf from the
// host libm (glibc 2.39 on the build image; see SURVEY.md 8(c) "third-party arithmetic").
// Visible-normal sampling is ill-conditioned: a 1-ulp difference in one of these results moves
// the sampled direction by more than the 1e-6 tolerance in a few percent of samples, and it can
// flip a discrete decision (TIR, lobe, horizon).  CUDA's libdevice versions are 1-4 ulp off,
// and even a correctly rounded result differs from the host library in 1-25 % of calls
// (tanf/atan2f/acosf there are binary32 fdlibm ports with < 1 ulp error, logf/powf 0.82 ulp).
// So the kernels carry their own implementations of the PUBLISHED algorithms the host library
// ships, operation for operation:
//
//   tanf, atanf, atan2f, acosf : FreeBSD/Sun fdlibm binary32 ports (k_tanf.c, s_atanf.c,
//                                e_atan2f.c, e_acosf.c; "Copyright (C) 1993 by Sun Microsystems,
//                                Inc. ... Permission to use, copy, modify, and distribute this
//                                software is freely granted, provided that this notice is
//                                preserved."), with glibc 2.39's double-precision argument
//                                reduction for tanf.
//   sincosf, expf, logf, powf  : ARM Optimized Routines (Szabolcs Nagy, MIT licence) as adopted
//                                by glibc >= 2.27/2.28: binary64 polynomial evaluation, one final
//                                rounding.  The x86-64 host selects the FMA build of these
//                                (ifunc), so every a*b+c there is fused; fma() is spelled out
//                                here to match (DFMA on the device is IEEE-exact).
//
// Everything else is plain IEEE binary32/binary64 arithmetic; the translation unit is compiled
// with -fmad=false so nothing is contracted behind our back.  Coefficients and tables were
// checked against the host library's .rodata, and tests/test_libm_compat.py runs these very
// functions on the CPU (they are __host__ __device__) against the host libm over all 2^32
// arguments of the univariate functions and ~10^9 argument pairs of the bivariate ones.
//
// Domain notes: only finite arguments occur on the path; NaN/Inf fall through the generic
// code and produce NaN/Inf-like results without the host's errno/fenv side effects.
#pragma once
#include <stdint.h>
#include <string.h>
#include "rls_fp.cuh"

#if defined(__CUDACC__)
#define RLM_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define RLM_HD static inline
#endif

namespace rlm {

// ------------------------------------------------------------------ bit casts
RLM_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RLM_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
RLM_HD uint64_t d2u(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
RLM_HD double u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
RLM_HD double fma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
RLM_HD float fabsf_(float x) { return u2f(f2u(x) & 0x7fffffffu); }
RLM_HD float sqrtf_(float x)
{
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}
RLM_HD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ffu; }

// ------------------------------------------------------------------ tables
// Read through the read-only data path: a warp's divergent lookups into these 128-256 byte
// tables hit at most two L1 lines.
#if defined(__CUDA_ARCH__)
#define RLM_TABLE static __device__ const
#define RLM_LD(p) __ldg(&(p))
#else
#define RLM_TABLE static const
#define RLM_LD(p) (p)
#endif

// 2^(i/32), i = 0..31, stored as asuint64(2^(i/32)) - (i << 47)   (__exp2f_data.tab)
RLM_TABLE uint64_t kExp2Tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
// {1/c, log(c)} for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0)   (__logf_data.tab)
RLM_TABLE double kLogTab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0, 0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2, 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2,
};
// {1/c, log2(c)}   (__powf_log2_data.tab)
RLM_TABLE double kLog2Tab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,
    0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,
    0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,
    0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1p+0, 0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3,
    0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,
    0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2,
};

// Binary64 coefficients live in the constant bank on the device: DFMA / DMUL take a c[bank][offset]
// operand directly, whereas a literal costs two 32-bit moves into a register pair per use.
#if defined(__CUDA_ARCH__)
#define RLM_COEF static __constant__ double
#else
#define RLM_COEF static const double
#endif
RLM_COEF kExpC[5] = { 0x1.71547652b82fep+5, 0x1.8p+52,                       // InvLn2N, Shift   (__exp2f_data, N = 32)
                      0x1.c6af84b912394p-20, 0x1.ebfce50fac4f3p-13, 0x1.62e42ff0c52d6p-6 };   // C0, C1, C2
RLM_COEF kExp2C[4] = { 0x1.8p+47,                                                // ShiftScaled (powf's exp2_inline)
                       0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1 };
RLM_COEF kSinCosC[9] = { -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10,    // c1..c4
                         0x1.99343027bf8c3p-16,
                         -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13,    // s1..s3
                         0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0 };                           // 2/pi * 2^24, pi/2
RLM_COEF kLogC[4] = { 0x1.62e42fefa39efp-1, -0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2 };  // Ln2, A0..A2
RLM_COEF kLog2C[5] = { 0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2,
                       -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp+0 };                            // A0..A4 (__powf_log2_data)

// 4/pi in 32-bit words, for the |x| >= 120 argument reduction   (__inv_pio4)
RLM_TABLE uint32_t kInvPio4[24] = {
    0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
    0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu,
    0xf534ddc0u, 0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u,
    0x993c4390u, 0x3c439041u,
};

// The exp2 and log tables in SHARED memory, for kernels that call rlm::smem_tables_init() first
// (policy types with kSmemTables): a lookup is SHL + LOP3 + LDS at an immediate base instead of
// SHL + LOP3 + 64-bit add pair + LDG (15 lookups per skin-profile sample).
#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t *smem_tables()
{
    __shared__ uint64_t t[96];                 // [0, 32) kExp2Tab, [32, 64) kLogTab, [64, 96) kLog2Tab (bit patterns)
    return t;
}
__device__ __forceinline__ void smem_tables_init()          // every thread of the CTA, before any early exit
{
    uint64_t *t = smem_tables();
    if (threadIdx.x < 32) t[threadIdx.x] = kExp2Tab[threadIdx.x];
    else if (threadIdx.x < 64) t[threadIdx.x] = d2u(kLogTab[threadIdx.x - 32]);
    else if (threadIdx.x < 96) t[threadIdx.x] = d2u(kLog2Tab[threadIdx.x - 64]);
    __syncthreads();
}
#endif
template <bool kSmem> RLM_HD uint64_t exp2_tab(uint32_t i)
{
#if defined(__CUDA_ARCH__)
    if (kSmem) return smem_tables()[i];
#endif
    return RLM_LD(kExp2Tab[i]);
}
template <bool kSmem> RLM_HD double log2_tab(int i)
{
#if defined(__CUDA_ARCH__)
    if (kSmem) return u2d(smem_tables()[64 + i]);
#endif
    return RLM_LD(kLog2Tab[i]);
}
template <bool kSmem> RLM_HD double log_tab(int i)
{
#if defined(__CUDA_ARCH__)
    if (kSmem) return u2d(smem_tables()[32 + i]);
#endif
    return RLM_LD(kLogTab[i]);
}

// ================================================================== sincosf
// glibc 2.39 sysdeps/ieee754/flt-32/s_sincosf.{c,h} (ARM Optimized Routines), FMA build.
// The cosine coefficients of quadrants 2,3 (__sincosf_table[1]) are the negated coefficients of
// quadrants 0,1: every fma/multiply below is odd in them under round-to-nearest, so the host's
// result there is exactly -(result with table[0]) -- evaluate once, flip the sign bit.  The
// cosine polynomial is >= 0.7 on the reduced range, so no signed-zero case arises.
RLM_HD void sincosf_poly(double x, double x2, bool neg_cos, int n, float *sinp, float *cosp)
{
    const double c0 = 0x1p0, c1 = kSinCosC[0], c2 = kSinCosC[1], c3 = kSinCosC[2], c4 = kSinCosC[3];
    const double s1 = kSinCosC[4], s2 = kSinCosC[5], s3 = kSinCosC[6];

    double x4 = x2 * x2;
    double x3 = x2 * x;
    double cc2 = fma_(x2, c4, c3);
    double ss1 = fma_(x2, s3, s2);
    double cc1 = fma_(x2, c1, c0);
    double x5 = x3 * x2;
    double x6 = x4 * x2;
    double s = fma_(x3, s1, x);
    double c = fma_(x4, c2, cc1);
    float sv = (float)fma_(x5, ss1, s);
    float cv = (float)fma_(x6, cc2, c);
    cv = u2f(f2u(cv) ^ (neg_cos ? 0x80000000u : 0u));
    // swap sin/cos result based on quadrant
    *sinp = (n & 1) ? cv : sv;
    *cosp = (n & 1) ? sv : cv;
}
// x * sign[q & 3] with sign = {1, -1, -1, 1}: a multiplication by +-1 is a sign-bit flip.
RLM_HD double sincosf_signed(double x, int q)
{
    return u2d(d2u(x) ^ ((uint64_t)((uint32_t)(q + 1) & 2u) << 62));
}
// Slow path for |y| >= 120: 192-bit 4/pi multiply (reduce_large).  Never taken on the path
// (arguments are angles in [-2pi, 2pi]); kept so the function is total.
RLM_HD double sincosf_reduce_large(uint32_t xi, int *np)
{
    const int j = (int)((xi >> 26) & 15u);
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    res0 = (uint64_t)(xi * RLM_LD(kInvPio4[j]));
    res1 = (uint64_t)xi * RLM_LD(kInvPio4[j + 4]);
    res2 = (uint64_t)xi * RLM_LD(kInvPio4[j + 8]);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * 0x1.921FB54442D18p-62;
}
// The |y| < 120 path alone, WITHOUT the host's |y| < 2^-12 shortcut (sin = y, cos = 1): the
// polynomial rounds to exactly those values there except for sin(-0) (tests/native/libm_check
// "sincoslean": equal to the host for every |y| < 120), so the fast policy carries neither the range
// branch nor the two selects of the shortcut.
RLM_HD void sincosf_main_(float y, float *sinp, float *cosp)
{
    double x = (double)y;
    double r = x * kSinCosC[7];            // 2/pi * 2^24
    int n = ((int32_t)r + 0x800000) >> 24;
    x = fma_(-(double)n, kSinCosC[8], x);
    float sv;
    sincosf_poly(sincosf_signed(x, n), x * x, (n & 2) != 0, n, &sv, cosp);
    *sinp = (y == 0.0f) ? y : sv;          // sin(-0) = -0: the one argument the polynomial gets wrong (+0)
}
RLM_HD void sincosf_(float y, float *sinp, float *cosp)
{
    double x = (double)y;
    // abstop12(y) < 0x42f  <=>  |y| < 120 (one FSETP with an |.| operand instead of shift/mask/compare)
    if (fabsf_(y) < 120.0f) {
        // For |y| < pi/4 the host takes a shortcut that is bit-identical to reduce_fast with
        // n = 0 (x - 0*hpi == x, sign 1), so one uniform path serves both -- no divergence.
        double r = x * kSinCosC[7];            // 2/pi * 2^24
        int n = ((int32_t)r + 0x800000) >> 24;
        x = fma_(-(double)n, kSinCosC[8], x);
        float sv, cv;
        sincosf_poly(sincosf_signed(x, n), x * x, (n & 2) != 0, n, &sv, &cv);
        const bool tiny = fabsf_(y) < 0x1p-12f;  // abstop12(y) < 0x398: sin = y, cos = 1
        *sinp = tiny ? y : sv;
        *cosp = tiny ? 1.0f : cv;
    } else if (abstop12(y) < 0x7f8u) {
        int n;
        uint32_t xi = f2u(y);
        int sign = (int)(xi >> 31);
        x = sincosf_reduce_large(xi, &n);
        int q = n + sign;
        sincosf_poly(sincosf_signed(x, q), x * x, (q & 2) != 0, n, sinp, cosp);
    } else {
        *sinp = *cosp = y - y;
    }
}
// Policy form: |y| >= 120, Inf and NaN go to the exact re-run.
template <class Fp>
RLM_HD void sincosf_(Fp &fp, float y, float *sinp, float *cosp)
{
    // Measured on B200: the branch-free form gains 2 % in the rlDisney unit (three calls) and LOSES
    // 1-3 % in the two rlGgx units, so it is a property of the policy type (FpFastLeanTrig).
    if (!(Fp::kFast && Fp::kLeanTrig)) { sincosf_(y, sinp, cosp); return; }
    fp.require(fabsf_(y) < 120.0f);
    sincosf_main_(y, sinp, cosp);
}

// ===================================================================== tanf
// fdlibm k_tanf.c (binary32), written with selects instead of branches: every lane runs the
// same instruction stream (one polynomial, one division) whichever of the routine's three
// regimes it is in.  Each regime's arithmetic is operation-for-operation the original.
template <class Fp>
RLM_HD float kernel_tanf(Fp &fp, float x, float y, int iy)
{
    const float pio4 = 7.8539812565e-01f, pio4lo = 3.7748947079e-08f;
    const float T0 = 3.3333334327e-01f, T1 = 1.3333334029e-01f, T2 = 5.3968254477e-02f,
                T3 = 2.1869488060e-02f, T4 = 8.8632395491e-03f, T5 = 3.5920790397e-03f,
                T6 = 1.4562094584e-03f, T7 = 5.8804126456e-04f, T8 = 2.4646313977e-04f,
                T9 = 7.8179444245e-05f, T10 = 7.1407252108e-05f, T11 = -1.8558637748e-05f,
                T12 = 2.5907305826e-05f;
    const float x_in = x;
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    const bool big = ix >= 0x3f2ca140;         // |x| >= 0.6744
    const float sgn_big = (float)(1 - ((hx >> 30) & 2));
    const float fiy = (float)iy;
    bool big_tiny = false;
    if (big) {
        float xa = (hx < 0) ? -x : x, ya = (hx < 0) ? -y : y;
        float z0 = pio4 - xa;
        float w0 = pio4lo - ya;
        x = z0 + w0;
        y = 0.0f;
        big_tiny = fabsf_(x) < 0x1p-13f;
    }
    float z = x * x;
    float w = z * z;
    float r = T1 + w * (T3 + w * (T5 + w * (T7 + w * (T9 + w * T11))));
    float v = z * (T2 + w * (T4 + w * (T6 + w * (T8 + w * (T10 + w * T12)))));
    float s = z * x;
    r = y + z * (s * (r + v) + y);
    r += T0 * s;
    w = x + r;
    // the one division: w*w/(w+iy) (big) or -1/w (small, iy == -1)
    float q = fp.div(big ? w * w : -1.0f, big ? w + fiy : w);
    float res_big = sgn_big * (fiy - 2.0f * (x - (q - r)));
    // -1/(x+r), accurately
    float zt = u2f(f2u(w) & 0xfffff000u);
    float vt = r - (zt - x);
    float t = u2f(f2u(q) & 0xfffff000u);
    float st = 1.0f + t * zt;
    float res_inv = t + q * (st + t * vt);
    float res = big ? res_big : ((iy == 1) ? w : res_inv);
    if (Fp::kFast) {
        // The two special returns of the original (|pi/4 - |x|| < 2^-13; |x| < 2^-13) are not
        // carried by the fast stream: such arguments go to the exact re-run.
        fp.require(!big_tiny && ix >= 0x39000000);
        return res;
    }
    if (big_tiny) res = (float)((1 - ((hx >> 30) & 2)) * iy) * (1.0f - (float)(2 * iy) * x);
    if (ix < 0x39000000) {                     // |x| < 2^-13
        if ((int)x_in == 0) {
            if ((ix | (iy + 1)) == 0) res = 1.0f / fabsf_(x_in);
            else if (iy == 1) res = x_in;
            else res = -1.0f / x_in;
        }
    }
    return res;
}
// glibc 2.39 s_tanf.c with the binary64 reduce_fast of e_rem_pio2f.c.  For |x| <= pi/4 the host
// skips the reduction; reducing anyway gives n = 0, y0 = x, y1 = 0 -- the same kernel call -- so
// one path serves both.
template <class Fp>
RLM_HD float tanf_(Fp &fp, float x)
{
    uint32_t ix = f2u(x) & 0x7fffffffu;
    if (ix >= 0x7f800000u) return x - x;
    double dx = (double)x;
    int n;
    if (fabsf_(x) < 120.0f) {                               // abstop12(x) < 0x42f
        double r = dx * kSinCosC[7];
        n = ((int32_t)r + 0x800000) >> 24;
        dx = dx - (double)n * kSinCosC[8];                  // not fused in this routine
    } else {
        uint32_t xi = f2u(x);
        dx = sincosf_reduce_large(xi, &n);
        dx = (xi >> 31) ? -dx : dx;
    }
    float y0 = (float)dx;
    float y1 = (float)(dx - (double)y0);
    return kernel_tanf(fp, y0, y1, 1 - ((n & 1) << 1));
}
RLM_HD float tanf_(float x) { rls::FpExact fp; return tanf_(fp, x); }

// ==================================================================== atanf
// fdlibm s_atanf.c (binary32) with the five argument-reduction branches folded into selects:
// one division (x/1 == x in the unreduced range), one polynomial.
template <class Fp>
RLM_HD float atanf_(Fp &fp, float x)
{
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
                aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
                aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
                aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix >= 0x4c000000) {                    // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return 1.5707962513e+00f + 7.5497894159e-08f;
        else return -1.5707962513e+00f - 7.5497894159e-08f;
    }
    const float ax = fabsf_(x);
    const bool r0 = ix < 0x3ee00000;           // |x| < 0.4375: no reduction
    const bool r1 = ix < 0x3f300000;           // < 11/16
    const bool r2 = ix < 0x3f980000;           // < 19/16
    const bool r3 = ix < 0x401c0000;           // < 2.4375
    float num = r0 ? x : (r1 ? 2.0f * ax - 1.0f : (r2 ? ax - 1.0f : (r3 ? ax - 1.5f : -1.0f)));
    float den = r0 ? 1.0f : (r1 ? 2.0f + ax : (r2 ? ax + 1.0f : (r3 ? 1.0f + 1.5f * ax : ax)));
    float hi = r1 ? 4.6364760399e-01f : (r2 ? 7.8539812565e-01f : (r3 ? 9.8279368877e-01f : 1.5707962513e+00f));
    float lo = r1 ? 5.0121582440e-09f : (r2 ? 3.7748947079e-08f : (r3 ? 3.4473217170e-08f : 7.5497894159e-08f));
    float t = fp.div_pz(num, den);               // den >= 1; num is 0 for |x| == 0, 0.5, 1, 1.5
    float z = t * t;
    float w = z * z;
    float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    float ts = t * (s1 + s2);
    if (r0) return (ix < 0x31000000) ? x : t - ts;   // |x| < 2^-29 returns x
    float zz = hi - ((ts - lo) - t);
    return (hx < 0) ? -zz : zz;
}
RLM_HD float atanf_(float x) { rls::FpExact fp; return atanf_(fp, x); }

// Fast-policy form of atanf_ for a finite argument ax >= 0 (the |y/x| of atan2f_): the same
// five regimes; the |x| >= 2^25 shortcut of the original is not carried (fp.require).
// The five argument reductions are ONE expression (A*ax + B) / (-B*ax + A) with per-regime
// constants, each regime's operations being the original's up to exact steps:
//   regime            original                      A    B      A*ax + B          -B*ax + A
//   |x| < 7/16        x                             1    0      1*x + 0 = x       (-0)*x + 1 = 1
//   < 11/16           (2x - 1) / (2 + x)            2   -1      2x - 1            1*x + 2
//   < 19/16           (x - 1) / (x + 1)             1   -1      1*x - 1           1*x + 1
//   < 39/16           (x - 1.5) / (1 + 1.5x)        1   -1.5    1*x - 1.5         1.5x + 1
//   else              -1 / x                        0   -1      0*x - 1 = -1      1*x + 0 = x
// (a product by 1, 2 or 0 and a sum with 0 are exact; IEEE addition commutes.)
template <class Fp>
RLM_HD float atanf_nonneg_(Fp &fp, float ax)
{
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
                aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
                aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
                aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const int32_t ix = (int32_t)f2u(ax);
    fp.require((uint32_t)ix < 0x4c000000u);    // also excludes NaN / Inf / negative arguments
    const bool r0 = ix < 0x3ee00000;
    const bool r1 = ix < 0x3f300000;
    const bool r2 = ix < 0x3f980000;
    const bool r3 = ix < 0x401c0000;
    const float A = r0 ? 1.0f : (r1 ? 2.0f : (r3 ? 1.0f : 0.0f));
    const float B = r0 ? 0.0f : ((r3 && !r2) ? -1.5f : -1.0f);
    const float num = A * ax + B;
    const float den = (-B) * ax + A;
    float hi = r1 ? 4.6364760399e-01f : (r2 ? 7.8539812565e-01f : (r3 ? 9.8279368877e-01f : 1.5707962513e+00f));
    float lo = r1 ? 5.0121582440e-09f : (r2 ? 3.7748947079e-08f : (r3 ? 3.4473217170e-08f : 7.5497894159e-08f));
    float t = fp.div_pz(num, den);
    float z = t * t;
    float w = z * z;
    float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    float ts = t * (s1 + s2);
    float small = (ix < 0x31000000) ? ax : t - ts;
    return r0 ? small : hi - ((ts - lo) - t);
}

// =================================================================== atan2f
// fdlibm e_atan2f.c (binary32)
template <class Fp>
RLM_HD float atan2f_(Fp &fp, float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    if (Fp::kFast) {
        // Main path only.  The operand window of div_z (x finite, non-zero; y finite) and the
        // fp.require of atanf_nonneg_ exclude every special case except y == +-0, x == 1 and
        // |k| > 60 with a quotient below 2^25, for which the generic formulas below give the
        // original's bits: atanf(0) = 0; y/1 == y; pi - (0 - pi_lo) == pi + tiny == pi;
        // a quotient < 2^-60 does not change z - pi_lo; (z - pi_lo) - pi == -(pi - (z - pi_lo)).
        float z0 = atanf_nonneg_(fp, fabsf_(fp.div_z(y, x)));      // only |y/x| is used
        float w0 = ((int32_t)f2u(x) < 0) ? pi - (z0 - pi_lo) : z0; // >= +0
        return u2f(f2u(w0) | (f2u(y) & 0x80000000u));
    }
    float z;
    int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
    int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_(fp, y);
    int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi + tiny;
        default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return pi_o_4 + tiny;
            case 1: return -pi_o_4 - tiny;
            case 2: return 3.0f * pi_o_4 + tiny;
            default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    int32_t k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_(fp, fabsf_(fp.div_z(y, x)));     // only |y/x| is used
    switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}
RLM_HD float atan2f_(float y, float x) { rls::FpExact fp; return atan2f_(fp, y, x); }

// ==================================================================== acosf
// fdlibm e_acosf.c (binary32) as shipped by glibc 2.39 (six-term P, four-term Q), with the
// three ranges sharing one P/Q evaluation: z = x*x for |x| < 0.5, else (1 - |x|)/2
// ((1 + x)*0.5 for x < -0.5 is the same IEEE operation as (1 - |x|)*0.5).
template <class Fp>
RLM_HD float acosf_(Fp &fp, float x)
{
    const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f,
                pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f,
                qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f,
                qS4 = 7.7038154006e-02f;
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix >= 0x3f800000) {
        if (ix > 0x3f800000) return (x - x) / (x - x);
        return (hx > 0) ? 0.0f : pi + 2.0f * pio2_lo;
    }
    const bool small = ix < 0x3f000000;        // |x| < 0.5
    const float z = small ? x * x : (1.0f - fabsf_(x)) * 0.5f;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = fp.div_pz(p, q);           // q in (0.2, 1]; p == 0 for x == 0
    if (small) {
        if (ix <= 0x32800000) return pio2_hi + pio2_lo;
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    const float s = fp.sqrt(z);
    if (hx < 0) {                              // x < -0.5
        float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    const float df = u2f(f2u(s) & 0xfffff000u);
    const float c = fp.div_pz(z - df * df, s + df);
    const float w = r * s + c;
    return 2.0f * (df + w);
}
RLM_HD float acosf_(float x) { rls::FpExact fp; return acosf_(fp, x); }

// ===================================================================== expf
// glibc 2.39 e_expf.c (ARM Optimized Routines), FMA build
template <bool kSmem> RLM_HD float expf_main_(float x);
RLM_HD float expf_(float x)
{
    if (!(fabsf_(x) < 88.0f)) {                // abstop12(x) >= 0x42b: |x| >= 88 or NaN
        if (f2u(x) == 0xff800000u) return 0.0f;
        if (abstop12(x) >= 0x7f8u) return x + x;
        if (x > 0x1.62e42ep6f) return u2f(0x7f800000u);             // overflow
        if (x < -0x1.9fe368p6f) return 0.0f;                        // underflow
        if (x < -0x1.9d1d9ep6f) return 0x1.4p-75f * 0x1.4p-75f;     // may-underflow value
    }
    return expf_main_<false>(x);
}
// The main path of expf_ alone.  With the argument clamped from below at -104.5 it returns the host's
// bits for EVERY x < 88 (tests/native/libm_check "explean", exhaustive): on [-103.97, -88) the host
// itself falls through to this path, and below it both round to the values the host's underflow
// returns produce (2^-149 down to -103.97, then 0).
template <bool kSmem = false>
RLM_HD float expf_main_(float x)
{
    double xd = (double)x;
    const double InvLn2N = kExpC[0], Shift = kExpC[1];
    const double C0 = kExpC[2], C1 = kExpC[3], C2 = kExpC[4];
    double kd = fma_(InvLn2N, xd, Shift);
    uint64_t ki = d2u(kd);
    kd -= Shift;
    double r = fma_(InvLn2N, xd, -kd);
    uint64_t t = exp2_tab<kSmem>((uint32_t)ki & 31u);
    t += ki << 47;
    double s = u2d(t);
    double z = fma_(C0, r, C1);
    double r2 = r * r;
    double y = fma_(C2, r, 1.0);
    y = fma_(z, r2, y);
    y = y * s;
    return (float)y;
}
// Policy form: the fast stream carries neither the range ladder nor its branch -- one clamp, and
// x >= 88 / NaN go to the exact re-run.
template <class Fp>
RLM_HD float expf_(Fp &fp, float x)
{
    if (!Fp::kFast) return expf_(x);
    fp.require(x < 88.0f);
#if defined(__CUDA_ARCH__)
    return expf_main_<Fp::kSmemTables>(fmaxf(x, -104.5f));
#else
    return expf_main_<false>(x > -104.5f ? x : -104.5f);
#endif
}


// ===================================================================== logf
// glibc 2.39 e_logf.c (ARM Optimized Routines), FMA build
template <bool kSmem> RLM_HD float logf_main_(uint32_t ix);
RLM_HD float logf_(float x)
{
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -u2f(0x7f800000u);          // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                    // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return (x - x) / (x - x);
        ix = f2u(x * 0x1p23f);                              // subnormal: normalise
        ix -= 23u << 23;
    }
    return logf_main_<false>(ix);
}
// The main path of logf_ for the bits ix of a positive normal number.  It also returns the host's +0
// for x == 1 (the host's shortcut is a speed-up, not a special value): tests/native/libm_check
// "loglean" walks every positive normal binary32.
template <bool kSmem = false>
RLM_HD float logf_main_(uint32_t ix)
{
    const double Ln2 = kLogC[0];
    const double A0 = kLogC[1], A1 = kLogC[2], A2 = kLogC[3];
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = log_tab<kSmem>(2 * i);
    double logc = log_tab<kSmem>(2 * i + 1);
    double z = (double)u2f(iz);
    double r = fma_(z, invc, -1.0);
    double y0 = fma_((double)k, Ln2, logc);
    double r2 = r * r;
    double y = fma_(A1, r, A2);
    y = fma_(A0, r2, y);
    y = fma_(y, r2, y0 + r);
    return (float)y;
}
// Policy form: zero, subnormal, negative, Inf and NaN arguments go to the exact re-run.
template <class Fp>
RLM_HD float logf_(Fp &fp, float x)
{
    if (!Fp::kFast) return logf_(x);
    const uint32_t ix = f2u(x);
    fp.require(ix - 0x00800000u < 0x7f800000u - 0x00800000u);
    return logf_main_<Fp::kSmemTables>(ix);
}


// ===================================================================== powf
// glibc 2.39 e_powf.c (ARM Optimized Routines), FMA build.  Main path for x > 0; the
// IEEE special cases the reference can reach (x == 0, x == 1, y == 0) are handled; negative
// bases do not occur on the path (arguments are squares or clamped to [0, 1]).
template <bool kSmem = false>
RLM_HD float powf_(float x, float y)
{
    uint32_t ix = f2u(x), iy = f2u(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || ((2 * iy - 1) >= (2u * 0x7f800000u - 1))) {
        // x is 0, subnormal, inf, nan or negative; or y is 0, inf or nan
        if ((2 * iy - 1) >= (2u * 0x7f800000u - 1)) {
            if (2 * iy == 0) return 1.0f;                                   // pow(x, +-0) = 1
            if (ix == 0x3f800000u) return 1.0f;                             // pow(1, y) = 1
            if (2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return x + y;
            if (2 * ix == 2u * 0x3f800000u) return 1.0f;
            if ((2 * ix < 2u * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;   // |x|<1 && y==inf etc.
            return y * y;
        }
        if ((2 * ix - 1) >= (2u * 0x7f800000u - 1)) {
            float x2 = x * x;                                               // x is +-0, +-inf, nan
            return (iy & 0x80000000u) ? 1.0f / x2 : x2;
        }
        if (ix & 0x80000000u) return (x - x) / (x - x);                     // negative base: not on the path
        if (ix < 0x00800000u) {                                             // subnormal x
            ix = f2u(x * 0x1p23f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    // log2_inline
    const double A0 = kLog2C[0], A1 = kLog2C[1], A2 = kLog2C[2], A3 = kLog2C[3], A4 = kLog2C[4];
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc = log2_tab<kSmem>(2 * i);
    double logc = log2_tab<kSmem>(2 * i + 1);
    double z = (double)u2f(iz);
    double r = fma_(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double yy = fma_(A0, r, A1);
    double p = fma_(A2, r, A3);
    double r4 = r2 * r2;
    double q = fma_(A4, r, y0);
    q = fma_(p, r2, q);
    double logx = fma_(yy, r4, q);

    double ylogx = (double)y * logx;
    if (((d2u(ylogx) >> 47) & 0xffffu) >= (d2u(126.0) >> 47)) {
        if (ylogx > 0x1.fffffffd1d571p+6) return u2f(0x7f800000u);          // overflow
        if (ylogx <= -150.0) return 0.0f;                                   // underflow
        if (ylogx < -149.0) return 0x1.4p-75f * 0x1.4p-75f;                 // may-underflow value
    }
    // exp2_inline
    const double ShiftScaled = kExp2C[0];
    const double C0 = kExp2C[1], C1 = kExp2C[2], C2 = kExp2C[3];
    double kd = ylogx + ShiftScaled;
    uint64_t ki = d2u(kd);
    kd -= ShiftScaled;
    double rr = ylogx - kd;
    uint64_t t = exp2_tab<kSmem>((uint32_t)ki & 31u);
    t += ki << 47;
    double s = u2d(t);
    double zz = fma_(C0, rr, C1);
    double rr2 = rr * rr;
    double out = fma_(C2, rr, 1.0);
    out = fma_(zz, rr2, out);
    out = out * s;
    return (float)out;
}

// powf(x, 5.0f) for x in [0, 1] or NaN -- the Schlick weights pow(clamp(1 - cos, 0, 1), 5) of
// rlDisney (src/rlDisney.cpp:218-219,335).  The main path of powf_ without the tests that cannot
// fire on this domain: x normal (1 - cos is 0 or >= 2^-24), 5 log2(x) in [-120, 0] (no overflow /
// underflow range test).  x == 0 and NaN (and subnormals, for totality) go through powf_.
template <bool kSmem = false>
RLM_HD float pow5_unit_(float x)
{
    uint32_t ix = f2u(x);
    if (!(ix - 0x00800000u < 0x3f800000u - 0x00800000u + 1u)) return powf_<kSmem>(x, 5.0f);   // not a normal number in (0, 1]
    const double A0 = kLog2C[0], A1 = kLog2C[1], A2 = kLog2C[2], A3 = kLog2C[3], A4 = kLog2C[4];
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc = log2_tab<kSmem>(2 * i);
    double logc = log2_tab<kSmem>(2 * i + 1);
    double z = (double)u2f(iz);
    double r = fma_(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double yy = fma_(A0, r, A1);
    double p = fma_(A2, r, A3);
    double r4 = r2 * r2;
    double q = fma_(A4, r, y0);
    q = fma_(p, r2, q);
    double logx = fma_(yy, r4, q);
    double ylogx = 5.0 * logx;
    const double ShiftScaled = kExp2C[0];
    const double C0 = kExp2C[1], C1 = kExp2C[2], C2 = kExp2C[3];
    double kd = ylogx + ShiftScaled;
    uint64_t ki = d2u(kd);
    kd -= ShiftScaled;
    double rr = ylogx - kd;
    uint64_t t = exp2_tab<kSmem>((uint32_t)ki & 31u);
    t += ki << 47;
    double s = u2d(t);
    double zz = fma_(C0, rr, C1);
    double rr2 = rr * rr;
    double out = fma_(C2, rr, 1.0);
    out = fma_(zz, rr2, out);
    out = out * s;
    return (float)out;
}
// (A step-wise N-chain form that fetched every binary64 coefficient once for the two Schlick weights
// FL, FV of rlDisney was measured on B200: the interleaved chains spill at 64 registers, -1 %.)

} // namespace rlm
