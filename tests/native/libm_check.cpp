// tests/native/libm_check.cpp -- TEST INFRASTRUCTURE.
// Runs the product's device libm (rlshaders_b200/csrc/rls_libm.cuh, which is __host__
// __device__) on the CPU against the host C library, bit for bit.
//   libm_check <function> [stride]     function: sincos cos tan atan acos exp log atan2 pow pow5 pow5unit
// Univariate functions walk every binary32 bit pattern (stride 1) or every stride-th one;
// bivariate ones draw (stride-scaled) pseudo-random pairs from the path's domains.
// Prints "<function> checked N mismatches M first <hex args>" and exits 0.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "../../rlshaders_b200/csrc/rls_libm.cuh"

static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline bool same(float a, float b) { return fbits(a) == fbits(b) || (a != a && b != b); }
static inline uint64_t mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char **argv)
{
    std::string fn = argc > 1 ? argv[1] : "sincos";
    uint64_t stride = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
    uint64_t checked = 0, bad = 0;
    uint64_t first = ~0ull;
    if (fn == "atan2" || fn == "pow" || fn == "pow5") {
        const uint64_t n = (1ull << 30) / stride;
#pragma omp parallel for reduction(+ : checked, bad) reduction(min : first)
        for (uint64_t i = 0; i < n; i++) {
            uint64_t h = mix(i * 0x9E3779B97F4A7C15ull + 12345);
            float a = bitsf((uint32_t)h), b = bitsf((uint32_t)(h >> 32));
            float want, got;
            if (fn == "atan2") {
                // any two finite floats; half the draws are unit-range magnitudes like the path's
                if (i & 1) { a = (float)((int32_t)(uint32_t)h) * 0x1p-31f; b = (float)((int32_t)(uint32_t)(h >> 32)) * 0x1p-31f; }
                if (!(a == a) || !(b == b) || isinf(a) || isinf(b)) continue;
                want = atan2f(a, b); got = rlm::atan2f_(a, b);
            } else if (fn == "pow") {
                // base in [0, 1], exponent in (0, 1): the GTR1 sampler's pow(a2, 1 - ry)
                a = (float)((uint32_t)h >> 8) * 0x1p-24f; b = (float)((uint32_t)(h >> 40) | 1u) * 0x1p-24f;
                if (i % 7 == 0) { a = fabsf(bitsf((uint32_t)h)); b = bitsf((uint32_t)(h >> 32)); if (!(a == a) || !(b == b)) continue; }
                want = powf(a, b); got = rlm::powf_(a, b);
            } else {
                a = (float)((uint32_t)h >> 8) * 0x1p-24f; b = 5.0f;      // Schlick weights: pow(x, 5)
                want = powf(a, b); got = rlm::powf_(a, b);
            }
            checked++;
            if (!same(want, got)) { bad++; uint64_t key = ((uint64_t)fbits(a) << 32) | fbits(b); if (key < first) first = key; }
        }
    } else {
#pragma omp parallel for reduction(+ : checked, bad) reduction(min : first)
        for (uint64_t u = 0; u < (1ull << 32); u += stride) {
            float x = bitsf((uint32_t)u);
            bool ok = true;
            if (fn == "sincos") {
                float s0, c0, s1, c1;
                sincosf(x, &s0, &c0);
                rlm::sincosf_(x, &s1, &c1);
                ok = same(s0, s1) && same(c0, c1);
            } else if (fn == "sincoslean") {            // fast-policy sincosf: main path only, for every |x| < 120
                if (fabsf(x) < 120.0f) {
                    float s0, c0, s1, c1;
                    sincosf(x, &s0, &c0);
                    rlm::sincosf_main_(x, &s1, &c1);
                    ok = same(s0, s1) && same(c0, c1);
                }
            } else if (fn == "cos") {                   // SampleWriter's cosf(theta) is served by sincosf_
                float s1, c1;
                rlm::sincosf_(x, &s1, &c1);
                ok = same(cosf(x), c1);
            } else if (fn == "tan") ok = same(tanf(x), rlm::tanf_(x));
            else if (fn == "atan") ok = same(atanf(x), rlm::atanf_(x));
            else if (fn == "acos") ok = same(acosf(x), rlm::acosf_(x));
            else if (fn == "exp") ok = same(expf(x), rlm::expf_(x));
            else if (fn == "log") ok = same(logf(x), rlm::logf_(x));
            else if (fn == "loglean") {                 // fast-policy logf: main path only, every positive normal x
                if ((uint32_t)u - 0x00800000u < 0x7f800000u - 0x00800000u) ok = same(logf(x), rlm::logf_main_<false>((uint32_t)u));
            }
            else if (fn == "explean") {                 // fast-policy expf: clamp + main path, for every x < 88
                if (x < 88.0f) ok = same(expf(x), rlm::expf_main_<false>(x > -104.5f ? x : -104.5f));
            }
            else if (fn == "pow5unit") {                // Schlick weights: every x in [0, 1] (subnormals too) and NaN
                // (1 - c is never -0 in round-to-nearest, so the sign bit is clear on the path)
                if ((x != x) || (x >= 0.0f && x <= 1.0f && !(fbits(x) >> 31))) ok = same(powf(x, 5.0f), rlm::pow5_unit_(x));
            }
            else { fprintf(stderr, "unknown function %s\n", fn.c_str()); exit(2); }
            checked++;
            if (!ok) { bad++; if (u < first) first = u; }
        }
    }
    printf("%s checked %llu mismatches %llu first %016llx\n", fn.c_str(), (unsigned long long)checked,
           (unsigned long long)bad, (unsigned long long)first);
    return 0;
}
