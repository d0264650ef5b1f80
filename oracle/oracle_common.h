/*
 * oracle/oracle_common.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Pieces shared by both oracle libraries that are NOT reference arithmetic: parameter
 * fetch from the ABI structs, the synthetic-input integer hash, and the sweep's cell
 * geometry.  (The product restates the hash and the cell geometry on the device in
 * rlshaders_b200/csrc/; tests compare the two.)
 */
#ifndef RLS_ORACLE_COMMON_H
#define RLS_ORACLE_COMMON_H

#include <stdint.h>
#include <stddef.h>
#include "oracle_api.h"

static inline float orc_p1(const rls_param1 *p, size_t i) { return p->array ? p->array[i] : p->value; }
static inline void  orc_p3(const rls_param3 *p, size_t i, float out[3])
{
    out[0] = p->array.x ? p->array.x[i] : p->value[0];
    out[1] = p->array.y ? p->array.y[i] : p->value[1];
    out[2] = p->array.z ? p->array.z[i] : p->value[2];
}

/* SplitMix64-style counter hash: (seed, stream, index) -> 64 random bits. */
static inline uint64_t orc_hash64(uint64_t seed, uint32_t stream, uint64_t index)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1ull) + 0xD1B54A32D192ED03ull * (uint64_t)stream;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* 24-bit uniform in [2^-24, 1 - 2^-24]; never 0 or 1 (rlGgx.cpp:20 divides by 1-rx,
 * rlSss.cpp:83 divides by r). */
static inline float orc_uniform(uint64_t seed, uint32_t stream, uint64_t index)
{
    uint32_t k = (uint32_t)(orc_hash64(seed, stream, index) >> 40);
    if (k == 0u) k = 1u;
    return (float)k * 5.9604644775390625e-8f;   /* 2^-24, exact */
}

/* RLS_FLAG_SLOPE_EARLY_OUT probe.  The reference's visible-normal sampler does not report which branch it took, so
 * the flag is derived by re-evaluating the expressions that decide its two early-outs -- the float operations of
 * src/rlGgx.cpp:66-84 (view -> stretched polar angle theta; theta is left 0 unless V.z < 1 - AI_EPSILON) and :27-38
 * (theta < AI_EPSILON, or |A^2 - 1| < AI_EPSILON) -- on the same inputs.  src/rlDisney.cpp:467-489,416-438 is the same
 * code.  view, U, V, N: 3 floats each; ax, ay: the sampler's alphas; rx: the first uniform as the sampler receives it. */
#include <math.h>
static inline int orc_vndf_early_out(const float *view, const float *U, const float *V, const float *N,
                                     float ax, float ay, float rx)
{
    float d = N[0] * view[0] + N[1] * view[1] + N[2] * view[2];
    float cosThetaV = d < -1.0f ? -1.0f : (d > 1.0f ? 1.0f : d);
    float phiV = atan2f(V[0] * view[0] + V[1] * view[1] + V[2] * view[2], U[0] * view[0] + U[1] * view[1] + U[2] * view[2]);
    float r = sqrtf(1.0f - cosThetaV * cosThetaV);
    float x = r * cosf(phiV), y = r * sinf(phiV), z = cosThetaV;
    x *= ax;
    y *= ay;
    float len = sqrtf(x * x + y * y + z * z);
    if (len != 0.0f) { float inv = 1.0f / len; x *= inv; y *= inv; z *= inv; }
    if (!(z < (1.0f - 1.0e-4f))) return 1;              /* theta = 0 -> sampleSlope's first early-out */
    float theta = acosf(z);
    if (theta < 1.0e-4f) return 1;
    float B = tanf(theta);
    float B2 = B * B;
    float G1 = 2.0f / (1.0f + sqrtf(1.0f + B2));
    float A = 2.0f * rx / G1 - 1.0f;
    float A2 = A * A;
    float dA = A2 - 1.0f;
    return ((dA < 0) ? -dA : dA) < 1.0e-4f;
}

/* Sweep cell -> (roughness, cos theta_v, ior); see include/rls_b200.h rls_albedo_sweep. */
static inline void orc_sweep_cell(const rls_sweep_grid *g, uint32_t cell, float *roughness, float *cosv, float *ior)
{
    uint32_t ie = cell % (uint32_t)g->n_ior;
    uint32_t ic = (cell / (uint32_t)g->n_ior) % (uint32_t)g->n_cos;
    uint32_t ir = cell / (uint32_t)(g->n_ior * g->n_cos);
    float tr = g->n_rough > 1 ? (float)ir / (float)(g->n_rough - 1) : 0.0f;
    float te = g->n_ior > 1 ? (float)ie / (float)(g->n_ior - 1) : 0.0f;
    *roughness = g->roughness_lo + (g->roughness_hi - g->roughness_lo) * tr;
    *ior = g->ior_lo + (g->ior_hi - g->ior_lo) * te;
    *cosv = (float)(ic + 1u) / (float)g->n_cos;
}

#endif
