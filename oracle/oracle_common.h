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
