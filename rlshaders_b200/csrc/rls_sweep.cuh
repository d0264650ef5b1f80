// rls_sweep.cuh -- pieces of the directional-albedo / white-furnace sweep (BASELINE config 5, include/rls_b200.h
// rls_albedo_sweep) shared by its bit-exact kernel (rls_b200.cu) and its tolerance kernel (rls_tol.cu): the grid, the
// counter hash of the in-kernel uniforms, the cell geometry and the per-cell accumulation.
//
// Work distribution: ONE WARP PER CELL.  A lane takes samples k0 + lane, k0 + lane + 32, ... of the cell's spp range and
// keeps five FP64 partial sums; the warp combines them with shuffles and lanes 0..4 write the cell's five values.  No
// shared memory, no CTA barrier, and the fixed cost per cell (cell set-up + 25 64-bit shuffles) is amortised over
// spp / 32 samples per lane whatever the CTA size -- with the spp range sharded over 8 GPUs (512 samples per cell and
// GPU) a CTA per cell left 4 samples per thread under a barrier and five shared-memory reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rls {

struct SweepGridDev { int n_rough, n_cos, n_ior; float rlo, rhi, ilo, ihi; };
constexpr int kSweepValues = 5;          // RLS_SWEEP_VALUES_PER_CELL
constexpr int kSweepBlock = 128;         // 4 cells per CTA

__device__ __forceinline__ uint64_t sweep_hash64(uint64_t seed, uint32_t stream, uint64_t index)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1ull) + 0xD1B54A32D192ED03ull * (uint64_t)stream;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float sweep_uniform24(uint64_t seed, uint32_t stream, uint64_t index)
{
    uint32_t k = (uint32_t)(sweep_hash64(seed, stream, index) >> 40);
    if (k == 0u) k = 1u;
    return (float)k * 5.9604644775390625e-8f;
}
// cell -> (roughness, cos theta_v, ior); the oracle restates this in oracle_common.h (orc_sweep_cell)
__device__ __forceinline__ void sweep_cell(const SweepGridDev &g, uint32_t cell, float &rough, float &cosv, float &ior)
{
    const uint32_t ie = cell % (uint32_t)g.n_ior;
    const uint32_t ic = (cell / (uint32_t)g.n_ior) % (uint32_t)g.n_cos;
    const uint32_t ir = cell / (uint32_t)(g.n_ior * g.n_cos);
    const float tr = g.n_rough > 1 ? (float)ir / (float)(g.n_rough - 1) : 0.0f;
    const float te = g.n_ior > 1 ? (float)ie / (float)(g.n_ior - 1) : 0.0f;
    rough = g.rlo + (g.rhi - g.rlo) * tr;
    ior = g.ilo + (g.ihi - g.ilo) * te;
    cosv = (float)(ic + 1u) / (float)g.n_cos;
}
// [0] sum f_r / pdf_r (valid)  [1] sum weight_t (refracted)  [2] sum fresnel  [3] valid count  [4] TIR count
__device__ __forceinline__ void sweep_accumulate(double (&acc)[kSweepValues], uint32_t flags, float f_r, float pdf_r, float w_t, float F)
{
    const bool valid = !(flags & 0x0003u);                  // !(ZERO_L | BELOW_HORIZON)
    if (valid) { acc[0] += (double)(f_r / pdf_r); acc[3] += 1.0; }
    if (flags & 0x0020u) acc[4] += 1.0; else acc[1] += (double)w_t;      // TIR
    acc[2] += (double)F;
}
// Warp sum of the five partials; lane j < 5 then holds value j of the cell and stores (or adds) it.
__device__ __forceinline__ void sweep_store(double (&acc)[kSweepValues], uint32_t cell, double *table)
{
    const uint32_t lane = threadIdx.x & 31u;
    double mine = 0.0;
#pragma unroll
    for (int j = 0; j < kSweepValues; j++) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == (uint32_t)j) mine = v;
    }
    if (lane < (uint32_t)kSweepValues) table[(size_t)cell * kSweepValues + lane] = mine;
}

} // namespace rls
