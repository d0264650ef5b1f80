// rls_tol.cu -- the fused kernels of the TOLERANCE arithmetic policy (RLS_ARITH_TOLERANT, rls_tol.cuh).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=true -c rls_tol.cu   (FMA contraction ON:
// this is the one translation unit of the library that is not bit-exact by construction; see __graft_entry__.build).
//
// One thread = one sample, SoA loads / stores exactly as the bit-exact kernels (rls_b200.cu).  Per 32 samples the
// dielectric unit is ~1/4 of the bit-exact kernel's instructions, so these kernels are bound by HBM, not by issue.
// A sample whose band tracker fired is appended to the re-run list (rls_tol_launch.cuh) and evaluated again by the
// bit-exact policy in a second kernel: flags, lobes and discontinuous choices are therefore the reference's, bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rls_tol.cuh"
#include "rls_tol_launch.cuh"

namespace rls {
namespace tol {

#ifndef RLS_TOL_BLOCK
#define RLS_TOL_BLOCK 256
#endif
#ifndef RLS_TOL_MIN_BLOCKS
#define RLS_TOL_MIN_BLOCKS 4
#endif
static constexpr int kBlockTol = RLS_TOL_BLOCK;

static __device__ __forceinline__ v3 tv(f3 a) { return mk(a.x, a.y, a.z); }
static __device__ __forceinline__ void st3(const V3 &v, uint32_t i, v3 a) { v.x[i] = a.x; v.y[i] = a.y; v.z[i] = a.z; }

// Returns the flags word to store: the sample's own, or the sentinel when the list is full.
static __device__ __forceinline__ uint32_t enlist(bool rerun, uint32_t i, uint32_t flags, const Worklist &wl)
{
    if (rerun) {
        const unsigned k = atomicAdd(wl.count, 1u);
        if (k < wl.cap) wl.list[k] = i; else flags = kRerunSentinel;
    }
    return flags;
}

#define RLS_TOL_INDEX()                                                        \
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;                  \
    if (i >= n) return;

__global__ void __launch_bounds__(kBlockTol, RLS_TOL_MIN_BLOCKS)
k_ggx_sample_eval_pdf_tol(uint32_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, BsdfOutDev o, Worklist wl)
{
    RLS_TOL_INDEX();
    const Shading s = load_shading(sg, i);
    Bands bd;
    const GgxBsdfT r = ggx_unit(bd, tv(s.U), tv(s.V), tv(s.N), tv(s.wo), s.backfacing, tv(fetch(p.ks, i)), fetch(p.ior, i),
                                fetch(p.rough, i), fetch(p.aniso, i), __ldg(rx + i), __ldg(ry + i));
    st3(o.wi, i, r.L);
    st3(o.f, i, r.f);
    o.pdf[i] = r.pdf;
    if (o.fresnel) o.fresnel[i] = r.fresnel;
    o.flags[i] = enlist(bd.rerun, i, r.flags, wl);
}

template <bool kArrays>       // kArrays: ior and specularRoughness are per-sample arrays
__global__ void __launch_bounds__(kBlockTol, RLS_TOL_MIN_BLOCKS)
k_ggx_dielectric_tol(uint32_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, DielectricOutDev o, Worklist wl)
{
    RLS_TOL_INDEX();
    const Shading s = load_shading(sg, i);
    const float ior = fetch_t<kArrays>(p.ior, i), rough = fetch_t<kArrays>(p.rough, i), aniso = fetch(p.aniso, i);
    Bands bd;
    const DielectricT r = dielectric_unit(bd, tv(s.U), tv(s.V), tv(s.N), tv(s.wo), s.backfacing, ior, rough, aniso,
                                          __ldg(rx + i), __ldg(ry + i));
    o.fresnel[i] = r.F;
    st3(o.wi_r, i, r.wi_r);
    o.f_r[i] = r.f_r;
    o.pdf_r[i] = r.pdf_r;
    st3(o.wi_t, i, r.wi_t);
    o.f_t[i] = r.f_t;
    o.weight_t[i] = r.w_t;
    o.flags[i] = enlist(bd.rerun, i, r.flags, wl);
}

template <bool kArrays>       // kArrays: every rlDisney parameter is a per-sample array
__global__ void __launch_bounds__(kBlockTol, RLS_TOL_MIN_BLOCKS)
k_disney_sample_eval_pdf_tol(uint32_t n, ShadingSoA sg, DisneyParamsDev p, const float *rx_s, const float *ry_s,
                             const float *rx_d, const float *ry_d, DisneyOutDev o, Worklist wl)
{
    RLS_TOL_INDEX();
    const Shading s = load_shading(sg, i);
    DisneyIn in;
    in.base = tv(fetch_t<kArrays>(p.base_color, i));
    in.subsurface = fetch_t<kArrays>(p.subsurface, i); in.metallic = fetch_t<kArrays>(p.metallic, i);
    in.specular = fetch_t<kArrays>(p.specular, i); in.specular_tint = fetch_t<kArrays>(p.specular_tint, i);
    in.roughness = fetch_t<kArrays>(p.roughness, i); in.anisotropic = fetch_t<kArrays>(p.anisotropic, i);
    in.sheen = fetch_t<kArrays>(p.sheen, i); in.sheen_tint = fetch_t<kArrays>(p.sheen_tint, i);
    in.clearcoat = fetch_t<kArrays>(p.clearcoat, i); in.clearcoat_gloss = fetch_t<kArrays>(p.clearcoat_gloss, i);
    Bands bd;
    const DisneyT r = disney_unit(bd, tv(s.U), tv(s.V), tv(s.N), tv(s.wo), in, p.sample_from_visible_normal != 0,
                                  __ldg(rx_s + i), __ldg(ry_s + i), __ldg(rx_d + i), __ldg(ry_d + i));
    st3(o.wi_s, i, r.Ls); st3(o.f_s, i, r.fs); o.pdf_s[i] = r.ps;
    st3(o.wi_d, i, r.Ld); st3(o.f_d, i, r.fd); o.pdf_d[i] = r.pd;
    o.flags[i] = enlist(bd.rerun, i, r.flags, wl);
}

__global__ void __launch_bounds__(kBlockTol, RLS_TOL_MIN_BLOCKS)
k_skin_profile_tol(uint32_t n, SkinParamsDev sp, const float *rx, ProfileOutDev o, Worklist wl)
{
    RLS_TOL_INDEX();
    const f3 d = fetch(sp.sss_scatter_dist, i);
    const float mult = fetch(sp.sss_dist_multiplier, i);
    Bands bd;
    // src/rlSkin.cpp:236: scatterDist = sss_scatter_dist * sss_dist_multiplier (one rounding each, as the reference)
    const ProfileT r = skin_profile_unit(bd, mk(mul_rn(d.x, mult), mul_rn(d.y, mult), mul_rn(d.z, mult)), __ldg(rx + i));
    o.r[i] = r.r;
    o.pdf[i] = r.pdf;
    st3(o.Rd, i, r.Rd);
    o.flags[i] = enlist(bd.rerun, i, r.flags, wl);
}

// Two consecutive samples per thread with 64-bit loads / stores: the skin unit is ~100 FP instructions under ~130 of
// address arithmetic, constant loads and pointer tests for its ten arrays -- at one sample per thread it is ISSUE bound
// (239 instructions per 32 samples, 83 % issue slots, DRAM 62 %) although it moves only 40 B per sample.  A pair shares
// the address arithmetic, and the launch site guarantees what the pointer tests asked (sss_scatter_dist is three
// per-sample arrays, the multiplier is uniform, every pointer is 8-byte aligned).  The odd last sample takes scalar
// accesses in the same kernel.
__global__ void __launch_bounds__(kBlockTol, RLS_TOL_MIN_BLOCKS)
k_skin_profile_tol_x2(uint32_t n, const float *dx, const float *dy, const float *dz, float mult, const float *rx,
                      ProfileOutDev o, Worklist wl)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = 2u * t;
    if (i >= n) return;
    if (i + 1u < n) {
        const float2 ax = __ldg((const float2 *)(dx + i)), ay = __ldg((const float2 *)(dy + i)), az = __ldg((const float2 *)(dz + i));
        const float2 u = __ldg((const float2 *)(rx + i));
        Bands b0, b1;
        const ProfileT r0 = skin_profile_unit(b0, mk(mul_rn(ax.x, mult), mul_rn(ay.x, mult), mul_rn(az.x, mult)), u.x);
        const ProfileT r1 = skin_profile_unit(b1, mk(mul_rn(ax.y, mult), mul_rn(ay.y, mult), mul_rn(az.y, mult)), u.y);
        *(float2 *)(o.r + i) = make_float2(r0.r, r1.r);
        *(float2 *)(o.pdf + i) = make_float2(r0.pdf, r1.pdf);
        *(float2 *)(o.Rd.x + i) = make_float2(r0.Rd.x, r1.Rd.x);
        *(float2 *)(o.Rd.y + i) = make_float2(r0.Rd.y, r1.Rd.y);
        *(float2 *)(o.Rd.z + i) = make_float2(r0.Rd.z, r1.Rd.z);
        const uint32_t f0 = enlist(b0.rerun, i, r0.flags, wl), f1 = enlist(b1.rerun, i + 1u, r1.flags, wl);
        *(uint2 *)(o.flags + i) = make_uint2(f0, f1);
    } else {
        Bands bd;
        const ProfileT r = skin_profile_unit(bd, mk(mul_rn(__ldg(dx + i), mult), mul_rn(__ldg(dy + i), mult), mul_rn(__ldg(dz + i), mult)),
                                             __ldg(rx + i));
        o.r[i] = r.r; o.pdf[i] = r.pdf; st3(o.Rd, i, r.Rd);
        o.flags[i] = enlist(bd.rerun, i, r.flags, wl);
    }
}

// The albedo sweep (rls_sweep.cuh: one warp per cell) on the tolerance-policy unit.  Band samples are listed for the
// exact re-run and left out of the sums.
__global__ void __launch_bounds__(kSweepBlock, 8)
k_albedo_sweep_tol(SweepGridDev g, uint32_t n_cells, uint64_t seed, uint32_t k0, uint32_t k1, double *table, SweepWorklist wl)
{
    const uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= n_cells) return;
    float rough, cosv, ior;
    sweep_cell(g, cell, rough, cosv, ior);
    const v3 U = mk(1.0f, 0.0f, 0.0f), V = mk(0.0f, 1.0f, 0.0f), N = mk(0.0f, 0.0f, 1.0f);
    const v3 wo = mk(sqrtf(1.0f - cosv * cosv), 0.0f, cosv);
    double acc[kSweepValues] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
    for (uint32_t k = k0 + (threadIdx.x & 31u); k < k1; k += 32u) {
        const uint64_t idx = ((uint64_t)cell << 32) | (uint64_t)k;
        Bands bd;
        const DielectricT r = dielectric_unit(bd, U, V, N, wo, false, ior, rough, 0.0f, sweep_uniform24(seed, 0u, idx),
                                              sweep_uniform24(seed, 1u, idx), /* ior_band = */ false);
        bool mine = true;
        if (bd.rerun) {
            const unsigned j = atomicAdd(wl.count, 1u);
            if (j < wl.cap) { wl.list[j] = make_uint2(cell, k); mine = false; }
        }
        if (mine) sweep_accumulate(acc, r.flags, r.f_r, r.pdf_r, r.w_t, r.F);
    }
    sweep_store(acc, cell, table);
}

static inline unsigned grid_for(size_t n) { return (unsigned)((n + kBlockTol - 1) / kBlockTol); }

cudaError_t launch_ggx_sample_eval_pdf(cudaStream_t st, size_t n, const ShadingSoA &sg, const GgxParamsDev &p, const float *rx,
                                       const float *ry, const BsdfOutDev &o, const Worklist &wl)
{
    k_ggx_sample_eval_pdf_tol<<<grid_for(n), kBlockTol, 0, st>>>((uint32_t)n, sg, p, rx, ry, o, wl);
    return cudaGetLastError();
}
cudaError_t launch_ggx_dielectric(cudaStream_t st, size_t n, const ShadingSoA &sg, const GgxParamsDev &p, const float *rx,
                                  const float *ry, const DielectricOutDev &o, const Worklist &wl)
{
    if (p.ior.array && p.rough.array) k_ggx_dielectric_tol<true><<<grid_for(n), kBlockTol, 0, st>>>((uint32_t)n, sg, p, rx, ry, o, wl);
    else k_ggx_dielectric_tol<false><<<grid_for(n), kBlockTol, 0, st>>>((uint32_t)n, sg, p, rx, ry, o, wl);
    return cudaGetLastError();
}
cudaError_t launch_disney(cudaStream_t st, size_t n, const ShadingSoA &sg, const DisneyParamsDev &p, bool all_arrays,
                          const float *rx_s, const float *ry_s, const float *rx_d, const float *ry_d, const DisneyOutDev &o,
                          const Worklist &wl)
{
    if (all_arrays) k_disney_sample_eval_pdf_tol<true><<<grid_for(n), kBlockTol, 0, st>>>((uint32_t)n, sg, p, rx_s, ry_s, rx_d, ry_d, o, wl);
    else k_disney_sample_eval_pdf_tol<false><<<grid_for(n), kBlockTol, 0, st>>>((uint32_t)n, sg, p, rx_s, ry_s, rx_d, ry_d, o, wl);
    return cudaGetLastError();
}
cudaError_t launch_albedo_sweep(cudaStream_t st, const SweepGridDev &g, uint32_t n_cells, uint64_t seed, uint32_t k0, uint32_t k1,
                                double *table, const SweepWorklist &wl)
{
    const unsigned blocks = (unsigned)(((uint64_t)n_cells * 32ull + kSweepBlock - 1) / kSweepBlock);
    k_albedo_sweep_tol<<<blocks, kSweepBlock, 0, st>>>(g, n_cells, seed, k0, k1, table, wl);
    return cudaGetLastError();
}
cudaError_t launch_skin_profile(cudaStream_t st, size_t n, const SkinParamsDev &p, const float *rx, const ProfileOutDev &o,
                                const Worklist &wl)
{
    const auto al8 = [](const void *q) { return ((uintptr_t)q & 7u) == 0; };
    const bool pairs = n >= 2 && p.sss_scatter_dist.x && p.sss_scatter_dist.y && p.sss_scatter_dist.z && !p.sss_dist_multiplier.array &&
                       al8(p.sss_scatter_dist.x) && al8(p.sss_scatter_dist.y) && al8(p.sss_scatter_dist.z) && al8(rx) && al8(o.r) &&
                       al8(o.pdf) && al8(o.Rd.x) && al8(o.Rd.y) && al8(o.Rd.z) && al8(o.flags);
    if (pairs)
        k_skin_profile_tol_x2<<<grid_for((n + 1) / 2), kBlockTol, 0, st>>>((uint32_t)n, p.sss_scatter_dist.x, p.sss_scatter_dist.y,
                                                                          p.sss_scatter_dist.z, p.sss_dist_multiplier.value, rx, o, wl);
    else
        k_skin_profile_tol<<<grid_for(n), kBlockTol, 0, st>>>((uint32_t)n, p, rx, o, wl);
    return cudaGetLastError();
}

} // namespace tol
} // namespace rls
