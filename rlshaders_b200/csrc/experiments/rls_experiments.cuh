// rls_experiments.cuh -- kernels that were built, verified bit-exact, MEASURED SLOWER on B200 and therefore do
// not ship: the product library (librls_b200.so) is compiled without RLS_EXPERIMENTS and contains none of this.
// tools/build_experiments.sh builds librls_b200_experiments.so (the same translation unit with -DRLS_EXPERIMENTS)
// for the A/B runs and the parity tests of these forms; the environment switches below exist only in that build.
// Negative results: profiles/r01_ncu_summary.md (last table), profiles/r01_packed_experiment.md, DESIGN.md 8.
//
// Included by rls_b200.cu after the product kernels (it reuses dielectric_sample / disney_sample and the launch
// helpers) -- not a stand-alone header.
#pragma once
#include <stdlib.h>
#include "rls_pair.cuh"
#include "rls_tile.cuh"
#include "rls_packed.cuh"

// ---- environment switches (experiments build only), read once per context at rls_init
static int experiments_configure(rls_context *ctx)
{
    rls_experiments &x = ctx->exp;
    auto flag = [](const char *name) { const char *v = getenv(name); return v && atoi(v) != 0; };
    x.packed = flag("RLS_PACKED");
    x.paired = flag("RLS_PAIRED");
    x.tma = flag("RLS_TMA");
    if (const char *v = getenv("RLS_PERSISTENT")) { int k = atoi(v); x.persistent = k < -32 ? -32 : (k > 32 ? 32 : k); }
    if (const char *v = getenv("RLS_STAGGER_NS")) { int k = atoi(v); x.stagger_ns = (unsigned)(k < 0 ? 0 : (k > 1000000 ? 1000000 : k)); }
    x.disney_lobe_sort = flag("RLS_DISNEY_LOBE_SORT");
    x.gauss_scalar = flag("RLS_GAUSS_SCALAR");
    const unsigned long long negzero2 = 0x8000000080000000ull;   // {-0, -0} for the packed multiplies: a run-time value on purpose
    cudaError_t e = cudaMemcpyToSymbol(pk::c_negzero2, &negzero2, sizeof(negzero2));
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaMemcpyToSymbol(c_negzero2)");
    if (cudaMalloc((void **)&x.chunk_counter, sizeof(unsigned)) != cudaSuccess) x.chunk_counter = nullptr;
    return RLS_OK;
}
static void experiments_release(rls_context *ctx) { if (ctx->exp.chunk_counter) cudaFree(ctx->exp.chunk_counter); }

// ============================================================ rlGgx dielectric: persistent forms
// dielectric_sample (rls_b200.cu) with the lane-paired evaluation as an option (rls_pair.cuh, RLS_PAIRED=1)
template <bool kFast, bool kArrays, bool kPair>
RLS_DEV void exp_dielectric_sample(uint32_t i, const ShadingSoA &sg, const GgxParamsDev &p, const float *rx, const float *ry,
                                   const DielectricOutDev &o, unsigned long long *fallbacks)
{
    if (!kPair || !kFast) { dielectric_sample<kFast, kArrays>(i, sg, p, rx, ry, o, fallbacks); return; }
    Dielectric r;
    bool ok;
    {
        const Shading s = load_shading(sg, i);
        const float ior = fetch_t<kArrays>(p.ior, i), rough = fetch_t<kArrays>(p.rough, i), aniso = fetch(p.aniso, i);
        FpFast fp;
        r = pk::dielectric_unit_paired(fp, s, ior, rough, aniso, __ldg(rx + i), __ldg(ry + i), p.ndf != 0);
        ok = fp.ok();
    }
    if (!ok) {
        const Shading s = load_shading<true>(sg, i);
        FpExact fp;
        r = dielectric_unit(fp, s, fetch_t<kArrays, true>(p.ior, i), fetch_t<kArrays, true>(p.rough, i), fetch<true>(p.aniso, i),
                            __ldcg(rx + i), __ldcg(ry + i), p.ndf != 0);
        atomicAdd(fallbacks, 1ull);
    }
    o.fresnel[i] = r.F;
    store3(o.wi_r, i, r.wi_r);
    o.f_r[i] = r.f_r;
    o.pdf_r[i] = r.pdf_r;
    store3(o.wi_t, i, r.wi_t);
    o.f_t[i] = r.f_t;
    o.weight_t[i] = r.w_t;
    o.flags[i] = r.flags;
}
template <bool kFast, bool kArrays, bool kPair>
__global__ void __launch_bounds__(kBlockGgx, RLS_GGX_MIN_BLOCKS)
k_ggx_dielectric_exp(size_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, DielectricOutDev o,
                     unsigned long long *fallbacks)
{
    RLS_INDEX();
    exp_dielectric_sample<kFast, kArrays, kPair>(i, sg, p, rx, ry, o, fallbacks);
}

// Persistent form: grid = resident CTAs (a multiple of the SM count), every thread strides over the
// batch.  No CTA turnover (a CTA slot of the plain kernel stays partly empty until its slowest warp
// retires).  Measured SLOWER (static split, see k_ggx_dielectric_dynamic below); kept for A/B runs.
template <bool kFast, bool kArrays, bool kPair>
__global__ void __launch_bounds__(kBlockGgx, RLS_GGX_MIN_BLOCKS)
k_ggx_dielectric_persistent(size_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, DielectricOutDev o,
                            unsigned long long *fallbacks, unsigned stagger_ns, unsigned sm_count)
{
    // phase-shift the CTAs that share an SM (the first wave is dealt round-robin: CTA b runs on SM b % sm_count)
    if (stagger_ns) __nanosleep((blockIdx.x / sm_count) * stagger_ns);
    const uint32_t stride = gridDim.x * blockDim.x;
#pragma unroll 1
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += stride)
        exp_dielectric_sample<kFast, kArrays, kPair>(i, sg, p, rx, ry, o, fallbacks);
}

// Persistent form with DYNAMIC work distribution: every warp draws 128-sample chunks from a global counter
// (the next index is fetched while the current chunk computes).  The static grid-stride form above loses
// 13 %: the warp arbiter is unfair, CTAs in favoured slots finish their share early and the SM runs its tail at
// low occupancy (ncu: 40.5 % average active warps with 9 resident CTAs per SM, 49.1 % for the plain kernel).
template <bool kFast, bool kArrays, bool kPair>
__global__ void __launch_bounds__(kBlockGgx, RLS_GGX_MIN_BLOCKS)
k_ggx_dielectric_dynamic(size_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, DielectricOutDev o,
                         unsigned long long *fallbacks, unsigned *counter)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_chunks = (uint32_t)((n + 127u) / 128u);
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(counter, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    while (c < n_chunks) {
        uint32_t nxt = 0;
        if (lane == 0) nxt = atomicAdd(counter, 1u);
#pragma unroll 1
        for (uint32_t j = 0; j < 4u; j++) {
            const uint32_t i = c * 128u + j * 32u + lane;
            if (i < (uint32_t)n) exp_dielectric_sample<kFast, kArrays, kPair>(i, sg, p, rx, ry, o, fallbacks);
        }
        c = __shfl_sync(0xffffffffu, nxt, 0);
    }
}

// Persistent, TMA-staged form of k_ggx_dielectric (rls_tile.cuh): grid = resident CTAs, every CTA
// walks whole 256-sample tiles; the inputs of the next tile are in flight while this one computes.
// Fast policy only (the launch site keeps the plain kernel for the exact policy, unaligned arrays
// and the ragged tail); a sample whose operands left the window reloads its inputs from global
// memory and is re-run with FpExact, as in the plain kernel.
namespace dielectric_slots {
enum In { U = 0, V = 3, N = 6, WO = 9, RX = 12, RY = 13, ROUGH = 14, IOR = 15, ANISO = 16, BACK = 17, kIn = 18 };
enum Out { FRESNEL = 0, WI_R = 1, F_R = 4, PDF_R = 5, WI_T = 6, F_T = 9, WEIGHT_T = 10, FLAGS = 11, kOut = 12 };
}
template <bool kArrays, bool kPair>
__global__ void __launch_bounds__(tile::kTile, 4)
k_ggx_dielectric_tma(uint32_t n_tiles, const __grid_constant__ tile::Arrays arr, ShadingSoA sg, GgxParamsDev p,
                     const float *rx, const float *ry, unsigned long long *fallbacks)
{
    using namespace dielectric_slots;
    __shared__ __align__(128) unsigned char smem[tile::Pipe<kIn, kOut>::kSmemBytes];
    __shared__ uint64_t bar;
    tile::Pipe<kIn, kOut> pipe;
    pipe.init(&arr, smem, &bar);
    const uint32_t tid = threadIdx.x;
    uint32_t t = blockIdx.x;
    if (t < n_tiles) pipe.issue_loads(t);
#pragma unroll 1
    for (; t < n_tiles; t += gridDim.x) {
        pipe.wait_inputs();
#define RLS_IN(k) pipe.in_slot(k)[tid]
        Shading s;
        s.U = mk3(RLS_IN(U), RLS_IN(U + 1), RLS_IN(U + 2));
        s.V = mk3(RLS_IN(V), RLS_IN(V + 1), RLS_IN(V + 2));
        s.N = mk3(RLS_IN(N), RLS_IN(N + 1), RLS_IN(N + 2));
        s.wo = mk3(RLS_IN(WO), RLS_IN(WO + 1), RLS_IN(WO + 2));
        s.backfacing = sg.backfacing ? (reinterpret_cast<const uint8_t *>(pipe.in_slot(BACK))[tid] != 0) : false;
        const float ior = (kArrays || p.ior.array) ? RLS_IN(IOR) : p.ior.value;
        const float rough = (kArrays || p.rough.array) ? RLS_IN(ROUGH) : p.rough.value;
        const float aniso = p.aniso.array ? RLS_IN(ANISO) : p.aniso.value;
        const float u1 = RLS_IN(RX), u2 = RLS_IN(RY);
#undef RLS_IN
        pipe.inputs_consumed(t + gridDim.x, n_tiles);
        Dielectric r;
        bool ok;
        {
            FpFast fp;
            r = kPair ? pk::dielectric_unit_paired(fp, s, ior, rough, aniso, u1, u2, p.ndf != 0)
                      : dielectric_unit(fp, s, ior, rough, aniso, u1, u2, p.ndf != 0);
            ok = fp.ok();
        }
        if (!ok) {
            const uint32_t i = t * tile::kTile + tid;
            const Shading se = load_shading<true>(sg, i);
            FpExact fp;
            r = dielectric_unit(fp, se, fetch<true>(p.ior, i), fetch<true>(p.rough, i), fetch<true>(p.aniso, i),
                                __ldcg(rx + i), __ldcg(ry + i), p.ndf != 0);
            atomicAdd(fallbacks, 1ull);
        }
        pipe.begin_store();
#define RLS_OUT(k) pipe.out_slot(k)[tid]
        RLS_OUT(FRESNEL) = r.F;
        RLS_OUT(WI_R) = r.wi_r.x; RLS_OUT(WI_R + 1) = r.wi_r.y; RLS_OUT(WI_R + 2) = r.wi_r.z;
        RLS_OUT(F_R) = r.f_r;
        RLS_OUT(PDF_R) = r.pdf_r;
        RLS_OUT(WI_T) = r.wi_t.x; RLS_OUT(WI_T + 1) = r.wi_t.y; RLS_OUT(WI_T + 2) = r.wi_t.z;
        RLS_OUT(F_T) = r.f_t;
        RLS_OUT(WEIGHT_T) = r.w_t;
        RLS_OUT(FLAGS) = __uint_as_float(r.flags);
#undef RLS_OUT
        pipe.end_store(t);
    }
    pipe.finish();
}

// Two samples per thread, packed f32x2 arithmetic (rls_packed.cuh): thread t owns samples 2t and
// 2t + 1 of every SoA array (one 64-bit load / store each).  Fast policy only; a pair whose
// tracker left the window is re-run lane by lane with the scalar FpExact unit.
// EXPERIMENT, off by default (RLS_PACKED=1): 22 % fewer issue slots per sample (1092 vs 1396 per
// 32 samples) but 128 registers/thread leave 4 warps per scheduler, issue utilisation drops from
// 85 % to 48 % and the kernel runs at 16.2 instead of 22.6 G samples/s; capping registers at
// 80 / 64 spills and is slower still (15.9 / 14.5).  profiles/r01_packed_experiment.md.
#ifndef RLS_PACKED_MIN_BLOCKS
#define RLS_PACKED_MIN_BLOCKS 2
#endif
template <bool kArrays>
__global__ void __launch_bounds__(kBlock, RLS_PACKED_MIN_BLOCKS)
k_ggx_dielectric2(size_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, DielectricOutDev o,
                  unsigned long long *fallbacks)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)((n + 1) / 2)) return;
    const bool tail = 2u * t + 1u >= (uint32_t)n;
    const pk::V2 U = pk::ld2(sg.U, t, tail), V = pk::ld2(sg.V, t, tail), N = pk::ld2(sg.N, t, tail), wo = pk::ld2(sg.wo, t, tail);
    pk::B2 back; back.a = false; back.b = false;
    if (sg.backfacing) {
        back.a = __ldg(sg.backfacing + 2u * t) != 0;
        back.b = tail ? back.a : (__ldg(sg.backfacing + 2u * t + 1u) != 0);
    }
    const pk::F2 ior = kArrays ? pk::ld2(p.ior.array, t, tail) : pk::fetch2(p.ior, t, tail);
    const pk::F2 rough = kArrays ? pk::ld2(p.rough.array, t, tail) : pk::fetch2(p.rough, t, tail);
    const pk::F2 aniso = pk::fetch2(p.aniso, t, tail);
    const pk::F2 u1 = pk::ld2(rx, t, tail), u2 = pk::ld2(ry, t, tail);
    pk::Fp2 fp;
    pk::Dielectric2 r = pk::dielectric_unit(fp, U, V, N, wo, back, ior, rough, aniso, u1, u2);
    if (!fp.ok()) {
        atomicAdd(fallbacks, tail ? 1ull : 2ull);
#pragma unroll 1
        for (int l = 0; l < 2; l++) {
            Shading s;
            s.U = l ? pk::lane1(U) : pk::lane0(U); s.V = l ? pk::lane1(V) : pk::lane0(V);
            s.N = l ? pk::lane1(N) : pk::lane0(N); s.wo = l ? pk::lane1(wo) : pk::lane0(wo);
            s.backfacing = l ? back.b : back.a;
            FpExact fe;
            const Dielectric e = dielectric_unit(fe, s, l ? pk::hi(ior) : pk::lo(ior), l ? pk::hi(rough) : pk::lo(rough),
                                                 l ? pk::hi(aniso) : pk::lo(aniso), l ? pk::hi(u1) : pk::lo(u1),
                                                 l ? pk::hi(u2) : pk::lo(u2), false);
#define RLS_SET_LANE(dst, val) dst = l ? pk::mk(pk::lo(dst), (val)) : pk::mk((val), pk::hi(dst))
            RLS_SET_LANE(r.F, e.F); RLS_SET_LANE(r.f_r, e.f_r); RLS_SET_LANE(r.pdf_r, e.pdf_r);
            RLS_SET_LANE(r.f_t, e.f_t); RLS_SET_LANE(r.w_t, e.w_t);
            RLS_SET_LANE(r.wi_r.x, e.wi_r.x); RLS_SET_LANE(r.wi_r.y, e.wi_r.y); RLS_SET_LANE(r.wi_r.z, e.wi_r.z);
            RLS_SET_LANE(r.wi_t.x, e.wi_t.x); RLS_SET_LANE(r.wi_t.y, e.wi_t.y); RLS_SET_LANE(r.wi_t.z, e.wi_t.z);
#undef RLS_SET_LANE
            if (l) r.flags1 = e.flags; else r.flags0 = e.flags;
        }
    }
    pk::st2(o.fresnel, t, tail, r.F);
    pk::st2(o.wi_r, t, tail, r.wi_r);
    pk::st2(o.f_r, t, tail, r.f_r);
    pk::st2(o.pdf_r, t, tail, r.pdf_r);
    pk::st2(o.wi_t, t, tail, r.wi_t);
    pk::st2(o.f_t, t, tail, r.f_t);
    pk::st2(o.weight_t, t, tail, r.w_t);
    if (tail) o.flags[2u * t] = r.flags0;
    else reinterpret_cast<uint2 *>(o.flags)[t] = make_uint2(r.flags0, r.flags1);
}

// ============================================================ rlDisney: CTA-level lobe partition
// Warp-uniform lobe selection (BASELINE north_star): a stable partition of the CTA's samples by specular lobe, so that
// the ~10 % of samples that take the GTR1 (clearcoat) lobe sit together in the CTA's last warp(s) instead of making
// almost every warp (1 - 0.9^32 = 97 %) run the GTR1-only code (general powf) for three lanes AND the GTR2-only code
// (visible-normal sampling) for the rest.  Thread t then works on sample base + perm[t]; GTR2 samples keep their order,
// so a warp's loads span ~36 consecutive samples instead of 32.  The predicate only steers the grouping (results do
// not depend on it), so it is the approximate rx (c' + 1) < 1 rather than the unit's exact rx < 1 / (c' + 1).
// MEASURED (tools/disney_ab.py, 2^26 samples, every parameter per sample): 16.02 G samples/s against 16.68 without it --
// the GTR1-only code the other warps skip (98 slots of 1768 per warp, ncu) is worth less than the partition costs (two CTA barriers
// before the first load of the unit can issue, ~50 slots, 29 loads and 15 stores per sample over two cache lines).
// Kept behind RLS_DISNEY_LOBE_SORT=1 with its test (tests/test_gpu_parity.py::test_disney_lobe_partition_is_invisible).
template <bool kArrays>
RLS_DEV uint32_t disney_lobe_partition(size_t n, const DisneyParamsDev &p, const float *rx_s)
{
    __shared__ uint16_t perm[kBlock];
    __shared__ uint32_t gtr1_in_warp[kBlock / 32];
    const uint32_t t = threadIdx.x, lane = t & 31u, w = t >> 5, base = blockIdx.x * blockDim.x;
    bool gtr1 = true;                                   // samples past the end are grouped with the last warp
    if (base + t < (uint32_t)n) {
        const float c = fetch_t<kArrays>(p.clearcoat, base + t) * 0.25f;
        gtr1 = !(__ldg(rx_s + base + t) * (c + 1.0f) < 1.0f);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, gtr1);
    if (lane == 0) gtr1_in_warp[w] = __popc(b);
    __syncthreads();
    uint32_t before = 0, total = 0;                     // GTR1 samples in the warps before this one / in the CTA
#pragma unroll
    for (uint32_t k = 0; k < kBlock / 32; k++) {
        const uint32_t c = gtr1_in_warp[k];
        before += k < w ? c : 0u;
        total += c;
    }
    const uint32_t mine = __popc(b & ((1u << lane) - 1u));
    const uint32_t pos = gtr1 ? (kBlock - total) + before + mine : (t - before - mine);
    perm[pos] = (uint16_t)t;
    __syncthreads();
    return base + perm[t];
}

// ============================================================ launch hooks
// ---- persistent TMA-staged launches (rls_tile.cuh): helpers shared by the fused entry points
static inline rls_cvec3 adv(rls_cvec3 v, size_t k) { rls_cvec3 o = { v.x ? v.x + k : nullptr, v.y ? v.y + k : nullptr, v.z ? v.z + k : nullptr }; return o; }
static inline rls_vec3 adv(rls_vec3 v, size_t k) { rls_vec3 o = { v.x ? v.x + k : nullptr, v.y ? v.y + k : nullptr, v.z ? v.z + k : nullptr }; return o; }
template <typename T> static inline T *adv(T *q, size_t k) { return q ? q + k : nullptr; }
static inline rls_param1 adv(rls_param1 q, size_t k) { q.array = adv(q.array, k); return q; }
static inline rls_param3 adv(rls_param3 q, size_t k) { q.array = adv(q.array, k); return q; }
static inline rls_shading_soa adv(const rls_shading_soa &s, size_t k)
{
    rls_shading_soa o; o.U = adv(s.U, k); o.V = adv(s.V, k); o.N = adv(s.N, k); o.wo = adv(s.wo, k);
    o.backfacing = adv(s.backfacing, k); return o;
}
struct TileArrays {
    tile::Arrays a;
    bool aligned = true;
    TileArrays() { memset(&a, 0, sizeof(a)); }
    void in(int slot, const void *q, int elem = 4)
    {
        a.in[slot] = q; a.in_elem[slot] = (uint8_t)elem;
        if (q) { a.in_bytes_per_tile += (uint32_t)(tile::kTile * elem); aligned = aligned && !((uintptr_t)q & 15u); }
    }
    void in3(int slot, const rls_cvec3 &v) { in(slot, v.x); in(slot + 1, v.y); in(slot + 2, v.z); }
    void out(int slot, void *q) { a.out[slot] = q; if (q) aligned = aligned && !((uintptr_t)q & 15u); }
    void out3(int slot, const rls_vec3 &v) { out(slot, v.x); out(slot + 1, v.y); out(slot + 2, v.z); }
};
// grid of a persistent kernel: every CTA resident at once (a multiple of the SM count), never more CTAs than tiles
template <typename K> static unsigned persistent_grid(rls_context *ctx, K kernel, size_t n_tiles)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, tile::kTile, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const size_t resident = (size_t)per_sm * (size_t)ctx->sm_count;
    return (unsigned)(n_tiles < resident ? n_tiles : resident);
}

static int launch_ggx_dielectric(rls_context *ctx, cudaStream_t st, size_t n, const rls_shading_soa *sg,
                                 const rls_ggx_params *p, const float *rx, const float *ry, const rls_ggx_dielectric_out *o);
static int launch_ggx_dielectric_tma(rls_context *ctx, cudaStream_t st, size_t n, const rls_shading_soa *sg,
                                     const rls_ggx_params *p, const float *rx, const float *ry, const rls_ggx_dielectric_out *o,
                                     bool *taken)
{
    using namespace dielectric_slots;
    *taken = false;
    if (!ctx->exp.tma || ctx->arith != RLS_ARITH_FAST || n < (size_t)tile::kTile) return RLS_OK;
    const GgxParamsDev pd = dev(*p);
    TileArrays ta;
    ta.in3(U, sg->U); ta.in3(V, sg->V); ta.in3(N, sg->N); ta.in3(WO, sg->wo);
    ta.in(RX, rx); ta.in(RY, ry); ta.in(ROUGH, pd.rough.array); ta.in(IOR, pd.ior.array); ta.in(ANISO, pd.aniso.array);
    ta.in(BACK, sg->backfacing, 1);
    ta.out(FRESNEL, o->fresnel); ta.out3(WI_R, o->wi_r); ta.out(F_R, o->f_r); ta.out(PDF_R, o->pdf_r);
    ta.out3(WI_T, o->wi_t); ta.out(F_T, o->f_t); ta.out(WEIGHT_T, o->weight_t); ta.out(FLAGS, o->flags);
    if (!ta.aligned) return RLS_OK;
    *taken = true;
    const size_t n_tiles = n / tile::kTile;
    const bool arrays = pd.ior.array && pd.rough.array;
#define RLS_DIELECTRIC_TMA(A, P) \
    k_ggx_dielectric_tma<A, P><<<persistent_grid(ctx, k_ggx_dielectric_tma<A, P>, n_tiles), tile::kTile, 0, st>>>( \
        (uint32_t)n_tiles, ta.a, sh(*sg), pd, rx, ry, ctx->fallbacks)
    if (arrays && ctx->exp.paired) RLS_DIELECTRIC_TMA(true, true);
    else if (arrays) RLS_DIELECTRIC_TMA(true, false);
    else if (ctx->exp.paired) RLS_DIELECTRIC_TMA(false, true);
    else RLS_DIELECTRIC_TMA(false, false);
#undef RLS_DIELECTRIC_TMA
    RLS_LAUNCH_CHECK(ctx);
    const size_t done = n_tiles * tile::kTile;
    if (done == n) return RLS_OK;
    // ragged tail: the plain kernel on the remaining n - done < 256 samples
    const rls_shading_soa sg2 = adv(*sg, done);
    rls_ggx_params p2 = *p;
    p2.specularRoughness = adv(p->specularRoughness, done); p2.ior = adv(p->ior, done); p2.anisotropic = adv(p->anisotropic, done);
    p2.KsColor = adv(p->KsColor, done);
    rls_ggx_dielectric_out o2;
    o2.fresnel = adv(o->fresnel, done); o2.wi_r = adv(o->wi_r, done); o2.f_r = adv(o->f_r, done); o2.pdf_r = adv(o->pdf_r, done);
    o2.wi_t = adv(o->wi_t, done); o2.f_t = adv(o->f_t, done); o2.weight_t = adv(o->weight_t, done); o2.flags = adv(o->flags, done);
    return launch_ggx_dielectric(ctx, st, n - done, &sg2, &p2, rx + done, ry + done, &o2);
}

// The dielectric launch of the experiments build: TMA-staged / packed / persistent / dynamic / paired forms when
// their switch is on; *taken = false sends the call to the product kernel.
static int experiments_launch_ggx_dielectric(rls_context *ctx, cudaStream_t st, size_t n, const rls_shading_soa *sg,
                                             const rls_ggx_params *p, const float *rx, const float *ry,
                                             const rls_ggx_dielectric_out *o, bool *taken)
{
    const rls_experiments &x = ctx->exp;
    int rc = launch_ggx_dielectric_tma(ctx, st, n, sg, p, rx, ry, o, taken);
    if (*taken || rc != RLS_OK) return rc;
    const DielectricOutDev d = dev(*o);
    const GgxParamsDev pd = dev(*p);
    const bool arrays = pd.ior.array && pd.rough.array;
    const bool fast = ctx->arith == RLS_ARITH_FAST;
    // Packed two-samples-per-thread kernel: fast policy, shipped sampler, every array 8-byte aligned.
    if (fast && !pd.ndf && x.packed &&
        aligned8({ sg->U.x, sg->U.y, sg->U.z, sg->V.x, sg->V.y, sg->V.z, sg->N.x, sg->N.y, sg->N.z, sg->wo.x, sg->wo.y, sg->wo.z,
                   pd.ior.array, pd.rough.array, pd.aniso.array, rx, ry, d.fresnel, d.wi_r.x, d.wi_r.y, d.wi_r.z, d.f_r, d.pdf_r,
                   d.wi_t.x, d.wi_t.y, d.wi_t.z, d.f_t, d.weight_t, d.flags })) {
        const unsigned grid = (unsigned)(((n + 1) / 2 + kBlock - 1) / kBlock);
        if (arrays) k_ggx_dielectric2<true><<<grid, kBlock, 0, st>>>(n, sh(*sg), pd, rx, ry, d, ctx->fallbacks);
        else k_ggx_dielectric2<false><<<grid, kBlock, 0, st>>>(n, sh(*sg), pd, rx, ry, d, ctx->fallbacks);
        *taken = true;
        RLS_LAUNCH_CHECK(ctx);
        return RLS_OK;
    }
    if (!fast || (!x.paired && x.persistent == 0)) return RLS_OK;          // nothing experimental requested
    *taken = true;
#define RLS_DIELECTRIC_LAUNCH(A, P) \
    do { if (x.persistent < 0 && x.chunk_counter) { \
             cudaMemsetAsync(x.chunk_counter, 0, sizeof(unsigned), st); \
             k_ggx_dielectric_dynamic<true, A, P><<<ctx->sm_count * (-x.persistent), kBlockGgx, 0, st>>>(n, sh(*sg), pd, rx, ry, d, ctx->fallbacks, x.chunk_counter); } \
         else if (x.persistent > 0 && grid_for(n, kBlockGgx) > (unsigned)(ctx->sm_count * x.persistent)) \
             k_ggx_dielectric_persistent<true, A, P><<<ctx->sm_count * x.persistent, kBlockGgx, 0, st>>>(n, sh(*sg), pd, rx, ry, d, ctx->fallbacks, x.stagger_ns, (unsigned)ctx->sm_count); \
         else k_ggx_dielectric_exp<true, A, P><<<grid_for(n, kBlockGgx), kBlockGgx, 0, st>>>(n, sh(*sg), pd, rx, ry, d, ctx->fallbacks); } while (0)
    if (arrays && x.paired) RLS_DIELECTRIC_LAUNCH(true, true);
    else if (arrays) RLS_DIELECTRIC_LAUNCH(true, false);
    else if (x.paired) RLS_DIELECTRIC_LAUNCH(false, true);
    else RLS_DIELECTRIC_LAUNCH(false, false);
#undef RLS_DIELECTRIC_LAUNCH
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

// rlDisney with the CTA-level lobe partition (RLS_DISNEY_LOBE_SORT=1, fast policy): pays only when lobes are mixed
// inside a CTA, i.e. clearcoat varies per sample or is non-zero.
template <bool kArrays>
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_disney_sample_eval_pdf_sorted(size_t n, ShadingSoA sg, DisneyParamsDev p, const float *rx_s, const float *ry_s,
                                const float *rx_d, const float *ry_d, DisneyOutDev o, unsigned long long *fallbacks)
{
    rlm::smem_tables_init();
    const uint32_t i = disney_lobe_partition<kArrays>(n, p, rx_s);
    if (i >= (uint32_t)n) return;
    disney_sample<true, kArrays>(i, sg, p, rx_s, ry_s, rx_d, ry_d, o, fallbacks);
}
static int experiments_launch_disney(rls_context *ctx, cudaStream_t st, size_t n, const rls_shading_soa *sg,
                                     const DisneyParamsDev &pd, bool arrays, const float *rx_s, const float *ry_s,
                                     const float *rx_d, const float *ry_d, const DisneyOutDev &d, bool *taken)
{
    *taken = ctx->exp.disney_lobe_sort && ctx->arith == RLS_ARITH_FAST && (pd.clearcoat.array || pd.clearcoat.value != 0.0f);
    if (!*taken) return RLS_OK;
    if (arrays) k_disney_sample_eval_pdf_sorted<true><<<grid_for(n), kBlock, 0, st>>>(n, sh(*sg), pd, rx_s, ry_s, rx_d, ry_d, d, ctx->fallbacks);
    else k_disney_sample_eval_pdf_sorted<false><<<grid_for(n), kBlock, 0, st>>>(n, sh(*sg), pd, rx_s, ry_s, rx_d, ry_d, d, ctx->fallbacks);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
