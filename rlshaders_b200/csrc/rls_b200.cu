// rls_b200.cu -- kernels and C ABI (include/rls_b200.h) of the B200-native rlShaders BSDF path.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -Xcompiler -fPIC -shared -o rlshaders_b200/librls_b200.so rls_b200.cu
//
// Layout: one thread = one shading sample; all per-sample state lives in registers; inputs
// and outputs are structure-of-arrays so that a warp's 32 loads/stores of one component are
// one fully coalesced 128-byte line.  There is no CPU path in this library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <initializer_list>
#include <new>

#include "../../include/rls_b200.h"
#include "rls_math.cuh"
#include "rls_libm.cuh"
#include "rls_ggx.cuh"
#include "rls_disney.cuh"
#include "rls_profile.cuh"
#include "rls_fused.cuh"
#include "rls_callers.cuh"
#include "rls_kernel_args.cuh"
#include "rls_tol_launch.cuh"
#include "rls_sweep.cuh"

using namespace rls;

// ============================================================== context
#ifndef RLS_STAGES
#define RLS_STAGES 3
#endif
static constexpr int kStages = RLS_STAGES;  // host-staging pipeline depth
// Launch shape: 256 threads x >= 4 resident CTAs per SM (<= 64 registers/thread).  Chosen from
// sweeps on B200 (tools/sweep_variants.sh; profiles/r01_launch_sweep.txt): the kernels are
// instruction-issue bound, so occupancy beyond ~50 % does not help and tighter register caps
// spill; among the 64-register shapes the smaller CTA wins by 2-7 % (finer-grained tail).
#ifndef RLS_BLOCK
#define RLS_BLOCK 256
#endif
#ifndef RLS_MIN_BLOCKS
#define RLS_MIN_BLOCKS 4
#endif
static constexpr int kBlock = RLS_BLOCK;
// The two fused rlGgx kernels fit 56 registers once the exact re-run reloads its inputs, so 9 CTAs of
// 128 threads (36 warps / SM) are resident instead of 4 x 256 (32 warps); smaller CTAs also turn over
// more evenly.  B200 sweep (profiles/r01_launch_sweep.txt, third sweep): dielectric 22.66 -> 22.96, conductor
// 23.37 -> 24.07 G samples/s; rlDisney and the skin profile keep 256 x 4.
#ifndef RLS_GGX_BLOCK
#define RLS_GGX_BLOCK 128
#endif
#ifndef RLS_GGX_MIN_BLOCKS
#define RLS_GGX_MIN_BLOCKS 9
#endif
static constexpr int kBlockGgx = RLS_GGX_BLOCK;
// The fused skin-profile kernel: 128 x 9 as well (fourth sweep: 36.8 vs 35.5 G samples/s at 256 x 4).
#ifndef RLS_SKIN_BLOCK
#define RLS_SKIN_BLOCK 128
#endif
#ifndef RLS_SKIN_MIN_BLOCKS
#define RLS_SKIN_MIN_BLOCKS 9
#endif
static constexpr int kBlockSkin = RLS_SKIN_BLOCK;
// ... and the albedo sweep (128 threads = 4 warps = 4 cells per CTA, rls_sweep.cuh).
#ifndef RLS_SWEEP_MIN_BLOCKS
#define RLS_SWEEP_MIN_BLOCKS 9
#endif
static_assert(kBlock >= 96 && kBlockSkin >= 96, "rlm::smem_tables_init() fills 96 table entries with one thread each");

#ifdef RLS_EXPERIMENTS
// Only in librls_b200_experiments.so (tools/build_experiments.sh): A/B switches of kernels that measured slower.
struct rls_experiments {
    bool paired = false, tma = false, packed = false, disney_lobe_sort = false, gauss_scalar = false;
    int persistent = 0;              // CTAs per SM of the persistent dielectric kernel (negative: dynamic chunks)
    unsigned stagger_ns = 0;
    unsigned *chunk_counter = nullptr;
};
#endif
struct rls_context {
    int          device = 0;
    cudaStream_t stream = nullptr;
    bool         own_stream = false;
    std::string  err;
    uint64_t     launches = 0;
    int          arith = RLS_ARITH_FAST;          // policy of the fused kernels (rls_fp.cuh)
    int          sm_count = 0;
    unsigned long long *fallbacks = nullptr;      // device counters: [0] samples re-run (FpFast -> FpExact; tolerance policy ->
                                                  // bit-exact policy), [1] scratch (FpExact re-runs inside a tolerance re-run)
    // RLS_ARITH_TOLERANT: one re-run list per stream the fused kernels are launched on (slot 0 = `stream`,
    // 1.. = the host-staging streams), allocated on first use and grown when a larger batch arrives
    tol::Worklist worklist[kStages + 1] = {};
    tol::SweepWorklist sweep_worklist = {};       // RLS_ARITH_TOLERANT sweep: (cell, k) of the band samples
#ifdef RLS_EXPERIMENTS
    rls_experiments exp;                          // switches of the measured-slower forms (experiments/rls_experiments.cuh)
#endif
    // host-staging resources (lazily created by the *_host entry points)
    cudaStream_t stage_stream[kStages] = {};
    cudaEvent_t  stage_done[kStages] = {};
    void        *stage_buf[kStages] = {};
    size_t       stage_bytes = 0;
};

#ifdef RLS_EXPERIMENTS
static int experiments_configure(rls_context *ctx);      // experiments/rls_experiments.cuh
static void experiments_release(rls_context *ctx);
#endif

static thread_local std::string g_init_error;
// The kernels that have no tolerance form run the fast (bit-exact) policy under RLS_ARITH_TOLERANT.
static inline bool uses_fast_policy(const rls_context *ctx) { return ctx->arith != RLS_ARITH_EXACT; }

static int fail(rls_context *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg; else g_init_error = msg;
    return code;
}
static int cuda_fail(rls_context *ctx, cudaError_t e, const char *what)
{
    return fail(ctx, e == cudaErrorMemoryAllocation ? RLS_ERR_OUT_OF_MEMORY : RLS_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}
#define RLS_CUDA(ctx, call)                                            \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" int rls_abi_version(void) { return RLS_B200_ABI_VERSION; }

extern "C" const char *rls_node_name(int i)
{
    static const char *names[] = { "rlGgx", "rlDisney", "rlSkin" };   // src/_PluginMain.cpp:8-13
    return (i >= 0 && i < 3) ? names[i] : nullptr;
}

extern "C" int rls_init(int device, void *stream, rls_context **out_ctx)
{
    if (!out_ctx) return fail(nullptr, RLS_ERR_INVALID_ARGUMENT, "rls_init: out_ctx is NULL");
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, RLS_ERR_NO_DEVICE, std::string("rls_init: no CUDA device (") +
                    cudaGetErrorString(e) + "); this library has no CPU path");
    if (device < 0 || device >= count)
        return fail(nullptr, RLS_ERR_INVALID_ARGUMENT, "rls_init: device index out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, RLS_ERR_NO_DEVICE, std::string("rls_init: device '") + prop.name +
                    "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                    "; kernels are built for sm_100a only");
    rls_context *ctx = new (std::nothrow) rls_context();
    if (!ctx) return fail(nullptr, RLS_ERR_OUT_OF_MEMORY, "rls_init: host allocation failed");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    DeviceGuard guard(device);
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaStreamCreateWithFlags"); }
        ctx->own_stream = true;
    }
#ifdef RLS_EXPERIMENTS
    if (experiments_configure(ctx) != RLS_OK) {
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return RLS_ERR_CUDA;
    }
#endif
    e = cudaMalloc((void **)&ctx->fallbacks, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->fallbacks, 0, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return cuda_fail(nullptr, e, "cudaMalloc(fallback counter)");
    }
    *out_ctx = ctx;
    return RLS_OK;
}

extern "C" int rls_set_arith_policy(rls_context *ctx, int policy)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (policy != RLS_ARITH_FAST && policy != RLS_ARITH_EXACT && policy != RLS_ARITH_TOLERANT)
        return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "rls_set_arith_policy: policy must be RLS_ARITH_FAST, RLS_ARITH_EXACT or RLS_ARITH_TOLERANT");
    ctx->arith = policy;
    return RLS_OK;
}

extern "C" int rls_get_arith_policy(const rls_context *ctx) { return ctx ? ctx->arith : RLS_ERR_INVALID_ARGUMENT; }
extern "C" void *rls_stream(const rls_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int rls_device_alloc(rls_context *ctx, size_t bytes, void **out_ptr)
{
    if (!ctx || !out_ptr) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    RLS_CUDA(ctx, cudaMalloc(out_ptr, bytes));
    return RLS_OK;
}
extern "C" int rls_device_free(rls_context *ctx, void *ptr)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    RLS_CUDA(ctx, cudaFree(ptr));
    return RLS_OK;
}
extern "C" int rls_memcpy_to_host(rls_context *ctx, void *host_dst, const void *device_src, size_t bytes)
{
    if (!ctx || (bytes && (!host_dst || !device_src))) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    RLS_CUDA(ctx, cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RLS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RLS_OK;
}

extern "C" int rls_fallback_count(rls_context *ctx, uint64_t *out_count, int reset)
{
    if (!ctx || !out_count) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    unsigned long long v = 0;
    RLS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < kStages; b++)
        if (ctx->stage_stream[b]) RLS_CUDA(ctx, cudaStreamSynchronize(ctx->stage_stream[b]));
    RLS_CUDA(ctx, cudaMemcpy(&v, ctx->fallbacks, sizeof(v), cudaMemcpyDeviceToHost));
    if (reset) RLS_CUDA(ctx, cudaMemset(ctx->fallbacks, 0, sizeof(v)));
    *out_count = (uint64_t)v;
    return RLS_OK;
}

extern "C" int rls_shutdown(rls_context *ctx)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int b = 0; b < kStages; b++) {
        if (ctx->stage_stream[b]) { cudaStreamSynchronize(ctx->stage_stream[b]); cudaStreamDestroy(ctx->stage_stream[b]); }
        if (ctx->stage_done[b]) cudaEventDestroy(ctx->stage_done[b]);
        if (ctx->stage_buf[b]) cudaFree(ctx->stage_buf[b]);
    }
    if (ctx->fallbacks) cudaFree(ctx->fallbacks);
    for (tol::Worklist &w : ctx->worklist) { if (w.list) cudaFree(w.list); if (w.count) cudaFree(w.count); }
    if (ctx->sweep_worklist.list) cudaFree(ctx->sweep_worklist.list);
    if (ctx->sweep_worklist.count) cudaFree(ctx->sweep_worklist.count);
#ifdef RLS_EXPERIMENTS
    experiments_release(ctx);
#endif
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RLS_OK;
}

extern "C" int rls_synchronize(rls_context *ctx)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    RLS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RLS_OK;
}

extern "C" const char *rls_last_error_string(const rls_context *ctx)
{
    return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" uint64_t rls_kernel_launch_count(const rls_context *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int rls_host_alloc(rls_context *ctx, size_t bytes, void **out_ptr)
{
    if (!ctx || !out_ptr) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    RLS_CUDA(ctx, cudaHostAlloc(out_ptr, bytes, cudaHostAllocDefault));
    return RLS_OK;
}
extern "C" int rls_host_free(rls_context *ctx, void *ptr)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    RLS_CUDA(ctx, cudaFreeHost(ptr));
    return RLS_OK;
}

// ==================================================== ABI struct -> device struct
static inline CV3 cv(const rls_cvec3 &v) { CV3 o; o.x = v.x; o.y = v.y; o.z = v.z; return o; }
static inline V3  mv(const rls_vec3 &v) { V3 o; o.x = v.x; o.y = v.y; o.z = v.z; return o; }
static inline P1  p1(const rls_param1 &p) { P1 o; o.value = p.value; o.array = p.array; return o; }
static inline P3  p3(const rls_param3 &p)
{
    P3 o; o.value[0] = p.value[0]; o.value[1] = p.value[1]; o.value[2] = p.value[2];
    o.x = p.array.x; o.y = p.array.y; o.z = p.array.z; return o;
}
static inline ShadingSoA sh(const rls_shading_soa &s)
{
    ShadingSoA o; o.U = cv(s.U); o.V = cv(s.V); o.N = cv(s.N); o.wo = cv(s.wo); o.backfacing = s.backfacing; return o;
}
static inline bool has3(const rls_cvec3 &v) { return v.x && v.y && v.z; }
static inline bool has3(const rls_vec3 &v) { return v.x && v.y && v.z; }
static inline bool ok_shading(const rls_shading_soa *s) { return s && has3(s->U) && has3(s->V) && has3(s->N) && has3(s->wo); }
static inline bool ok_p3(const rls_param3 &p) { return (!p.array.x && !p.array.y && !p.array.z) || has3(p.array); }

static inline GgxParamsDev dev(const rls_ggx_params &p)
{
    GgxParamsDev o; o.ks = p3(p.KsColor); o.rough = p1(p.specularRoughness); o.ior = p1(p.ior); o.aniso = p1(p.anisotropic);
    o.ndf = p.normal_sampler == RLS_GGX_SAMPLER_NDF; return o;
}
static inline DisneyParamsDev dev(const rls_disney_params &p)
{
    DisneyParamsDev o;
    o.base_color = p3(p.base_color); o.subsurface = p1(p.subsurface); o.metallic = p1(p.metallic);
    o.specular = p1(p.specular); o.specular_tint = p1(p.specular_tint); o.roughness = p1(p.roughness);
    o.anisotropic = p1(p.anisotropic); o.sheen = p1(p.sheen); o.sheen_tint = p1(p.sheen_tint);
    o.clearcoat = p1(p.clearcoat); o.clearcoat_gloss = p1(p.clearcoat_gloss);
    o.sample_from_visible_normal = p.sample_from_visible_normal;
    return o;
}
static inline SkinParamsDev dev(const rls_skin_params &p)
{
    SkinParamsDev o;
    o.sss_color = p3(p.sss_color); o.sss_scatter_dist = p3(p.sss_scatter_dist);
    o.sss_weight = p1(p.sss_weight); o.sss_dist_multiplier = p1(p.sss_dist_multiplier);
    o.specular_weight = p1(p.specular_weight); o.sheen_weight = p1(p.sheen_weight);
    return o;
}

static inline BsdfOutDev dev(const rls_bsdf_out &o)
{
    BsdfOutDev d; d.wi = mv(o.wi); d.f = mv(o.f); d.pdf = o.pdf; d.fresnel = o.fresnel; d.flags = o.flags; return d;
}
static inline DielectricOutDev dev(const rls_ggx_dielectric_out &o)
{
    DielectricOutDev d; d.fresnel = o.fresnel; d.wi_r = mv(o.wi_r); d.f_r = o.f_r; d.pdf_r = o.pdf_r;
    d.wi_t = mv(o.wi_t); d.f_t = o.f_t; d.weight_t = o.weight_t; d.flags = o.flags; return d;
}
static inline DisneyOutDev dev(const rls_disney_out &o)
{
    DisneyOutDev d; d.wi_s = mv(o.wi_s); d.f_s = mv(o.f_s); d.pdf_s = o.pdf_s;
    d.wi_d = mv(o.wi_d); d.f_d = mv(o.f_d); d.pdf_d = o.pdf_d; d.flags = o.flags; return d;
}
static inline ProfileOutDev dev(const rls_profile_out &o)
{
    ProfileOutDev d; d.r = o.r; d.pdf = o.pdf; d.Rd = mv(o.Rd); d.flags = o.flags; return d;
}

static inline unsigned grid_for(size_t n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }
#ifdef RLS_EXPERIMENTS
static inline bool aligned8(std::initializer_list<const void *> ptrs)
{
    for (const void *q : ptrs) if ((uintptr_t)q & 7u) return false;
    return true;
}
#endif
// 32-bit sample index: one IMAD.WIDE per array address instead of a 64-bit add pair; the entry
// points reject n >= 2^32 (that many samples would not fit in HBM anyway).
#define RLS_INDEX()                                                            \
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;                  \
    if (i >= (uint32_t)n) return;

// ================================================================ rlGgx kernels
// Fused kernels: run the sample with the guard-free FpFast sequences; if any operand left the
// window in which they equal the IEEE operators, run it again with FpExact (rls_fp.cuh).
#define RLS_FAST_THEN_EXACT(kFast, result, expr)                                   \
    do {                                                                            \
        bool ok_ = false;                                                           \
        if (kFast) { FpFast fp; result = (expr); ok_ = fp.ok(); }                   \
        if (!ok_) {                                                                 \
            FpExact fp; result = (expr);                                            \
            if (kFast) atomicAdd(fallbacks, 1ull);                                  \
        }                                                                           \
    } while (0)

template <class Fp>
RLS_DEV Ggx ggx_make(Fp &fp, const ShadingSoA &sg, const GgxParamsDev &p, uint32_t i)
{
    Shading s = load_shading(sg, i);
    Ggx g;
    ggx_init(fp, g, s, fetch(p.ks, i), fetch(p.ior, i), fetch(p.rough, i), fetch(p.aniso, i));
    g.ndf = p.ndf != 0;
    return g;
}

__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_ggx_eval_sample(size_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, V3 wi, float *fresnel)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, i);
    f3 M = ggx_sample_normal(fp, g, __ldg(rx + i), __ldg(ry + i));
    f3 L = reflect_direction(g.wo, M);                 // src/rlGgx.h:100-101
    store3(wi, i, L);
    if (fresnel) fresnel[i] = ggx_fresnel(fp, g, L, M);    // :103-104,181-184 with one sample
}

__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_ggx_eval_brdf(size_t n, ShadingSoA sg, GgxParamsDev p, CV3 wi, V3 f)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, i);
    store3(f, i, ggx_eval_brdf(fp, g, load3(wi, i)));
}

__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_ggx_eval_pdf(size_t n, ShadingSoA sg, GgxParamsDev p, CV3 wi, float *pdf)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, i);
    pdf[i] = ggx_eval_pdf(fp, g, load3(wi, i));
}

// The refraction half of the path at caller-supplied directions (src/rlGgx.h:277-328): the pieces the fused dielectric
// unit composes, as entry points of their own so that each can be checked at the oracle's direction bits.
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_ggx_refract_direction(size_t n, ShadingSoA sg, GgxParamsDev p, CV3 m, V3 wi, uint32_t *flags)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, i);
    f3 dir = mk3(0.0f, 0.0f, 0.0f);
    const bool ok = ggx_refract_direction(fp, g, load3(m, i), g.wo, dir);      // getRefractDirection(m, V, dir)
    store3(wi, i, ok ? dir : mk3(0.0f, 0.0f, 0.0f));
    if (flags) flags[i] = (ok ? 0u : RLS_FLAG_TIR) | (g.entering ? RLS_FLAG_ENTERING : 0u);
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_ggx_eval_btdf(size_t n, ShadingSoA sg, GgxParamsDev p, CV3 wi, float *ft)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, i);
    ft[i] = ggx_refraction(fp, g, g.wo, load3(wi, i), g.N);                     // refraction(V, wi, N)
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_ggx_sample_weight(size_t n, ShadingSoA sg, GgxParamsDev p, CV3 wi, CV3 m, float *w)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, i);
    w[i] = ggx_sample_weight(fp, g, g.wo, load3(wi, i), load3(m, i));          // getSampleWeight(V, wi, m)
}

template <class Fp>
RLS_DEV GgxBsdf ggx_unit_from(Fp &fp, const Shading &s, f3 ks, float ior, float rough, float aniso, bool ndf, float rx, float ry)
{
    Ggx g;
    ggx_init(fp, g, s, ks, ior, rough, aniso);
    g.ndf = ndf;
    return ggx_unit(fp, g, rx, ry);
}
template <bool kFast>
RLS_DEV void ggx_sample(uint32_t i, const ShadingSoA &sg, const GgxParamsDev &p, const float *rx, const float *ry,
                        const V3 &wi, const V3 &f, float *pdf, float *fresnel, uint32_t *flags, unsigned long long *fallbacks)
{
    GgxBsdf o;
    bool ok = false;
    if (kFast) {
        const Shading s = load_shading(sg, i);
        FpFast fp;
        o = ggx_unit_from(fp, s, fetch(p.ks, i), fetch(p.ior, i), fetch(p.rough, i), fetch(p.aniso, i), p.ndf != 0,
                          __ldg(rx + i), __ldg(ry + i));
        ok = fp.ok();
    }
    if (!ok) {          // the exact re-run reloads its inputs (see k_ggx_dielectric)
        const Shading s = load_shading<true>(sg, i);
        FpExact fp;
        o = ggx_unit_from(fp, s, fetch<true>(p.ks, i), fetch<true>(p.ior, i), fetch<true>(p.rough, i), fetch<true>(p.aniso, i),
                          p.ndf != 0, __ldcg(rx + i), __ldcg(ry + i));
        if (kFast) atomicAdd(fallbacks, 1ull);
    }
    store3(wi, i, o.L);
    store3(f, i, o.f);
    pdf[i] = o.pdf;
    if (fresnel) fresnel[i] = o.fresnel;
    flags[i] = o.flags;
}
template <bool kFast>
__global__ void __launch_bounds__(kBlockGgx, RLS_GGX_MIN_BLOCKS)
k_ggx_sample_eval_pdf(size_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry,
                      V3 wi, V3 f, float *pdf, float *fresnel, uint32_t *flags, unsigned long long *fallbacks)
{
    RLS_INDEX();
    ggx_sample<kFast>(i, sg, p, rx, ry, wi, f, pdf, fresnel, flags, fallbacks);
}

// ---- RLS_ARITH_TOLERANT: the bit-exact re-run of the samples a tolerance kernel listed (rls_tol_launch.cuh).
// `body(i)` evaluates sample i with the fast policy (+ its own exact re-run) and overwrites its outputs.
// kGroup = 8: EIGHT lanes per listed sample, the aligned group of 8 consecutive samples that shares its 32-byte sectors.
// A lone sample costs one scattered 4-byte access per array, each moving a whole sector (the stores as read-modify-write:
// ncu measured 2.8 kB of DRAM traffic per listed rlGgx sample and a kernel stalled on the L1 miss queue); the group reads
// and writes the same sectors whole, and its seven neighbours simply get the bit-exact result as well (the list is a
// function of the inputs alone, so the output stays deterministic).  Measured: rlGgx dielectric re-run 45 -> 37 us per
// 2^26 samples; rlDisney (44 arrays, 64 registers, 1800 instructions per sample) LOSES with it -- 2.7 instead of 1.75 ns
// per listed sample, the eightfold exact work costs more than the whole sectors save -- and keeps kGroup = 1.
template <int kGroup, class Body>
RLS_DEV void rerun_listed(uint32_t n, const tol::Worklist &wl, const uint32_t *flags, unsigned long long *fallbacks, Body body)
{
    const unsigned c = *wl.count;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (tid == 0 && c) atomicAdd(fallbacks, (unsigned long long)c);
    const unsigned listed = c < wl.cap ? c : wl.cap;
    if (kGroup == 1) {
        for (uint32_t k = tid; k < listed; k += stride) body(wl.list[k]);
    } else {
        for (uint64_t t = tid; t < (uint64_t)listed * (uint64_t)kGroup; t += stride) {
            const uint32_t i = (wl.list[t / kGroup] & ~(uint32_t)(kGroup - 1)) + (uint32_t)(t % kGroup);
            if (i < n) body(i);
        }
    }
    if (c > wl.cap)                      // list overflow: the remaining samples carry the sentinel in their flags word
        for (uint32_t i = tid; i < n; i += stride)
            if (flags[i] == tol::kRerunSentinel) body(i);
}
__global__ void __launch_bounds__(kBlockGgx, RLS_GGX_MIN_BLOCKS)
k_ggx_sample_eval_pdf_rerun(uint32_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, BsdfOutDev o,
                            tol::Worklist wl, unsigned long long *fallbacks)
{
    rerun_listed<8>(n, wl, o.flags, fallbacks, [&](uint32_t i) {
        ggx_sample<true>(i, sg, p, rx, ry, o.wi, o.f, o.pdf, o.fresnel, o.flags, fallbacks + 1);
    });
}


template <bool kFast, bool kArrays>   // kArrays: ior and specularRoughness are per-sample arrays
RLS_DEV void dielectric_sample(uint32_t i, const ShadingSoA &sg, const GgxParamsDev &p, const float *rx, const float *ry,
                               const DielectricOutDev &o, unsigned long long *fallbacks)
{
    Dielectric r;
    bool ok = false;
    if (kFast) {
        const Shading s = load_shading(sg, i);
        const float ior = fetch_t<kArrays>(p.ior, i), rough = fetch_t<kArrays>(p.rough, i), aniso = fetch(p.aniso, i);
        const float u1 = __ldg(rx + i), u2 = __ldg(ry + i);
        FpFast fp;
        // (dielectric_unit<kFlat = true>, the select form of the refraction branch, measured neutral: +-0.3 %)
        r = dielectric_unit(fp, s, ior, rough, aniso, u1, u2, p.ndf != 0);
        ok = fp.ok();
    }
    if (!ok) {
        // The exact re-run RELOADS its inputs (ld.global.cg, which the compiler cannot merge with the
        // ld.global.nc above): the fast path then need not keep 17 input registers alive for it.
        const Shading s = load_shading<true>(sg, i);
        const float ior = fetch_t<kArrays, true>(p.ior, i), rough = fetch_t<kArrays, true>(p.rough, i), aniso = fetch<true>(p.aniso, i);
        const float u1 = __ldcg(rx + i), u2 = __ldcg(ry + i);
        FpExact fp;
        r = dielectric_unit(fp, s, ior, rough, aniso, u1, u2, p.ndf != 0);
        if (kFast) atomicAdd(fallbacks, 1ull);
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
template <bool kFast, bool kArrays>
__global__ void __launch_bounds__(kBlockGgx, RLS_GGX_MIN_BLOCKS)
k_ggx_dielectric(size_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, DielectricOutDev o,
                 unsigned long long *fallbacks)
{
    RLS_INDEX();
    dielectric_sample<kFast, kArrays>(i, sg, p, rx, ry, o, fallbacks);
}
template <bool kArrays>
__global__ void __launch_bounds__(kBlockGgx, RLS_GGX_MIN_BLOCKS)
k_ggx_dielectric_rerun(uint32_t n, ShadingSoA sg, GgxParamsDev p, const float *rx, const float *ry, DielectricOutDev o,
                       tol::Worklist wl, unsigned long long *fallbacks)
{
    rerun_listed<8>(n, wl, o.flags, fallbacks, [&](uint32_t i) {
        dielectric_sample<true, kArrays>(i, sg, p, rx, ry, o, fallbacks + 1);
    });
}

// ============================================================= rlDisney kernels
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_disney_eval_sample(size_t n, ShadingSoA sg, DisneyParamsDev p, int type, const float *rx, const float *ry, V3 wi, uint32_t *flags)
{
    RLS_INDEX();
    FpExact fp;
    Disney d; disney_init(fp, d, load_shading(sg, i), p, i);
    uint32_t lobe = 0;
    bool early = false;
    f3 L = (type == kRayDiffuse) ? disney_sample_diffuse(fp, d, __ldg(rx + i), __ldg(ry + i))
                                 : disney_sample_specular(fp, d, __ldg(rx + i), __ldg(ry + i), lobe, &early);
    store3(wi, i, L);
    if (flags) {
        uint32_t fl = (lobe << RLS_FLAG_LOBE_SHIFT) | (early ? RLS_FLAG_SLOPE_EARLY_OUT : 0u);
        if (is_zero(L)) fl |= RLS_FLAG_ZERO_L;
        if (dot(L, d.N) <= 0.0f) fl |= RLS_FLAG_BELOW_HORIZON;
        flags[i] = fl;
    }
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_disney_eval_brdf(size_t n, ShadingSoA sg, DisneyParamsDev p, int type, CV3 wi, V3 f)
{
    RLS_INDEX();
    FpExact fp;
    Disney d; disney_init(fp, d, load_shading(sg, i), p, i);
    store3(f, i, disney_eval_brdf(fp, d, type, load3(wi, i)));
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_disney_eval_pdf(size_t n, ShadingSoA sg, DisneyParamsDev p, int type, CV3 wi, float *pdf)
{
    RLS_INDEX();
    FpExact fp;
    Disney d; disney_init(fp, d, load_shading(sg, i), p, i);
    pdf[i] = disney_eval_pdf(fp, d, type, load3(wi, i));
}


template <bool kArrays, bool kReload, class Fp>
RLS_DEV DisneyOut1 disney_unit_from(Fp &fp, const Shading &s, const DisneyParamsDev &p, uint32_t i,
                                    float rx_s, float ry_s, float rx_d, float ry_d)
{
    Disney d; disney_init<kArrays, kReload>(fp, d, s, p, i);
    return disney_unit(fp, d, rx_s, ry_s, rx_d, ry_d);
}
// One rlDisney sample: fast policy, exact re-run when an operand left the window.
template <bool kFast, bool kArrays>
RLS_DEV void disney_sample(uint32_t i, const ShadingSoA &sg, const DisneyParamsDev &p, const float *rx_s, const float *ry_s,
                           const float *rx_d, const float *ry_d, const DisneyOutDev &o, unsigned long long *fallbacks)
{
    DisneyOut1 r;
    bool ok = false;
    if (kFast) {
        const Shading s = load_shading(sg, i);
        FpFastLeanTrig fp;
        r = disney_unit_from<kArrays, false>(fp, s, p, i, __ldg(rx_s + i), __ldg(ry_s + i), __ldg(rx_d + i), __ldg(ry_d + i));
        ok = fp.ok();
    }
    if (!ok) {          // the exact re-run reloads its 29 inputs (see k_ggx_dielectric)
        const Shading s = load_shading<true>(sg, i);
        FpExact fp;
        r = disney_unit_from<kArrays, true>(fp, s, p, i, __ldcg(rx_s + i), __ldcg(ry_s + i), __ldcg(rx_d + i), __ldcg(ry_d + i));
        if (kFast) atomicAdd(fallbacks, 1ull);
    }
    store3(o.wi_s, i, r.Ls); store3(o.f_s, i, r.fs); o.pdf_s[i] = r.ps;
    store3(o.wi_d, i, r.Ld); store3(o.f_d, i, r.fd); o.pdf_d[i] = r.pd;
    o.flags[i] = r.flags;
}
template <bool kFast, bool kArrays>
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_disney_sample_eval_pdf(size_t n, ShadingSoA sg, DisneyParamsDev p, const float *rx_s, const float *ry_s,
                         const float *rx_d, const float *ry_d, DisneyOutDev o, unsigned long long *fallbacks)
{
    if (kFast) rlm::smem_tables_init();      // exp2 / log / log2 tables in shared memory (before the early exit below)
    RLS_INDEX();
    disney_sample<kFast, kArrays>(i, sg, p, rx_s, ry_s, rx_d, ry_d, o, fallbacks);
}
template <bool kArrays>
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_disney_sample_eval_pdf_rerun(uint32_t n, ShadingSoA sg, DisneyParamsDev p, const float *rx_s, const float *ry_s,
                               const float *rx_d, const float *ry_d, DisneyOutDev o, tol::Worklist wl,
                               unsigned long long *fallbacks)
{
    rlm::smem_tables_init();
    rerun_listed<1>(n, wl, o.flags, fallbacks, [&](uint32_t i) {
        disney_sample<true, kArrays>(i, sg, p, rx_s, ry_s, rx_d, ry_d, o, fallbacks + 1);
    });
}

// ================================================================ profile kernels
struct NdProfileSoADev { V3 distance, C1, C2; float *max_radius; };
RLS_DEV NdProfile nd_load(const NdProfileSoADev &s, uint32_t i)
{
    NdProfile p;
    p.d[0] = s.distance.x[i]; p.d[1] = s.distance.y[i]; p.d[2] = s.distance.z[i];
    p.C1[0] = s.C1.x[i]; p.C1[1] = s.C1.y[i]; p.C1[2] = s.C1.z[i];
    p.C2[0] = s.C2.x[i]; p.C2[1] = s.C2.y[i]; p.C2[2] = s.C2.z[i];
    p.R = s.max_radius[i];
    return p;
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_nd_set_distance(size_t n, CV3 dist, NdProfileSoADev o)
{
    RLS_INDEX();
    FpExact fp;
    NdProfile p; nd_set_distance(fp, p, load3(dist, i));
    store3(o.distance, i, mk3(p.d[0], p.d[1], p.d[2]));
    store3(o.C1, i, mk3(p.C1[0], p.C1[1], p.C1[2]));
    store3(o.C2, i, mk3(p.C2[0], p.C2[1], p.C2[2]));
    o.max_radius[i] = p.R;
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_nd_get_radius(size_t n, NdProfileSoADev s, const float *rx, float *r, uint32_t *flags)
{
    RLS_INDEX();
    NdProfile p = nd_load(s, i);
    uint32_t fl;
    FpExact fp;
    r[i] = nd_get_radius(fp, p, __ldg(rx + i), fl);
    if (flags) flags[i] = fl;
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_nd_get_pdf(size_t n, NdProfileSoADev s, const float *r, float *pdf)
{
    RLS_INDEX();
    NdProfile p = nd_load(s, i);
    FpExact fp;
    pdf[i] = nd_get_pdf(fp, p, __ldg(r + i));
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_nd_eval_profile(size_t n, NdProfileSoADev s, const float *r, V3 rd)
{
    RLS_INDEX();
    NdProfile p = nd_load(s, i);
    FpExact fp;
    store3(rd, i, nd_eval_profile(fp, p, __ldg(r + i)));
}
// GaussianProfile (src/rlSss.h:63-97): 3 floats of state, 1 float per result; exact policy (variant row)
struct GaussProfileSoADev { float *variance, *max_radius, *norm; };
RLS_DEV GaussProfile gauss_load(const GaussProfileSoADev &s, uint32_t i)
{
    GaussProfile p; p.var = s.variance[i]; p.R = s.max_radius[i]; p.norm = s.norm[i]; return p;
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_gauss_set_distance(size_t n, const float *dist_x, GaussProfileSoADev o)
{
    RLS_INDEX();
    FpExact fp;
    GaussProfile p; gauss_set_distance(fp, p, __ldg(dist_x + i));
    o.variance[i] = p.var; o.max_radius[i] = p.R; o.norm[i] = p.norm;
}
template <int kOp>      // 0 getRadius(rx), 1 getPdf(r), 2 evalProfile(r)
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_gauss_unary(size_t n, GaussProfileSoADev s, const float *x, float *out)
{
    RLS_INDEX();
    FpExact fp;
    const GaussProfile p = gauss_load(s, i);
    const float v = __ldg(x + i);
    out[i] = kOp == 0 ? gauss_get_radius(fp, p, v) : (kOp == 1 ? gauss_get_pdf(fp, p, v) : gauss_eval_profile(fp, p, v));
}
template <bool kFast>
__global__ void __launch_bounds__(kBlockSkin, RLS_SKIN_MIN_BLOCKS)
k_gauss_profile(size_t n, const float *dist_x, const float *rx, float *r, float *pdf, float *rd, unsigned long long *fallbacks)
{
    if (kFast) rlm::smem_tables_init();
    RLS_INDEX();
    Gauss1 o;
    bool ok = false;
    if (kFast) {
        FpFastSmemTab fp;
        o = gauss_profile_unit(fp, __ldg(dist_x + i), __ldg(rx + i));
        ok = fp.ok();
    }
    if (!ok) {
        FpExact fp;
        o = gauss_profile_unit(fp, __ldcg(dist_x + i), __ldcg(rx + i));
        if (kFast) atomicAdd(fallbacks, 1ull);
    }
    r[i] = o.r; rd[i] = o.rd; pdf[i] = o.pdf;
}
// Four consecutive samples per thread, 128-bit loads and stores (all five arrays 16-byte aligned; the host picks this
// form and runs the scalar kernel on the n % 4 tail): the unit is ~170 warp instructions per 32 samples, so with one
// sample per thread the CTA turnover and the load-to-use latency bound it (ncu: 66 % issue slots, 65 % occupancy, DRAM 33 %).
// One operand tracker for the quad; an out-of-window operand re-runs the four samples exactly.
template <bool kFast>
__global__ void __launch_bounds__(kBlockSkin, 4)
k_gauss_profile_x4(size_t n4, const float4 *dist_x, const float4 *rx, float4 *r, float4 *pdf, float4 *rd,
                   unsigned long long *fallbacks)
{
    if (kFast) rlm::smem_tables_init();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t)n4) return;
    Gauss1 o[4];
    bool ok = false;
    if (kFast) {
        FpFastSmemTab fp;
        const float4 d = __ldg(dist_x + i), x = __ldg(rx + i);
        o[0] = gauss_profile_unit(fp, d.x, x.x); o[1] = gauss_profile_unit(fp, d.y, x.y);
        o[2] = gauss_profile_unit(fp, d.z, x.z); o[3] = gauss_profile_unit(fp, d.w, x.w);
        ok = fp.ok();
    }
    if (!ok) {
        FpExact fp;
        const float4 d = __ldcg(dist_x + i), x = __ldcg(rx + i);
        o[0] = gauss_profile_unit(fp, d.x, x.x); o[1] = gauss_profile_unit(fp, d.y, x.y);
        o[2] = gauss_profile_unit(fp, d.z, x.z); o[3] = gauss_profile_unit(fp, d.w, x.w);
        if (kFast) atomicAdd(fallbacks, 4ull);
    }
    r[i] = make_float4(o[0].r, o[1].r, o[2].r, o[3].r);
    rd[i] = make_float4(o[0].rd, o[1].rd, o[2].rd, o[3].rd);
    pdf[i] = make_float4(o[0].pdf, o[1].pdf, o[2].pdf, o[3].pdf);
}
struct Profile1 { float r, pdf; f3 Rd; uint32_t flags; };
template <class Fp>
RLS_DEV Profile1 skin_profile_unit(Fp &fp, f3 dist, float rx)
{
    Profile1 o;
    NdProfile p; nd_set_distance(fp, p, dist);
    o.r = nd_get_radius(fp, p, rx, o.flags);
    nd_pdf_and_profile(fp, p, o.r, o.pdf, o.Rd);
    return o;
}
template <bool kFast>
RLS_DEV void skin_sample(uint32_t i, const SkinParamsDev &sp, const float *rx, const ProfileOutDev &o, unsigned long long *fallbacks)
{
    Profile1 r;
    bool ok = false;
    if (kFast) {
        FpFastSmemTab fp;
        r = skin_profile_unit(fp, skin_scatter_dist(sp, i), __ldg(rx + i));
        ok = fp.ok();
    }
    if (!ok) {          // the exact re-run reloads its inputs (see k_ggx_dielectric)
        FpExact fp;
        r = skin_profile_unit(fp, skin_scatter_dist<true>(sp, i), __ldcg(rx + i));
        if (kFast) atomicAdd(fallbacks, 1ull);
    }
    o.r[i] = r.r;
    o.pdf[i] = r.pdf;
    store3(o.Rd, i, r.Rd);
    o.flags[i] = r.flags;
}
template <bool kFast>
__global__ void __launch_bounds__(kBlockSkin, RLS_SKIN_MIN_BLOCKS)
k_skin_profile(size_t n, SkinParamsDev sp, const float *rx, ProfileOutDev o, unsigned long long *fallbacks)
{
    if (kFast) rlm::smem_tables_init();      // exp2 / log tables in shared memory (before the early exit below)
    RLS_INDEX();
    skin_sample<kFast>(i, sp, rx, o, fallbacks);
}
__global__ void __launch_bounds__(kBlockSkin, RLS_SKIN_MIN_BLOCKS)
k_skin_profile_rerun(uint32_t n, SkinParamsDev sp, const float *rx, ProfileOutDev o, tol::Worklist wl, unsigned long long *fallbacks)
{
    rlm::smem_tables_init();
    rerun_listed<1>(n, wl, o.flags, fallbacks, [&](uint32_t i) { skin_sample<true>(i, sp, rx, o, fallbacks + 1); });
}
// src/rlSkin.cpp:191,204,214,228,231,238
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_skin_layer_weights(size_t n, SkinParamsDev sp, const float *avgSheen, const float *avgSpec, float *specScale, float *sssWeight)
{
    RLS_INDEX();
    float sheenWeight = fetch(sp.sheen_weight, i);
    float specularWeight = fetch(sp.specular_weight, i);
    float w = fetch(sp.sss_weight, i);
    float sheenFresnel = 0.0f, specularFresnel = 0.0f;
    if (sheenWeight > kEps) sheenFresnel = __ldg(avgSheen + i) * sheenWeight;
    if (specularWeight > kEps) specularFresnel = __ldg(avgSpec + i) * specularWeight;
    specScale[i] = specularWeight * (1.0f - sheenFresnel);
    w *= 1.0f - specularFresnel * (1.0f - sheenFresnel);
    sssWeight[i] = w;
}

// src/rlSss.h:487-533 (getProbeRay) with origin = 0
struct ProbeOutDev { float *r; V3 origin, dir; float *maxdist; uint32_t *flags; };
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_skin_probe_ray(size_t n, ShadingSoA sg, SkinParamsDev sp, const float *rx_in, const float *ry_in, ProbeOutDev o)
{
    RLS_INDEX();
    Shading s = load_shading(sg, i);
    FpExact fp;
    NdProfile p; nd_set_distance(fp, p, skin_scatter_dist(sp, i));
    float rx = __ldg(rx_in + i), ry = __ldg(ry_in + i);
    int idx;
    if (rx < 0.5f) { idx = 0; rx = linearstep_m(fp, 0.0f, 0.5f, rx); }
    else if (rx < 0.75f) { idx = 2; rx = linearstep_m(fp, 0.5f, 0.75f, rx); }
    else { idx = 3; rx = linearstep_m(fp, 0.75f, 1.0f, rx); }
    uint32_t fl;
    float r = nd_get_radius(fp, p, rx, fl);
    float rmax = p.R;
    float phi = kTwoPi * ry;
    float sn, cs;
    rlm::sincosf_(phi, &sn, &cs);
    f3 offset = mk3(cs * r, sqrtf(rmax * rmax - r * r), sn * r);
    float maxdist = offset.y * 2.0f;
    f3 dir;
    if (idx < 2) { dir = -s.N; offset = rotate_to_frame(offset, s.U, -dir, s.V); }
    else if (idx == 2) { dir = s.U; offset = rotate_to_frame(offset, s.V, -dir, s.N); }
    else { dir = s.V; offset = rotate_to_frame(offset, s.N, -dir, s.U); }
    o.r[i] = r;
    store3(o.origin, i, mk3(0.0f + offset.x, 0.0f + offset.y, 0.0f + offset.z));
    store3(o.dir, i, dir);
    o.maxdist[i] = maxdist;
    o.flags[i] = fl | ((uint32_t)idx << RLS_FLAG_PROBE_AXIS_SHIFT);
}
// src/rlSss.h:252-263: 3-axis MIS pdf of one probe hit
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_skin_probe_mis_pdf(size_t n, ShadingSoA sg, SkinParamsDev sp, CV3 disp, CV3 hitN, float *pdf)
{
    RLS_INDEX();
    Shading s = load_shading(sg, i);
    FpExact fp;
    NdProfile p; nd_set_distance(fp, p, skin_scatter_dist(sp, i));
    f3 dp = load3(disp, i), hn = load3(hitN, i);
    f3 off = mk3(dot(dp, s.U), dot(dp, s.V), dot(dp, s.N));     // world -> local (AiM4VectorByMatrixMult)
    off = mk3(off.x * off.x, off.y * off.y, off.z * off.z);
    float rr0 = sqrtf(off.y + off.z), rr1 = sqrtf(off.x + off.z), rr2 = sqrtf(off.x + off.y);
    pdf[i] = nd_get_pdf(fp, p, rr0) * abs_m(dot(s.U, hn)) * 0.25f
           + nd_get_pdf(fp, p, rr1) * abs_m(dot(s.V, hn)) * 0.25f
           + nd_get_pdf(fp, p, rr2) * abs_m(dot(s.N, hn)) * 0.5f;
}

// ============================================ callers of the triple (8(f) rows f2-f4)
struct SkinLayersOutDev { V3 sheen, spec; float *sheenF, *specF, *sssW; uint32_t *flags; };
template <bool kFast>
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_skin_glossy_layers(size_t n, uint32_t K, ShadingSoA sg, SkinLayersDev sp, const float *rx_a, const float *ry_a,
                     const float *rx_b, const float *ry_b, CV3 li_a, CV3 li_b, SkinLayersOutDev o,
                     unsigned long long *fallbacks)
{
    RLS_INDEX();
    const Shading s = load_shading(sg, i);
    SkinLayers1 r;
    RLS_FAST_THEN_EXACT(kFast, r, skin_layers_unit(fp, s, sp, K, n, i, rx_a, ry_a, rx_b, ry_b, li_a, li_b));
    store3(o.sheen, i, r.sheen);
    store3(o.spec, i, r.spec);
    o.sheenF[i] = r.sheenF;
    o.specF[i] = r.specF;
    o.sssW[i] = r.sssW;
    o.flags[i] = r.flags;
}

__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_ggx_light_sample(size_t n, ShadingSoA sg, GgxParamsDev p, LightDev light, const float *rx, const float *ry,
                   LightDev at_l, V3 out, float *w_light, float *w_brdf)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, i);
    const f3 Ld = load3(light.dir, i), Li = load3(light.radiance, i);
    const float pl = __ldg(light.pdf + i);
    const f3 f_l = ggx_eval_brdf(fp, g, Ld);
    const float p_bl = ggx_eval_pdf(fp, g, Ld);
    const bool have = at_l.pdf != nullptr;
    f3 L = mk3(0.0f, 0.0f, 0.0f), f_b = L, Li_b = L;
    float p_b = 0.0f, p_lb = 0.0f;
    if (have) {
        GgxBsdf u = ggx_unit(fp, g, __ldg(rx + i), __ldg(ry + i));
        L = u.L; f_b = u.f; p_b = u.pdf;
        Li_b = load3(at_l.radiance, i);
        p_lb = __ldg(at_l.pdf + i);
    }
    Mis1 m = mis_combine(Ld, Li, pl, f_l, p_bl, have, L, f_b, p_b, Li_b, p_lb);
    store3(out, i, m.rgb);
    if (w_light) w_light[i] = m.w_light;
    if (w_brdf) w_brdf[i] = m.w_brdf;
}

__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_disney_light_sample(size_t n, ShadingSoA sg, DisneyParamsDev p, int type, LightDev light, const float *rx,
                      const float *ry, LightDev at_l, V3 out, float *w_light, float *w_brdf)
{
    RLS_INDEX();
    FpExact fp;
    Disney d; disney_init(fp, d, load_shading(sg, i), p, i);
    const f3 Ld = load3(light.dir, i), Li = load3(light.radiance, i);
    const float pl = __ldg(light.pdf + i);
    const f3 f_l = disney_eval_brdf(fp, d, type, Ld);
    const float p_bl = disney_eval_pdf(fp, d, type, Ld);
    const bool have = at_l.pdf != nullptr;
    f3 L = mk3(0.0f, 0.0f, 0.0f), f_b = L, Li_b = L;
    float p_b = 0.0f, p_lb = 0.0f;
    if (have) {
        uint32_t lobe = 0;
        L = (type == kRayDiffuse) ? disney_sample_diffuse(fp, d, __ldg(rx + i), __ldg(ry + i))
                                  : disney_sample_specular(fp, d, __ldg(rx + i), __ldg(ry + i), lobe);
        f_b = disney_eval_brdf(fp, d, type, L);
        p_b = disney_eval_pdf(fp, d, type, L);
        Li_b = load3(at_l.radiance, i);
        p_lb = __ldg(at_l.pdf + i);
    }
    Mis1 m = mis_combine(Ld, Li, pl, f_l, p_bl, have, L, f_b, p_b, Li_b, p_lb);
    store3(out, i, m.rgb);
    if (w_light) w_light[i] = m.w_light;
    if (w_brdf) w_brdf[i] = m.w_brdf;
}

// SampleWriter::writeRadiance (src/rlUtil.h:98-114): one thread per pixel of one shading point.
RLS_DEV f3 writer_direction(int i, int j, int W, int H)
{
    FpExact fp;
    float theta = kHalfPi * (float)j / (float)H;
    float phi = kTwoPi * (float)i / (float)W;
    float sn, cs;
    rlm::sincosf_(theta, &sn, &cs);                    // cosf(theta): same binary64 kernel as sincosf
    return spherical_direction(fp, cs, phi);
}
RLS_DEV void write_pixel(float *image, int W, int H, int x, int y, f3 rgb)     // writePixel :158-163
{
    const size_t stride = (size_t)W * H, at = (size_t)x + (size_t)y * W;
    image[at] = rgb.z; image[at + stride] = rgb.y; image[at + stride * 2] = rgb.x;
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_writer_radiance_ggx(size_t n, ShadingSoA sg, GgxParamsDev p, uint32_t point, int W, int H, float *image)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, point);
    const int x = (int)(i % (uint32_t)W), y = (int)(i / (uint32_t)W);
    write_pixel(image, W, H, x, y, ggx_eval_brdf(fp, g, writer_direction(x, y, W, H)));
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_writer_radiance_disney(size_t n, ShadingSoA sg, DisneyParamsDev p, int type, uint32_t point, int W, int H, float *image)
{
    RLS_INDEX();
    FpExact fp;
    Disney d; disney_init(fp, d, load_shading(sg, point), p, point);
    const int x = (int)(i % (uint32_t)W), y = (int)(i / (uint32_t)W);
    write_pixel(image, W, H, x, y, disney_eval_brdf(fp, d, type, writer_direction(x, y, W, H)));
}
// SampleWriter::writeSample (:116-156).  scratch[pixel] = max over samples of ((k + 1) << 1 | red):
// the sample with the largest index wins, as the last write does in the sequential loop.
RLS_DEV void scatter_mark(f3 L, uint32_t k, int W, int H, uint32_t *scratch, uint32_t *missing)
{
    int x, y; bool red;
    if (!scatter_pixel(L, W, H, x, y, red)) return;
    atomicMax(scratch + (size_t)x + (size_t)y * W, ((k + 1u) << 1) | (red ? 1u : 0u));
    if (red && missing) atomicAdd(missing, 1u);
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_writer_scatter_ggx(size_t n, ShadingSoA sg, GgxParamsDev p, uint32_t point, const float *rx, const float *ry,
                     int W, int H, uint32_t *scratch, uint32_t *missing)
{
    RLS_INDEX();
    FpExact fp;
    Ggx g = ggx_make(fp, sg, p, point);
    f3 M = ggx_sample_normal(fp, g, __ldg(rx + i), __ldg(ry + i));
    scatter_mark(reflect_direction(g.wo, M), i, W, H, scratch, missing);
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_writer_scatter_disney(size_t n, ShadingSoA sg, DisneyParamsDev p, int type, uint32_t point, const float *rx,
                        const float *ry, int W, int H, uint32_t *scratch, uint32_t *missing)
{
    RLS_INDEX();
    FpExact fp;
    Disney d; disney_init(fp, d, load_shading(sg, point), p, point);
    uint32_t lobe = 0;
    f3 L = (type == kRayDiffuse) ? disney_sample_diffuse(fp, d, __ldg(rx + i), __ldg(ry + i))
                                 : disney_sample_specular(fp, d, __ldg(rx + i), __ldg(ry + i), lobe);
    scatter_mark(L, i, W, H, scratch, missing);
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_writer_scatter_paint(size_t n, int W, int H, const uint32_t *scratch, float *image)
{
    RLS_INDEX();
    const uint32_t v = scratch[i];
    if (v == 0u) return;
    const f3 c = (v & 1u) ? mk3(1.0f, 0.0f, 0.0f) : mk3(0.0f, 1.0f, 0.0f);     // AI_RGB_RED / AI_RGB_GREEN
    write_pixel(image, W, H, (int)(i % (uint32_t)W), (int)(i / (uint32_t)W), c);
}

// ============================================================ synthetic generators
RLS_DEV float uniform24(uint64_t seed, uint32_t stream, uint64_t index) { return sweep_uniform24(seed, stream, index); }
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_synth_uniform(size_t n, uint64_t seed, uint32_t stream, uint64_t first, float lo, float hi, float *out)
{
    RLS_INDEX();
    out[i] = lo + (hi - lo) * uniform24(seed, stream, first + i);
}
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_synth_shading(size_t n, uint64_t seed, uint64_t first, float cos_lo, float cos_hi, float back_frac,
                V3 U, V3 V, V3 N, V3 wo, uint8_t *backfacing)
{
    RLS_INDEX();
    FpExact fp;
    uint64_t idx = first + i;
    float u1 = uniform24(seed, 10, idx), u2 = uniform24(seed, 11, idx), u3 = uniform24(seed, 12, idx);
    float u4 = uniform24(seed, 13, idx), u5 = uniform24(seed, 14, idx), u6 = uniform24(seed, 15, idx);
    float nz = 1.0f - 2.0f * u1;
    float rn = sqrtf(fmaxf(0.0f, 1.0f - nz * nz));
    float sn, cn; sincosf(kTwoPi * u2, &sn, &cn);
    f3 Nn = normalize(fp, mk3(rn * cn, rn * sn, nz));
    f3 A = fabsf(Nn.x) < 0.9f ? mk3(1.0f, 0.0f, 0.0f) : mk3(0.0f, 1.0f, 0.0f);
    f3 T = normalize(fp, A - Nn * dot(A, Nn));
    f3 B = mk3(Nn.y * T.z - Nn.z * T.y, Nn.z * T.x - Nn.x * T.z, Nn.x * T.y - Nn.y * T.x);
    float st, ct; sincosf(kTwoPi * u3, &st, &ct);
    f3 Uu = normalize(fp, T * ct + B * st);
    f3 Vv = mk3(Nn.y * Uu.z - Nn.z * Uu.y, Nn.z * Uu.x - Nn.x * Uu.z, Nn.x * Uu.y - Nn.y * Uu.x);
    float cz = cos_lo + (cos_hi - cos_lo) * u4;
    float sr = sqrtf(fmaxf(0.0f, 1.0f - cz * cz));
    float sv, cvv; sincosf(kTwoPi * u5, &sv, &cvv);
    f3 w = normalize(fp, Uu * (sr * cvv) + Vv * (sr * sv) + Nn * cz);
    store3(U, i, Uu); store3(V, i, Vv); store3(N, i, Nn); store3(wo, i, w);
    if (backfacing) backfacing[i] = (u6 < back_frac) ? 1 : 0;
}

// ================================================================= albedo sweep
// Bit-exact policy: one warp per cell (rls_sweep.cuh), the rough-dielectric unit on the canonical frame.  All samples of
// a cell share (roughness, cos, ior), so everything that depends on the shading point only is loop invariant and is
// hoisted by the compiler.  The guarded operators run this kernel FASTER than the fast policy (30.3 vs 26.0 G samples/s
// on B200: the FP64 accumulators and the counter hash leave no registers for a second code path), so the bit-exact sweep
// uses FpExact throughout; RLS_ARITH_TOLERANT has its own kernel (rls_tol.cu).
__global__ void __launch_bounds__(kSweepBlock, RLS_SWEEP_MIN_BLOCKS)
k_albedo_sweep(SweepGridDev g, uint32_t n_cells, uint64_t seed, uint32_t k0, uint32_t k1, double *table)
{
    const uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= n_cells) return;                         // whole warps
    float rough, cosv, ior;
    sweep_cell(g, cell, rough, cosv, ior);
    Shading s;
    s.U = mk3(1.0f, 0.0f, 0.0f); s.V = mk3(0.0f, 1.0f, 0.0f); s.N = mk3(0.0f, 0.0f, 1.0f);
    s.wo = mk3(sqrtf(1.0f - cosv * cosv), 0.0f, cosv);
    s.backfacing = false;
    double acc[kSweepValues] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
    for (uint32_t k = k0 + (threadIdx.x & 31u); k < k1; k += 32u) {      // the entry point rejects k1 > 2^32 - 32
        const uint64_t idx = ((uint64_t)cell << 32) | (uint64_t)k;
        FpExact fp;
        const Dielectric r = dielectric_unit(fp, s, ior, rough, 0.0f, sweep_uniform24(seed, 0u, idx), sweep_uniform24(seed, 1u, idx));
        sweep_accumulate(acc, r.flags, r.f_r, r.pdf_r, r.w_t, r.F);
    }
    sweep_store(acc, cell, table);
}
// RLS_ARITH_TOLERANT: the samples the tolerance sweep kernel listed (cell, k) evaluated exactly and ADDED to the table
// (FP64 atomics: the table then differs from run to run in the last bits of the sums, never in the counts).
__global__ void __launch_bounds__(kSweepBlock, RLS_SWEEP_MIN_BLOCKS)
k_albedo_sweep_rerun(SweepGridDev g, uint64_t seed, tol::SweepWorklist wl, double *table, unsigned long long *fallbacks)
{
    const unsigned c = *wl.count;
    const unsigned listed = c < wl.cap ? c : wl.cap;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (tid == 0 && listed) atomicAdd(fallbacks, (unsigned long long)listed);
    for (uint32_t j = tid; j < listed; j += stride) {
        const uint2 e = wl.list[j];
        float rough, cosv, ior;
        sweep_cell(g, e.x, rough, cosv, ior);
        Shading s;
        s.U = mk3(1.0f, 0.0f, 0.0f); s.V = mk3(0.0f, 1.0f, 0.0f); s.N = mk3(0.0f, 0.0f, 1.0f);
        s.wo = mk3(sqrtf(1.0f - cosv * cosv), 0.0f, cosv);
        s.backfacing = false;
        const uint64_t idx = ((uint64_t)e.x << 32) | (uint64_t)e.y;
        FpExact fp;
        const Dielectric r = dielectric_unit(fp, s, ior, rough, 0.0f, sweep_uniform24(seed, 0u, idx), sweep_uniform24(seed, 1u, idx));
        double acc[kSweepValues] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
        sweep_accumulate(acc, r.flags, r.f_r, r.pdf_r, r.w_t, r.F);
#pragma unroll
        for (int v = 0; v < kSweepValues; v++)
            if (acc[v] != 0.0) atomicAdd(table + (size_t)e.x * kSweepValues + v, acc[v]);
    }
}

// ====================================================== launch helpers (one stream)
#define RLS_LAUNCH_CHECK(ctx)                                                    \
    do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, "kernel launch"); \
         (ctx)->launches++; } while (0)
#define RLS_REQUIRE(ctx, cond, msg)                                              \
    do { if (!(cond)) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, msg); } while (0)

// ---- RLS_ARITH_TOLERANT plumbing: the re-run list of the stream a fused kernel is launched on
static int worklist_for(rls_context *ctx, cudaStream_t st, size_t n, tol::Worklist *out)
{
    int slot = 0;
    for (int b = 0; b < kStages; b++) if (ctx->stage_stream[b] && st == ctx->stage_stream[b]) slot = b + 1;
    tol::Worklist &w = ctx->worklist[slot];
    // capacity: 1/16 of the batch (the bands list ~5e-4 of the samples), at least 64 Ki entries; a fuller list falls
    // back to the sentinel scan (rls_tol_launch.cuh), so the capacity is a performance choice, not a correctness one
    size_t want = n / 16 < 65536 ? 65536 : n / 16;
    if (want > n) want = n;
    if (!w.count) {
        RLS_CUDA(ctx, cudaMalloc((void **)&w.count, sizeof(unsigned)));
    }
    if (w.cap < want) {
        RLS_CUDA(ctx, cudaStreamSynchronize(st));        // an earlier launch on this stream may still use the old list
        if (w.list) { RLS_CUDA(ctx, cudaFree(w.list)); w.list = nullptr; w.cap = 0; }
        RLS_CUDA(ctx, cudaMalloc((void **)&w.list, want * sizeof(uint32_t)));
        w.cap = (uint32_t)want;
    }
    RLS_CUDA(ctx, cudaMemsetAsync(w.count, 0, sizeof(unsigned), st));
    *out = w;
    return RLS_OK;
}
// Grid of a re-run kernel: the list length is only known on the device, so the grid is sized for a fraction 1 / `per` of
// the batch (one list slot -- times the group size -- per thread up to that fraction; a thread strides over the list
// beyond it).  CTAs past the end of the list exit at once, at ~0.5 ns each: `per` = 64 for the rlGgx / rlDisney units
// (0.1 % of the samples listed, eight lanes each for rlGgx), 2048 for the skin profile (4e-6 listed: 32 768 idle CTAs cost
// 20 us per 2^28-sample launch).  A grid of 2 CTAs per SM made every thread walk ~4 exact samples one after the other.
static inline unsigned rerun_grid(const rls_context *ctx, size_t n, int block, size_t per = 64)
{
    size_t blocks = (n / per + block - 1) / block;
    const size_t lo = (size_t)ctx->sm_count * 2, hi = (size_t)1 << 20;
    return (unsigned)(blocks < lo ? lo : (blocks > hi ? hi : blocks));
}
#define RLS_TOL_CHECK(ctx, call)                                        \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, "tolerance-policy kernel launch"); \
         (ctx)->launches++; } while (0)

static int launch_ggx_sample_eval_pdf(rls_context *ctx, cudaStream_t st, size_t n, const rls_shading_soa *sg,
                                      const rls_ggx_params *p, const float *rx, const float *ry, const rls_bsdf_out *o)
{
    if (ctx->arith == RLS_ARITH_TOLERANT && p->normal_sampler == RLS_GGX_SAMPLER_VNDF) {
        tol::Worklist wl;
        const int rc = worklist_for(ctx, st, n, &wl);
        if (rc != RLS_OK) return rc;
        RLS_TOL_CHECK(ctx, tol::launch_ggx_sample_eval_pdf(st, n, sh(*sg), dev(*p), rx, ry, dev(*o), wl));
        k_ggx_sample_eval_pdf_rerun<<<rerun_grid(ctx, n, kBlockGgx), kBlockGgx, 0, st>>>((uint32_t)n, sh(*sg), dev(*p), rx, ry, dev(*o), wl, ctx->fallbacks);
        RLS_LAUNCH_CHECK(ctx);
        return RLS_OK;
    }
    if (uses_fast_policy(ctx))
        k_ggx_sample_eval_pdf<true><<<grid_for(n, kBlockGgx), kBlockGgx, 0, st>>>(n, sh(*sg), dev(*p), rx, ry, mv(o->wi), mv(o->f), o->pdf, o->fresnel, o->flags, ctx->fallbacks);
    else
        k_ggx_sample_eval_pdf<false><<<grid_for(n, kBlockGgx), kBlockGgx, 0, st>>>(n, sh(*sg), dev(*p), rx, ry, mv(o->wi), mv(o->f), o->pdf, o->fresnel, o->flags, ctx->fallbacks);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
#ifdef RLS_EXPERIMENTS
#include "experiments/rls_experiments.cuh"
#endif
static int launch_ggx_dielectric(rls_context *ctx, cudaStream_t st, size_t n, const rls_shading_soa *sg,
                                 const rls_ggx_params *p, const float *rx, const float *ry, const rls_ggx_dielectric_out *o)
{
#ifdef RLS_EXPERIMENTS
    {
        bool taken = false;
        const int rc = experiments_launch_ggx_dielectric(ctx, st, n, sg, p, rx, ry, o, &taken);
        if (taken || rc != RLS_OK) return rc;
    }
#endif
    const DielectricOutDev d = dev(*o);
    const GgxParamsDev pd = dev(*p);
    const bool arrays = pd.ior.array && pd.rough.array;
    const bool fast = uses_fast_policy(ctx);
    if (ctx->arith == RLS_ARITH_TOLERANT && !pd.ndf) {
        tol::Worklist wl;
        const int rc = worklist_for(ctx, st, n, &wl);
        if (rc != RLS_OK) return rc;
        RLS_TOL_CHECK(ctx, tol::launch_ggx_dielectric(st, n, sh(*sg), pd, rx, ry, d, wl));
        if (arrays) k_ggx_dielectric_rerun<true><<<rerun_grid(ctx, n, kBlockGgx), kBlockGgx, 0, st>>>((uint32_t)n, sh(*sg), pd, rx, ry, d, wl, ctx->fallbacks);
        else k_ggx_dielectric_rerun<false><<<rerun_grid(ctx, n, kBlockGgx), kBlockGgx, 0, st>>>((uint32_t)n, sh(*sg), pd, rx, ry, d, wl, ctx->fallbacks);
        RLS_LAUNCH_CHECK(ctx);
        return RLS_OK;
    }
#define RLS_DIELECTRIC_LAUNCH(F, A) \
    k_ggx_dielectric<F, A><<<grid_for(n, kBlockGgx), kBlockGgx, 0, st>>>(n, sh(*sg), pd, rx, ry, d, ctx->fallbacks)
    if (fast && arrays) RLS_DIELECTRIC_LAUNCH(true, true);
    else if (fast) RLS_DIELECTRIC_LAUNCH(true, false);
    else if (arrays) RLS_DIELECTRIC_LAUNCH(false, true);
    else RLS_DIELECTRIC_LAUNCH(false, false);
#undef RLS_DIELECTRIC_LAUNCH
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
static bool disney_all_arrays(const DisneyParamsDev &pd)       // every parameter spatially varying?
{
    const P1 *scalars[] = { &pd.subsurface, &pd.metallic, &pd.specular, &pd.specular_tint, &pd.roughness, &pd.anisotropic,
                            &pd.sheen, &pd.sheen_tint, &pd.clearcoat, &pd.clearcoat_gloss };
    bool arrays = pd.base_color.x && pd.base_color.y && pd.base_color.z;
    for (const P1 *q : scalars) arrays = arrays && q->array;
    return arrays;
}
static int launch_disney_sample_eval_pdf(rls_context *ctx, cudaStream_t st, size_t n, const rls_shading_soa *sg,
                                         const rls_disney_params *p, const float *rx_s, const float *ry_s,
                                         const float *rx_d, const float *ry_d, const rls_disney_out *o)
{
    const DisneyOutDev d = dev(*o);
    const DisneyParamsDev pd = dev(*p);
    const bool arrays = disney_all_arrays(pd);
    const bool fast = uses_fast_policy(ctx);
#ifdef RLS_EXPERIMENTS
    {
        bool taken = false;
        const int rc = experiments_launch_disney(ctx, st, n, sg, pd, arrays, rx_s, ry_s, rx_d, ry_d, d, &taken);
        if (taken || rc != RLS_OK) return rc;
    }
#endif
    if (ctx->arith == RLS_ARITH_TOLERANT) {
        tol::Worklist wl;
        const int rc = worklist_for(ctx, st, n, &wl);
        if (rc != RLS_OK) return rc;
        RLS_TOL_CHECK(ctx, tol::launch_disney(st, n, sh(*sg), pd, arrays, rx_s, ry_s, rx_d, ry_d, d, wl));
        if (arrays) k_disney_sample_eval_pdf_rerun<true><<<rerun_grid(ctx, n, kBlock), kBlock, 0, st>>>((uint32_t)n, sh(*sg), pd, rx_s, ry_s, rx_d, ry_d, d, wl, ctx->fallbacks);
        else k_disney_sample_eval_pdf_rerun<false><<<rerun_grid(ctx, n, kBlock), kBlock, 0, st>>>((uint32_t)n, sh(*sg), pd, rx_s, ry_s, rx_d, ry_d, d, wl, ctx->fallbacks);
        RLS_LAUNCH_CHECK(ctx);
        return RLS_OK;
    }
#define RLS_DISNEY_LAUNCH(F, A) \
    k_disney_sample_eval_pdf<F, A><<<grid_for(n), kBlock, 0, st>>>(n, sh(*sg), pd, rx_s, ry_s, rx_d, ry_d, d, ctx->fallbacks)
    if (fast && arrays) RLS_DISNEY_LAUNCH(true, true);
    else if (fast) RLS_DISNEY_LAUNCH(true, false);
    else if (arrays) RLS_DISNEY_LAUNCH(false, true);
    else RLS_DISNEY_LAUNCH(false, false);
#undef RLS_DISNEY_LAUNCH
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
static int launch_skin_profile(rls_context *ctx, cudaStream_t st, size_t n, const rls_skin_params *p,
                               const float *rx, const rls_profile_out *o)
{
    const ProfileOutDev d = dev(*o);
    if (ctx->arith == RLS_ARITH_TOLERANT) {
        tol::Worklist wl;
        const int rc = worklist_for(ctx, st, n, &wl);
        if (rc != RLS_OK) return rc;
        RLS_TOL_CHECK(ctx, tol::launch_skin_profile(st, n, dev(*p), rx, d, wl));
        k_skin_profile_rerun<<<rerun_grid(ctx, n, kBlockSkin, 2048), kBlockSkin, 0, st>>>((uint32_t)n, dev(*p), rx, d, wl, ctx->fallbacks);
        RLS_LAUNCH_CHECK(ctx);
        return RLS_OK;
    }
    if (uses_fast_policy(ctx))
        k_skin_profile<true><<<grid_for(n, kBlockSkin), kBlockSkin, 0, st>>>(n, dev(*p), rx, d, ctx->fallbacks);
    else
        k_skin_profile<false><<<grid_for(n, kBlockSkin), kBlockSkin, 0, st>>>(n, dev(*p), rx, d, ctx->fallbacks);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

static bool ok_ggx_params(const rls_ggx_params *p) { return p && ok_p3(p->KsColor); }
static bool ok_bsdf_out(const rls_bsdf_out *o) { return o && has3(o->wi) && has3(o->f) && o->pdf && o->flags; }
static bool ok_dielectric_out(const rls_ggx_dielectric_out *o)
{
    return o && o->fresnel && has3(o->wi_r) && o->f_r && o->pdf_r && has3(o->wi_t) && o->f_t && o->weight_t && o->flags;
}
static bool ok_disney_out(const rls_disney_out *o)
{
    return o && has3(o->wi_s) && has3(o->f_s) && o->pdf_s && has3(o->wi_d) && has3(o->f_d) && o->pdf_d && o->flags;
}
static bool ok_profile_out(const rls_profile_out *o) { return o && o->r && o->pdf && has3(o->Rd) && o->flags; }
static bool ok_sample_type(int t) { return t == RLS_RAY_DIFFUSE || t == RLS_RAY_GLOSSY; }

// =================================================================== C ABI: rlGgx
extern "C" int rls_ggx_eval_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                   const float *rx, const float *ry, rls_vec3 out_wi, float *out_fresnel)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && rx && ry && has3(out_wi), "rls_ggx_eval_sample: NULL argument");
    DeviceGuard guard(ctx->device);
    k_ggx_eval_sample<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), rx, ry, mv(out_wi), out_fresnel);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ggx_eval_brdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                 rls_cvec3 wi, rls_vec3 out_f)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && has3(wi) && has3(out_f), "rls_ggx_eval_brdf: NULL argument");
    DeviceGuard guard(ctx->device);
    k_ggx_eval_brdf<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), cv(wi), mv(out_f));
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ggx_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                rls_cvec3 wi, float *out_pdf)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && has3(wi) && out_pdf, "rls_ggx_eval_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    k_ggx_eval_pdf<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), cv(wi), out_pdf);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ggx_refract_direction(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                         rls_cvec3 m, rls_vec3 out_wi, uint32_t *out_flags)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && has3(m) && has3(out_wi), "rls_ggx_refract_direction: NULL argument");
    DeviceGuard guard(ctx->device);
    k_ggx_refract_direction<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), cv(m), mv(out_wi), out_flags);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ggx_eval_btdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                 rls_cvec3 wi, float *out_ft)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && has3(wi) && out_ft, "rls_ggx_eval_btdf: NULL argument");
    DeviceGuard guard(ctx->device);
    k_ggx_eval_btdf<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), cv(wi), out_ft);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ggx_sample_weight(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                     rls_cvec3 wi, rls_cvec3 m, float *out_weight)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && has3(wi) && has3(m) && out_weight, "rls_ggx_sample_weight: NULL argument");
    DeviceGuard guard(ctx->device);
    k_ggx_sample_weight<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), cv(wi), cv(m), out_weight);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ggx_sample_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                       const float *rx, const float *ry, const rls_bsdf_out *out)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && rx && ry && ok_bsdf_out(out), "rls_ggx_sample_eval_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    return launch_ggx_sample_eval_pdf(ctx, ctx->stream, n, sg, p, rx, ry, out);
}
extern "C" int rls_ggx_dielectric_sample_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                                  const rls_ggx_params *p, const float *rx, const float *ry,
                                                  const rls_ggx_dielectric_out *out)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && rx && ry && ok_dielectric_out(out),
                "rls_ggx_dielectric_sample_eval_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    return launch_ggx_dielectric(ctx, ctx->stream, n, sg, p, rx, ry, out);
}

// ================================================================ C ABI: rlDisney
static bool ok_disney_params(const rls_disney_params *p) { return p && ok_p3(p->base_color); }

extern "C" int rls_disney_eval_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                                      int sample_type, const float *rx, const float *ry, rls_vec3 out_wi, uint32_t *out_flags)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_disney_params(p) && rx && ry && has3(out_wi), "rls_disney_eval_sample: NULL argument");
    RLS_REQUIRE(ctx, ok_sample_type(sample_type), "rls_disney_eval_sample: sample_type must be RLS_RAY_DIFFUSE or RLS_RAY_GLOSSY");
    DeviceGuard guard(ctx->device);
    k_disney_eval_sample<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), sample_type, rx, ry, mv(out_wi), out_flags);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_disney_eval_brdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                                    int sample_type, rls_cvec3 wi, rls_vec3 out_f)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_disney_params(p) && has3(wi) && has3(out_f), "rls_disney_eval_brdf: NULL argument");
    RLS_REQUIRE(ctx, ok_sample_type(sample_type), "rls_disney_eval_brdf: sample_type must be RLS_RAY_DIFFUSE or RLS_RAY_GLOSSY");
    DeviceGuard guard(ctx->device);
    k_disney_eval_brdf<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), sample_type, cv(wi), mv(out_f));
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_disney_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                                   int sample_type, rls_cvec3 wi, float *out_pdf)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_disney_params(p) && has3(wi) && out_pdf, "rls_disney_eval_pdf: NULL argument");
    RLS_REQUIRE(ctx, ok_sample_type(sample_type), "rls_disney_eval_pdf: sample_type must be RLS_RAY_DIFFUSE or RLS_RAY_GLOSSY");
    DeviceGuard guard(ctx->device);
    k_disney_eval_pdf<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), sample_type, cv(wi), out_pdf);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_disney_sample_eval_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                                          const float *rx_s, const float *ry_s, const float *rx_d, const float *ry_d,
                                          const rls_disney_out *out)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_disney_params(p) && rx_s && ry_s && rx_d && ry_d && ok_disney_out(out),
                "rls_disney_sample_eval_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    return launch_disney_sample_eval_pdf(ctx, ctx->stream, n, sg, p, rx_s, ry_s, rx_d, ry_d, out);
}

// ========================================================== C ABI: rlSss / rlSkin
static bool ok_profile(const rls_ndprofile_soa *s) { return s && has3(s->distance) && has3(s->C1) && has3(s->C2) && s->max_radius; }
static NdProfileSoADev dev(const rls_ndprofile_soa &s)
{
    NdProfileSoADev o; o.distance = mv(s.distance); o.C1 = mv(s.C1); o.C2 = mv(s.C2); o.max_radius = s.max_radius; return o;
}
static bool ok_skin_params(const rls_skin_params *p) { return p && ok_p3(p->sss_color) && ok_p3(p->sss_scatter_dist); }

extern "C" int rls_ndprofile_set_distance(rls_context *ctx, size_t n, rls_cvec3 dist, rls_cvec3 albedo, const rls_ndprofile_soa *out)
{
    (void)albedo;   // only feeds the dead `s` of src/rlSss.cpp:22-23
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, has3(dist) && ok_profile(out), "rls_ndprofile_set_distance: NULL argument");
    DeviceGuard guard(ctx->device);
    k_nd_set_distance<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, cv(dist), dev(*out));
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ndprofile_get_radius(rls_context *ctx, size_t n, const rls_ndprofile_soa *profile, const float *rx,
                                        float *out_r, uint32_t *out_flags)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_profile(profile) && rx && out_r, "rls_ndprofile_get_radius: NULL argument");
    DeviceGuard guard(ctx->device);
    k_nd_get_radius<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, dev(*profile), rx, out_r, out_flags);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ndprofile_get_pdf(rls_context *ctx, size_t n, const rls_ndprofile_soa *profile, const float *r, float *out_pdf)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_profile(profile) && r && out_pdf, "rls_ndprofile_get_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    k_nd_get_pdf<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, dev(*profile), r, out_pdf);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_ndprofile_eval_profile(rls_context *ctx, size_t n, const rls_ndprofile_soa *profile, const float *r, rls_vec3 out_rd)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_profile(profile) && r && has3(out_rd), "rls_ndprofile_eval_profile: NULL argument");
    DeviceGuard guard(ctx->device);
    k_nd_eval_profile<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, dev(*profile), r, mv(out_rd));
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
// ---- GaussianProfile (src/rlSss.h:63-97)
static bool ok_gauss(const rls_gaussprofile_soa *s) { return s && s->variance && s->max_radius && s->norm; }
static GaussProfileSoADev dev(const rls_gaussprofile_soa &s)
{
    GaussProfileSoADev o; o.variance = s.variance; o.max_radius = s.max_radius; o.norm = s.norm; return o;
}
extern "C" int rls_gaussprofile_set_distance(rls_context *ctx, size_t n, rls_cvec3 dist, rls_cvec3 albedo,
                                             const rls_gaussprofile_soa *out)
{
    (void)albedo;   // unused by the reference too (src/rlSss.h:71-76)
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, dist.x && ok_gauss(out), "rls_gaussprofile_set_distance: NULL argument");
    DeviceGuard guard(ctx->device);
    k_gauss_set_distance<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, dist.x, dev(*out));
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
template <int kOp>
static int gauss_unary(rls_context *ctx, size_t n, const rls_gaussprofile_soa *profile, const float *x, float *out,
                       const char *what)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_gauss(profile) && x && out, what);
    DeviceGuard guard(ctx->device);
    k_gauss_unary<kOp><<<grid_for(n), kBlock, 0, ctx->stream>>>(n, dev(*profile), x, out);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_gaussprofile_get_radius(rls_context *ctx, size_t n, const rls_gaussprofile_soa *profile,
                                           const float *rx, float *out_r)
{
    return gauss_unary<0>(ctx, n, profile, rx, out_r, "rls_gaussprofile_get_radius: NULL argument");
}
extern "C" int rls_gaussprofile_get_pdf(rls_context *ctx, size_t n, const rls_gaussprofile_soa *profile,
                                        const float *r, float *out_pdf)
{
    return gauss_unary<1>(ctx, n, profile, r, out_pdf, "rls_gaussprofile_get_pdf: NULL argument");
}
extern "C" int rls_gaussprofile_eval_profile(rls_context *ctx, size_t n, const rls_gaussprofile_soa *profile,
                                             const float *r, float *out_rd)
{
    return gauss_unary<2>(ctx, n, profile, r, out_rd, "rls_gaussprofile_eval_profile: NULL argument");
}
extern "C" int rls_gaussprofile_sample_eval_pdf(rls_context *ctx, size_t n, const float *dist_x, const float *rx,
                                                float *out_r, float *out_pdf, float *out_rd)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, dist_x && rx && out_r && out_pdf && out_rd, "rls_gaussprofile_sample_eval_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    const bool fast = uses_fast_policy(ctx);
    bool aligned = (((uintptr_t)dist_x | (uintptr_t)rx | (uintptr_t)out_r | (uintptr_t)out_pdf | (uintptr_t)out_rd) & 15u) == 0;
#ifdef RLS_EXPERIMENTS
    aligned = aligned && !ctx->exp.gauss_scalar;      // A/B: one sample per thread
#endif
    const size_t n4 = aligned ? n / 4 : 0, done = n4 * 4;
    if (n4) {
        if (fast)
            k_gauss_profile_x4<true><<<grid_for(n4, kBlockSkin), kBlockSkin, 0, ctx->stream>>>(
                n4, (const float4 *)dist_x, (const float4 *)rx, (float4 *)out_r, (float4 *)out_pdf, (float4 *)out_rd, ctx->fallbacks);
        else
            k_gauss_profile_x4<false><<<grid_for(n4, kBlockSkin), kBlockSkin, 0, ctx->stream>>>(
                n4, (const float4 *)dist_x, (const float4 *)rx, (float4 *)out_r, (float4 *)out_pdf, (float4 *)out_rd, ctx->fallbacks);
        RLS_LAUNCH_CHECK(ctx);
    }
    if (done == n) return RLS_OK;
    const size_t m = n - done;
    if (fast)
        k_gauss_profile<true><<<grid_for(m, kBlockSkin), kBlockSkin, 0, ctx->stream>>>(m, dist_x + done, rx + done, out_r + done, out_pdf + done, out_rd + done, ctx->fallbacks);
    else
        k_gauss_profile<false><<<grid_for(m, kBlockSkin), kBlockSkin, 0, ctx->stream>>>(m, dist_x + done, rx + done, out_r + done, out_pdf + done, out_rd + done, ctx->fallbacks);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_skin_profile_sample_eval_pdf(rls_context *ctx, size_t n, const rls_skin_params *p, const float *rx,
                                                const rls_profile_out *out)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_skin_params(p) && rx && ok_profile_out(out), "rls_skin_profile_sample_eval_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    return launch_skin_profile(ctx, ctx->stream, n, p, rx, out);
}
extern "C" int rls_skin_layer_weights(rls_context *ctx, size_t n, const rls_skin_params *p, const float *avg_sheen,
                                      const float *avg_spec, float *out_spec_scale, float *out_sss_weight)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, p && avg_sheen && avg_spec && out_spec_scale && out_sss_weight, "rls_skin_layer_weights: NULL argument");
    DeviceGuard guard(ctx->device);
    k_skin_layer_weights<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, dev(*p), avg_sheen, avg_spec, out_spec_scale, out_sss_weight);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

extern "C" int rls_skin_probe_ray(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_skin_params *p,
                                  const float *rx, const float *ry, const rls_probe_out *out)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_skin_params(p) && rx && ry && out && out->r && has3(out->origin) && has3(out->dir) &&
                out->maxdist && out->flags, "rls_skin_probe_ray: NULL argument");
    DeviceGuard guard(ctx->device);
    ProbeOutDev d; d.r = out->r; d.origin = mv(out->origin); d.dir = mv(out->dir); d.maxdist = out->maxdist; d.flags = out->flags;
    k_skin_probe_ray<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), rx, ry, d);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_skin_probe_mis_pdf(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_skin_params *p,
                                      rls_cvec3 disp, rls_cvec3 hit_normal, float *out_pdf)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_skin_params(p) && has3(disp) && has3(hit_normal) && out_pdf,
                "rls_skin_probe_mis_pdf: NULL argument");
    DeviceGuard guard(ctx->device);
    k_skin_probe_mis_pdf<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), cv(disp), cv(hit_normal), out_pdf);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

// ================================================================= C ABI: sweep
// Validates the grid and enqueues the sweep of samples [spp_begin, spp_end) of every cell on `st` (the entry point below
// and the multi-device form rls_multi_albedo_sweep share it).  `slot` names the re-run list (RLS_ARITH_TOLERANT).
static int launch_albedo_sweep(rls_context *ctx, cudaStream_t st, const rls_sweep_grid *grid, uint64_t seed,
                               uint32_t spp_begin, uint32_t spp_end, double *table)
{
    RLS_REQUIRE(ctx, grid && table, "rls_albedo_sweep: NULL argument");
    RLS_REQUIRE(ctx, grid->n_rough > 0 && grid->n_cos > 0 && grid->n_ior > 0 && spp_end >= spp_begin, "rls_albedo_sweep: bad grid");
    const uint64_t cells64 = (uint64_t)grid->n_rough * (uint64_t)grid->n_cos * (uint64_t)grid->n_ior;
    RLS_REQUIRE(ctx, cells64 <= 0x7fffffffull / 32ull, "rls_albedo_sweep: more than 2^26 - 1 cells (one warp per cell)");
    RLS_REQUIRE(ctx, spp_end <= 0xffffffffu - 32u, "rls_albedo_sweep: spp_end must be at most 2^32 - 33");
    SweepGridDev g; g.n_rough = grid->n_rough; g.n_cos = grid->n_cos; g.n_ior = grid->n_ior;
    g.rlo = grid->roughness_lo; g.rhi = grid->roughness_hi; g.ilo = grid->ior_lo; g.ihi = grid->ior_hi;
    const uint32_t cells = (uint32_t)cells64;
    const unsigned blocks = (unsigned)((cells64 * 32ull + kSweepBlock - 1) / kSweepBlock);
    if (ctx->arith == RLS_ARITH_TOLERANT) {
        // the band samples go to a (cell, k) list and are added to the table by the exact policy afterwards
        tol::SweepWorklist &w = ctx->sweep_worklist;
        const uint32_t want = 1u << 22;
        if (!w.count) RLS_CUDA(ctx, cudaMalloc((void **)&w.count, sizeof(unsigned)));
        if (!w.list) { RLS_CUDA(ctx, cudaMalloc((void **)&w.list, (size_t)want * sizeof(uint2))); w.cap = want; }
        RLS_CUDA(ctx, cudaMemsetAsync(w.count, 0, sizeof(unsigned), st));
        RLS_TOL_CHECK(ctx, tol::launch_albedo_sweep(st, g, cells, seed, spp_begin, spp_end, table, w));
        k_albedo_sweep_rerun<<<rerun_grid(ctx, (size_t)cells * (spp_end - spp_begin), kSweepBlock), kSweepBlock, 0, st>>>(g, seed, w, table, ctx->fallbacks);
        RLS_LAUNCH_CHECK(ctx);
        return RLS_OK;
    }
    k_albedo_sweep<<<blocks, kSweepBlock, 0, st>>>(g, cells, seed, spp_begin, spp_end, table);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_albedo_sweep(rls_context *ctx, const rls_sweep_grid *grid, uint64_t seed, uint32_t spp_begin,
                                uint32_t spp_end, double *table)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    return launch_albedo_sweep(ctx, ctx->stream, grid, seed, spp_begin, spp_end, table);
}

// ================================================================= C ABI: synth
extern "C" int rls_synth_uniform(rls_context *ctx, size_t n, uint64_t seed, uint32_t stream, uint64_t first_index,
                                 float lo, float hi, float *out)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, out, "rls_synth_uniform: NULL argument");
    DeviceGuard guard(ctx->device);
    k_synth_uniform<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, seed, stream, first_index, lo, hi, out);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_synth_shading(rls_context *ctx, size_t n, uint64_t seed, uint64_t first_index, float cos_lo, float cos_hi,
                                 float backfacing_fraction, const rls_shading_soa *sg)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg), "rls_synth_shading: NULL argument");
    DeviceGuard guard(ctx->device);
    V3 U = { (float *)sg->U.x, (float *)sg->U.y, (float *)sg->U.z };
    V3 V = { (float *)sg->V.x, (float *)sg->V.y, (float *)sg->V.z };
    V3 N = { (float *)sg->N.x, (float *)sg->N.y, (float *)sg->N.z };
    V3 W = { (float *)sg->wo.x, (float *)sg->wo.y, (float *)sg->wo.z };
    k_synth_shading<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, seed, first_index, cos_lo, cos_hi, backfacing_fraction,
                                                            U, V, N, W, (uint8_t *)sg->backfacing);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

// ===================================================== C ABI: callers (8(f) f2-f4)
static inline bool none3(const rls_cvec3 &v) { return !v.x && !v.y && !v.z; }
static inline SkinLayerDev layer_dev(const rls_param3 &c, const rls_param1 &w, const rls_param1 &r, const rls_param1 &ior)
{
    SkinLayerDev o; o.color = p3(c); o.weight = p1(w); o.roughness = p1(r); o.ior = p1(ior); return o;
}
extern "C" int rls_skin_glossy_layers(rls_context *ctx, size_t n, uint32_t k, const rls_shading_soa *sg,
                                      const rls_skin_params *p, const float *rx_a, const float *ry_a,
                                      const float *rx_b, const float *ry_b, rls_cvec3 li_a, rls_cvec3 li_b,
                                      const rls_skin_layers_out *out)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, k >= 1, "rls_skin_glossy_layers: k must be at least 1");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_skin_params(p) && ok_p3(p->sheen_color) && ok_p3(p->specular_color) &&
                rx_a && ry_a && rx_b && ry_b && out && has3(out->sheen) && has3(out->specular) && out->sheen_fresnel &&
                out->specular_fresnel && out->sss_weight && out->flags, "rls_skin_glossy_layers: NULL argument");
    RLS_REQUIRE(ctx, (none3(li_a) || has3(li_a)) && (none3(li_b) || has3(li_b)),
                "rls_skin_glossy_layers: a radiance input must have all three channels or none");
    DeviceGuard guard(ctx->device);
    SkinLayersDev sp;
    sp.sheen = layer_dev(p->sheen_color, p->sheen_weight, p->sheen_roughness, p->sheen_ior);
    sp.spec = layer_dev(p->specular_color, p->specular_weight, p->specular_roughness, p->specular_ior);
    sp.sss_weight = p1(p->sss_weight);
    SkinLayersOutDev o; o.sheen = mv(out->sheen); o.spec = mv(out->specular); o.sheenF = out->sheen_fresnel;
    o.specF = out->specular_fresnel; o.sssW = out->sss_weight; o.flags = out->flags;
    if (uses_fast_policy(ctx))
        k_skin_glossy_layers<true><<<grid_for(n), kBlock, 0, ctx->stream>>>(n, k, sh(*sg), sp, rx_a, ry_a, rx_b, ry_b, cv(li_a), cv(li_b), o, ctx->fallbacks);
    else
        k_skin_glossy_layers<false><<<grid_for(n), kBlock, 0, ctx->stream>>>(n, k, sh(*sg), sp, rx_a, ry_a, rx_b, ry_b, cv(li_a), cv(li_b), o, ctx->fallbacks);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

static inline bool ok_light(const rls_light_sample *l, bool need_dir)
{
    return l && (!need_dir || has3(l->dir)) && has3(l->radiance) && l->pdf;
}
static inline LightDev light_dev(const rls_light_sample *l)
{
    LightDev o; o.dir = CV3{ nullptr, nullptr, nullptr }; o.radiance = o.dir; o.pdf = nullptr;
    if (l) { o.dir = cv(l->dir); o.radiance = cv(l->radiance); o.pdf = l->pdf; }
    return o;
}
extern "C" int rls_ggx_evaluate_light_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                             const rls_light_sample *light, const float *rx, const float *ry,
                                             const rls_light_sample *at_l, rls_vec3 out_rgb, float *w_light, float *w_brdf)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && ok_light(light, true) && has3(out_rgb) &&
                (!at_l || (ok_light(at_l, false) && rx && ry)), "rls_ggx_evaluate_light_sample: NULL argument");
    DeviceGuard guard(ctx->device);
    k_ggx_light_sample<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), light_dev(light), rx, ry, light_dev(at_l),
                                                               mv(out_rgb), w_light, w_brdf);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_disney_evaluate_light_sample(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                                                int sample_type, const rls_light_sample *light, const float *rx, const float *ry,
                                                const rls_light_sample *at_l, rls_vec3 out_rgb, float *w_light, float *w_brdf)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_disney_params(p) && ok_light(light, true) && has3(out_rgb) &&
                (!at_l || (ok_light(at_l, false) && rx && ry)), "rls_disney_evaluate_light_sample: NULL argument");
    RLS_REQUIRE(ctx, ok_sample_type(sample_type), "rls_disney_evaluate_light_sample: sample_type must be RLS_RAY_DIFFUSE or RLS_RAY_GLOSSY");
    DeviceGuard guard(ctx->device);
    k_disney_light_sample<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), sample_type, light_dev(light), rx, ry,
                                                                  light_dev(at_l), mv(out_rgb), w_light, w_brdf);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

extern "C" int rls_sample_writer_radiance(rls_context *ctx, int node, const rls_shading_soa *sg, const void *params, size_t point,
                                          int sample_type, int width, int height, float *image)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    RLS_REQUIRE(ctx, (node == RLS_NODE_GGX || node == RLS_NODE_DISNEY) && ok_shading(sg) && params && image,
                "rls_sample_writer_radiance: bad argument");
    RLS_REQUIRE(ctx, width > 0 && height > 0 && (size_t)width * height < (1ull << 31) && !(point >> 32),
                "rls_sample_writer_radiance: bad image size or point index");
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)width * height;
    if (node == RLS_NODE_GGX) {
        const rls_ggx_params *p = (const rls_ggx_params *)params;
        RLS_REQUIRE(ctx, ok_ggx_params(p), "rls_sample_writer_radiance: bad rlGgx parameters");
        k_writer_radiance_ggx<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), (uint32_t)point, width, height, image);
    } else {
        const rls_disney_params *p = (const rls_disney_params *)params;
        RLS_REQUIRE(ctx, ok_disney_params(p) && ok_sample_type(sample_type), "rls_sample_writer_radiance: bad rlDisney parameters");
        k_writer_radiance_disney<<<grid_for(n), kBlock, 0, ctx->stream>>>(n, sh(*sg), dev(*p), sample_type, (uint32_t)point, width, height, image);
    }
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_sample_writer_scatter(rls_context *ctx, int node, const rls_shading_soa *sg, const void *params, size_t point,
                                         int sample_type, size_t n_samples, const float *rx, const float *ry, int width,
                                         int height, float *image, uint32_t *scratch, uint32_t *out_missing)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    RLS_REQUIRE(ctx, (node == RLS_NODE_GGX || node == RLS_NODE_DISNEY) && ok_shading(sg) && params && image && scratch &&
                (n_samples == 0 || (rx && ry)), "rls_sample_writer_scatter: bad argument");
    RLS_REQUIRE(ctx, width > 0 && height > 0 && (size_t)width * height < (1ull << 31) && !(point >> 32) && n_samples < (1ull << 31),
                "rls_sample_writer_scatter: bad image size, point index or sample count");
    DeviceGuard guard(ctx->device);
    const size_t npix = (size_t)width * height;
    RLS_CUDA(ctx, cudaMemsetAsync(scratch, 0, npix * sizeof(uint32_t), ctx->stream));
    if (out_missing) RLS_CUDA(ctx, cudaMemsetAsync(out_missing, 0, sizeof(uint32_t), ctx->stream));
    if (n_samples) {
        if (node == RLS_NODE_GGX) {
            const rls_ggx_params *p = (const rls_ggx_params *)params;
            RLS_REQUIRE(ctx, ok_ggx_params(p), "rls_sample_writer_scatter: bad rlGgx parameters");
            k_writer_scatter_ggx<<<grid_for(n_samples), kBlock, 0, ctx->stream>>>(n_samples, sh(*sg), dev(*p), (uint32_t)point, rx, ry,
                                                                                  width, height, scratch, out_missing);
        } else {
            const rls_disney_params *p = (const rls_disney_params *)params;
            RLS_REQUIRE(ctx, ok_disney_params(p) && ok_sample_type(sample_type), "rls_sample_writer_scatter: bad rlDisney parameters");
            k_writer_scatter_disney<<<grid_for(n_samples), kBlock, 0, ctx->stream>>>(n_samples, sh(*sg), dev(*p), sample_type, (uint32_t)point,
                                                                                     rx, ry, width, height, scratch, out_missing);
        }
        RLS_LAUNCH_CHECK(ctx);
    }
    k_writer_scatter_paint<<<grid_for(npix), kBlock, 0, ctx->stream>>>(npix, width, height, scratch, image);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

// ================================================================= C ABI: diagnostics
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_debug_libm(int fn, size_t n, const float *a, const float *b, float *out0, float *out1)
{
    RLS_INDEX();
    float x = a[i], y = b ? b[i] : 0.0f, r0 = 0.0f, r1 = 0.0f;
    switch (fn) {
    case 0: rlm::sincosf_(x, &r0, &r1); break;
    case 1: r0 = rlm::tanf_(x); break;
    case 2: r0 = rlm::atanf_(x); break;
    case 3: r0 = rlm::acosf_(x); break;
    case 4: r0 = rlm::expf_(x); break;
    case 5: r0 = rlm::logf_(x); break;
    case 6: r0 = rlm::atan2f_(x, y); break;
    default: r0 = rlm::powf_(x, y); break;
    }
    out0[i] = r0;
    if (out1) out1[i] = r1;
}
extern "C" int rls_debug_libm(rls_context *ctx, int fn, size_t n, const float *a, const float *b, float *out0, float *out1)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, fn >= 0 && fn <= 7 && a && out0 && (fn < 6 || b), "rls_debug_libm: bad argument");
    DeviceGuard guard(ctx->device);
    k_debug_libm<<<grid_for(n), kBlock, 0, ctx->stream>>>(fn, n, a, b, out0, out1);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

// Fast-policy == exact-policy over RANGES OF BIT PATTERNS (no input arrays): argument k of the range
// is the binary32 with bits first + k * stride.  For every argument whose fast-policy evaluation
// leaves the operand tracker satisfied (FpFast::ok) the result must equal the exact policy's bit
// for bit; counts[0] = such arguments, counts[1] = mismatches among them, counts[2] = arguments sent
// to the exact re-run.  fn: 0 sqrt(a), 1 1/a, 2 a/b, 3 tanf(a), 4 acosf(a), 5 atan2f(a, b),
// 6 atan2f(b, a), 7 a/b with a zero-tolerant numerator (div_pz, b > 0), 8 b/a, 9 a/3 (div3),
// 10 a/b through the shared refined reciprocal of b (shared_rcp + div_by), 11 expf(a), 12 sincosf(a).
__global__ void __launch_bounds__(kBlock, RLS_MIN_BLOCKS)
k_debug_policy_check(int fn, uint32_t first, uint64_t count, uint32_t stride, float b, unsigned long long *counts)
{
    rlm::smem_tables_init();                 // FpFastLeanTrig reads its tables from shared memory
    unsigned long long okc = 0, bad = 0, rerun = 0;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x) {
        const float a = __uint_as_float(first + (uint32_t)k * stride);
        FpFastLeanTrig ff; FpExact fe;       // the lean-trigonometry flavour, so that fn 12 checks the branch-free sincosf
        float rf, re;
        switch (fn) {
        case 0: rf = ff.sqrt(a); re = fe.sqrt(a); break;
        case 1: rf = ff.rcp(a); re = fe.rcp(a); break;
        case 2: rf = ff.div(a, b); re = fe.div(a, b); break;
        case 3: rf = rlm::tanf_(ff, a); re = rlm::tanf_(fe, a); break;
        case 4: rf = rlm::acosf_(ff, a); re = rlm::acosf_(fe, a); break;
        case 5: rf = rlm::atan2f_(ff, a, b); re = rlm::atan2f_(fe, a, b); break;
        case 6: rf = rlm::atan2f_(ff, b, a); re = rlm::atan2f_(fe, b, a); break;
        case 7: rf = ff.div_pz(a, b); re = fe.div_pz(a, b); break;
        case 8: rf = ff.div(b, a); re = fe.div(b, a); break;
        case 9: rf = ff.div3(a); re = fe.div3(a); break;
        case 10: rf = ff.div_by(a, b, ff.shared_rcp(b)); re = fe.div(a, b); break;
        case 11: rf = rlm::expf_(ff, a); re = rlm::expf_(fe, a); break;
        default: {      // sincosf: both results must match (the sine is returned, the cosine is folded in)
            float sf, cf, se, ce;
            rlm::sincosf_(ff, a, &sf, &cf); rlm::sincosf_(fe, a, &se, &ce);
            const bool cos_same = __float_as_uint(cf) == __float_as_uint(ce) || (cf != cf && ce != ce);
            rf = sf; re = cos_same ? se : __uint_as_float(__float_as_uint(sf) ^ 1u);
            break; }
        }
        if (ff.ok()) {
            okc++;
            const bool both_nan = (rf != rf) && (re != re);
            if (__float_as_uint(rf) != __float_as_uint(re) && !both_nan) bad++;
        } else {
            rerun++;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        okc += __shfl_down_sync(0xffffffffu, okc, o);
        bad += __shfl_down_sync(0xffffffffu, bad, o);
        rerun += __shfl_down_sync(0xffffffffu, rerun, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counts + 0, okc);
        atomicAdd(counts + 1, bad);
        atomicAdd(counts + 2, rerun);
    }
}
extern "C" int rls_debug_policy_check(rls_context *ctx, int fn, uint32_t first_bits, uint64_t count, uint32_t stride, float b,
                                      unsigned long long *counts)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    RLS_REQUIRE(ctx, fn >= 0 && fn <= 12 && counts && stride >= 1, "rls_debug_policy_check: bad argument");
    if (count == 0) return RLS_OK;
    DeviceGuard guard(ctx->device);
    const uint64_t blocks = (count + kBlock - 1) / kBlock;
    const unsigned grid = (unsigned)(blocks < (uint64_t)ctx->sm_count * 32 ? blocks : (uint64_t)ctx->sm_count * 32);
    k_debug_policy_check<<<grid, kBlock, 0, ctx->stream>>>(fn, first_bits, count, stride, b, counts);
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}

// ===================================================== compact frames (include/rls_b200.h rls_shading_quat_soa)
// One rounding per written operation (this translation unit is compiled -fmad=false): the frame bits are a function of
// the quaternion bits alone, the same on the host (the test suite restates the decode in numpy).
__global__ void __launch_bounds__(256)
k_frame_from_quat(uint32_t n, const float *qx, const float *qy, const float *qz, const float *qw, V3 U, V3 V, V3 N)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(qx + i), y = __ldg(qy + i), z = __ldg(qz + i), w = __ldg(qw + i);
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x * x2, yy = y * y2, zz = z * z2, xy = x * y2, xz = x * z2, yz = y * z2;
    const float wx = w * x2, wy = w * y2, wz = w * z2;
    U.x[i] = 1.0f - (yy + zz); U.y[i] = xy + wz;          U.z[i] = xz - wy;
    V.x[i] = xy - wz;          V.y[i] = 1.0f - (xx + zz); V.z[i] = yz + wx;
    N.x[i] = xz + wy;          N.y[i] = yz - wx;          N.z[i] = 1.0f - (xx + yy);
}
static int launch_frame_from_quat(rls_context *ctx, cudaStream_t st, size_t n, const float *qx, const float *qy, const float *qz,
                                  const float *qw, const rls_vec3 &U, const rls_vec3 &V, const rls_vec3 &N)
{
    k_frame_from_quat<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((uint32_t)n, qx, qy, qz, qw, mv(U), mv(V), mv(N));
    RLS_LAUNCH_CHECK(ctx);
    return RLS_OK;
}
extern "C" int rls_frame_from_quaternion(rls_context *ctx, size_t n, const float *qx, const float *qy, const float *qz,
                                         const float *qw, rls_vec3 out_U, rls_vec3 out_V, rls_vec3 out_N)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, qx && qy && qz && qw && out_U.x && out_U.y && out_U.z && out_V.x && out_V.y && out_V.z &&
                out_N.x && out_N.y && out_N.z, "rls_frame_from_quaternion: NULL argument");
    DeviceGuard guard(ctx->device);
    return launch_frame_from_quat(ctx, ctx->stream, n, qx, qy, qz, qw, out_U, out_V, out_N);
}

// ===================================================== host-buffer (end-to-end) forms
// A chunk of samples is staged through one of kStages device buffers: H2D of every input
// slice, the kernel, D2H of every output slice, all on that stage's stream, so that chunk
// k+1's upload and chunk k-1's download overlap chunk k's kernel (PCIe is full duplex).
namespace {

struct Stager {
    rls_context *ctx;
    size_t chunk;        // samples per chunk (capacity)
    size_t first = 0;    // first sample of the current chunk
    size_t count = 0;    // samples in the current chunk
    int    stage = 0;
    size_t offset = 0;   // bump pointer into the stage buffer
    cudaError_t err = cudaSuccess;
    struct Pending { void *host; const void *dev; size_t bytes; };
    std::vector<Pending> downloads;

    static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

    void *slot(size_t elem)
    {
        void *p = (char *)ctx->stage_buf[stage] + offset;
        offset += align256(chunk * elem);
        return p;
    }
    template <typename T> const T *in(const T *host)
    {
        if (!host) return nullptr;
        T *d = (T *)slot(sizeof(T));
        cudaError_t e = cudaMemcpyAsync(d, host + first, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stage_stream[stage]);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        return d;
    }
    template <typename T> T *out(T *host)
    {
        if (!host) return nullptr;
        T *d = (T *)slot(sizeof(T));
        downloads.push_back({ (void *)(host + first), d, count * sizeof(T) });
        return d;
    }
    rls_cvec3 in3(const rls_cvec3 &v) { rls_cvec3 o; o.x = in(v.x); o.y = in(v.y); o.z = in(v.z); return o; }
    rls_vec3 out3(const rls_vec3 &v) { rls_vec3 o; o.x = out(v.x); o.y = out(v.y); o.z = out(v.z); return o; }
    rls_param1 in1(const rls_param1 &p) { rls_param1 o = p; o.array = in(p.array); return o; }
    rls_param3 inp3(const rls_param3 &p) { rls_param3 o = p; o.array = in3(p.array); return o; }
    rls_shading_soa shading(const rls_shading_soa &s)
    {
        rls_shading_soa o; o.U = in3(s.U); o.V = in3(s.V); o.N = in3(s.N); o.wo = in3(s.wo); o.backfacing = in(s.backfacing); return o;
    }
    template <typename T> T *scratch() { return (T *)slot(sizeof(T)); }
    // Compact frame: q, wo, backfacing are uploaded; U, V, N are decoded into this stage's scratch on the stage's stream.
    int shading_from_quat(const rls_shading_quat_soa &q, rls_shading_soa *o)
    {
        const float *x = in(q.qx), *y = in(q.qy), *z = in(q.qz), *w = in(q.qw);
        rls_vec3 U = { scratch<float>(), scratch<float>(), scratch<float>() };
        rls_vec3 V = { scratch<float>(), scratch<float>(), scratch<float>() };
        rls_vec3 N = { scratch<float>(), scratch<float>(), scratch<float>() };
        o->U = { U.x, U.y, U.z }; o->V = { V.x, V.y, V.z }; o->N = { N.x, N.y, N.z };
        o->wo = in3(q.wo); o->backfacing = in(q.backfacing);
        return launch_frame_from_quat(ctx, ctx->stage_stream[stage], count, x, y, z, w, U, V, N);
    }
    void finish()
    {
        for (const Pending &d : downloads) {
            cudaError_t e = cudaMemcpyAsync(d.host, d.dev, d.bytes, cudaMemcpyDeviceToHost, ctx->stage_stream[stage]);
            if (e != cudaSuccess && err == cudaSuccess) err = e;
        }
        downloads.clear();
        cudaError_t e = cudaEventRecord(ctx->stage_done[stage], ctx->stage_stream[stage]);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
    }
};

int stage_prepare(rls_context *ctx, size_t bytes_per_stage)
{
    for (int b = 0; b < kStages; b++) {
        if (!ctx->stage_stream[b]) RLS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stage_stream[b], cudaStreamNonBlocking));
        if (!ctx->stage_done[b]) RLS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_done[b], cudaEventDisableTiming));
    }
    if (bytes_per_stage > ctx->stage_bytes) {
        for (int b = 0; b < kStages; b++) {
            if (ctx->stage_buf[b]) { RLS_CUDA(ctx, cudaFree(ctx->stage_buf[b])); ctx->stage_buf[b] = nullptr; }
        }
        ctx->stage_bytes = 0;
        for (int b = 0; b < kStages; b++) RLS_CUDA(ctx, cudaMalloc(&ctx->stage_buf[b], bytes_per_stage));
        ctx->stage_bytes = bytes_per_stage;
    }
    return RLS_OK;
}

// Runs `body(stager)` once per chunk; `arrays_bytes_per_sample` bounds the stage size.
template <typename Body>
int run_staged(rls_context *ctx, size_t n, size_t chunk, size_t slots, Body body)
{
    if (chunk == 0) chunk = (size_t)1 << 21;     // 8 MiB per array copy: +1-5 % over 4 MiB on PCIe Gen5 (tools/pcie_granularity_probe.py)
    if (chunk > n) chunk = n;
    chunk = (chunk + 63) & ~(size_t)63;
    DeviceGuard guard(ctx->device);
    int rc = stage_prepare(ctx, slots * Stager::align256(chunk * sizeof(float)));
    if (rc != RLS_OK) return rc;
    // The staging streams are private: order them after whatever the caller already queued on the context's stream.
    cudaEvent_t entry = nullptr;
    RLS_CUDA(ctx, cudaEventCreateWithFlags(&entry, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(entry, ctx->stream);
    for (int b = 0; b < kStages && e == cudaSuccess; b++) e = cudaStreamWaitEvent(ctx->stage_stream[b], entry, 0);
    cudaEventDestroy(entry);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "host staging: ordering after the context stream");
    Stager st{ ctx, chunk };
    size_t c = 0;
    // On ANY failure every staging stream is drained before returning: copies of earlier chunks may still be reading
    // or writing the caller's host buffers.
    auto drain = [&]() { for (int b = 0; b < kStages; b++) if (ctx->stage_stream[b]) cudaStreamSynchronize(ctx->stage_stream[b]); };
    for (size_t first = 0; first < n; first += chunk, c++) {
        st.stage = (int)(c % kStages);
        st.first = first;
        st.count = (n - first < chunk) ? (n - first) : chunk;
        st.offset = 0;
        if (c >= (size_t)kStages) {
            e = cudaEventSynchronize(ctx->stage_done[st.stage]);
            if (e != cudaSuccess) { drain(); return cuda_fail(ctx, e, "host staging: waiting for a stage"); }
        }
        rc = body(st);
        st.finish();                   // the downloads of this chunk are queued even when its launch failed: nothing is dropped
        if (rc != RLS_OK) { drain(); return rc; }
        if (st.err != cudaSuccess) { drain(); return cuda_fail(ctx, st.err, "host staging copy"); }
    }
    for (int b = 0; b < kStages; b++) {
        e = cudaStreamSynchronize(ctx->stage_stream[b]);
        if (e != cudaSuccess) { drain(); return cuda_fail(ctx, e, "host staging: final synchronisation"); }
    }
    return RLS_OK;
}

} // namespace

extern "C" int rls_ggx_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_ggx_params *p,
                                            const float *rx, const float *ry, const rls_bsdf_out *out, size_t chunk)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && rx && ry && ok_bsdf_out(out), "rls_ggx_sample_eval_pdf_host: NULL argument");
    return run_staged(ctx, n, chunk, 32, [&](Stager &st) {
        rls_shading_soa s = st.shading(*sg);
        rls_ggx_params q = *p;
        q.KsColor = st.inp3(p->KsColor); q.specularRoughness = st.in1(p->specularRoughness);
        q.ior = st.in1(p->ior); q.anisotropic = st.in1(p->anisotropic);
        const float *drx = st.in(rx), *dry = st.in(ry);
        rls_bsdf_out o; o.wi = st.out3(out->wi); o.f = st.out3(out->f); o.pdf = st.out(out->pdf);
        o.fresnel = st.out(out->fresnel); o.flags = st.out(out->flags);
        return launch_ggx_sample_eval_pdf(ctx, ctx->stage_stream[st.stage], st.count, &s, &q, drx, dry, &o);
    });
}
extern "C" int rls_ggx_dielectric_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_shading_soa *sg,
                                                       const rls_ggx_params *p, const float *rx, const float *ry,
                                                       const rls_ggx_dielectric_out *out, size_t chunk)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_ggx_params(p) && rx && ry && ok_dielectric_out(out),
                "rls_ggx_dielectric_sample_eval_pdf_host: NULL argument");
    return run_staged(ctx, n, chunk, 36, [&](Stager &st) {
        rls_shading_soa s = st.shading(*sg);
        rls_ggx_params q = *p;
        q.KsColor = st.inp3(p->KsColor); q.specularRoughness = st.in1(p->specularRoughness);
        q.ior = st.in1(p->ior); q.anisotropic = st.in1(p->anisotropic);
        const float *drx = st.in(rx), *dry = st.in(ry);
        rls_ggx_dielectric_out o;
        o.fresnel = st.out(out->fresnel); o.wi_r = st.out3(out->wi_r); o.f_r = st.out(out->f_r); o.pdf_r = st.out(out->pdf_r);
        o.wi_t = st.out3(out->wi_t); o.f_t = st.out(out->f_t); o.weight_t = st.out(out->weight_t); o.flags = st.out(out->flags);
        return launch_ggx_dielectric(ctx, ctx->stage_stream[st.stage], st.count, &s, &q, drx, dry, &o);
    });
}
extern "C" int rls_disney_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_shading_soa *sg, const rls_disney_params *p,
                                               const float *rx_s, const float *ry_s, const float *rx_d, const float *ry_d,
                                               const rls_disney_out *out, size_t chunk)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading(sg) && ok_disney_params(p) && rx_s && ry_s && rx_d && ry_d && ok_disney_out(out),
                "rls_disney_sample_eval_pdf_host: NULL argument");
    return run_staged(ctx, n, chunk, 48, [&](Stager &st) {
        rls_shading_soa s = st.shading(*sg);
        rls_disney_params q = *p;
        q.base_color = st.inp3(p->base_color); q.subsurface = st.in1(p->subsurface); q.metallic = st.in1(p->metallic);
        q.specular = st.in1(p->specular); q.specular_tint = st.in1(p->specular_tint); q.roughness = st.in1(p->roughness);
        q.anisotropic = st.in1(p->anisotropic); q.sheen = st.in1(p->sheen); q.sheen_tint = st.in1(p->sheen_tint);
        q.clearcoat = st.in1(p->clearcoat); q.clearcoat_gloss = st.in1(p->clearcoat_gloss);
        const float *a = st.in(rx_s), *b = st.in(ry_s), *c = st.in(rx_d), *d = st.in(ry_d);
        rls_disney_out o;
        o.wi_s = st.out3(out->wi_s); o.f_s = st.out3(out->f_s); o.pdf_s = st.out(out->pdf_s);
        o.wi_d = st.out3(out->wi_d); o.f_d = st.out3(out->f_d); o.pdf_d = st.out(out->pdf_d); o.flags = st.out(out->flags);
        return launch_disney_sample_eval_pdf(ctx, ctx->stage_stream[st.stage], st.count, &s, &q, a, b, c, d, &o);
    });
}
extern "C" int rls_skin_profile_sample_eval_pdf_host(rls_context *ctx, size_t n, const rls_skin_params *p, const float *rx,
                                                     const rls_profile_out *out, size_t chunk)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_skin_params(p) && rx && ok_profile_out(out), "rls_skin_profile_sample_eval_pdf_host: NULL argument");
    return run_staged(ctx, n, chunk, 16, [&](Stager &st) {
        rls_skin_params q = *p;
        // sss_color is NOT uploaded: the profile maths never reads it (it only feeds the dead `s` of src/rlSss.cpp:22-23),
        // and it was 12 of the 28 input bytes per sample
        q.sss_color.array.x = q.sss_color.array.y = q.sss_color.array.z = nullptr;
        q.sss_scatter_dist = st.inp3(p->sss_scatter_dist);
        q.sss_dist_multiplier = st.in1(p->sss_dist_multiplier);
        const float *drx = st.in(rx);
        rls_profile_out o; o.r = st.out(out->r); o.pdf = st.out(out->pdf); o.Rd = st.out3(out->Rd); o.flags = st.out(out->flags);
        return launch_skin_profile(ctx, ctx->stage_stream[st.stage], st.count, &q, drx, &o);
    });
}

// ---- the same three fused units with compact frames (rls_shading_quat_soa): 4 + 3 floats (+ 1 byte) of shading per sample
static inline bool ok_shading_quat(const rls_shading_quat_soa *s) { return s && s->qx && s->qy && s->qz && s->qw && has3(s->wo); }
extern "C" int rls_ggx_sample_eval_pdf_hostq(rls_context *ctx, size_t n, const rls_shading_quat_soa *sq, const rls_ggx_params *p,
                                             const float *rx, const float *ry, const rls_bsdf_out *out, size_t chunk)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading_quat(sq) && ok_ggx_params(p) && rx && ry && ok_bsdf_out(out), "rls_ggx_sample_eval_pdf_hostq: NULL argument");
    return run_staged(ctx, n, chunk, 36, [&](Stager &st) {
        rls_shading_soa s;
        int rc = st.shading_from_quat(*sq, &s);
        if (rc != RLS_OK) return rc;
        rls_ggx_params q = *p;
        q.KsColor = st.inp3(p->KsColor); q.specularRoughness = st.in1(p->specularRoughness);
        q.ior = st.in1(p->ior); q.anisotropic = st.in1(p->anisotropic);
        const float *drx = st.in(rx), *dry = st.in(ry);
        rls_bsdf_out o; o.wi = st.out3(out->wi); o.f = st.out3(out->f); o.pdf = st.out(out->pdf);
        o.fresnel = st.out(out->fresnel); o.flags = st.out(out->flags);
        return launch_ggx_sample_eval_pdf(ctx, ctx->stage_stream[st.stage], st.count, &s, &q, drx, dry, &o);
    });
}
extern "C" int rls_ggx_dielectric_sample_eval_pdf_hostq(rls_context *ctx, size_t n, const rls_shading_quat_soa *sq,
                                                        const rls_ggx_params *p, const float *rx, const float *ry,
                                                        const rls_ggx_dielectric_out *out, size_t chunk)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading_quat(sq) && ok_ggx_params(p) && rx && ry && ok_dielectric_out(out),
                "rls_ggx_dielectric_sample_eval_pdf_hostq: NULL argument");
    return run_staged(ctx, n, chunk, 40, [&](Stager &st) {
        rls_shading_soa s;
        int rc = st.shading_from_quat(*sq, &s);
        if (rc != RLS_OK) return rc;
        rls_ggx_params q = *p;
        q.KsColor = st.inp3(p->KsColor); q.specularRoughness = st.in1(p->specularRoughness);
        q.ior = st.in1(p->ior); q.anisotropic = st.in1(p->anisotropic);
        const float *drx = st.in(rx), *dry = st.in(ry);
        rls_ggx_dielectric_out o;
        o.fresnel = st.out(out->fresnel); o.wi_r = st.out3(out->wi_r); o.f_r = st.out(out->f_r); o.pdf_r = st.out(out->pdf_r);
        o.wi_t = st.out3(out->wi_t); o.f_t = st.out(out->f_t); o.weight_t = st.out(out->weight_t); o.flags = st.out(out->flags);
        return launch_ggx_dielectric(ctx, ctx->stage_stream[st.stage], st.count, &s, &q, drx, dry, &o);
    });
}
extern "C" int rls_disney_sample_eval_pdf_hostq(rls_context *ctx, size_t n, const rls_shading_quat_soa *sq, const rls_disney_params *p,
                                                const float *rx_s, const float *ry_s, const float *rx_d, const float *ry_d,
                                                const rls_disney_out *out, size_t chunk)
{
    if (!ctx) return RLS_ERR_INVALID_ARGUMENT;
    if (n == 0) return RLS_OK;
    if (n >> 32) return fail(ctx, RLS_ERR_INVALID_ARGUMENT, "n must be below 2^32 samples per call");
    RLS_REQUIRE(ctx, ok_shading_quat(sq) && ok_disney_params(p) && rx_s && ry_s && rx_d && ry_d && ok_disney_out(out),
                "rls_disney_sample_eval_pdf_hostq: NULL argument");
    return run_staged(ctx, n, chunk, 52, [&](Stager &st) {
        rls_shading_soa s;
        int rc = st.shading_from_quat(*sq, &s);
        if (rc != RLS_OK) return rc;
        rls_disney_params q = *p;
        q.base_color = st.inp3(p->base_color); q.subsurface = st.in1(p->subsurface); q.metallic = st.in1(p->metallic);
        q.specular = st.in1(p->specular); q.specular_tint = st.in1(p->specular_tint); q.roughness = st.in1(p->roughness);
        q.anisotropic = st.in1(p->anisotropic); q.sheen = st.in1(p->sheen); q.sheen_tint = st.in1(p->sheen_tint);
        q.clearcoat = st.in1(p->clearcoat); q.clearcoat_gloss = st.in1(p->clearcoat_gloss);
        const float *a = st.in(rx_s), *b = st.in(ry_s), *c = st.in(rx_d), *d = st.in(ry_d);
        rls_disney_out o;
        o.wi_s = st.out3(out->wi_s); o.f_s = st.out3(out->f_s); o.pdf_s = st.out(out->pdf_s);
        o.wi_d = st.out3(out->wi_d); o.f_d = st.out3(out->f_d); o.pdf_d = st.out(out->pdf_d); o.flags = st.out(out->flags);
        return launch_disney_sample_eval_pdf(ctx, ctx->stage_stream[st.stage], st.count, &s, &q, a, b, c, d, &o);
    });
}
