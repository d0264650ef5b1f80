// rls_driver.cpp -- C++ host driver: stands where Arnold's shader_evaluate / BRDF-callback
// loop stood (reference src/rlGgx.cpp:248-327, src/rlDisney.cpp:677-729).  It owns the batch:
// builds explicit shading frames, view vectors, per-sample node parameters and uniform pairs
// in pinned host memory, calls the batched entry points, and integrates f / pdf the way
// AiBRDFIntegrate's caller would (a white-furnace style estimator), printing samples/s.
//
//   g++ -O2 -std=c++14 -o rls_driver rls_driver.cpp -L.. -lrls_b200 -Wl,-rpath,'$ORIGIN/..'
//   ./rls_driver [ggx|dielectric|disney|skin] [log2(samples)] [device]          host buffers, one device
//   ./rls_driver [ggx_q|dielectric_q|disney_q] ...                              the same with compact frames (quaternions)
//   ./rls_driver --gpus N [--policy fast|exact|tolerant] [--reps R] dielectric|disney|skin|sweep [log2(samples per GPU)]
//       ONE process, N devices (rls_multi): device-resident slices of the index-addressed synthetic stream, every device
//       launched then synchronised, throughput = samples / slowest device's time; `sweep` = config 5 with the spp range
//       sharded over the devices and the NCCL all-reduce (+ CUDA graph) behind rls_multi_albedo_sweep; its optional
//       argument is log2(spp) (default 12; below 10 the grid shrinks to 16 x 16 x 4 cells: sanitizer runs).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "rls_host.hpp"

using namespace rls::host;

namespace {

// Same counter hash as the device generators (rls_synth_uniform): uniforms in [2^-24, 1-2^-24].
inline float uniform24(uint64_t seed, uint32_t stream, uint64_t index)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1ull) + 0xD1B54A32D192ED03ull * (uint64_t)stream;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    uint32_t k = (uint32_t)(z >> 40);
    if (k == 0u) k = 1u;
    return (float)k * 5.9604644775390625e-8f;
}
void fill(float *a, size_t n, uint64_t seed, uint32_t stream, float lo, float hi)
{
    for (size_t i = 0; i < n; i++) a[i] = lo + (hi - lo) * uniform24(seed, stream, i);
}
struct V { float x, y, z; };
inline V cross(V a, V b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline V norm(V a) { float l = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return { a.x / l, a.y / l, a.z / l }; }

// Explicit frames + view vectors: N uniform on the sphere, cos(theta_v) ~ U[0.02, 1].
rls_shading_soa make_shading(Arena &arena, size_t n, uint64_t seed, float backfacing_fraction)
{
    rls_vec3 U = arena.vec3(n), Vv = arena.vec3(n), N = arena.vec3(n), W = arena.vec3(n);
    uint8_t *bf = backfacing_fraction > 0 ? arena.bytes(n) : nullptr;
    const float twoPi = 6.28318530717958647692f;
    for (size_t i = 0; i < n; i++) {
        float nz = 1.0f - 2.0f * uniform24(seed, 10, i), rn = std::sqrt(std::fmax(0.0f, 1.0f - nz * nz));
        float pn = twoPi * uniform24(seed, 11, i);
        V nn = norm({ rn * std::cos(pn), rn * std::sin(pn), nz });
        V a = std::fabs(nn.x) < 0.9f ? V{ 1, 0, 0 } : V{ 0, 1, 0 };
        float d = a.x * nn.x + a.y * nn.y + a.z * nn.z;
        V t = norm({ a.x - nn.x * d, a.y - nn.y * d, a.z - nn.z * d });
        V b = cross(nn, t);
        float pt = twoPi * uniform24(seed, 12, i), ct = std::cos(pt), st = std::sin(pt);
        V u = norm({ t.x * ct + b.x * st, t.y * ct + b.y * st, t.z * ct + b.z * st });
        V v = cross(nn, u);
        float cz = 0.02f + 0.98f * uniform24(seed, 13, i), sr = std::sqrt(std::fmax(0.0f, 1.0f - cz * cz));
        float pv = twoPi * uniform24(seed, 14, i), cv = sr * std::cos(pv), sv = sr * std::sin(pv);
        V w = norm({ u.x * cv + v.x * sv + nn.x * cz, u.y * cv + v.y * sv + nn.y * cz, u.z * cv + v.z * sv + nn.z * cz });
        U.x[i] = u.x; U.y[i] = u.y; U.z[i] = u.z; Vv.x[i] = v.x; Vv.y[i] = v.y; Vv.z[i] = v.z;
        N.x[i] = nn.x; N.y[i] = nn.y; N.z[i] = nn.z; W.x[i] = w.x; W.y[i] = w.y; W.z[i] = w.z;
        if (bf) bf[i] = uniform24(seed, 15, i) < backfacing_fraction;
    }
    rls_shading_soa sg = { as_const(U), as_const(Vv), as_const(N), as_const(W), bf };
    return sg;
}

// The same batch with compact frames: q per sample (rls_host.hpp quaternion_from_frame), wo and backfacing shared.
rls_shading_quat_soa to_quat(Arena &arena, size_t n, const rls_shading_soa &sg)
{
    float *q[4] = { arena.floats(n), arena.floats(n), arena.floats(n), arena.floats(n) };
    for (size_t i = 0; i < n; i++) {
        const float U[3] = { sg.U.x[i], sg.U.y[i], sg.U.z[i] }, V[3] = { sg.V.x[i], sg.V.y[i], sg.V.z[i] }, N[3] = { sg.N.x[i], sg.N.y[i], sg.N.z[i] };
        float o[4];
        quaternion_from_frame(U, V, N, o);
        for (int k = 0; k < 4; k++) q[k][i] = o[k];
    }
    rls_shading_quat_soa sq = { q[0], q[1], q[2], q[3], sg.wo, sg.backfacing };
    return sq;
}

double seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

} // namespace

// ------------------------------------------------------------------ one process, several devices
namespace {

struct DeviceArena {                       // device arrays of one context, freed together
    rls_context *ctx;
    std::vector<void *> blocks;
    explicit DeviceArena(rls_context *c) : ctx(c) {}
    ~DeviceArena() { for (void *p : blocks) rls_device_free(ctx, p); }
    void *raw(size_t bytes)
    {
        void *p = nullptr;
        if (rls_device_alloc(ctx, bytes, &p) != RLS_OK) throw Error(rls_last_error_string(ctx));
        blocks.push_back(p);
        return p;
    }
    float *floats(size_t n) { return (float *)raw(n * sizeof(float)); }
    uint32_t *words(size_t n) { return (uint32_t *)raw(n * sizeof(uint32_t)); }
    rls_vec3 vec3(size_t n) { rls_vec3 v = { floats(n), floats(n), floats(n) }; return v; }
};
void must(rls_context *c, int rc) { if (rc != RLS_OK) throw Error(rls_last_error_string(c)); }
float *synth(DeviceArena &a, size_t n, uint64_t seed, uint32_t stream, uint64_t first, float lo, float hi)
{
    float *p = a.floats(n);
    must(a.ctx, rls_synth_uniform(a.ctx, n, seed, stream, first, lo, hi, p));
    return p;
}
// mean over the first m samples of a device array pair a / b where valid (flags clear of ZERO_L | BELOW_HORIZON)
double host_mean_ratio(rls_context *c, const float *num, const float *den, const uint32_t *flags, size_t m)
{
    std::vector<float> a(m), b(m);
    std::vector<uint32_t> f(m, 0);
    must(c, rls_memcpy_to_host(c, a.data(), num, m * sizeof(float)));
    must(c, rls_memcpy_to_host(c, b.data(), den, m * sizeof(float)));
    if (flags) must(c, rls_memcpy_to_host(c, f.data(), flags, m * sizeof(uint32_t)));
    double s = 0.0; size_t k = 0;
    for (size_t i = 0; i < m; i++) if (!(f[i] & (RLS_FLAG_ZERO_L | RLS_FLAG_BELOW_HORIZON)) && b[i] > 0) { s += (double)a[i] / b[i]; k++; }
    return k ? s / (double)k : 0.0;
}

int multi_main(int gpus, const std::string &policy, int reps, const std::string &what, int log2n)
{
    rls_multi *m = nullptr;
    if (rls_multi_init(gpus, nullptr, &m) != RLS_OK) {
        std::fprintf(stderr, "rls_driver: %s\n", rls_multi_last_error_string(nullptr));
        return 1;
    }
    const int G = rls_multi_device_count(m);
    const int pol = policy == "tolerant" ? RLS_ARITH_TOLERANT : (policy == "exact" ? RLS_ARITH_EXACT : RLS_ARITH_FAST);
    int status = 0;
    try {
        for (int k = 0; k < G; k++) must(rls_multi_context(m, k), rls_set_arith_policy(rls_multi_context(m, k), pol));
        std::vector<std::unique_ptr<DeviceArena>> arena;
        for (int k = 0; k < G; k++) arena.emplace_back(new DeviceArena(rls_multi_context(m, k)));
        float ms = 0.0f;
        std::vector<float> per(G, 0.0f);
        std::vector<double> check(G, 0.0);
        double total = 0.0;
        if (what == "sweep") {
            // optional argument: log2(spp), default 12 (BASELINE configs[4]: 65 536 cells x 4096 spp); below 2^10 spp the
            // grid shrinks to 16 x 16 x 4 cells as well (sanitizer runs, tools/sanitize.sh)
            const int log2spp = (log2n >= 1 && log2n <= 16) ? log2n : 12;
            const uint32_t spp = 1u << log2spp;
            rls_sweep_grid grid = { 64, 64, 16, 0.02f, 1.0f, 1.0f, 2.5f };
            if (log2spp < 10) { grid.n_rough = 16; grid.n_cos = 16; grid.n_ior = 4; }
            const size_t cells = (size_t)grid.n_rough * grid.n_cos * grid.n_ior, count = cells * RLS_SWEEP_VALUES_PER_CELL;
            std::vector<double *> tables(G);
            for (int k = 0; k < G; k++) tables[k] = (double *)arena[k]->raw(count * sizeof(double));
            for (int w = 0; w < 2; w++)              // the second identical call captures the graphs
                if (rls_multi_albedo_sweep(m, &grid, 0x5EED0005, spp, tables.data(), 0) != RLS_OK) throw Error(rls_multi_last_error_string(m));
            if (rls_multi_synchronize(m) != RLS_OK) throw Error(rls_multi_last_error_string(m));
            rls_multi_timer_begin(m);
            for (int r = 0; r < reps; r++)
                if (rls_multi_albedo_sweep(m, &grid, 0x5EED0005, spp, tables.data(), 0) != RLS_OK) throw Error(rls_multi_last_error_string(m));
            if (rls_multi_timer_end(m, &ms, per.data()) != RLS_OK) throw Error(rls_multi_last_error_string(m));
            total = (double)cells * spp * reps;
            std::vector<double> t(count);
            for (int k = 0; k < G; k++) {           // every device holds the full table: checksum of each column
                must(rls_multi_context(m, k), rls_memcpy_to_host(rls_multi_context(m, k), t.data(), tables[k], count * sizeof(double)));
                double s = 0.0;
                for (size_t i = 0; i < count; i++) s += t[i] * (double)(1 + i % RLS_SWEEP_VALUES_PER_CELL);
                check[k] = s;
            }
        } else {
            const size_t n = (size_t)1 << log2n;
            struct Launch { std::function<void()> go; };
            std::vector<Launch> launch(G);
            std::vector<std::function<double()>> probe(G);
            for (int k = 0; k < G; k++) {
                rls_context *c = rls_multi_context(m, k);
                DeviceArena &a = *arena[k];
                const uint64_t first = (uint64_t)k * n;              // contiguous index ranges of ONE synthetic stream
                const size_t mcheck = n < 65536 ? n : 65536;
                if (what == "dielectric" || what == "disney") {
                    const uint64_t seed = what == "dielectric" ? 0x5EED0002 : 0x5EED0003;
                    rls_vec3 U = a.vec3(n), V = a.vec3(n), N = a.vec3(n), W = a.vec3(n);
                    uint8_t *bf = what == "dielectric" ? (uint8_t *)a.raw(n) : nullptr;
                    rls_shading_soa sg = { as_const(U), as_const(V), as_const(N), as_const(W), bf };
                    must(c, rls_synth_shading(c, n, seed, first, 0.02f, 1.0f, what == "dielectric" ? 0.25f : 0.0f, &sg));
                    if (what == "dielectric") {
                        rls_ggx_params p = ggx_defaults();
                        p.specularRoughness = varying(synth(a, n, seed, 2, first, 0.05f, 1.0f));
                        p.ior = varying(synth(a, n, seed, 3, first, 1.05f, 2.5f));
                        const float *rx = synth(a, n, seed, 0, first, 0, 1), *ry = synth(a, n, seed, 1, first, 0, 1);
                        rls_ggx_dielectric_out o = { a.floats(n), a.vec3(n), a.floats(n), a.floats(n), a.vec3(n), a.floats(n), a.floats(n), a.words(n) };
                        launch[k].go = [=]() { must(c, rls_ggx_dielectric_sample_eval_pdf(c, n, &sg, &p, rx, ry, &o)); };
                        probe[k] = [=]() { return host_mean_ratio(c, o.f_r, o.pdf_r, o.flags, mcheck); };
                    } else {
                        rls_disney_params p = disney_defaults();
                        rls_param1 *scalars[10] = { &p.subsurface, &p.metallic, &p.specular, &p.specular_tint, &p.roughness,
                                                    &p.anisotropic, &p.sheen, &p.sheen_tint, &p.clearcoat, &p.clearcoat_gloss };
                        for (int j = 0; j < 10; j++) *scalars[j] = varying(synth(a, n, seed, 20 + j, first, 0, 1));
                        rls_vec3 base = { synth(a, n, seed, 30, first, 0, 1), synth(a, n, seed, 31, first, 0, 1), synth(a, n, seed, 32, first, 0, 1) };
                        p.base_color.array = as_const(base);
                        const float *u0 = synth(a, n, seed, 0, first, 0, 1), *u1 = synth(a, n, seed, 1, first, 0, 1);
                        const float *u2 = synth(a, n, seed, 2, first, 0, 1), *u3 = synth(a, n, seed, 3, first, 0, 1);
                        rls_disney_out o = { a.vec3(n), a.vec3(n), a.floats(n), a.vec3(n), a.vec3(n), a.floats(n), a.words(n) };
                        launch[k].go = [=]() { must(c, rls_disney_sample_eval_pdf(c, n, &sg, &p, u0, u1, u2, u3, &o)); };
                        probe[k] = [=]() { return host_mean_ratio(c, o.f_d.x, o.pdf_d, nullptr, mcheck); };
                    }
                } else if (what == "skin") {
                    const uint64_t seed = 0x5EED0004;
                    rls_skin_params p = skin_defaults();
                    rls_vec3 dist = { synth(a, n, seed, 50, first, 0.05f, 2.0f), synth(a, n, seed, 51, first, 0.05f, 2.0f), synth(a, n, seed, 52, first, 0.05f, 2.0f) };
                    p.sss_scatter_dist.array = as_const(dist);
                    const float *rx = synth(a, n, seed, 0, first, 0, 1);
                    rls_profile_out o = { a.floats(n), a.floats(n), a.vec3(n), a.words(n) };
                    launch[k].go = [=]() { must(c, rls_skin_profile_sample_eval_pdf(c, n, &p, rx, &o)); };
                    probe[k] = [=]() { return host_mean_ratio(c, o.Rd.x, o.pdf, nullptr, mcheck); };
                } else {
                    throw Error("unknown workload '" + what + "' (dielectric | disney | skin | sweep)");
                }
            }
            for (int w = 0; w < 3; w++) for (int k = 0; k < G; k++) launch[k].go();       // warm-up
            if (rls_multi_synchronize(m) != RLS_OK) throw Error(rls_multi_last_error_string(m));
            rls_multi_timer_begin(m);
            for (int r = 0; r < reps; r++) for (int k = 0; k < G; k++) launch[k].go();   // every device launched, then synchronised
            if (rls_multi_timer_end(m, &ms, per.data()) != RLS_OK) throw Error(rls_multi_last_error_string(m));
            total = (double)n * G * reps;
            for (int k = 0; k < G; k++) check[k] = probe[k]();
        }
        std::printf("{\"workload\": \"%s\", \"gpus\": %d, \"policy\": \"%s\", \"reps\": %d, \"samples_per_s\": %.6g, \"ms_slowest_device\": %.4f, "
                    "\"graph_replays\": %llu, \"ms_per_device\": [", what.c_str(), G, policy.c_str(), reps, total / (ms * 1e-3), ms,
                    (unsigned long long)rls_multi_graph_replays(m));
        for (int k = 0; k < G; k++) std::printf("%s%.4f", k ? ", " : "", per[k]);
        std::printf("], \"check\": [");
        for (int k = 0; k < G; k++) std::printf("%s%.17g", k ? ", " : "", check[k]);
        std::printf("]}\n");
    } catch (const Error &e) {
        std::fprintf(stderr, "rls_driver: %s\n", e.what());
        status = 1;
    }
    rls_multi_shutdown(m);
    return status;
}

} // namespace

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "--gpus") {
        int gpus = argc > 2 ? atoi(argv[2]) : 0, reps = 10, i = 3;
        std::string policy = "fast";
        while (i + 1 < argc && argv[i][0] == '-') {
            if (std::string(argv[i]) == "--policy") policy = argv[i + 1];
            else if (std::string(argv[i]) == "--reps") reps = atoi(argv[i + 1]);
            i += 2;
        }
        const std::string w = i < argc ? argv[i] : "dielectric";
        return multi_main(gpus, policy, reps < 1 ? 1 : reps, w, i + 1 < argc ? atoi(argv[i + 1]) : 24);
    }
    std::string what = argc > 1 ? argv[1] : "dielectric";
    // a trailing "_q" selects compact frames: unit quaternions uploaded, frames decoded on the device (rls_*_hostq)
    const bool quat = what.size() > 2 && what.compare(what.size() - 2, 2, "_q") == 0;
    const std::string label = what;
    if (quat) what.resize(what.size() - 2);
    size_t n = (size_t)1 << (argc > 2 ? atoi(argv[2]) : 22);
    int device = argc > 3 ? atoi(argv[3]) : 0;
    try {
        Context ctx(device);
        Arena arena(ctx);
        const int reps = 3;
        double best = 1e30, estimate = 0.0;
        size_t counted = 0;
        if (what == "ggx" || what == "dielectric") {
            rls_shading_soa sg = make_shading(arena, n, 0x5EED0002, what == "dielectric" ? 0.25f : 0.0f);
            float *rx = arena.floats(n), *ry = arena.floats(n);
            fill(rx, n, 0x5EED0002, 0, 0, 1); fill(ry, n, 0x5EED0002, 1, 0, 1);
            rls_ggx_params p = ggx_defaults();
            if (what == "ggx") {                       // gold fixture, testsuite/mtoa/0002
                p.specularRoughness = uniform(0.3f); p.ior = uniform(0.47f);
                rls_bsdf_out out = { arena.vec3(n), arena.vec3(n), arena.floats(n), nullptr, arena.words(n) };
                GgxSampler s = quat ? GgxSampler(ctx, n, to_quat(arena, n, sg), p) : GgxSampler(ctx, n, sg, p);
                for (int r = 0; r < reps; r++) { double t0 = seconds(); s.sampleEvalPdf(rx, ry, out); best = std::fmin(best, seconds() - t0); }
                for (size_t i = 0; i < n; i++) if (!(out.flags[i] & (RLS_FLAG_ZERO_L | RLS_FLAG_BELOW_HORIZON))) { estimate += out.f.x[i] / out.pdf[i]; counted++; }
            } else {
                float *rough = arena.floats(n), *ior = arena.floats(n);
                fill(rough, n, 0x5EED0002, 2, 0.05f, 1.0f); fill(ior, n, 0x5EED0002, 3, 1.05f, 2.5f);
                p.specularRoughness = varying(rough); p.ior = varying(ior);
                rls_ggx_dielectric_out out = { arena.floats(n), arena.vec3(n), arena.floats(n), arena.floats(n),
                                               arena.vec3(n), arena.floats(n), arena.floats(n), arena.words(n) };
                GgxSampler s = quat ? GgxSampler(ctx, n, to_quat(arena, n, sg), p) : GgxSampler(ctx, n, sg, p);
                for (int r = 0; r < reps; r++) { double t0 = seconds(); s.dielectricSampleEvalPdf(rx, ry, out); best = std::fmin(best, seconds() - t0); }
                for (size_t i = 0; i < n; i++) if (!(out.flags[i] & (RLS_FLAG_ZERO_L | RLS_FLAG_BELOW_HORIZON))) { estimate += out.f_r[i] / out.pdf_r[i]; counted++; }
            }
        } else if (what == "disney") {
            rls_shading_soa sg = make_shading(arena, n, 0x5EED0003, 0.0f);
            float *u[4];
            for (int j = 0; j < 4; j++) { u[j] = arena.floats(n); fill(u[j], n, 0x5EED0003, j, 0, 1); }
            rls_disney_params p = disney_defaults();
            rls_param1 *scalars[10] = { &p.subsurface, &p.metallic, &p.specular, &p.specular_tint, &p.roughness,
                                        &p.anisotropic, &p.sheen, &p.sheen_tint, &p.clearcoat, &p.clearcoat_gloss };
            for (int j = 0; j < 10; j++) { float *a = arena.floats(n); fill(a, n, 0x5EED0003, 20 + j, 0, 1); *scalars[j] = varying(a); }
            rls_vec3 base = arena.vec3(n);
            fill(base.x, n, 0x5EED0003, 30, 0, 1); fill(base.y, n, 0x5EED0003, 31, 0, 1); fill(base.z, n, 0x5EED0003, 32, 0, 1);
            p.base_color.array = as_const(base);
            rls_disney_out out = { arena.vec3(n), arena.vec3(n), arena.floats(n), arena.vec3(n), arena.vec3(n), arena.floats(n), arena.words(n) };
            DisneySampler s = quat ? DisneySampler(ctx, n, to_quat(arena, n, sg), p) : DisneySampler(ctx, n, sg, p);
            for (int r = 0; r < reps; r++) { double t0 = seconds(); s.sampleEvalPdf(u[0], u[1], u[2], u[3], out); best = std::fmin(best, seconds() - t0); }
            for (size_t i = 0; i < n; i++) if (out.pdf_d[i] > 0) { estimate += out.f_d.x[i] / out.pdf_d[i]; counted++; }
        } else if (what == "skin") {
            rls_skin_params p = skin_defaults();
            rls_vec3 color = arena.vec3(n), dist = arena.vec3(n);
            float *cc[3] = { color.x, color.y, color.z }, *dd[3] = { dist.x, dist.y, dist.z };
            for (int j = 0; j < 3; j++) { fill(cc[j], n, 0x5EED0004, 40 + j, 0.05f, 1.0f); fill(dd[j], n, 0x5EED0004, 50 + j, 0.05f, 2.0f); }
            p.sss_color.array = as_const(color); p.sss_scatter_dist.array = as_const(dist);
            float *rx = arena.floats(n);
            fill(rx, n, 0x5EED0004, 0, 0, 1);
            rls_profile_out out = { arena.floats(n), arena.floats(n), arena.vec3(n), arena.words(n) };
            SkinProfile s(ctx, n, p);
            for (int r = 0; r < reps; r++) { double t0 = seconds(); s.sampleEvalPdf(rx, out); best = std::fmin(best, seconds() - t0); }
            for (size_t i = 0; i < n; i++) { estimate += out.Rd.x[i] / out.pdf[i]; counted++; }
        } else {
            std::fprintf(stderr, "unknown workload '%s' (ggx | dielectric | disney | skin, ggx_q | dielectric_q | disney_q)\n", what.c_str());
            return 2;
        }
        std::printf("{\"workload\": \"%s\", \"samples\": %zu, \"host_to_host_samples_per_s\": %.4g, \"mean_f_over_pdf\": %.6f, \"nodes\": [\"%s\", \"%s\", \"%s\"]}\n",
                    label.c_str(), n, (double)n / best, counted ? estimate / (double)counted : 0.0,
                    rls_node_name(0), rls_node_name(1), rls_node_name(2));
    } catch (const Error &e) {
        std::fprintf(stderr, "rls_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
