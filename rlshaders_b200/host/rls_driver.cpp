// rls_driver.cpp -- C++ host driver: stands where Arnold's shader_evaluate / BRDF-callback
// loop stood (reference src/rlGgx.cpp:248-327, src/rlDisney.cpp:677-729).  It owns the batch:
// builds explicit shading frames, view vectors, per-sample node parameters and uniform pairs
// in pinned host memory, calls the batched entry points, and integrates f / pdf the way
// AiBRDFIntegrate's caller would (a white-furnace style estimator), printing samples/s.
//
//   g++ -O2 -std=c++14 -o rls_driver rls_driver.cpp -L.. -lrls_b200 -Wl,-rpath,'$ORIGIN/..'
//   ./rls_driver [ggx|dielectric|disney|skin] [log2(samples)] [device]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

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

double seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

} // namespace

int main(int argc, char **argv)
{
    std::string what = argc > 1 ? argv[1] : "dielectric";
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
                GgxSampler s(ctx, n, sg, p);
                for (int r = 0; r < reps; r++) { double t0 = seconds(); s.sampleEvalPdf(rx, ry, out); best = std::fmin(best, seconds() - t0); }
                for (size_t i = 0; i < n; i++) if (!(out.flags[i] & (RLS_FLAG_ZERO_L | RLS_FLAG_BELOW_HORIZON))) { estimate += out.f.x[i] / out.pdf[i]; counted++; }
            } else {
                float *rough = arena.floats(n), *ior = arena.floats(n);
                fill(rough, n, 0x5EED0002, 2, 0.05f, 1.0f); fill(ior, n, 0x5EED0002, 3, 1.05f, 2.5f);
                p.specularRoughness = varying(rough); p.ior = varying(ior);
                rls_ggx_dielectric_out out = { arena.floats(n), arena.vec3(n), arena.floats(n), arena.floats(n),
                                               arena.vec3(n), arena.floats(n), arena.floats(n), arena.words(n) };
                GgxSampler s(ctx, n, sg, p);
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
            DisneySampler s(ctx, n, sg, p);
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
            std::fprintf(stderr, "unknown workload '%s' (ggx | dielectric | disney | skin)\n", what.c_str());
            return 2;
        }
        std::printf("{\"workload\": \"%s\", \"samples\": %zu, \"host_to_host_samples_per_s\": %.4g, \"mean_f_over_pdf\": %.6f, \"nodes\": [\"%s\", \"%s\", \"%s\"]}\n",
                    what.c_str(), n, (double)n / best, counted ? estimate / (double)counted : 0.0,
                    rls_node_name(0), rls_node_name(1), rls_node_name(2));
    } catch (const Error &e) {
        std::fprintf(stderr, "rls_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
