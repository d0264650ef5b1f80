// rls_multi.cu -- ONE process, SEVERAL devices: the multi-GPU half of the host driver that replaces Arnold's bucket /
// thread parallelism around shader_evaluate (reference src/rlGgx.cpp:248-327).
//
//   * configs 1-4 shard by contiguous index ranges with no exchange: rls_multi hands out one rls_context (own stream) per
//     device; the caller enqueues each device's slice through the ordinary entry points and times the slowest device
//     with rls_multi_timer_begin / _end (CUDA events on every device's stream).
//   * config 5 (albedo sweep): the spp range of every cell is split over the devices, every device writes its partial
//     table, and ONE ncclAllReduce(sum, FP64) per device combines them in place.  Single process, ncclCommInitAll,
//     one stream per device; sweep kernel(s) + all-reduce are captured into one CUDA graph per device on the second
//     identical call and replayed from then on (the table is 2.6 MB: the collective is latency bound, so what matters
//     is not paying a host round trip between the kernel and the collective).
//
// NCCL is not a link-time dependency: libnccl.so.2 is dlopen()ed at the first collective (the copy already loaded into
// the process -- torch's -- if there is one, else the path given to rls_multi_set_nccl_library, else the system's).  Compiled with the bit-exact flags
// like everything but rls_tol.cu; this file contains no kernels.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/rls_b200.h"

namespace {

// The subset of nccl.h this file needs (types and enum values are part of NCCL's stable ABI).
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;                      // ncclSuccess == 0
enum { kNcclFloat64 = 8, kNcclSum = 0 };       // ncclDouble / ncclFloat64 = 8, ncclSum = 0 (nccl.h ncclDataType_t, ncclRedOp_t)
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok() const { return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd && GetErrorString; }
};

std::string g_nccl_path;        // rls_multi_set_nccl_library

bool load_nccl(NcclApi &api, std::string &why)
{
    const char *cands[] = { g_nccl_path.c_str(), "libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so" };
    // a copy already mapped into the process (torch's) wins: both sides of the process then share one NCCL
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    for (const char *c : cands) {
        if (h) break;
        if (c && *c) h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) { why = std::string("cannot load libnccl.so.2 (name it with rls_multi_set_nccl_library): ") + (dlerror() ? dlerror() : "not found"); return false; }
    api.lib = h;
#define RLS_SYM(field, name) *(void **)(&api.field) = dlsym(h, name)
    RLS_SYM(CommInitAll, "ncclCommInitAll"); RLS_SYM(CommDestroy, "ncclCommDestroy"); RLS_SYM(AllReduce, "ncclAllReduce");
    RLS_SYM(GroupStart, "ncclGroupStart"); RLS_SYM(GroupEnd, "ncclGroupEnd"); RLS_SYM(GetErrorString, "ncclGetErrorString");
    RLS_SYM(GetVersion, "ncclGetVersion");
#undef RLS_SYM
    if (!api.ok()) { why = "libnccl.so.2 lacks a required symbol"; return false; }
    return true;
}

struct SweepKey {
    rls_sweep_grid grid; uint64_t seed; uint32_t spp; int flags; std::vector<double *> tables; std::vector<int> arith;
    bool operator==(const SweepKey &o) const
    {
        return memcmp(&grid, &o.grid, sizeof(grid)) == 0 && seed == o.seed && spp == o.spp && flags == o.flags &&
               tables == o.tables && arith == o.arith;
    }
};

} // namespace

struct rls_multi {
    std::vector<int> devices;
    std::vector<rls_context *> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev0, ev1;
    NcclApi nccl;
    std::vector<ncclComm_t> comm;
    bool nccl_tried = false;
    // graph cache of the sweep: captured on the second call with the same arguments
    SweepKey last_key; int same_key_calls = 0; bool graph_failed = false;
    std::vector<cudaGraphExec_t> graph;
    uint64_t graph_replays = 0;
    std::string err;
};

static thread_local std::string g_multi_init_error;
static int mfail(rls_multi *m, int code, const std::string &msg) { if (m) m->err = msg; else g_multi_init_error = msg; return code; }
#define RLS_MCUDA(m, call) \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return mfail(m, RLS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" int rls_multi_set_nccl_library(const char *path) { g_nccl_path = path ? path : ""; return RLS_OK; }
extern "C" const char *rls_multi_last_error_string(const rls_multi *m) { return m ? m->err.c_str() : g_multi_init_error.c_str(); }

static void drop_graphs(rls_multi *m)
{
    for (size_t k = 0; k < m->graph.size(); k++)
        if (m->graph[k]) { cudaSetDevice(m->devices[k]); cudaGraphExecDestroy(m->graph[k]); }
    m->graph.clear();
}

extern "C" int rls_multi_shutdown(rls_multi *m)
{
    if (!m) return RLS_ERR_INVALID_ARGUMENT;
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t k = 0; k < m->ctx.size(); k++) if (m->ctx[k]) rls_synchronize(m->ctx[k]);
    drop_graphs(m);
    for (size_t k = 0; k < m->comm.size(); k++) if (m->comm[k] && m->nccl.CommDestroy) m->nccl.CommDestroy(m->comm[k]);
    for (size_t k = 0; k < m->ctx.size(); k++) {
        cudaSetDevice(m->devices[k]);
        if (k < m->ev0.size() && m->ev0[k]) cudaEventDestroy(m->ev0[k]);
        if (k < m->ev1.size() && m->ev1[k]) cudaEventDestroy(m->ev1[k]);
        if (m->ctx[k]) rls_shutdown(m->ctx[k]);
    }
    cudaSetDevice(prev);
    delete m;
    return RLS_OK;
}

extern "C" int rls_multi_init(int n_devices, const int *devices, rls_multi **out)
{
    if (!out) return mfail(nullptr, RLS_ERR_INVALID_ARGUMENT, "rls_multi_init: out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return mfail(nullptr, RLS_ERR_NO_DEVICE, "rls_multi_init: no CUDA device; this library has no CPU path");
    if (n_devices <= 0) n_devices = count;                       // 0 = every visible device
    if (n_devices > count) return mfail(nullptr, RLS_ERR_INVALID_ARGUMENT, "rls_multi_init: more devices requested than visible");
    rls_multi *m = new (std::nothrow) rls_multi();
    if (!m) return mfail(nullptr, RLS_ERR_OUT_OF_MEMORY, "rls_multi_init: host allocation failed");
    int prev = 0;
    cudaGetDevice(&prev);
    for (int k = 0; k < n_devices; k++) {
        const int dev = devices ? devices[k] : k;
        rls_context *c = nullptr;
        const int rc = rls_init(dev, nullptr, &c);               // own non-blocking stream
        if (rc != RLS_OK) {
            const std::string why = rls_last_error_string(nullptr);
            rls_multi_shutdown(m);
            cudaSetDevice(prev);
            return mfail(nullptr, rc, "rls_multi_init: device " + std::to_string(dev) + ": " + why);
        }
        m->devices.push_back(dev);
        m->ctx.push_back(c);
        m->stream.push_back((cudaStream_t)rls_stream(c));
        cudaSetDevice(dev);
        cudaEvent_t a = nullptr, b = nullptr;
        cudaEventCreate(&a); cudaEventCreate(&b);
        m->ev0.push_back(a); m->ev1.push_back(b);
    }
    cudaSetDevice(prev);
    *out = m;
    return RLS_OK;
}

extern "C" int rls_multi_device_count(const rls_multi *m) { return m ? (int)m->ctx.size() : 0; }
extern "C" rls_context *rls_multi_context(rls_multi *m, int k) { return (m && k >= 0 && k < (int)m->ctx.size()) ? m->ctx[k] : nullptr; }
extern "C" uint64_t rls_multi_graph_replays(const rls_multi *m) { return m ? m->graph_replays : 0; }

extern "C" int rls_multi_synchronize(rls_multi *m)
{
    if (!m) return RLS_ERR_INVALID_ARGUMENT;
    for (rls_context *c : m->ctx) { const int rc = rls_synchronize(c); if (rc != RLS_OK) return mfail(m, rc, rls_last_error_string(c)); }
    return RLS_OK;
}

extern "C" int rls_multi_timer_begin(rls_multi *m)
{
    if (!m) return RLS_ERR_INVALID_ARGUMENT;
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t k = 0; k < m->ctx.size(); k++) { cudaSetDevice(m->devices[k]); RLS_MCUDA(m, cudaEventRecord(m->ev0[k], m->stream[k])); }
    cudaSetDevice(prev);
    return RLS_OK;
}
// Waits for every device; out_ms_max = the slowest device's elapsed time (the job's time), out_ms_per_device optional.
extern "C" int rls_multi_timer_end(rls_multi *m, float *out_ms_max, float *out_ms_per_device)
{
    if (!m || !out_ms_max) return RLS_ERR_INVALID_ARGUMENT;
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t k = 0; k < m->ctx.size(); k++) { cudaSetDevice(m->devices[k]); RLS_MCUDA(m, cudaEventRecord(m->ev1[k], m->stream[k])); }
    float mx = 0.0f;
    for (size_t k = 0; k < m->ctx.size(); k++) {
        cudaSetDevice(m->devices[k]);
        RLS_MCUDA(m, cudaEventSynchronize(m->ev1[k]));
        float ms = 0.0f;
        RLS_MCUDA(m, cudaEventElapsedTime(&ms, m->ev0[k], m->ev1[k]));
        if (out_ms_per_device) out_ms_per_device[k] = ms;
        if (ms > mx) mx = ms;
    }
    cudaSetDevice(prev);
    *out_ms_max = mx;
    return RLS_OK;
}

static int ensure_nccl(rls_multi *m)
{
    if (!m->comm.empty()) return RLS_OK;
    if (m->nccl_tried) return mfail(m, RLS_ERR_NCCL, m->err.empty() ? "NCCL unavailable" : m->err);
    m->nccl_tried = true;
    std::string why;
    if (!load_nccl(m->nccl, why)) return mfail(m, RLS_ERR_NCCL, why);
    m->comm.assign(m->ctx.size(), nullptr);
    const ncclResult_t r = m->nccl.CommInitAll(m->comm.data(), (int)m->devices.size(), m->devices.data());
    if (r != 0) { m->comm.clear(); return mfail(m, RLS_ERR_NCCL, std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(r)); }
    return RLS_OK;
}

extern "C" int rls_multi_partition(uint64_t total, int parts, int k, uint64_t *out_begin, uint64_t *out_end)
{
    if (parts <= 0 || k < 0 || k >= parts || !out_begin || !out_end) return RLS_ERR_INVALID_ARGUMENT;
    // total * k can exceed 64 bits for totals near 2^64: split into quotient and remainder of total / parts
    const uint64_t q = total / (uint64_t)parts, r = total % (uint64_t)parts;
    *out_begin = q * (uint64_t)k + r * (uint64_t)k / (uint64_t)parts;
    *out_end = q * (uint64_t)(k + 1) + r * (uint64_t)(k + 1) / (uint64_t)parts;
    return RLS_OK;
}

// Enqueues, on every device's stream: its share of the sweep, then (unless RLS_MULTI_NO_REDUCE) the all-reduce.
static int enqueue_sweep(rls_multi *m, const rls_sweep_grid *grid, uint64_t seed, uint32_t spp, double *const *tables, int flags)
{
    const int G = (int)m->ctx.size();
    const size_t count = (size_t)grid->n_rough * grid->n_cos * grid->n_ior * RLS_SWEEP_VALUES_PER_CELL;
    for (int k = 0; k < G; k++) {
        uint64_t k0 = 0, k1 = 0;
        rls_multi_partition(spp, G, k, &k0, &k1);
        const int rc = rls_albedo_sweep(m->ctx[k], grid, seed, (uint32_t)k0, (uint32_t)k1, tables[k]);
        if (rc != RLS_OK) return mfail(m, rc, rls_last_error_string(m->ctx[k]));
    }
    if (G > 1 && !(flags & RLS_MULTI_NO_REDUCE)) {
        ncclResult_t r = m->nccl.GroupStart();
        for (int k = 0; k < G && r == 0; k++)
            r = m->nccl.AllReduce(tables[k], tables[k], count, kNcclFloat64, kNcclSum, m->comm[k], m->stream[k]);
        const ncclResult_t e = m->nccl.GroupEnd();
        if (r == 0) r = e;
        if (r != 0) return mfail(m, RLS_ERR_NCCL, std::string("ncclAllReduce: ") + m->nccl.GetErrorString(r));
    }
    return RLS_OK;
}

extern "C" int rls_multi_albedo_sweep(rls_multi *m, const rls_sweep_grid *grid, uint64_t seed, uint32_t spp,
                                      double *const *tables, int flags)
{
    if (!m) return RLS_ERR_INVALID_ARGUMENT;
    if (!grid || !tables) return mfail(m, RLS_ERR_INVALID_ARGUMENT, "rls_multi_albedo_sweep: NULL argument");
    const int G = (int)m->ctx.size();
    for (int k = 0; k < G; k++) if (!tables[k]) return mfail(m, RLS_ERR_INVALID_ARGUMENT, "rls_multi_albedo_sweep: NULL table");
    if (G > 1 && !(flags & RLS_MULTI_NO_REDUCE)) { const int rc = ensure_nccl(m); if (rc != RLS_OK) return rc; }
    int prev = 0;
    cudaGetDevice(&prev);

    SweepKey key; key.grid = *grid; key.seed = seed; key.spp = spp; key.flags = flags;
    key.tables.assign(tables, tables + G);
    for (int k = 0; k < G; k++) key.arith.push_back(rls_get_arith_policy(m->ctx[k]));
    if (key == m->last_key) m->same_key_calls++; else { m->last_key = key; m->same_key_calls = 1; drop_graphs(m); m->graph_failed = false; }

    const bool want_graph = !(flags & RLS_MULTI_NO_GRAPH) && !m->graph_failed;
    if (want_graph && !m->graph.empty()) {                     // replay
        for (int k = 0; k < G; k++) { cudaSetDevice(m->devices[k]); RLS_MCUDA(m, cudaGraphLaunch(m->graph[k], m->stream[k])); }
        m->graph_replays++;
        cudaSetDevice(prev);
        return RLS_OK;
    }
    if (want_graph && m->same_key_calls >= 2) {                // capture (the first call ran plainly: every lazy allocation is done)
        bool ok = true;
        int begun = 0;
        for (int k = 0; k < G && ok; k++) {
            cudaSetDevice(m->devices[k]);
            ok = cudaStreamBeginCapture(m->stream[k], cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) begun++;
        }
        int rc = ok ? enqueue_sweep(m, grid, seed, spp, tables, flags) : RLS_ERR_CUDA;
        std::vector<cudaGraph_t> g(G, nullptr);
        for (int k = 0; k < begun; k++) {
            cudaSetDevice(m->devices[k]);
            if (cudaStreamEndCapture(m->stream[k], &g[k]) != cudaSuccess) { ok = false; g[k] = nullptr; }
        }
        if (ok && rc == RLS_OK) {
            m->graph.assign(G, nullptr);
            for (int k = 0; k < G && ok; k++) {
                cudaSetDevice(m->devices[k]);
                ok = cudaGraphInstantiate(&m->graph[k], g[k], 0) == cudaSuccess;
            }
        } else {
            ok = false;
        }
        for (int k = 0; k < G; k++) if (g[k]) cudaGraphDestroy(g[k]);
        if (!ok) {                                             // this NCCL / driver cannot capture it: run plainly from now on
            cudaGetLastError();
            drop_graphs(m);
            m->graph_failed = true;
        } else {
            for (int k = 0; k < G; k++) { cudaSetDevice(m->devices[k]); RLS_MCUDA(m, cudaGraphLaunch(m->graph[k], m->stream[k])); }
            m->graph_replays++;
            cudaSetDevice(prev);
            return RLS_OK;
        }
    }
    const int rc = enqueue_sweep(m, grid, seed, spp, tables, flags);
    cudaSetDevice(prev);
    return rc;
}
